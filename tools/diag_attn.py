"""GPU diagnostic for the tensor-core attention kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200 import _abi, ops
dev = torch.device("cuda")
torch.manual_seed(0)

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

def ref_attn(q, k, v):
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    w = qh @ kh.transpose(2, 3)
    return (w.softmax(-1) @ vh).permute(0, 2, 1, 3), torch.logsumexp(w, -1)

def run(B, H, Lq, Lk, scale=0.5, bwd=False):
    d = H * 64
    q = (torch.randn(B * Lq, d, device=dev) * scale).bfloat16()
    kv = (torch.randn(B * Lk, 2 * d, device=dev) * scale).bfloat16()
    k, v = kv[:, :d], kv[:, d:]
    shp = ops.attn_shape(B, H, Lq, Lk, 64, False, Lq * d, d, Lk * 2 * d, 2 * d, Lk * 2 * d, 2 * d, Lq * d, d)
    o = torch.zeros(B * Lq, d, device=dev, dtype=torch.bfloat16); lse = torch.zeros(B, H, Lq, device=dev)
    _abi.reset_counters()
    ops.attention_fwd(shp, q, k, v, o, lse)
    torch.cuda.synchronize()
    qf = q.float().view(B, Lq, H, 64).requires_grad_(True); kf = k.float().reshape(B, Lk, H, 64).requires_grad_(True); vf = v.float().reshape(B, Lk, H, 64).requires_grad_(True)
    r, lr = ref_attn(qf, kf, vf)
    e = rel(o.float().view(B, Lq, H, 64), r); el = rel(lse, lr)
    print(f"FWD B={B} H={H} Lq={Lq} Lk={Lk} scale={scale}: o rel {e:.3e} lse rel {el:.3e} {_abi.counters()['attn_tc']} tc", "OK" if e < 2e-2 else "MISMATCH")
    if e >= 2e-2:
        err = (o.float().view(B, Lq, H, 64) - r).abs()
        print("   per-row max err (b0,h0) first 8:", err[0, :8, 0].max(-1).values.tolist(), " rows bad:", int((err.amax(dim=(0, 2, 3)) > 0.05).sum()), "/", Lq)
        print("   out", o.float().view(B, Lq, H, 64)[0, 0, 0, :6].tolist(), "\n   ref", r[0, 0, 0, :6].tolist())
    if bwd:
        do = torch.randn(B * Lq, d, device=dev).bfloat16()
        r.backward(do.float().view(B, Lq, H, 64))
        dq = torch.zeros(B * Lq, d, device=dev, dtype=torch.bfloat16); dkv = torch.zeros(B * Lk, 2 * d, device=dev, dtype=torch.bfloat16)
        delta = torch.zeros(B * H * Lq, device=dev)
        _abi.reset_counters()
        ops.attention_bwd(shp, q, k, v, o, do, lse, delta, dq, dkv[:, :d], dkv[:, d:])
        torch.cuda.synchronize()
        for name, got, want in (("dq", dq.float().view(B, Lq, H, 64), qf.grad), ("dk", dkv[:, :d].float().reshape(B, Lk, H, 64), kf.grad), ("dv", dkv[:, d:].float().reshape(B, Lk, H, 64), vf.grad)):
            e = rel(got, want)
            print(f"   BWD {name} rel {e:.3e}", "OK" if e < 3e-2 else "MISMATCH", _abi.counters())

if __name__ == "__main__":
    bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
    for cfg in [(1, 1, 128, 128), (1, 1, 128, 256), (2, 2, 128, 92), (2, 3, 200, 300), (2, 8, 1500, 1500), (4, 8, 32, 1500)]:
        run(*cfg, bwd=bwd)
    run(1, 2, 256, 512, scale=2.0, bwd=bwd)     # large score range: exercises the lazy rescale
    # timing at the benchmark shape
    B, H, S = 64, 8, 1500
    d = H * 64
    qkv = (torch.randn(B * S, 3 * d, device=dev) * 0.5).bfloat16()
    o = torch.empty(B * S, d, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, S, device=dev)
    shp = ops.attn_shape(B, H, S, S, 64, False, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * d, d)
    for _ in range(3):
        ops.attention_fwd(shp, qkv, qkv[:, d:], qkv[:, 2 * d:], o, lse)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.attention_fwd(shp, qkv, qkv[:, d:], qkv[:, 2 * d:], o, lse)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"fwd B=64 H=8 S=1500: {ms:.3f} ms  {4.0 * B * H * S * S * 64 / ms / 1e9:.1f} TFLOP/s")
    if bwd:
        do = torch.randn(B * S, d, device=dev).bfloat16(); dqkv = torch.empty_like(qkv); delta = torch.empty(B * H * S, device=dev)
        for _ in range(2):
            ops.attention_bwd(shp, qkv, qkv[:, d:], qkv[:, 2 * d:], o, do, lse, delta, dqkv, dqkv[:, d:], dqkv[:, 2 * d:])
        s.record()
        for _ in range(5):
            ops.attention_bwd(shp, qkv, qkv[:, d:], qkv[:, 2 * d:], o, do, lse, delta, dqkv, dqkv[:, d:], dqkv[:, 2 * d:])
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        print(f"bwd B=64 H=8 S=1500: {ms:.3f} ms  {10.0 * B * H * S * S * 64 / ms / 1e9:.1f} TFLOP/s (algorithmic 2.5x fwd)")

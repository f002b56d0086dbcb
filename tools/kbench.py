"""Kernel micro-benchmarks at the encoder shapes of the headline workload (B=64, S=1500, d=512, H=8, F=2048, r=32).

    python tools/kbench.py [attn] [gemm] [ln] [--B 64] [--iters 20]

Every timing is CUDA events on the launching stream after 3 warm-up calls; inputs (hundreds of MB) exceed the L2.  Each
attention variant is first checked against a torch fp32 restatement on a B=2 slice."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from neuspeech1_b200 import ops
from neuspeech1_b200._abi import ACT_DGELU, ACT_GELU

DEV = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False


NCU = False          # --ncu: every target kernel exactly once (no warm-up, no correctness pre-pass): the capture list stays short


def timeit(fn, iters):
    if NCU:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def attn_ref(q, k, v):
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    w = qh @ kh.transpose(2, 3)
    return (w.softmax(-1) @ vh).permute(0, 2, 1, 3), torch.logsumexp(w, dim=-1)


def bench_attn(B, iters, out):
    H, S, Dh = 8, 1500, 64
    d = H * Dh
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(B * S, 3 * d, generator=g) * 0.5).to(DEV, torch.bfloat16)
    do = torch.randn(B * S, d, generator=g).to(DEV, torch.bfloat16)
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    shp = ops.attn_shape(B, H, S, S, Dh, False, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * d, d)
    o = torch.empty(B * S, d, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, S, device=DEV)
    delta = torch.empty(B * H * S, device=DEV)
    dqkv = torch.zeros(B * S, 3 * d, dtype=torch.bfloat16, device=DEV)
    dq, dk, dv = dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:]
    ws = torch.empty(ops.attention_bwd_workspace_bytes(shp) + 1024, dtype=torch.uint8, device=DEV)
    off = (-ws.data_ptr()) % 1024
    ws = ws[off: off + ops.attention_bwd_workspace_bytes(shp)]
    ops.attention_fwd(shp, q, k, v, o, lse)
    if NCU:
        ops.attention_bwd_ws(shp, q, k, v, o, do, lse, delta, dq, dk, dv, ws)
        torch.cuda.synchronize()
        out["attn"] = {}
        return
    # ---- correctness on the first 2 batch entries
    nb = min(B, 2)
    qf = q[: nb * S].float().reshape(nb, S, H, Dh).requires_grad_(True)
    kf = k[: nb * S].float().reshape(nb, S, H, Dh).requires_grad_(True)
    vf = v[: nb * S].float().reshape(nb, S, H, Dh).requires_grad_(True)
    ref, lse_ref = attn_ref(qf, kf, vf)
    ref.backward(do[: nb * S].float().view(nb, S, H, Dh))
    res = {"fwd_o": rel(o[: nb * S].float().view(nb, S, H, Dh), ref), "fwd_lse": rel(lse[:nb], lse_ref)}
    for name, fn in (("bwd2k", lambda: ops.attention_bwd(shp, q, k, v, o, do, lse, delta, dq, dk, dv)),
                     ("bwdfused", lambda: ops.attention_bwd_ws(shp, q, k, v, o, do, lse, delta, dq, dk, dv, ws))):
        dqkv.zero_()
        fn()
        torch.cuda.synchronize()
        res[name + "_dq"] = rel(dq[: nb * S].float().reshape(nb, S, H, Dh), qf.grad)
        res[name + "_dk"] = rel(dk[: nb * S].float().reshape(nb, S, H, Dh), kf.grad)
        res[name + "_dv"] = rel(dv[: nb * S].float().reshape(nb, S, H, Dh), vf.grad)
    fl = 4.0 * B * H * S * S * Dh
    t = timeit(lambda: ops.attention_fwd(shp, q, k, v, o, lse), iters)
    res["fwd_ms"] = t; res["fwd_tflops"] = fl / t / 1e9
    t = timeit(lambda: ops.attention_bwd(shp, q, k, v, o, do, lse, delta, dq, dk, dv), iters)
    res["bwd2k_ms"] = t; res["bwd2k_tflops"] = 2.5 * fl / t / 1e9
    t = timeit(lambda: ops.attention_bwd_ws(shp, q, k, v, o, do, lse, delta, dq, dk, dv, ws), iters)
    res["bwdfused_ms"] = t; res["bwdfused_tflops"] = 2.5 * fl / t / 1e9
    out["attn"] = res


def bench_gemm(B, iters, out):
    S, d, F, r = 1500, 512, 2048, 32
    M = B * S
    g = torch.Generator().manual_seed(1)
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(DEV, torch.bfloat16)
    x512, x2048 = mk(M, d), mk(M, F)
    res = {}

    def run(name, a, w, N, ep_kw=None, a2=None, w2=None, k2=0, outbuf=None):
        o = outbuf if outbuf is not None else torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        ep = ops.epilogue(**(ep_kw or {}))
        fn = lambda: ops.gemm_nt(a, w, o, ep, a2=a2, w2=w2, k2=k2)
        t = timeit(fn, iters)
        fl = 2.0 * M * N * (a.shape[1] + k2)
        res[name] = {"ms": t, "tflops": fl / t / 1e9}

    bias512 = torch.zeros(d, device=DEV); bias2048 = torch.zeros(F, device=DEV); bias1536 = torch.zeros(3 * d, device=DEV)
    t32, t96 = mk(M, r), mk(M, 3 * r)
    z1 = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    res_h = mk(M, d)
    run("qkv+lora", x512, mk(3 * d, d), 3 * d, dict(bias=bias1536, alpha=0.125, alpha_cols=d, a2_group_cols=d), a2=t96, w2=mk(3 * d, r), k2=r)
    run("out+lora+res", x512, mk(d, d), d, dict(bias=bias512, residual=res_h, ldr=d), a2=t32, w2=mk(d, r), k2=r)
    run("fc1+lora+gelu+aux", x512, mk(F, d), F, dict(bias=bias2048, act=ACT_GELU, aux_out=z1, ldaux=F), a2=t32, w2=mk(F, r), k2=r)
    run("fc1+lora+gelu", x512, mk(F, d), F, dict(bias=bias2048, act=ACT_GELU), a2=t32, w2=mk(F, r), k2=r)
    run("fc1 plain", x512, mk(F, d), F)
    run("fc2+lora+res", x2048, mk(d, F), d, dict(bias=bias512, residual=res_h, ldr=d), a2=t32, w2=mk(d, r), k2=r)
    run("dfc2+dgelu", x512, mk(F, d), F, dict(act=ACT_DGELU, aux_in=z1, ldaux=F), a2=t32, w2=mk(F, r), k2=r)
    run("dqkv", mk(M, 3 * d), mk(d, 3 * d), d, None, a2=t96, w2=mk(d, 3 * r), k2=3 * r)
    run("kv_all", x512, mk(6 * 2 * d, d), 6 * 2 * d, dict(bias=torch.zeros(6 * 2 * d, device=DEV)))
    run("lora_t 512->32", x512, mk(r, d), r, dict(alpha=2.0, alpha_cols=r))
    run("lora_t 512->96", x512, mk(3 * r, d), 3 * r, dict(alpha=2.0, alpha_cols=3 * r))
    run("lora_t 2048->32", x2048, mk(r, F), r, dict(alpha=2.0, alpha_cols=r))
    # wgrad
    G = torch.zeros(d * r, device=DEV)
    t = timeit(lambda: ops.gemm_tn(x512, t32, G, r, 1), iters)
    res["wgrad 512x32"] = {"ms": t, "tflops": 2.0 * M * d * r / t / 1e9}
    G2 = torch.zeros(F * r, device=DEV)
    t = timeit(lambda: ops.gemm_tn(x2048, t32, G2, r, 1), iters)
    res["wgrad 2048x32"] = {"ms": t, "tflops": 2.0 * M * F * r / t / 1e9}
    G3 = torch.zeros(r * d, device=DEV)
    t = timeit(lambda: ops.gemm_tn(t32, x512, G3, d, 1), iters)
    res["wgrad 32x512 (dA)"] = {"ms": t, "tflops": 2.0 * M * d * r / t / 1e9}
    # ---- LoRA-branch dropout (finetune.py:210, p = 0.05): bit plane, mask stages, masked second product, q/k/v correction pass
    seed = torch.tensor([1234], dtype=torch.int32, device=DEV)
    bits512 = torch.empty(3, M, d // 32, dtype=torch.int32, device=DEV)
    bits2048 = torch.empty(1, M, F // 32, dtype=torch.int32, device=DEV)
    t = timeit(lambda: ops.dropout_bits(M, d, seed, [11, 22, 33], 0.05, bits512), iters)
    res["dropout_bits 3x(M,512)"] = {"ms": t, "gbs": bits512.numel() * 4 / t / 1e6}
    ops.dropout_bits(M, F, seed, [44], 0.05, bits2048)
    t1 = torch.empty(M, r, dtype=torch.bfloat16, device=DEV); t3 = torch.empty(M, 3 * r, dtype=torch.bfloat16, device=DEV)
    A1, A3, A4 = mk(r, d), mk(3 * r, d), mk(r, F)
    t = timeit(lambda: ops.gemm_nt(x512, A1, t1, ops.epilogue(alpha=2.0, alpha_cols=r, drop_a=bits512[:1])), iters)
    res["lora_t 512->32 masked"] = {"ms": t, "gbs": M * d * 2 / t / 1e6}
    t = timeit(lambda: ops.gemm_nt(x512, A3, t3, ops.epilogue(alpha=2.0, alpha_cols=3 * r, drop_a=bits512)), iters)
    res["lora_t 512->96 masked"] = {"ms": t, "gbs": M * d * 2 / t / 1e6}
    t = timeit(lambda: ops.gemm_nt(x2048, A4, t1, ops.epilogue(alpha=2.0, alpha_cols=r, drop_a=bits2048)), iters)
    res["lora_t 2048->32 masked"] = {"ms": t, "gbs": M * F * 2 / t / 1e6}
    t = timeit(lambda: ops.gemm_tn_masked(x512, t32, G3, 1, d, bits512[0]), iters)
    res["wgrad dA 512 masked"] = {"ms": t, "gbs": M * d * 2 / t / 1e6}
    G4 = torch.zeros(r * F, device=DEV)
    t = timeit(lambda: ops.gemm_tn_masked(x2048, t32, G4, 1, F, bits2048[0]), iters)
    res["wgrad dA 2048 masked"] = {"ms": t, "gbs": M * F * 2 / t / 1e6}
    dz1 = torch.empty(M, F, dtype=torch.bfloat16, device=DEV)
    w2t, a2t = mk(F, d), mk(F, r)
    t = timeit(lambda: ops.gemm_nt(x512, w2t, dz1, ops.epilogue(act=ACT_DGELU, aux_in=z1, ldaux=F, drop_bits=bits2048[0]), a2=t32, w2=a2t, k2=r), iters)
    res["dfc2+dgelu masked product"] = {"ms": t, "tflops": 2.0 * M * F * (d + r) / t / 1e9}
    do_ = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    wot, aot = mk(d, d), mk(d, r)
    t = timeit(lambda: ops.gemm_nt(x512, wot, do_, ops.epilogue(drop_bits=bits512[0]), a2=t32, w2=aot, k2=r), iters)
    res["dout masked product"] = {"ms": t, "tflops": 2.0 * M * d * (d + r) / t / 1e9}
    dA3 = torch.zeros(3 * r, d, device=DEV)
    At3 = mk(d, 3 * r)
    t = timeit(lambda: ops.lora_da(x512, t96, dA3, 3, bits512, dx=do_, At=At3), iters)
    res["lora_da qkv + dx correction"] = {"ms": t, "gbs": M * d * 2 / t / 1e6}
    # ---- round 2, late: the plane drawn in the mask stage (drop_mode 2), the one-pass B-side backward (dt = dy B, dB = dy^T t)
    salts = [11, 22, 33]
    t = timeit(lambda: ops.gemm_nt(x512, A3, t3, ops.epilogue(alpha=2.0, alpha_cols=3 * r, drop_a=bits512, drop_gen=(seed, salts, 0.05))), iters)
    res["lora_t 512->96 masked, plane drawn in the stage"] = {"ms": t, "gbs": M * d * 2 / t / 1e6}
    t = timeit(lambda: ops.gemm_nt(x2048, A4, t1, ops.epilogue(alpha=2.0, alpha_cols=r, drop_a=bits2048, drop_gen=(seed, [44], 0.05))), iters)
    res["lora_t 2048->32 masked, plane drawn in the stage"] = {"ms": t, "gbs": M * F * 2 / t / 1e6}
    for N, Gn in ((d, 1), (d, 3), (F, 1)):
        dy = mk(M, Gn * N); Bt = mk(Gn * r, N); tt = mk(M, Gn * r)
        dtb = torch.empty(M, Gn * r, dtype=torch.bfloat16, device=DEV); dB = torch.zeros(Gn * N, r, device=DEV)
        nws = ops.lora_bwd_b_workspace_bytes(M, N, r, Gn)
        wsb = torch.zeros(nws, dtype=torch.uint8, device=DEV) if nws > 0 else None
        t = timeit(lambda: ops.lora_bwd_b(dy, Bt, tt, dtb, dB, N, r, [2.0] * Gn, [1.0] * Gn, workspace=wsb), iters)
        res[f"lora_bwd_b N={N} groups={Gn}"] = {"ms": t, "gbs": M * Gn * N * 2 / t / 1e6}
    out["gemm"] = res


def bench_head(B, iters, out):
    """loss head: tied vocabulary projection (utils/load_model.py:1047), cross-entropy, dlogits @ E."""
    L, d, V = 32, 512, 51865
    Vp = (V + 15) // 16 * 16
    M = B * L
    g = torch.Generator().manual_seed(2)
    y = (torch.randn(M, d, generator=g)).to(DEV, torch.bfloat16)
    E = (torch.randn(V, d, generator=g) * 0.02).to(DEV, torch.bfloat16)
    Et = torch.zeros(d, Vp, dtype=torch.bfloat16, device=DEV); Et[:, :V] = E.t()
    logits = torch.empty(M, Vp, dtype=torch.bfloat16, device=DEV)
    labels = torch.randint(0, V, (M,), generator=g).to(DEV)
    row_loss = torch.empty(M, device=DEV); loss_sum = torch.zeros(1, device=DEV); nv = torch.zeros(1, dtype=torch.int32, device=DEV)
    dy = torch.empty(M, d, dtype=torch.bfloat16, device=DEV)
    res = {}
    t = timeit(lambda: ops.gemm_nt(y, E, logits, ops.epilogue(), N=V), iters)
    res["logits"] = {"ms": t, "tflops": 2.0 * M * V * d / t / 1e9}
    t = timeit(lambda: ops.cross_entropy(logits, V, labels, row_loss, loss_sum, nv, write_grad=False), iters)
    res["ce_fwd"] = {"ms": t, "gbs": logits.numel() * 2 / t / 1e6}
    t = timeit(lambda: ops.cross_entropy(logits, V, labels, row_loss, None, nv, write_grad=True), iters)
    res["ce_bwd"] = {"ms": t, "gbs": 2 * logits.numel() * 2 / t / 1e6}
    t = timeit(lambda: ops.gemm_nt(logits, Et, dy, ops.epilogue()), iters)
    res["dlogits"] = {"ms": t, "tflops": 2.0 * M * Vp * d / t / 1e9}
    out["head"] = res


def bench_ln(B, iters, out):
    S, d = 1500, 512
    M = B * S
    x = torch.randn(M, d, device=DEV).to(torch.bfloat16)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    gm = torch.ones(d, device=DEV); bt = torch.zeros(d, device=DEV)
    mean = torch.empty(M, device=DEV); rstd = torch.empty(M, device=DEV)
    res = {}
    t = timeit(lambda: ops.layernorm_fwd(x, gm, bt, y, mean, rstd), iters)
    res["fwd"] = {"ms": t, "gbs": 2.0 * M * d * 2 / t / 1e6}
    t = timeit(lambda: ops.layernorm_bwd(y, x, gm, mean, rstd, dx, dres=x), iters)
    res["bwd"] = {"ms": t, "gbs": 4.0 * M * d * 2 / t / 1e6}
    out["ln"] = res


def bench_aug(B, iters, out):
    """K0: (B, C, 6000) fp32 -> (B, 6000, Cp) bf16 channels-last; algorithmic bytes B*C*T*(4 + 2)."""
    C, T = 208, 6000
    g = torch.Generator().manual_seed(5)
    x = (0.3 * torch.randn(B, C, T, generator=g)).clamp_(-1, 1)
    lens = torch.randint(400, 5000, (B,), generator=g)
    for b in range(B):
        x[b, :, int(lens[b]):] = 0
    x = x.to(DEV)
    y = torch.empty(B, T, C, dtype=torch.bfloat16, device=DEV)
    n = lens.to(torch.int32).to(DEV)
    res = {}
    byts = B * C * T * 6.0
    t = timeit(lambda: ops.aug_pass(x, y, 1), iters)
    res["identity, full length"] = {"ms": t, "gbs": byts / t / 1e6}
    t = timeit(lambda: ops.aug_pass(x, y, 1, n=n), iters)
    res["identity, ragged lengths"] = {"ms": t, "gbs": byts / t / 1e6}
    gl = (lens + 39) // 40
    gmax = int(gl.max()) * C
    grid = (torch.rand(B, gmax, generator=g) >= 0.25).to(torch.uint8).to(DEV)
    i32 = lambda v: v.to(torch.int32).to(DEV)
    kw = dict(n=n, flags=i32(torch.ones(B)), grid=grid, grid_stride=gmax, gl=i32(gl), rep_c=i32(torch.ones(B)), rep_t=i32(torch.full((B,), 40)))
    t = timeit(lambda: ops.aug_pass(x, y, 1, **kw), iters)
    res["block mask"] = {"ms": t, "gbs": byts / t / 1e6}
    out["aug"] = res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["attn", "gemm", "ln"])
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--lib", default=None, help="A/B aid: load this build of the shared library instead of the in-tree one")
    ap.add_argument("--ncu", action="store_true", help="launch every target kernel exactly once (for ncu --set full captures)")
    a = ap.parse_args()
    global NCU
    NCU = a.ncu
    if a.lib:
        from neuspeech1_b200 import _abi
        _abi.LIB_PATH = os.path.abspath(a.lib)
    out = {}
    if "attn" in a.what:
        bench_attn(a.B, a.iters, out)
    if "gemm" in a.what:
        bench_gemm(a.B, a.iters, out)
    if "head" in a.what:
        bench_head(a.B, a.iters, out)
    if "ln" in a.what:
        bench_ln(a.B, a.iters, out)
    if "aug" in a.what:
        bench_aug(a.B, a.iters, out)
    s = json.dumps(out, indent=1)
    print(s)
    for fam, d in out.items():   # compact one-line-per-kernel summary (what gets read back from a gpurun tail)
        for k, v in d.items():
            if not isinstance(v, dict):
                if k.endswith("_ms"):
                    print(f"## {fam:6s} {k[:-3]:28s} {v * 1e3:8.1f} us {d.get(k[:-3] + '_tflops', 0):7.0f} TF/s")
                else:
                    if not k.endswith("_tflops"): print(f"## {fam:6s} {k:28s} {v:.3e}")
                continue
            print(f"## {fam:6s} {k:28s} {v['ms'] * 1e3:8.1f} us" + (f" {v['tflops']:7.0f} TF/s" if "tflops" in v else "")
                  + (f" {v['gbs']:7.0f} GB/s" if "gbs" in v else ""))
    if a.out:
        open(a.out, "w").write(s)


if __name__ == "__main__":
    main()

"""GPU diagnostic ladder for the tcgen05 GEMM kernels: prints rel. error per case and, on mismatch, where the error sits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200 import _abi, ops

dev = torch.device("cuda")
torch.manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def describe(out, ref, name):
    err = (out.double() - ref.double()).abs()
    M, N = err.shape
    print(f"   [{name}] max err {float(err.max()):.4g}  ref absmax {float(ref.abs().max()):.4g}  nan {int(torch.isnan(out).sum())}")
    rows = err.max(dim=1).values; cols = err.max(dim=0).values
    badr = (rows > 0.05 * float(ref.abs().max())).nonzero().flatten().tolist()
    badc = (cols > 0.05 * float(ref.abs().max())).nonzero().flatten().tolist()
    print(f"   bad rows {len(badr)}/{M} first {badr[:12]} ; bad cols {len(badc)}/{N} first {badc[:12]}")
    print("   out[0,:8]", out[0, :8].float().tolist()); print("   ref[0,:8]", ref[0, :8].float().tolist())


def nt(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm_nt(a, w, out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    e = rel(out.float(), ref)
    print(f"NT M={M} N={N} K={K}: rel {e:.3e}", "OK" if e < 2e-2 else "MISMATCH")
    if e >= 2e-2:
        describe(out.float(), ref, "nt")
    return e


def tn(M, I, J):
    x = torch.randn(M, I, device=dev).bfloat16(); y = torch.randn(M, J, device=dev).bfloat16()
    g = torch.zeros(I, J, device=dev)
    ops.gemm_tn(x, y, g, J, 1)
    torch.cuda.synchronize()
    ref = x.float().t() @ y.float()
    e = rel(g, ref)
    print(f"TN M={M} I={I} J={J}: rel {e:.3e}", "OK" if e < 2e-2 else "MISMATCH")
    if e >= 2e-2:
        describe(g, ref, "tn")
    return e


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), _abi.load().ns_version())
    _abi.set_path(_abi.PATH_FAST)
    if which in ("all", "nt"):
        for (M, N, K) in [(128, 256, 16), (128, 256, 64), (128, 256, 128), (128, 256, 512), (128, 32, 64), (128, 64, 64), (128, 128, 64),
                          (256, 512, 512), (1000, 512, 512), (96000, 512, 512), (4096, 2048, 512), (4096, 512, 2048)]:
            nt(M, N, K)
    if which in ("all", "tn"):
        for (M, I, J) in [(64, 128, 64), (64, 128, 32), (64, 128, 256), (128, 128, 64), (256, 256, 64), (4096, 512, 32), (4096, 512, 208), (6000, 512, 512)]:
            tn(M, I, J)
    print(_abi.counters())

"""Decode-step cross-attention at B=128, S=1500, Whisper-base: absorbed form (two block-diagonal GEMMs + ns_cross_attention_absorbed)
against single-query attention over cached K|V.  CUDA events, L2 flushed between launches.  -> stdout"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuspeech1_b200 import ops

DEV = torch.device("cuda")
B, S, H, d = 128, 1500, 8, 512
Dh = d // H
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1000.0


q = (torch.randn(B, d, device=DEV) * 0.3).to(bf)
enc = torch.randn(B, S, d, device=DEV).to(bf)
wk_abs = (torch.randn(H * d, Dh, device=DEV) * 0.05).to(bf)
wv = (torch.randn(d, d, device=DEV) * 0.05).to(bf)
bv = torch.randn(d, device=DEV)
qp = torch.empty(B, H, d, dtype=bf, device=DEV); cp = torch.empty(B, H, d, dtype=bf, device=DEV); o = torch.empty(B, d, dtype=bf, device=DEV)
kv = torch.randn(B * S, 2 * d, device=DEV).to(bf)
shp = ops.attn_shape(B, H, 1, S, Dh, False, d, d, S * 2 * d, 2 * d, S * 2 * d, 2 * d, d, d)
print("Q' gemm      %.1f us" % timeit(lambda: ops.gemm_nt(q, wk_abs, qp.view(B, H * d), ops.epilogue(a_group_cols=d), K=Dh)))
print("absorbed attn %.1f us  (%.2f TB/s of encoder rows)" % ((t := timeit(lambda: ops.cross_attention_absorbed(qp, enc, cp))), B * S * d * 2 / t / 1e6))
print("out gemm     %.1f us" % timeit(lambda: ops.gemm_nt(cp.view(B, H * d), wv, o, ops.epilogue(bias=bv, a_group_cols=Dh), K=d)))
print("cached K|V attn %.1f us (%.2f TB/s)" % ((t := timeit(lambda: ops.attention_fwd(shp, q, kv, kv[:, d:], o))), B * S * 2 * d * 2 / t / 1e6))

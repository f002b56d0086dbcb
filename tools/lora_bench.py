"""Times of the LoRA side kernels at the benchmark shapes (M = 64*1500 rows) next to the GEMM-kernel passes they replace.
Developer tool: python tools/lora_bench.py -> gpurun_out/lora_bench.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuspeech1_b200 import ops

DEV = torch.device("cuda")
M, r = 96000, 32


def timeit(fn, n=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / n * 1000.0   # us


def once():
    """One launch of every kernel variant at K=512 (G=1, 3): the ncu target (tools/gpu_r02_d.sh)."""
    seed = torch.tensor([1234], dtype=torch.int32, device=DEV)
    for K, G in ((512, 1), (512, 3)):
        x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        A = (torch.randn(G * r, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
        At = A.t().contiguous()
        t = torch.empty(M, G * r, dtype=torch.bfloat16, device=DEV)
        dt = (torch.randn(M, G * r, device=DEV) * 0.1).to(torch.bfloat16)
        dA = torch.zeros(G * r, K, dtype=torch.float32, device=DEV)
        dx = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        salts = [11, 22, 33][:G]
        bits = torch.empty(G, M, K // 32, dtype=torch.int32, device=DEV)
        ops.dropout_bits(M, K, seed, salts, 0.05, bits)
        ops.lora_down(x, A, t, 2.0, G)
        ops.lora_down(x, A, t, 2.0, G, bits)
        ops.lora_da(x, dt, dA, G)
        ops.lora_da(x, dt, dA, G, bits)
        ops.lora_da(x, dt, dA, G, bits, dx=dx, At=At)
    torch.cuda.synchronize()


def bwd_b_only():
    """python tools/lora_bench.py --bwd-b: the one-pass B-side kernel alone (NS_LB_DEBUG experiments)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for N, G in ((512, 1), (512, 3), (2048, 1)):
        dy = torch.randn(M, G * N, device=DEV).to(torch.bfloat16)
        Bt = (torch.randn(G * r, N, device=DEV) * N ** -0.5).to(torch.bfloat16)
        t = torch.randn(M, G * r, device=DEV).to(torch.bfloat16)
        dt = torch.empty(M, G * r, dtype=torch.bfloat16, device=DEV)
        dB = torch.zeros(G * N, r, dtype=torch.float32, device=DEV)
        nws = ops.lora_bwd_b_workspace_bytes(M, N, r, G)
        wsb = torch.zeros(nws, dtype=torch.uint8, device=DEV) if nws > 0 else None
        us = timeit(lambda: ops.lora_bwd_b(dy, Bt, t, dt, dB, N, r, [2.0] * G, [1.0] * G, workspace=wsb), flush=flush)
        us_hot = timeit(lambda: ops.lora_bwd_b(dy, Bt, t, dt, dB, N, r, [2.0] * G, [1.0] * G, workspace=wsb))
        print(f"bwd_b N={N} G={G} dbg={os.environ.get('NS_LB_DEBUG', '0')}: {us:.1f} us flushed ({M * G * N * 2 / us / 1e6:.2f} TB/s), {us_hot:.1f} us back to back")


def main():
    if "--once" in sys.argv:
        return once()
    if "--bwd-b" in sys.argv:
        return bwd_b_only()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    seed = torch.tensor([1234], dtype=torch.int32, device=DEV)
    out = {}
    for K, G in ((512, 1), (512, 3), (2048, 1)):
        x = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        A = (torch.randn(G * r, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
        At = A.t().contiguous()
        t = torch.empty(M, G * r, dtype=torch.bfloat16, device=DEV)
        dt = (torch.randn(M, G * r, device=DEV) * 0.1).to(torch.bfloat16)
        dA = torch.zeros(G * r, K, dtype=torch.float32, device=DEV)
        dx = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        z = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        salts = [11, 22, 33][:G]
        bits = torch.empty(G, M, K // 32, dtype=torch.int32, device=DEV)
        ops.dropout_bits(M, K, seed, salts, 0.05, bits)
        key = f"K{K}_G{G}"
        out[key] = {
            "dropout_bits": timeit(lambda: ops.dropout_bits(M, K, seed, salts, 0.05, bits), flush=flush),
            "gemm_nt_thin": timeit(lambda: ops.gemm_nt(x, A, t, ops.epilogue(alpha=2.0, alpha_cols=G * r)), flush=flush),
            "lora_down_p0": timeit(lambda: ops.lora_down(x, A, t, 2.0, G), flush=flush),
            "lora_down_p05": timeit(lambda: ops.lora_down(x, A, t, 2.0, G, bits), flush=flush),
            "mask_stage_down_p05": timeit(lambda: ops.gemm_nt(x, A, t, ops.epilogue(alpha=2.0, alpha_cols=G * r, drop_a=bits)), flush=flush),
            "gemm_tn_dA": timeit(lambda: ops.gemm_tn(x, dt, dA, 1, K), flush=flush),
            "mask_stage_dA_p05_per_adapter": timeit(lambda: ops.gemm_tn_masked(x, dt[:, :r], dA[:r], 1, K, bits[0]), flush=flush),
            "lora_da_p0": timeit(lambda: ops.lora_da(x, dt, dA, G), flush=flush),
            "lora_da_p05": timeit(lambda: ops.lora_da(x, dt, dA, G, bits), flush=flush),
            "lora_da_fix_p05": timeit(lambda: ops.lora_da(x, dt, dA, G, bits, dx=dx, At=At), flush=flush),
            "lora_da_fix_gelu_p05": timeit(lambda: ops.lora_da(x, dt, dA, G, bits, dx=dx, At=At, z=z), flush=flush),
            "x_MB": M * K * 2 / 1e6,
        }
        print(key, {k: round(v, 1) for k, v in out[key].items()})
    # B side of the backward: dt = dy B and dB = dy^T t as two passes over dy and as one (ns_lora_bwd_b)
    for N, G in ((512, 1), (512, 3), (2048, 1)):
        dy = torch.randn(M, G * N, device=DEV).to(torch.bfloat16)
        Bt = (torch.randn(G * r, N, device=DEV) * N ** -0.5).to(torch.bfloat16)
        t = torch.randn(M, G * r, device=DEV).to(torch.bfloat16)
        dt = torch.empty(M, G * r, dtype=torch.bfloat16, device=DEV)
        dB = torch.zeros(G * N, r, dtype=torch.float32, device=DEV)
        nws = ops.lora_bwd_b_workspace_bytes(M, N, r, G)
        wsb = torch.zeros(nws, dtype=torch.uint8, device=DEV) if nws > 0 else None
        key = f"bwd_b_N{N}_G{G}"
        ep = ops.epilogue(alpha=2.0, alpha_cols=G * r, a_group_cols=r) if G > 1 else ops.epilogue(alpha=2.0, alpha_cols=r)
        out[key] = {
            "gemm_nt_dt": timeit(lambda: ops.gemm_nt(dy, Bt, dt, ep, K=N), flush=flush),
            "gemm_tn_dB": timeit(lambda: (ops.gemm_tn_grouped(dy, t, dB, N, r, r, 1, [1.0] * G) if G > 1 else ops.gemm_tn(dy, t, dB, r, 1)), flush=flush),
            "lora_bwd_b": timeit(lambda: ops.lora_bwd_b(dy, Bt, t, dt, dB, N, r, [2.0] * G, [1.0] * G, workspace=wsb), flush=flush),
            "dy_MB": M * G * N * 2 / 1e6,
        }
        print(key, {k: round(v, 1) for k, v in out[key].items()})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/lora_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()

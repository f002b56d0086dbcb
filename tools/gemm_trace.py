"""Timeline of the producer / MMA threads of CTAs 0 and 1 of one tcgen05 GEMM launch (developer aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200 import ops
DEV = torch.device("cuda")
M, N, K = 96000, 512, 2048
a = (torch.randn(M, K) * 0.05).to(DEV, torch.bfloat16); w = (torch.randn(N, K) * 0.05).to(DEV, torch.bfloat16)
out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
run = lambda: ops.gemm_nt(a, w, out, ops.epilogue())
run(); torch.cuda.synchronize()
tr = torch.zeros(4 * 512 * 2, dtype=torch.int64, device=DEV)
ops.lib().ns_debug_attn_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
ops.lib().ns_debug_attn_trace(None)
t = tr.cpu().view(4, 512, 2)
names = ["prod0", "mma0", "prod1", "x"]
ev = []
for r in range(3):
    for i in range(512):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0: break
        ev.append((clk, names[r], tag))
t0 = min(e[0] for e in ev)
ev.sort()
for e in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 120]:
    print(f"{e[0]-t0:8d} {e[1]:6s} {e[2]}")
ref = a[:256].float() @ w.float().t()
print("rel err first 256 rows", float((out[:256].float() - ref).norm() / ref.norm()), " last rows", float((out[-300:].float() - a[-300:].float() @ w.float().t()).norm() / (a[-300:].float() @ w.float().t()).norm()))

"""Timeline of epilogue warp 4 (CTA 0) of one tcgen05 GEMM launch next to the MMA thread (developer aid).
usage: gemm_epi_trace.py [plain|gelu|gelu_aux|dgelu] [n_events]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200 import ops, _abi
if os.environ.get("NS_LIB"):      # a build with -DNS_GEMM_EPI_TRACE (the probes are compiled out of the product library)
    _abi.LIB_PATH = os.path.abspath(os.environ["NS_LIB"])
DEV = torch.device("cuda")
kind = sys.argv[1] if len(sys.argv) > 1 else "gelu"
M, N, K = 96000, 2048, 512
a = (torch.randn(M, K) * 0.05).to(DEV, torch.bfloat16); w = (torch.randn(N, K) * 0.05).to(DEV, torch.bfloat16)
bias = torch.randn(N, device=DEV) * 0.1
out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
aux = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
if kind == "plain":
    ep = ops.epilogue(bias=bias)
elif kind == "gelu":
    ep = ops.epilogue(bias=bias, act=ops.ACT_GELU)
elif kind == "gelu_aux":
    ep = ops.epilogue(bias=bias, act=ops.ACT_GELU, aux_out=aux, ldaux=N)
else:
    aux.normal_()
    ep = ops.epilogue(act=ops.ACT_DGELU, aux_in=aux, ldaux=N)
run = lambda: ops.gemm_nt(a, w, out, ep)
run(); torch.cuda.synchronize()
tr = torch.zeros(4 * 512 * 2, dtype=torch.int64, device=DEV)
ops.lib().ns_debug_attn_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
ops.lib().ns_debug_attn_trace(None)
t = tr.cpu().view(4, 512, 2)
names = ["prod0", "mma0", "prod1", "epi"]
ev = []
for r in (1, 3):
    for i in range(512):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0: break
        ev.append((clk, names[r], tag))
t0 = min(e[0] for e in ev)
ev.sort()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
last = {}
print(f"# {kind}")
for e in ev[:n]:
    if e[1] == "mma0" and e[2] % 100 not in (0,):   # only stage-0 MMA events (one per ring revolution) to keep it short
        continue
    d = e[0] - last.get(e[1], e[0]); last[e[1]] = e[0]
    print(f"{e[0]-t0:8d} (+{d:6d}) {e[1]:6s} {e[2]}")

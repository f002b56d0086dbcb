"""Timeline of one CTA of the fused attention backward (developer aid; see ns_debug_attn_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200 import ops

DEV = torch.device("cuda")
B, H, S, Dh = 4, 8, 1500, 64
d = H * Dh
qkv = (torch.randn(B * S, 3 * d) * 0.5).to(DEV, torch.bfloat16)
do = torch.randn(B * S, d).to(DEV, torch.bfloat16)
q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
shp = ops.attn_shape(B, H, S, S, Dh, False, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * d, d)
o = torch.empty(B * S, d, dtype=torch.bfloat16, device=DEV); lse = torch.empty(B, H, S, device=DEV)
delta = torch.empty(B * H * S, device=DEV)
dqkv = torch.zeros(B * S, 3 * d, dtype=torch.bfloat16, device=DEV)
n = ops.attention_bwd_workspace_bytes(shp)
ws = torch.empty(n + 1024, dtype=torch.uint8, device=DEV); off = (-ws.data_ptr()) % 1024; ws = ws[off:off + n]
ops.attention_fwd(shp, q, k, v, o, lse)
run = lambda: ops.attention_bwd_ws(shp, q, k, v, o, do, lse, delta, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], ws)
run(); torch.cuda.synchronize()
tr = torch.zeros(4 * 512 * 2, dtype=torch.int64, device=DEV)
ops.lib().ns_debug_attn_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
ops.lib().ns_debug_attn_trace(None)
t = tr.cpu().view(4, 512, 2)
t0 = int(t[0, 0, 1])
names = ["mma", "dq", "g0", "g1"]
ev = []
for r in range(4):
    for i in range(512):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0:
            break
        ev.append((clk - t0, names[r], tag))
ev.sort()
lim = int(sys.argv[1]) if len(sys.argv) > 1 else 140
for e in ev[:lim]:
    print(f"{e[0]:8d} {e[1]:4s} {e[2]}")
print("... last:", ev[-1])

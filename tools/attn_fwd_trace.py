"""Timeline of one CTA of the ping-pong attention forward (developer aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200 import ops
DEV = torch.device("cuda")
B, H, S, Dh = 4, 8, 1500, 64
d = H * Dh
qkv = (torch.randn(B * S, 3 * d) * 0.5).to(DEV, torch.bfloat16)
q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
shp = ops.attn_shape(B, H, S, S, Dh, False, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * d, d)
o = torch.empty(B * S, d, dtype=torch.bfloat16, device=DEV); lse = torch.empty(B, H, S, device=DEV)
run = lambda: ops.attention_fwd(shp, q, k, v, o, lse)
run(); torch.cuda.synchronize()
tr = torch.zeros(4 * 512 * 2, dtype=torch.int64, device=DEV)
ops.lib().ns_debug_attn_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
ops.lib().ns_debug_attn_trace(None)
t = tr.cpu().view(4, 512, 2)
names = ["mma", "x", "g0", "g1"]
ev = []
for r in range(4):
    for i in range(512):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0: break
        ev.append((clk, names[r], tag))
t0 = min(e[0] for e in ev); ev.sort()
for e in ev: print(f"{e[0]-t0:8d} {e[1]:4s} {e[2]}")

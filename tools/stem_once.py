"""One launch of each stem weight-gradient product at the benchmark shapes (B=64): the ncu target for gemm_tn_kernel with wide Y.
    ncu --set full -k regex:gemm_tn_kernel python tools/stem_once.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuspeech1_b200 import ops

DEV = torch.device("cuda")
B, d = 64, 512
for T, Cp, stride in ((6000, 208, 1), (6000, 512, 2), (3000, 512, 2)):
    x = torch.randn(B, T, Cp, device=DEV).to(torch.bfloat16)
    dz = torch.randn(B, T // stride, d, device=DEV).to(torch.bfloat16)
    dw = torch.zeros(3, d, Cp, dtype=torch.float32, device=DEV)
    db = torch.zeros(d, dtype=torch.float32, device=DEV)
    ops.conv3_wgrad(dz, x, dw, db, stride)
torch.cuda.synchronize()

#!/bin/bash
# round 2, call R: source-level ncu of the two epilogue-bound GEMMs (fc1 + GELU + pre-activation output; masked dfc2 * dGELU)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 2 -c 1 -f -o gpurun_out/r02r_gemm_gelu_aux python tools/kbench.py gemm --ncu > gpurun_out/r02r_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 15 -c 1 -f -o gpurun_out/r02r_gemm_dgelu_pm python tools/kbench.py gemm --ncu > gpurun_out/r02r_ncu2.log 2>&1
ls -la gpurun_out/r02r_*
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | grep "rank 0"

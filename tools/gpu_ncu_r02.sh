#!/bin/bash
# round 2 evidence (run under gpurun, ONE GPU; numbers printed under ncu are never bench values):
#  (1) per-launch device time of every kernel of one EAGER training step of bench.py (cold-cache, serialised: compare SHARES
#      with kernel_shares of the bench line)                                  -> gpurun_out/r02_launches.csv
#  (2) ncu --set full, every hot kernel once at the bench shapes (tools/kbench.py --ncu): attention forward / fused backward
#      (+ prep, dQ convert), the tcgen05 GEMM family (plain, GELU, dGELU, masked product, mask stages, split-K wgrad), LoRA
#      dropout kernels, LayerNorm, cross-entropy, augmentation                -> gpurun_out/r02_full_*.csv (+ the attention .ncu-rep)
#  (3) a decode position (greedy, B = 128, eeg_ch = 273)                      -> gpurun_out/r02_full_decode.csv
# tools/ncu_summary_r02.py turns the CSVs into profiles/r02_*.txt / r02_traffic.json.
mkdir -p gpurun_out
TAG=${1:-r02}
NS_NO_TRAIN_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 1231 -c 361 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
full() {  # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -c 60 -f -o /tmp/${TAG}_full_$name "$@" > gpurun_out/${TAG}_ncu_$name.log 2>&1
  ncu -i /tmp/${TAG}_full_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_$name.csv 2>/dev/null
}
full attn 'attn_fwd_db_kernel|attn_bwd_fused_kernel|attn_bwd_prep_kernel|attn_bwd_dq_convert_kernel' python tools/kbench.py attn --ncu
cp /tmp/${TAG}_full_attn.ncu-rep gpurun_out/
full gemm 'gemm_nt_kernel|gemm_tn_kernel|lora_da_kernel|dropout_bits_kernel|lora_bwd_b_kernel' python tools/kbench.py gemm --ncu
full elem 'ln_|ce_|aug_b|adamw|sumsq' python tools/kbench.py head ln aug --ncu
NS_DECODE_NCU=1 ncu --set full --clock-control none -k regex:'attn_decode|cross_absorbed|gemm_nt_kernel|greedy|ln_fwd|embed' -s 1000 -c 80 -f -o /tmp/${TAG}_full_decode \
    python tools/bench_decode.py --B 128 --max-length 48 --batches 1 > gpurun_out/${TAG}_ncu_decode.log 2>&1
ncu -i /tmp/${TAG}_full_decode.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_decode.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*; du -sh gpurun_out

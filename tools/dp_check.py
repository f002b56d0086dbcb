"""Two-rank check of the data-parallel training step on real GPUs (run under torchrun, NCCL):
the step with the gradient all-reduce split in two buckets and overlapped with the stem backward (two CUDA graphs per step)
must produce the same reduced gradient as the step with one all-reduce after the whole backward (up to the summation order of
the split-K / dQ reductions: fp32 atomics, ~1e-6 relative), the same weights within what that noise does to AdamW's normalised
updates (measured against a second run of the single-bucket mode), and bit-identical weights on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from neuspeech1_b200 import engine as E
from neuspeech1_b200.engine import ModelDims, WhisperEEGEngine
from neuspeech1_b200.weights import random_lora, random_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
dims = ModelDims(d_model=256, enc_layers=2, dec_layers=2, enc_heads=4, dec_heads=4, enc_ffn=512, dec_ffn=512, vocab=2000,
                 max_source_positions=160, max_target_positions=32, eeg_ch=24, pad_token_id=1997, eos_token_id=1997,
                 decoder_start_token_id=1998, begin_suppress_tokens=(220, 1996))
P, lora = random_params(dims, seed=0), random_lora(dims, seed=1, b_std=0.05)
g = torch.Generator().manual_seed(100 + rank)                      # every rank its own batch shard
x = (0.3 * torch.randn(4, dims.eeg_ch, dims.T, generator=g)).clamp_(-1, 1).to(dev)
labels = torch.randint(0, 1990, (4, 8), generator=g).to(dev)


def hook(flat):
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / world)


res = {}
for name, no_overlap in (("overlap", False), ("single", True), ("single2", True)):
    E._NO_AR_OVERLAP = no_overlap
    eng = WhisperEEGEngine(dims, P, lora, dtype=torch.bfloat16, device=dev, lora_dropout=0.05, dropout_seed=3)
    losses = [float(eng.train_step(x, labels, lr=1e-3, all_reduce=hook))]
    g1 = eng.grad.clone()                                                                         # reduced gradient of step 1
    losses += [float(eng.train_step(x, labels, lr=1e-3, all_reduce=hook)) for _ in range(4)]    # capture, replays
    res[name] = (eng.flat.clone(), losses, eng.graph_launches, g1)
rel = lambda u, v: float((u - v).norm() / v.norm().clamp_min(1e-30))
gerr, gnoise = rel(res["overlap"][3], res["single"][3]), rel(res["single2"][3], res["single"][3])
werr = float((res["overlap"][0] - res["single"][0]).abs().max()); wnoise = float((res["single2"][0] - res["single"][0]).abs().max())
other = res["overlap"][0].clone()
dist.broadcast(other, src=0)
sync = float((res["overlap"][0] - other).abs().max())
ok = gerr <= max(10 * gnoise, 1e-5) and werr <= max(10 * wnoise, 1e-6) and sync == 0.0 and res["overlap"][2] > 0
print(f"rank {rank}: reduced gradient overlap vs single {gerr:.2e} (run-to-run {gnoise:.2e}), weights after 5 steps {werr:.2e} "
      f"(run-to-run {wnoise:.2e}), max |rank - rank0| = {sync:.1e}, graph launches {res['overlap'][2]}, "
      f"losses {['%.4f' % l for l in res['overlap'][1]]} -> {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

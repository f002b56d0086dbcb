"""Batched greedy decode benchmark (BASELINE.json configs[3]): Whisper-base, eeg_ch=273, B=128, merged weights (no LoRA at
inference: evaluation.py:88-89), KV cache, max_length 448, random-init weights (never emit EOS -> every row runs all 447
new tokens, the worst case).  Reports samples/s, tokens/s and p50 latency per batch, eager launches vs CUDA-graph replay.

    python tools/bench_decode.py [--B 128] [--eeg-ch 273] [--max-length 448] [--batches 3]"""
import argparse, json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuspeech1_b200.engine import ModelDims, WhisperEEGEngine
from neuspeech1_b200.weights import random_params

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=128); ap.add_argument("--eeg-ch", dest="eeg_ch", type=int, default=273)
ap.add_argument("--max-length", dest="max_length", type=int, default=448); ap.add_argument("--batches", type=int, default=3)
ap.add_argument("--beams", type=int, default=0, help="also time evaluation.py's beam search (num_beams, penalty 5.0, no-repeat 2) at --beam-B")
ap.add_argument("--beam-B", dest="beam_B", type=int, default=32)
a = ap.parse_args()
dev = torch.device("cuda")
dims = ModelDims(eeg_ch=a.eeg_ch)
eng = WhisperEEGEngine(dims, random_params(dims, seed=0), None, dtype=torch.bfloat16, device=dev)
g = torch.Generator().manual_seed(3)
x = (0.3 * torch.randn(a.B, dims.eeg_ch, dims.T, generator=g)).clamp_(-1, 1).to(dev)
res = {"config": {"workload": f"greedy decode, Whisper-base, eeg_ch={a.eeg_ch}, B={a.B}, max_length={a.max_length}, bf16, merged weights"}}
for name, graphs in (("eager", False), ("cuda_graphs", True)):
    ids = eng.greedy(x, max_length=a.max_length, use_graphs=graphs)          # warm-up (and graph capture)
    if graphs:
        ids = eng.greedy(x, max_length=a.max_length, use_graphs=True)        # second pass re-captures the early steps once
    torch.cuda.synchronize()
    lat = []
    for _ in range(a.batches):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); out = eng.greedy(x, max_length=a.max_length, use_graphs=graphs); e.record()
        torch.cuda.synchronize(); lat.append(s.elapsed_time(e))
    p50 = statistics.median(lat)
    res[name] = {"p50_ms_per_batch": p50, "samples_per_s": a.B * 1e3 / p50, "tokens_per_s": a.B * out.shape[1] * 1e3 / p50,
                 "ms_per_token_step": p50 / out.shape[1], "new_tokens": int(out.shape[1])}
    res[name + "_ids_head"] = out[0, :8].tolist()
res["graphs_match_eager"] = res["eager_ids_head"] == res["cuda_graphs_ids_head"]
if a.beams > 1:
    xb = x[:a.beam_B]
    run = lambda: eng.beam_search(xb, max_length=a.max_length, num_beams=a.beams, repetition_penalty=5.0, no_repeat_ngram_size=2,
                                  use_graphs=True)
    out = run(); out = run(); torch.cuda.synchronize()        # first call sizes the workspace and captures, second re-captures early steps
    lat = []
    for _ in range(a.batches):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); out = run(); e.record()
        torch.cuda.synchronize(); lat.append(s.elapsed_time(e))
    p50 = statistics.median(lat)
    res[f"beam{a.beams}"] = {"B": a.beam_B, "p50_ms_per_batch": p50, "samples_per_s": a.beam_B * 1e3 / p50, "new_tokens": int(out.shape[1]),
                             "ms_per_token_step": p50 / max(int(out.shape[1]), 1),
                             "note": "beams in the batch dimension (query dimension of the cross-attention), decoder pass replayed as a CUDA graph per position, fused scoring kernel (ns_beam_row_topk), cache-row table instead of cache copies"}
if os.environ.get("NS_DECODE_PROFILE"):
    from neuspeech1_b200 import ops
    ops.profile_begin()
    eng.greedy(x, max_length=a.max_length, use_graphs=False)
    prof = ops.profile_end()
    fam = {}
    for r in prof:
        f = fam.setdefault(r["name"], [0, 0.0]); f[0] += 1; f[1] += r["ms"]
    res["profile_ms"] = {k: [v[0], round(v[1], 2)] for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}
print(json.dumps(res))

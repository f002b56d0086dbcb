"""Is the training step launch-bound?  Times 10 eager steps against 10 replays of ONE CUDA graph of the same step (timing
probe only: the AdamW step counter is frozen inside the captured graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuspeech1_b200.engine import ModelDims, WhisperEEGEngine
from neuspeech1_b200.weights import random_lora, random_params
dev = torch.device("cuda")
dims = ModelDims()
eng = WhisperEEGEngine(dims, random_params(dims, seed=0), random_lora(dims, seed=1, b_std=0.01), dtype=torch.bfloat16, device=dev)
g = torch.Generator().manual_seed(7)
B, L = 64, 32
x = (0.3 * torch.randn(B, dims.eeg_ch, dims.T, generator=g)).clamp_(-1, 1).to(dev)
labels = torch.randint(0, 50257, (B, L), generator=g); labels[:, -4:] = -100; labels = labels.to(dev)
def timed(fn, n=10):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
step = lambda: eng.train_step(x, labels, lr=1e-3)
for _ in range(3): step()
print("## eager  ms/step", timed(step))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    loss = step()
print("## graph  ms/step", timed(graph.replay))
print("## eager again", timed(step))

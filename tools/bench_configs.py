"""Training-step throughput on the other measurement configurations of SURVEY.md section 8(d) (bench.py covers config #2 with
the identity augmentation):

  #2-mask  Whisper-base, eeg_ch=208, B=64, augmentation {"mask": {"prob": 1.0, "kwargs": {"unit": [1, 40], "mask_prob": 0.25,
           "random_type": 1}}} -- the augmentation pass expands a fresh Bernoulli grid every step (drawn on the host with the
           reference's RNG calls, neuspeech1_b200/augment_eeg.py)
  #3       Whisper-base, eeg_ch=273 (Schoffelen), B=64: only the first stem conv's K changes (3*273 -> padded 3*288)
  #5       large-v3 widths (d=1280, 20 heads, 32+32 layers, ffn 5120, vocab 51866), eeg_ch=273, B=16, LoRA r=32

Each: 3 warm-up steps, then `--steps` steps timed with CUDA events, inputs resident in HBM.  FLOP/sample from the model of
SURVEY.md section 8(d).

    python tools/bench_configs.py [--steps 5] [--only mask,c273,large] [--out gpurun_out/configs.json]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuspeech1_b200.augment_eeg import BatchAugmenter
from neuspeech1_b200.engine import ModelDims, WhisperEEGEngine
from neuspeech1_b200.weights import random_lora, random_params


def flop_per_sample(d: ModelDims, L: int) -> float:
    T, S, C, dm, r = d.T, d.max_source_positions, d.eeg_ch, d.d_model, d.lora_r
    F, Fd, V = d.enc_ffn, d.dec_ffn, d.vocab
    convA, convB, convC = 6 * T * C * dm, 3 * T * dm * dm, 1.5 * T * dm * dm
    qkvo, attn, mlp = 8 * S * dm * dm, 4 * S * S * dm, 4 * S * dm * F
    lora = 2 * S * r * (4 * (2 * dm) + 2 * (dm + F))
    dec_lin = 8 * L * dm * dm + 4 * L * dm * dm + 4 * S * dm * dm + 4 * L * dm * Fd
    dec_attn = 4 * L * L * dm + 4 * L * S * dm
    proj = 2 * L * dm * V
    fwd = convA + convB + convC + d.enc_layers * (qkvo + attn + mlp + lora) + d.dec_layers * (dec_lin + dec_attn) + proj
    bwd = convA + 2 * (convB + convC) + d.enc_layers * (qkvo + 2 * attn + mlp + 2 * lora) + d.dec_layers * (dec_lin + 2 * dec_attn) + proj
    return float(fwd + bwd)


def run(name, dims, B, L, steps, aug_cfg=None):
    dev = torch.device("cuda")
    eng = WhisperEEGEngine(dims, random_params(dims, seed=0), random_lora(dims, seed=1, b_std=0.01), dtype=torch.bfloat16, device=dev)
    g = torch.Generator().manual_seed(7)
    x = (0.3 * torch.randn(B, dims.eeg_ch, dims.T, generator=g)).clamp_(-1, 1)
    lens = [int(torch.randint(400, 5000, (1,), generator=g)) for _ in range(B)]
    for b, n in enumerate(lens):
        x[b, :, n:] = 0
    labels = torch.randint(0, 50257, (B, L), generator=g); labels[:, -4:] = -100
    x, labels = x.to(dev), labels.to(dev)
    augm = BatchAugmenter(aug_cfg) if aug_cfg else None
    shapes = [(dims.eeg_ch, n) for n in lens]

    def step():
        aug = augm.plan(shapes, dev) if augm else None          # fresh random decisions every step (host RNG, tiny H2D)
        return eng.train_step(x, labels, lr=1e-3, aug=aug)

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        loss = step()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    fl = flop_per_sample(dims, L)
    res = {"workload": name, "B": B, "L": L, "eeg_ch": dims.eeg_ch, "d_model": dims.d_model, "enc_layers": dims.enc_layers,
           "ms_per_step": ms, "samples_per_s": B * 1e3 / ms, "gflop_per_sample": fl / 1e9, "tflops": B * fl / ms / 1e9,
           "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
    del eng
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="mask,c273,large")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = []
    if "mask" in a.only:
        cfg = {"mask": {"prob": 1.0, "kwargs": {"unit": [1, 40], "mask_prob": 0.25, "random_type": 1}}}
        out.append(run("config #2 + block mask augmentation", ModelDims(eeg_ch=208), 64, 32, a.steps, cfg))
    if "c273" in a.only:
        out.append(run("config #3 (eeg_ch=273)", ModelDims(eeg_ch=273), 64, 32, a.steps))
    if "large" in a.only:
        big = ModelDims(d_model=1280, enc_layers=32, dec_layers=32, enc_heads=20, dec_heads=20, enc_ffn=5120, dec_ffn=5120,
                        vocab=51866, eeg_ch=273)
        out.append(run("config #5 (large-v3 widths, eeg_ch=273)", big, 16, 32, max(2, a.steps // 2)))
    for r in out:
        print("## " + json.dumps(r))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)

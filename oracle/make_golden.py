"""Generate tests/golden/*.npz by running the REFERENCE's own code  --  TEST INFRASTRUCTURE ONLY.

Run in the authoring container (needs /root/reference, which does not travel to the GPU box):
    python -m oracle.make_golden
It imports, from /root/reference (not copied): utils.model_utils.projection_module, utils.augment_eeg.*,
utils.utils.add_gaussian_noise, and builds the object evaluation.py:72-86 builds (stock HF Whisper + that stem).
The committed fixtures pin oracle/whisper_eeg.py and oracle/augment.py (tests/test_oracle.py, tests/test_augment_oracle.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("NEUSPEECH_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    from utils.model_utils import projection_module                      # reference
    from utils import augment_eeg as ref_aug                             # reference
    from utils.utils import add_gaussian_noise as ref_noise              # reference
    sys.path.pop(0)
    from oracle import whisper_eeg as O
    from oracle.hf_bridge import build_hf, inject_lora_hf

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- 1. stem: reference projection_module with oracle weights
    dims = O.TINY
    P = O.init_params(dims, seed=0)
    x, labels = O.synthetic_batch(dims, B=2, L=8, seed=1)
    ref_stem = projection_module("base", meg_ch=dims.eeg_ch, d_model=dims.d_model)
    assert ref_stem.stride == (2,)
    with torch.no_grad():
        ref_stem[0].weight.copy_(P["model.encoder.conv1.0.weight"]); ref_stem[0].bias.copy_(P["model.encoder.conv1.0.bias"])
        ref_stem[2].weight.copy_(P["model.encoder.conv1.2.weight"]); ref_stem[2].bias.copy_(P["model.encoder.conv1.2.bias"])
        stem_out = ref_stem(x)
    np.savez_compressed(os.path.join(OUT, "stem_ref.npz"), stem_out=stem_out.numpy())

    # ---- 2. tiny model: stock HF + reference stem (+ restated LoRA), fwd/loss/grads/greedy
    def ref_factory(ch, d):
        return projection_module("base", meg_ch=ch, d_model=d)

    lora = O.init_lora(dims, seed=1, b_std=0.05)
    m = build_hf(dims, P, stem_factory=ref_factory)
    with torch.no_grad():
        out = m(input_features=x, labels=labels)
        gen = m.generate(input_features=x, do_sample=False, num_beams=1, max_length=dims.max_target_positions)
    rec = dict(enc=out.encoder_last_hidden_state.numpy(), loss=np.float64(out.loss.item()),
               logits=out.logits.numpy(), greedy=gen.numpy())
    m = inject_lora_hf(m, lora, dims.lora_scale)
    for p in m.parameters():
        p.requires_grad_(False)
    train = []
    for n_, p in m.named_parameters():
        if "lora_" in n_ or n_.startswith("model.encoder.conv1.") or n_.startswith("model.encoder.conv2."):
            p.requires_grad_(True); train.append((n_, p))
    out = m(input_features=x, labels=labels)
    out.loss.backward()
    rec["lora_enc"] = out.encoder_last_hidden_state.detach().numpy()
    rec["lora_loss"] = np.float64(out.loss.item())
    for n_, p in train:
        key = n_.replace(".lora_A", ".lora_A.default.weight").replace(".lora_B", ".lora_B.default.weight")
        rec["grad:" + key] = p.grad.numpy()
    with torch.no_grad():
        gen = m.generate(input_features=x, do_sample=False, num_beams=1, max_length=dims.max_target_positions)
    rec["lora_greedy"] = gen.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_model.npz"), **rec)

    # ---- 3. Whisper-base (config #1 shape, B=2): subsampled encoder states, loss, 12 greedy tokens
    dims = O.WHISPER_BASE
    P = O.init_params(dims, seed=0)
    x, labels = O.synthetic_batch(dims, B=2, L=32, seed=1)
    m = build_hf(dims, P, stem_factory=ref_factory)
    with torch.no_grad():
        out = m(input_features=x, labels=labels)
        gen = m.generate(input_features=x, do_sample=False, num_beams=1, max_length=13)
    np.savez_compressed(os.path.join(OUT, "base_model.npz"),
                        enc_sub=out.encoder_last_hidden_state[:, ::50, ::8].numpy(), loss=np.float64(out.loss.item()),
                        logits_sub=out.logits[:, ::4, ::997].numpy(), greedy=gen.numpy())

    # ---- 4. augmentation: the reference's own generators under fixed seeds
    rec = {}
    cases = [((208, 6000), [1, 40], 0.25, 1), ((208, 6000), [1, 40], 0.25, 2), ((208, 6000), [1, 40], 0.25, 3),
             ((273, 6000), [4, 37], 0.5, 1), ((16, 777), [3, 50], 0.3, 1), ((16, 1234), [1, 40], 0.25, 1)]
    for i, (shape, unit, prob, rt) in enumerate(cases):
        torch.manual_seed(100 + i)
        msk = ref_aug.RandomShapeMasker(unit=list(unit), mask_prob=prob, random_type=rt)(shape)
        rec[f"mask{i}"] = np.packbits(msk.numpy().astype(np.uint8), axis=1)
        rec[f"mask{i}_meta"] = np.array([shape[0], shape[1], unit[0], unit[1], rt, 100 + i], dtype=np.int64)
        rec[f"mask{i}_prob"] = np.float64(prob)
    a = np.arange(30, dtype=np.float32).reshape(3, 10)
    rec["shift_in"] = a; rec["shift_out"] = ref_aug.shift_data(a, 4)
    np.random.seed(7)
    sig = (0.3 * np.random.randn(5, 400)).clip(-1, 1)
    np.random.seed(11)
    rec["noise_in"] = sig; rec["noise_out"] = ref_noise(sig, (20, 50))
    np.savez_compressed(os.path.join(OUT, "augment_ref.npz"), **rec)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

"""Pin the oracle restatement (oracle/whisper_eeg.py) against fixtures produced by the reference's own stem
inside stock HF Whisper (oracle/make_golden.py) and live against stock HF."""
import os

import numpy as np
import pytest
import torch

from oracle import whisper_eeg as O
from oracle.hf_bridge import build_hf, inject_lora_hf


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def tiny():
    dims = O.TINY
    P = O.init_params(dims, seed=0)
    x, labels = O.synthetic_batch(dims, B=2, L=8, seed=1)
    return dims, P, x, labels


def test_stem_matches_reference_projection_module(tiny, golden_dir):
    dims, P, x, _ = tiny
    g = np.load(os.path.join(golden_dir, "stem_ref.npz"))
    a = torch.nn.functional.gelu(torch.nn.functional.conv1d(x, P["model.encoder.conv1.0.weight"], P["model.encoder.conv1.0.bias"], padding=1))
    b = torch.nn.functional.conv1d(a, P["model.encoder.conv1.2.weight"], P["model.encoder.conv1.2.bias"], stride=2, padding=1)
    assert rel(b, g["stem_out"]) < 1e-6


def test_forward_loss_matches_golden(tiny, golden_dir):
    dims, P, x, labels = tiny
    g = np.load(os.path.join(golden_dir, "tiny_model.npz"))
    loss, logits, enc = O.forward_loss(x, labels, P, dims)
    assert rel(enc, g["enc"]) < 1e-5
    assert rel(logits, g["logits"]) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))


def test_greedy_matches_golden(tiny, golden_dir):
    dims, P, x, _ = tiny
    g = np.load(os.path.join(golden_dir, "tiny_model.npz"))
    ids = O.greedy_decode(x, P, dims, max_length=dims.max_target_positions)
    assert np.array_equal(ids.numpy(), g["greedy"])
    lora = O.init_lora(dims, seed=1, b_std=0.05)
    ids = O.greedy_decode(x, P, dims, max_length=dims.max_target_positions, lora=lora)
    assert np.array_equal(ids.numpy(), g["lora_greedy"])


def test_lora_grads_match_golden(tiny, golden_dir):
    dims, P, x, labels = tiny
    g = np.load(os.path.join(golden_dir, "tiny_model.npz"))
    lora = O.init_lora(dims, seed=1, b_std=0.05)
    loss, grads, enc = O.grads(x, labels, P, dims, lora)
    assert rel(enc, g["lora_enc"]) < 1e-5
    assert abs(float(loss) - float(g["lora_loss"])) < 1e-5 * abs(float(g["lora_loss"]))
    keys = [k for k in g.files if k.startswith("grad:")]
    assert len(keys) == len(grads) == dims.enc_layers * 12 + 6
    for k in keys:
        assert rel(grads[k[5:]], g[k]) < 1e-4, k


def test_live_against_stock_hf(tiny):
    """transformers ships in the image on both boxes: check the restatement live as well (different seed/batch)."""
    dims = O.TINY
    P = O.init_params(dims, seed=3)
    x, labels = O.synthetic_batch(dims, B=3, L=6, seed=4)
    lora = O.init_lora(dims, seed=5, b_std=0.05)
    m = inject_lora_hf(build_hf(dims, P), lora, dims.lora_scale)
    with torch.no_grad():
        out = m(input_features=x, labels=labels)
    loss, logits, enc = O.forward_loss(x, labels, P, dims, lora)
    assert rel(enc, out.encoder_last_hidden_state) < 1e-5
    assert rel(logits, out.logits) < 1e-5
    assert abs(float(loss) - float(out.loss)) < 1e-5


def test_wrong_length_raises(tiny):
    dims, P, x, _ = tiny
    with pytest.raises(ValueError):
        O.encoder(x[..., :-4], P, dims)


def test_base_model_golden(golden_dir):
    """Whisper-base shape (config #1, B=2) against the reference-stem HF run."""
    dims = O.WHISPER_BASE
    P = O.init_params(dims, seed=0)
    x, labels = O.synthetic_batch(dims, B=2, L=32, seed=1)
    g = np.load(os.path.join(golden_dir, "base_model.npz"))
    with torch.no_grad():
        loss, logits, enc = O.forward_loss(x, labels, P, dims)
    assert rel(enc[:, ::50, ::8], g["enc_sub"]) < 1e-5
    assert rel(logits[:, ::4, ::997], g["logits_sub"]) < 1e-4
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    ids = O.greedy_decode(x, P, dims, max_length=13)
    assert np.array_equal(ids.numpy(), g["greedy"])


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p = {"a": torch.randn(7, 5), "b": torch.randn(11)}
    q = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.AdamW(q.values(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    st = O.AdamWState()
    for it in range(3):
        g = {k: torch.randn_like(v) * 3 for k, v in p.items()}
        for k in q:
            q[k].grad = g[k].clone()
        n_ref = torch.nn.utils.clip_grad_norm_(list(q.values()), 1.0)
        opt.step()
        n = O.clip_and_adamw(p, g, st, lr=1e-2)
        assert abs(n - float(n_ref)) < 1e-5 * n
        for k in p:
            assert rel(p[k], q[k].detach()) < 1e-6


def test_lora_dropout_mask_spec():
    """The counter-hash keep mask of the LoRA-branch dropout (oracle.lora_dropout_keep, the specification for the device kernels):
    rate, independence between modules and seeds, p = 0 identical to no dropout, gradients flow through the kept elements."""
    import torch
    from oracle import whisper_eeg as O
    k1 = O.lora_dropout_keep(7, "model.encoder.layers.0.fc1", 4096, 512, 0.05)
    k2 = O.lora_dropout_keep(7, "model.encoder.layers.0.fc2", 4096, 512, 0.05)
    k3 = O.lora_dropout_keep(8, "model.encoder.layers.0.fc1", 4096, 512, 0.05)
    for k in (k1, k2, k3):
        assert abs(1.0 - k.float().mean().item() - 0.05) < 2e-3
    assert abs((k1 ^ k2).float().mean().item() - 2 * 0.05 * 0.95) < 3e-3 and abs((k1 ^ k3).float().mean().item() - 0.095) < 3e-3
    assert abs(k1.float().mean(dim=0).std().item() - (0.05 * 0.95 / 4096) ** 0.5) < 1e-3      # no column structure
    assert bool(O.lora_dropout_keep(7, "x", 64, 64, 0.0).all())
    dims = O.TINY
    P = O.init_params(dims, seed=0)
    lora = O.init_lora(dims, seed=1, b_std=0.05)
    x, labels = O.synthetic_batch(dims, B=2, L=8, seed=1)
    base, _, _ = O.forward_loss(x, labels, P, dims, lora)
    same, _, _ = O.forward_loss(x, labels, P, dims, {**lora, "__dropout__": (0.0, 3)})
    assert torch.equal(base, same)
    la = {k: v.clone().requires_grad_(True) for k, v in lora.items()}
    dropped, _, _ = O.forward_loss(x, labels, P, dims, {**la, "__dropout__": (0.3, 3)})
    assert abs(float(dropped) - float(base)) > 1e-6
    dropped.backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in la.values())

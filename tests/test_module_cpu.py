"""CPU-side checks of the drop-in module's object contract: module tree / parameter names as PEFT and the
reference's scripts expect them, checkpoint round trips, merge semantics.  No engine (no GPU) involved."""
import os

import pytest
import torch

from neuspeech1_b200 import lora as L
from neuspeech1_b200.engine import ModelDims, TrainableLayout
from neuspeech1_b200.load_model import WhisperForConditionalGeneration
from neuspeech1_b200.model_utils import projection_module
from neuspeech1_b200.weights import merge_lora, random_lora, random_params
from oracle import whisper_eeg as O

DIMS = ModelDims.from_any(O.TINY)


def make(lora=True):
    return WhisperForConditionalGeneration(DIMS, random_params(DIMS), random_lora(DIMS, b_std=0.05) if lora else None, device="cpu")


def test_projection_module_matches_reference_contract():
    m = projection_module("base", meg_ch=208, d_model=512)
    assert m.stride == (2,) and m[0].weight.shape == (512, 208, 3) and m[2].stride == (2,)
    with pytest.raises(NotImplementedError):
        projection_module("nope")


def test_parameter_names_match_hf_and_peft():
    m = make()
    names = dict(m.named_parameters())
    for k in ("model.encoder.conv1.0.weight", "model.encoder.conv1.2.bias", "model.encoder.conv2.weight",
              "model.encoder.layers.1.self_attn.q_proj.base_layer.weight", "model.encoder.layers.1.self_attn.q_proj.lora_A.default.weight",
              "model.encoder.layers.0.fc2.lora_B.default.weight", "model.decoder.layers.1.encoder_attn.k_proj.weight",
              "model.decoder.embed_tokens.weight"):
        assert k in names, k
    assert "model.encoder.layers.0.self_attn.k_proj.base_layer.bias" not in names          # k_proj has no bias (HF:279)
    trainable = {k for k, p in names.items() if p.requires_grad}
    lay = TrainableLayout(DIMS)
    norm = {k.replace(".base_layer.", ".") for k in trainable}
    assert norm == set(lay.entries), norm ^ set(lay.entries)
    assert m.model.encoder.conv2.in_channels == DIMS.d_model                               # finetune.py:140
    assert m.model.encoder.conv1.stride[0] * m.model.encoder.conv2.stride[0] == 4          # generation_whisper.py:653
    assert m.proj_out.weight is m.model.decoder.embed_tokens.weight


def test_match_modules_string_picks_the_36_encoder_linears():
    m = make(lora=False)
    names = L.match_modules_string(m.named_modules(), ["model.encoder"], ["k_proj", "q_proj", "v_proj", "out_proj", "fc1", "fc2"])
    assert len(names) == DIMS.enc_layers * 6 and all(n.startswith("model.encoder.layers.") for n in names)


def _peft_layout_fixture(dims, seed=3):
    """An adapter checkpoint written the way PEFT's `save_pretrained` lays it out (restated from PEFT's
    get_peft_model_state_dict: `base_model.model.` prefix, adapter name stripped from LoRA keys, `modules_to_save.<adapter>.`
    infix stripped from the saved stem convs) -- built by hand here, NOT through neuspeech1_b200.lora."""
    g = torch.Generator().manual_seed(seed)
    d, r, F = dims.d_model, dims.lora_r, dims.enc_ffn
    sd = {}
    for i in range(dims.enc_layers):
        for t, fin, fout in (("self_attn.q_proj", d, d), ("self_attn.k_proj", d, d), ("self_attn.v_proj", d, d),
                             ("self_attn.out_proj", d, d), ("fc1", d, F), ("fc2", F, d)):
            sd[f"base_model.model.model.encoder.layers.{i}.{t}.lora_A.weight"] = torch.randn(r, fin, generator=g)
            sd[f"base_model.model.model.encoder.layers.{i}.{t}.lora_B.weight"] = torch.randn(fout, r, generator=g)
    sd["base_model.model.model.encoder.conv1.0.weight"] = torch.randn(d, dims.eeg_ch, 3, generator=g)
    sd["base_model.model.model.encoder.conv1.0.bias"] = torch.randn(d, generator=g)
    sd["base_model.model.model.encoder.conv1.2.weight"] = torch.randn(d, d, 3, generator=g)
    sd["base_model.model.model.encoder.conv1.2.bias"] = torch.randn(d, generator=g)
    sd["base_model.model.model.encoder.conv2.weight"] = torch.randn(d, d, 3, generator=g)
    sd["base_model.model.model.encoder.conv2.bias"] = torch.randn(d, generator=g)
    targets = [f"model.encoder.layers.{i}.{t}" for i in range(dims.enc_layers)
               for t in ("self_attn.k_proj", "self_attn.v_proj", "self_attn.q_proj", "self_attn.out_proj", "fc1", "fc2")]
    cfg = {"peft_type": "LORA", "task_type": None, "base_model_name_or_path": "openai/whisper-base", "r": r,
           "lora_alpha": dims.lora_alpha, "lora_dropout": 0.05, "bias": "none", "fan_in_fan_out": False, "inference_mode": True,
           "target_modules": targets, "modules_to_save": ["model.encoder.conv1", "model.encoder.conv2"]}
    return sd, cfg


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
def test_adapter_reads_and_writes_pefts_on_disk_layout(tmp_path, fmt):
    """finetune.py:182-185 / evaluation.py:88-89 / merge_lora.py:43-44: `PeftModel.from_pretrained(model, dir)` on a PLAIN model.
    A hand-built checkpoint in PEFT's layout loads (wrapping the modules on the way, dropout taken from the config), every
    tensor lands in the right parameter, and what `save_adapter` writes back has exactly PEFT's keys and config fields."""
    import json
    from safetensors.torch import load_file, save_file
    sd, cfg = _peft_layout_fixture(DIMS)
    src = tmp_path / "ref_adapter"; src.mkdir()
    if fmt == "safetensors":
        save_file(sd, str(src / "adapter_model.safetensors"))
    else:
        torch.save(sd, str(src / "adapter_model.bin"))
    json.dump(cfg, open(src / "adapter_config.json", "w"))
    m = make(lora=False)                                                                   # plain module tree
    assert not any(isinstance(x, L.LoraLinear) for x in m.modules())
    L.load_adapter(m, str(src))
    own = dict(m.named_parameters())
    assert torch.equal(own["model.encoder.layers.1.self_attn.q_proj.lora_A.default.weight"],
                       sd["base_model.model.model.encoder.layers.1.self_attn.q_proj.lora_A.weight"])
    assert torch.equal(own["model.encoder.layers.0.fc2.lora_B.default.weight"], sd["base_model.model.model.encoder.layers.0.fc2.lora_B.weight"])
    assert torch.equal(own["model.encoder.conv1.modules_to_save.default.2.weight"], sd["base_model.model.model.encoder.conv1.2.weight"])
    assert torch.equal(own["model.encoder.conv2.modules_to_save.default.bias"], sd["base_model.model.model.encoder.conv2.bias"])
    assert not torch.equal(own["model.encoder.conv1.original_module.2.weight"], sd["base_model.model.model.encoder.conv1.2.weight"])
    assert m._lora_cfg == {"r": DIMS.lora_r, "lora_alpha": DIMS.lora_alpha, "lora_dropout": 0.05}
    assert m.model.encoder.conv1.stride == (2,)                                            # resolves through the wrapper
    dst = tmp_path / "ours"
    L.save_adapter(m, str(dst), safe_serialization=(fmt == "safetensors"))
    back = load_file(str(dst / "adapter_model.safetensors")) if fmt == "safetensors" else torch.load(str(dst / "adapter_model.bin"))
    assert set(back) == set(sd)                                                            # PEFT's keys, nothing else
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    cfg2 = json.load(open(dst / "adapter_config.json"))
    assert cfg2["peft_type"] == "LORA" and cfg2["r"] == DIMS.lora_r and cfg2["lora_alpha"] == DIMS.lora_alpha and cfg2["lora_dropout"] == 0.05
    assert sorted(cfg2["target_modules"]) == sorted(cfg["target_modules"]) and cfg2["modules_to_save"] == cfg["modules_to_save"]
    # a tensor without a home is an error, not a skip
    sd_bad = dict(sd); sd_bad["base_model.model.model.decoder.layers.0.fc1.lora_A.weight"] = torch.zeros(DIMS.lora_r, DIMS.d_model)
    bad = tmp_path / "bad"; bad.mkdir()
    save_file(sd_bad, str(bad / "adapter_model.safetensors")); json.dump(cfg, open(bad / "adapter_config.json", "w"))
    with pytest.raises(KeyError):
        L.load_adapter(make(lora=False), str(bad))


def test_adapter_round_trip_and_merge(tmp_path):
    m = make()
    L.lora_inject(m, r=DIMS.lora_r, lora_alpha=DIMS.lora_alpha, modules_to_save=["model.encoder.conv1", "model.encoder.conv2"])
    sd = L.adapter_state_dict(m)
    assert "base_model.model.model.encoder.conv1.0.weight" in sd and "base_model.model.model.encoder.layers.0.fc1.lora_A.weight" in sd
    L.save_adapter(m, str(tmp_path))
    m2 = make()
    L.lora_inject(m2, r=DIMS.lora_r, lora_alpha=DIMS.lora_alpha, modules_to_save=["model.encoder.conv1", "model.encoder.conv2"])
    for p in m2.parameters():
        if p.requires_grad:
            p.data.add_(1.0)
    L.load_adapter(m2, str(tmp_path))
    for (k, a), (_, b) in zip(sorted(L.adapter_state_dict(m).items()), sorted(L.adapter_state_dict(m2).items())):
        assert torch.equal(a, b), k
    # the round-1 layout of this package (`.default.` kept, adapter_model.bin) still loads
    old = {"base_model.model." + k: v.detach().clone() for k, v in m.state_dict().items()
           if ".lora_A." in k or ".lora_B." in k or ".modules_to_save." in k}
    legacy = tmp_path / "legacy"; legacy.mkdir()
    torch.save(old, str(legacy / "adapter_model.bin"))
    m4 = make()
    L.lora_inject(m4, r=DIMS.lora_r, lora_alpha=DIMS.lora_alpha, modules_to_save=["model.encoder.conv1", "model.encoder.conv2"])
    L.load_adapter(m4, str(legacy))
    for (k, a), (_, b) in zip(sorted(L.adapter_state_dict(m).items()), sorted(L.adapter_state_dict(m4).items())):
        assert torch.equal(a, b), k
    # merge_and_unload == W + (alpha/r) B A
    P = random_params(DIMS); lo = random_lora(DIMS, b_std=0.05)
    m3 = WhisperForConditionalGeneration(DIMS, P, lo, device="cpu")
    L.merge_and_unload(m3)
    ref = merge_lora(P, lo, DIMS.lora_scale)
    got = dict(m3.named_parameters())
    for k in ("model.encoder.layers.0.self_attn.q_proj.weight", "model.encoder.layers.1.fc2.weight"):
        assert torch.allclose(got[k], ref[k], atol=1e-6)
    assert not any(".lora_" in k for k in got)


def test_from_pretrained_reads_hf_checkpoint(tmp_path):
    """A stock HF Whisper checkpoint directory (mel stem) loads; the EEG stem is created fresh (finetune.py:127-148)."""
    from oracle.hf_bridge import hf_config
    from transformers import WhisperForConditionalGeneration as HF
    hf = HF(hf_config(O.TINY))
    hf.save_pretrained(str(tmp_path))
    m = WhisperForConditionalGeneration.from_pretrained(str(tmp_path), eeg_ch=16, device="cpu")
    sd = hf.state_dict()
    got = m.state_dict()
    for k in ("model.encoder.layers.0.fc1.weight", "model.decoder.embed_tokens.weight", "model.encoder.conv2.weight"):
        assert torch.equal(got[k], sd[k]), k
    assert got["model.encoder.conv1.0.weight"].shape == (O.TINY.d_model, 16, 3)


def test_forward_argument_errors():
    m = make()
    with pytest.raises(ValueError):
        m.forward(input_features=torch.zeros(1, 16, 256))
    with pytest.raises(NotImplementedError):
        m.generate(torch.zeros(1, 16, 256), do_sample=True)
    with pytest.raises(NotImplementedError):
        m.generate(torch.zeros(1, 16, 256), num_beams=4, num_beam_groups=2)


def test_absorbed_query_key_weight_matches_attention_scores():
    """engine.absorb_query_key: the decode step's absorbed cross-attention scores with the folded weight equal q . K of HF
    modeling_whisper.py:284-336 (q scaled by Dh^-0.5 with its bias, bias-free keys), head by head, and the value side identity
    softmax(S) (enc Wv^T + bv) = (softmax(S) enc) Wv^T + bv holds (rows of a softmax sum to one)."""
    import torch
    from neuspeech1_b200.engine import absorb_query_key
    g = torch.Generator().manual_seed(0)
    d, H, S, B = 64, 4, 37, 3
    dh = d // H
    wq, wk, wv = (torch.randn(d, d, generator=g, dtype=torch.float64) * 0.2 for _ in range(3))
    bq, bv = torch.randn(d, generator=g, dtype=torch.float64), torch.randn(d, generator=g, dtype=torch.float64)
    x = torch.randn(B, d, generator=g, dtype=torch.float64)
    enc = torch.randn(B, S, d, generator=g, dtype=torch.float64)
    q = ((x @ wq.t() + bq) * dh ** -0.5).view(B, H, dh)
    K = (enc @ wk.t()).view(B, S, H, dh)
    V = (enc @ wv.t() + bv).view(B, S, H, dh)
    scores = torch.einsum("bhc,bshc->bhs", q, K)
    out = torch.einsum("bhs,bshc->bhc", scores.softmax(-1), V).reshape(B, d)
    w_abs, b_abs = absorb_query_key(wq, bq, wk, H)
    assert w_abs.shape == (H * d, d) and b_abs.shape == (H * d,)
    qp = (x @ w_abs.t() + b_abs).view(B, H, d)
    scores2 = torch.einsum("bhn,bsn->bhs", qp, enc)
    assert torch.allclose(scores, scores2, atol=1e-10)
    cp = torch.einsum("bhs,bsn->bhn", scores2.softmax(-1), enc)                     # C'_h = P_h enc
    out2 = torch.stack([cp[:, h] @ wv[h * dh:(h + 1) * dh].t() + bv[h * dh:(h + 1) * dh] for h in range(H)], dim=1).reshape(B, d)
    assert torch.allclose(out, out2, atol=1e-10)

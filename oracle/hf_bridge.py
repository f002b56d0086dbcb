"""Bridge between the oracle's parameter dict and stock HF Whisper  --  TEST INFRASTRUCTURE ONLY.

evaluation.py:13,72-86 builds exactly this: stock `transformers.WhisperForConditionalGeneration` with the mel stem
replaced through `encoder.set_input_embeddings(projection_module(...))`.  `build_hf` rebuilds that object from a
hand-written WhisperConfig (no hub access) so the restatement in oracle/whisper_eeg.py can be pinned against it.
The stem here is a local re-statement of utils/model_utils.py:9-17 (the reference tree does not travel to the GPU box);
oracle/make_golden.py checks it against the reference's own `projection_module` and commits the fixtures.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .whisper_eeg import Dims


def hf_config(dims: Dims):
    from transformers import WhisperConfig
    cfg = WhisperConfig(
        vocab_size=dims.vocab, num_mel_bins=80, d_model=dims.d_model, encoder_layers=dims.enc_layers,
        decoder_layers=dims.dec_layers, encoder_attention_heads=dims.enc_heads, decoder_attention_heads=dims.dec_heads,
        encoder_ffn_dim=dims.enc_ffn, decoder_ffn_dim=dims.dec_ffn, max_source_positions=dims.max_source_positions,
        max_target_positions=dims.max_target_positions, pad_token_id=dims.pad_token_id, bos_token_id=dims.pad_token_id,
        eos_token_id=dims.eos_token_id, decoder_start_token_id=dims.decoder_start_token_id, dropout=0.0,
        attention_dropout=0.0, activation_dropout=0.0, suppress_tokens=None,
        begin_suppress_tokens=list(dims.begin_suppress_tokens))
    cfg._attn_implementation = "eager"
    return cfg


def local_projection_module(meg_ch: int, d_model: int) -> nn.Module:
    conv1 = nn.Sequential(nn.Conv1d(meg_ch, d_model, kernel_size=3, padding=1), nn.GELU(),
                          nn.Conv1d(d_model, d_model, kernel_size=3, stride=2, padding=1))
    conv1.stride = (2,)
    return conv1


def build_hf(dims: Dims, P=None, stem_factory=None):
    """Stock HF model + EEG stem; if `P` is given its tensors are loaded (strict)."""
    from transformers import WhisperForConditionalGeneration
    m = WhisperForConditionalGeneration(hf_config(dims)).eval()
    stem = (stem_factory or local_projection_module)(dims.eeg_ch, dims.d_model)
    m.model.encoder.set_input_embeddings(stem)
    if P is not None:
        sd = {k: v.clone() for k, v in P.items()}
        sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(k == "proj_out.weight" for k in missing), missing
        m.tie_weights()
    return m


def params_from_hf(m) -> dict:
    return {k: v.detach().clone() for k, v in m.state_dict().items() if k != "proj_out.weight"}


class LoraLinear(nn.Module):
    """Restatement of PEFT lora.Linear forward (dropout 0): base(x) + scale * B(A(x)).  finetune.py:210-211."""

    def __init__(self, base: nn.Linear, a: torch.Tensor, b: torch.Tensor, scale: float):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.Parameter(a.clone()); self.lora_B = nn.Parameter(b.clone()); self.scale = scale

    def forward(self, x):
        return self.base_layer(x) + self.scale * ((x @ self.lora_A.t()) @ self.lora_B.t())


def inject_lora_hf(m, lora: dict, scale: float):
    """Wrap the HF encoder linears named in `lora` (names as in oracle.init_lora)."""
    mods = sorted({k.split(".lora_")[0] for k in lora})
    for name in mods:
        parent_name, attr = name.rsplit(".", 1)
        parent = m.get_submodule(parent_name)
        base = getattr(parent, attr)
        setattr(parent, attr, LoraLinear(base, lora[name + ".lora_A.default.weight"],
                                         lora[name + ".lora_B.default.weight"], scale))
    return m

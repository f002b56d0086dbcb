"""Parameter dictionaries for the engine: random initialisation of the reference's architecture (no hub access needed) and
conversion from a HF / reference state_dict.  Names follow the reference's state_dict: stock HF Whisper names with the
`projection_module('base')` Sequential as `model.encoder.conv1` (utils/model_utils.py:9-17, finetune.py:138-148)."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .engine import ENC_LORA_TARGETS, ModelDims, lora_module_name


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Whisper's fixed encoder position table (HF modeling_whisper.py:55-65)."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def random_params(dims: ModelDims, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random-init weights of the architecture (HF init_std 0.02 for linears/embeddings, PyTorch default for the convs)."""
    g = torch.Generator().manual_seed(seed)
    d = dims.d_model
    P: Dict[str, torch.Tensor] = {}
    n = lambda *shape: torch.randn(*shape, generator=g) * std

    def conv(name, cout, cin):
        bound = 1.0 / math.sqrt(cin * 3)
        P[name + ".weight"] = (torch.rand(cout, cin, 3, generator=g) * 2 - 1) * bound
        P[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def ln(name):
        P[name + ".weight"] = torch.ones(d); P[name + ".bias"] = torch.zeros(d)

    def attn(prefix):
        for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
            P[f"{prefix}.{p}.weight"] = n(d, d)
            if p != "k_proj":
                P[f"{prefix}.{p}.bias"] = torch.zeros(d)

    conv("model.encoder.conv1.0", d, dims.eeg_ch); conv("model.encoder.conv1.2", d, d); conv("model.encoder.conv2", d, d)
    P["model.encoder.embed_positions.weight"] = sinusoids(dims.max_source_positions, d)
    for i in range(dims.enc_layers):
        pre = f"model.encoder.layers.{i}"
        attn(pre + ".self_attn"); ln(pre + ".self_attn_layer_norm"); ln(pre + ".final_layer_norm")
        P[pre + ".fc1.weight"] = n(dims.enc_ffn, d); P[pre + ".fc1.bias"] = torch.zeros(dims.enc_ffn)
        P[pre + ".fc2.weight"] = n(d, dims.enc_ffn); P[pre + ".fc2.bias"] = torch.zeros(d)
    ln("model.encoder.layer_norm")
    P["model.decoder.embed_tokens.weight"] = n(dims.vocab, d)
    P["model.decoder.embed_positions.weight"] = n(dims.max_target_positions, d)
    for i in range(dims.dec_layers):
        pre = f"model.decoder.layers.{i}"
        attn(pre + ".self_attn"); ln(pre + ".self_attn_layer_norm")
        attn(pre + ".encoder_attn"); ln(pre + ".encoder_attn_layer_norm"); ln(pre + ".final_layer_norm")
        P[pre + ".fc1.weight"] = n(dims.dec_ffn, d); P[pre + ".fc1.bias"] = torch.zeros(dims.dec_ffn)
        P[pre + ".fc2.weight"] = n(d, dims.dec_ffn); P[pre + ".fc2.bias"] = torch.zeros(d)
    ln("model.decoder.layer_norm")
    return P


def random_lora(dims: ModelDims, seed: int = 1, b_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """PEFT LoRA init (finetune.py:210-211): A ~ kaiming_uniform(a=sqrt(5)), B = 0 (b_std > 0 mimics a trained adapter)."""
    g = torch.Generator().manual_seed(seed)
    d, r = dims.d_model, dims.lora_r
    L: Dict[str, torch.Tensor] = {}
    for i in range(dims.enc_layers):
        for t in ENC_LORA_TARGETS:
            fin = dims.enc_ffn if t == "fc2" else d
            fout = dims.enc_ffn if t == "fc1" else d
            bound = 1.0 / math.sqrt(fin)
            L[lora_module_name(i, t) + ".lora_A.default.weight"] = (torch.rand(r, fin, generator=g) * 2 - 1) * bound
            L[lora_module_name(i, t) + ".lora_B.default.weight"] = torch.randn(fout, r, generator=g) * b_std
    return L


def merge_lora(P: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], scale: float) -> Dict[str, torch.Tensor]:
    """merge_and_unload (evaluation.py:88-89, merge_lora.py:43-44): W <- W + scale * B A.  Returns a new dict."""
    out = dict(P)
    for name in {k.split(".lora_")[0] for k in lora}:
        a = lora[name + ".lora_A.default.weight"].float(); b = lora[name + ".lora_B.default.weight"].float()
        out[name + ".weight"] = P[name + ".weight"].float() + scale * (b @ a)
    return out


def params_from_state_dict(sd: Dict[str, torch.Tensor], prefix: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """Accept a HF WhisperForConditionalGeneration state_dict (optionally PEFT-prefixed `base_model.model.`), with PEFT's
    `base_layer.` / `modules_to_save.default.` / `original_module.` infixes normalised away."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if prefix and k.startswith(prefix):
            k = k[len(prefix):]
        if k.startswith("base_model.model."):
            k = k[len("base_model.model."):]
        if ".original_module." in k or k == "proj_out.weight" or ".lora_" in k:
            continue
        k = k.replace(".base_layer.", ".").replace(".modules_to_save.default.", ".")
        out[k] = v.detach().float()
    return out


def lora_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if ".lora_A." in k or ".lora_B." in k:
            if k.startswith("base_model.model."):
                k = k[len("base_model.model."):]
            if ".default." not in k:                      # adapter_model.safetensors drops the adapter name
                k = k.replace(".lora_A.", ".lora_A.default.").replace(".lora_B.", ".lora_B.default.")
            out[k] = v.detach().float()
    return out

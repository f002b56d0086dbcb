"""LoRA adapters with PEFT's object layout (PEFT itself is absent from this image).

`lora_inject` reproduces what `get_peft_model(model, LoraConfig(r, lora_alpha, target_modules=..., modules_to_save=...))` does
to the module tree at finetune.py:194-212, so that parameter names match PEFT checkpoints:
    <module>.base_layer.weight / <module>.lora_A.default.weight (r,in) / <module>.lora_B.default.weight (out,r)
    model.encoder.conv1.modules_to_save.default.* (trainable copy) and .original_module.* (frozen)
The forward arithmetic  y = base(x) + (alpha/r) * B(A(dropout(x)))  runs fused in the tcgen05 GEMM epilogue path
(engine.py); these modules only own the fp32 master parameters.  lora_dropout is accepted for API parity and must be 0
for bit-parity runs (the engine applies no dropout on the LoRA branch yet -- listed in DESIGN.md).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

ENC_TARGETS = ("k_proj", "q_proj", "v_proj", "out_proj", "fc1", "fc2")


class LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, lora_alpha: int, lora_dropout: float = 0.0):
        super().__init__()
        self.base_layer = base
        self.in_features, self.out_features = base.in_features, base.out_features
        self.r = {"default": r}
        self.lora_alpha = {"default": lora_alpha}
        self.scaling = {"default": lora_alpha / r}
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)
        base.requires_grad_(False)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias


class ModulesToSaveWrapper(nn.Module):
    """PEFT's modules_to_save wrapper: a frozen `original_module` and the trainable copy actually used."""

    def __init__(self, module: nn.Module):
        super().__init__()
        import copy
        self.original_module = module
        self.modules_to_save = nn.ModuleDict({"default": copy.deepcopy(module)})
        self.original_module.requires_grad_(False)
        self.modules_to_save["default"].requires_grad_(True)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("modules_to_save")["default"], name)   # e.g. conv1.stride (HF length check)

    def __getitem__(self, i):
        return self.modules_to_save["default"][i]


def match_modules_string(named_modules: Iterable, start_prefixes, end_suffixes, mid_prefixes=()):
    """utils/load_model.py:48-85: names that start with one of start_prefixes and end with one of end_suffixes."""
    out = []
    for name, _ in named_modules:
        if any(name.startswith(s) for s in start_prefixes) and any(name.endswith(e) for e in end_suffixes):
            if not mid_prefixes or any(m in name for m in mid_prefixes):
                out.append(name)
    return out


def lora_inject(model: nn.Module, r: int = 32, lora_alpha: int = 64, lora_dropout: float = 0.0,
                target_modules: Optional[Iterable[str]] = None, modules_to_save: Optional[Iterable[str]] = None,
                state: Optional[Dict[str, torch.Tensor]] = None):
    """In-place equivalent of get_peft_model for the reference's configuration (finetune.py:194-212)."""
    if lora_dropout > 0:
        import warnings
        warnings.warn(f"neuspeech1_b200: lora_dropout={lora_dropout} is recorded but NOT applied by the B200 engine yet "
                      "(the LoRA branch sees the undropped input; DESIGN.md section 6). Training proceeds without it.")
    if target_modules is None:
        target_modules = match_modules_string(model.named_modules(), ["model.encoder"], list(ENC_TARGETS))
    for name in list(target_modules):
        parent_name, attr = name.rsplit(".", 1)
        parent = model.get_submodule(parent_name)
        base = getattr(parent, attr)
        if isinstance(base, LoraLinear):
            continue
        if not name.startswith("model.encoder.layers."):
            raise NotImplementedError(f"LoRA target {name}: the B200 engine adapts the encoder linears only (finetune.py:194)")
        setattr(parent, attr, LoraLinear(base, r, lora_alpha, lora_dropout))
    for name in (modules_to_save or ()):
        parent_name, attr = name.rsplit(".", 1)
        parent = model.get_submodule(parent_name)
        mod = getattr(parent, attr)
        if not isinstance(mod, ModulesToSaveWrapper):
            setattr(parent, attr, ModulesToSaveWrapper(mod))
    if state is not None:
        own = dict(model.named_parameters())
        for k, v in state.items():
            own[k].data.copy_(v)
    model._lora_cfg = {"r": r, "lora_alpha": lora_alpha, "lora_dropout": lora_dropout}
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
    return model


def merge_and_unload(model: nn.Module):
    """PeftModel.merge_and_unload (evaluation.py:88-89, merge_lora.py:43-44): W <- W + (alpha/r) B A, wrappers removed."""
    for name, mod in list(model.named_modules()):
        if isinstance(mod, LoraLinear):
            base = mod.base_layer
            a = mod.lora_A["default"].weight.data.float(); b = mod.lora_B["default"].weight.data.float()
            base.weight.data += (mod.scaling["default"] * (b @ a)).to(base.weight.device, base.weight.dtype)
            parent_name, attr = name.rsplit(".", 1)
            setattr(model.get_submodule(parent_name), attr, base)
        elif isinstance(mod, ModulesToSaveWrapper):
            parent_name, attr = name.rsplit(".", 1)
            setattr(model.get_submodule(parent_name), attr, mod.modules_to_save["default"])
    model._lora_cfg = None
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
    return model


def adapter_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """PEFT-format adapter checkpoint content: LoRA A/B + modules_to_save copies, `base_model.model.` prefixed."""
    out = {}
    for k, v in model.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k or ".modules_to_save." in k:
            out["base_model.model." + k] = v.detach().cpu().clone()
    return out


def save_adapter(model: nn.Module, path: str):
    os.makedirs(path, exist_ok=True)
    torch.save(adapter_state_dict(model), os.path.join(path, "adapter_model.bin"))
    import json
    json.dump({"peft_type": "LORA", **(model._lora_cfg or {}), "target_modules": list(ENC_TARGETS),
               "modules_to_save": ["model.encoder.conv1", "model.encoder.conv2"]}, open(os.path.join(path, "adapter_config.json"), "w"))


def load_adapter(model: nn.Module, path: str):
    sd = torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
    own = dict(model.named_parameters())
    for k, v in sd.items():
        k = k[len("base_model.model."):] if k.startswith("base_model.model.") else k
        own[k].data.copy_(v)
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
    return model

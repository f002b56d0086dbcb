"""LoRA adapters with PEFT's object layout (PEFT itself is absent from this image).

`lora_inject` reproduces what `get_peft_model(model, LoraConfig(r, lora_alpha, target_modules=..., modules_to_save=...))` does
to the module tree at finetune.py:194-212, so that parameter names match PEFT checkpoints:
    <module>.base_layer.weight / <module>.lora_A.default.weight (r,in) / <module>.lora_B.default.weight (out,r)
    model.encoder.conv1.modules_to_save.default.* (trainable copy) and .original_module.* (frozen)
The forward arithmetic  y = base(x) + (alpha/r) * B(A(dropout(x)))  runs fused in the tcgen05 GEMM epilogue path
(engine.py); these modules only own the fp32 master parameters.  lora_dropout (finetune.py:210: 0.05) is applied by the engine
in training mode as a counter-hash keep mask on the LoRA-branch input (csrc/ns_lora.cu); `model.eval()` switches it off.
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


PEFT_PREFIX = "base_model.model."      # PeftModel.base_model (LoraModel) .model (the wrapped WhisperForConditionalGeneration)


def adapter_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """What PEFT's `get_peft_model_state_dict` puts into `adapter_model.safetensors` / `.bin` (finetune.py:282 `save_pretrained`,
    utils/callback.py:11-22): LoRA A/B and the modules_to_save copies, keys prefixed `base_model.model.`, with the ADAPTER NAME
    STRIPPED -- `...q_proj.lora_A.weight` (not `.lora_A.default.weight`) and `...encoder.conv1.0.weight` (not
    `...conv1.modules_to_save.default.0.weight`).  (PEFT itself is absent from this image: layout restated, SURVEY appendix D.)"""
    out = {}
    for k, v in model.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k:
            out[PEFT_PREFIX + k.replace(".default.", ".")] = v.detach().cpu().clone().contiguous()
        elif ".modules_to_save.default." in k:
            out[PEFT_PREFIX + k.replace("modules_to_save.default.", "")] = v.detach().cpu().clone().contiguous()
    return out


def _adapter_config(model: nn.Module) -> Dict:
    cfg = model._lora_cfg or {}
    targets = [n for n, m in model.named_modules() if isinstance(m, LoraLinear)]
    saves = [n for n, m in model.named_modules() if isinstance(m, ModulesToSaveWrapper)]
    return {"peft_type": "LORA", "task_type": None, "base_model_name_or_path": getattr(model, "name_or_path", None),
            "r": cfg.get("r", 32), "lora_alpha": cfg.get("lora_alpha", 64), "lora_dropout": cfg.get("lora_dropout", 0.0),
            "bias": "none", "fan_in_fan_out": False, "inference_mode": True, "init_lora_weights": True,
            "target_modules": targets,                 # full module names, as finetune.py:194-202 passes them
            "modules_to_save": saves}


def save_adapter(model: nn.Module, path: str, safe_serialization: bool = True):
    """`PeftModel.save_pretrained(path)`: adapter_config.json + adapter_model.safetensors (or .bin for PEFT < 0.7 readers)."""
    import json
    os.makedirs(path, exist_ok=True)
    sd = adapter_state_dict(model)
    if safe_serialization:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(path, "adapter_model.safetensors"), metadata={"format": "pt"})
    else:
        torch.save(sd, os.path.join(path, "adapter_model.bin"))
    json.dump(_adapter_config(model), open(os.path.join(path, "adapter_config.json"), "w"), indent=2)


def _own_key(k: str, own: Dict[str, torch.Tensor]) -> str:
    """Map a checkpoint key (PEFT's stripped layout, or this package's round-1 `.default.` layout) to a parameter name here."""
    if k.startswith(PEFT_PREFIX):
        k = k[len(PEFT_PREFIX):]
    if k in own:
        return k
    if ".lora_A." in k or ".lora_B." in k:
        if ".default." not in k:
            k = k.replace(".lora_A.", ".lora_A.default.").replace(".lora_B.", ".lora_B.default.")
        return k
    # modules_to_save: `<module>.<rest>` -> `<module>.modules_to_save.default.<rest>` for the wrapped module that prefixes the key
    parts = k.split(".")
    for cut in range(len(parts) - 1, 0, -1):
        cand = ".".join(parts[:cut]) + ".modules_to_save.default." + ".".join(parts[cut:])
        if cand in own:
            return cand
    return k


def load_adapter(model: nn.Module, path: str, is_trainable: bool = True):
    """`PeftModel.from_pretrained(model, path)` (finetune.py:182-185, evaluation.py:88, merge_lora.py:43): reads PEFT's on-disk
    layout (adapter_config.json + adapter_model.safetensors | adapter_model.bin); wraps the target linears / modules_to_save first
    when the model is still plain.  Raises KeyError on a tensor that has no home (never drops one silently)."""
    import json
    cfg_path = os.path.join(path, "adapter_config.json")
    cfg = json.load(open(cfg_path)) if os.path.exists(cfg_path) else {}
    if cfg.get("peft_type", "LORA") != "LORA":
        raise NotImplementedError(f"adapter type {cfg.get('peft_type')}: load_adapter reads plain LoRA adapters "
                                  "(AdaLoRA state lives in neuspeech1_b200.adalora)")
    if not any(isinstance(m, LoraLinear) for m in model.modules()):
        tm = cfg.get("target_modules")
        if isinstance(tm, (list, tuple)) and tm and all("." not in t for t in tm):   # suffix list: the reference adapts the encoder only
            tm = match_modules_string(model.named_modules(), ["model.encoder"], list(tm))
        lora_inject(model, r=cfg.get("r", 32), lora_alpha=cfg.get("lora_alpha", 64), lora_dropout=cfg.get("lora_dropout", 0.0),
                    target_modules=tm, modules_to_save=cfg.get("modules_to_save") or ())
    st = os.path.join(path, "adapter_model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
    own = dict(model.named_parameters())
    for k, v in sd.items():
        name = _own_key(k, own)
        if name not in own:
            raise KeyError(f"adapter tensor {k!r} matches no parameter of the model (looked for {name!r})")
        if own[name].shape != v.shape:
            raise ValueError(f"adapter tensor {k!r}: shape {tuple(v.shape)} vs parameter {tuple(own[name].shape)}")
        own[name].data.copy_(v)
    if not is_trainable:
        for m in model.modules():
            if isinstance(m, (LoraLinear, ModulesToSaveWrapper)):
                m.requires_grad_(False)
    if hasattr(model, "invalidate_engine"):
        model.invalidate_engine()
    return model

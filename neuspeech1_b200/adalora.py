"""AdaLoRA adapters as the reference configures them (finetune.py:43,205-208: `AdaLoraConfig(init_r=12, target_r=4,
lora_alpha=32, lora_dropout=0.1, orth_reg_weight=0.5, ...)`, the CLI default) on top of the B200 engine.

PEFT is not installable here, so this is a restatement of its `SVDLinear` / `AdaLoraModel.forward` semantics from the
reference's call site -- **parity unpinned** (no PEFT source or golden vectors on this box):

    y    = base(x) + (x @ (A * E).T @ B.T) * lora_alpha / (ranknum + 1e-5)        A (r, in), E (r, 1), B (out, r), ranknum = r
    loss = CE + orth_reg_weight * mean over all A and B of  || A A^T - I ||_F  resp.  || B^T B - I ||_F
    init : A, B ~ N(0, 0.02), E = 0

The reference never calls `update_and_allocate`, so the rank budget stays at init_r: no pruning schedule is needed.

The tcgen05 path is the plain-LoRA one: the engine is built with rank 16 (init_r = 12 padded to the UMMA K step) and trains
on EFFECTIVE operands A_eff = [E * A; 0], B_eff = [B, 0]; this adapter keeps the master (A, E, B) in one flat fp32 buffer,
refreshes the effective operands before a step, maps the engine's gradients back by the chain rule
(dA = E * dA_eff, dE = rowsum(dA_eff * A), dB = dB_eff), adds the regulariser's gradient, and runs the same fused
clip + AdamW kernels over the master buffer; the stem convolutions stay with the engine's own optimizer step, both under ONE
global gradient norm.  lora_dropout (0.1 in the reference's AdaLoraConfig) is the engine's LoRA-branch dropout: the same
counter-based bit plane and kernels as plain LoRA (csrc/ns_lora.cu), applied to the branch input x.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from .engine import ENC_LORA_TARGETS, ModelDims, WhisperEEGEngine, lora_module_name

PAD_RANK = 16


def orth_regulariser(A: torch.Tensor, B: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """|| A A^T - I ||_F + || B^T B - I ||_F for stacks A (n, r, in), B (n, out, r) -> (sum of the 2n norms, dA, dB) with
    d||C||_F / dA = 2 C A / ||C||_F (C symmetric), d/dB = 2 B C / ||C||_F."""
    r = A.shape[1]
    eye = torch.eye(r, dtype=A.dtype, device=A.device)
    Ca = torch.bmm(A, A.transpose(1, 2)) - eye
    Cb = torch.bmm(B.transpose(1, 2), B) - eye
    na = Ca.flatten(1).norm(dim=1).clamp_min(1e-30)
    nb = Cb.flatten(1).norm(dim=1).clamp_min(1e-30)
    dA = 2.0 * torch.bmm(Ca, A) / na[:, None, None]
    dB = 2.0 * torch.bmm(B, Cb) / nb[:, None, None]
    return na.sum() + nb.sum(), dA, dB


class AdaLoraAdapter:
    def __init__(self, dims: ModelDims, params: Dict[str, torch.Tensor], init_r: int = 12, lora_alpha: float = 32.0,
                 orth_reg_weight: float = 0.5, dtype: torch.dtype = torch.bfloat16, device="cuda", seed: int = 0,
                 state: Optional[Dict[str, torch.Tensor]] = None, lora_dropout: float = 0.1, dropout_seed: int = 0):
        assert 0 < init_r <= PAD_RANK, "init_r is padded to one UMMA K step (16)"
        self.r, self.alpha, self.w = init_r, float(lora_alpha), float(orth_reg_weight)
        scale = self.alpha / (init_r + 1e-5)
        # the engine multiplies the LoRA branch by lora_alpha / lora_r of ITS dims: rank 16, alpha chosen to give `scale`
        self.dims = ModelDims(**{**dims.__dict__, "lora_r": PAD_RANK, "lora_alpha": scale * PAD_RANK})
        dev = torch.device(device)
        d, F = dims.d_model, dims.enc_ffn
        self.modules: List[Tuple[str, int, int]] = []            # (module name, in, out) in the engine's layout order
        for i in range(dims.enc_layers):
            for t in ENC_LORA_TARGETS:
                fin = F if t == "fc2" else d
                fout = F if t == "fc1" else d
                self.modules.append((lora_module_name(i, t), fin, fout))
        # master buffer: per module [A (r, in) | E (r) | B (out, r)], 64-float aligned entries
        self.entries: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, fin, fout in self.modules:
            for key, shape in ((".lora_A.default", (init_r, fin)), (".lora_E.default", (init_r, 1)), (".lora_B.default", (fout, init_r))):
                self.entries[name + key] = (off, shape)
                off += (int(math.prod(shape)) + 63) // 64 * 64
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.adam_m = torch.zeros_like(self.flat)
        self.adam_v = torch.zeros_like(self.flat)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        g = torch.Generator().manual_seed(seed)
        for name, fin, fout in self.modules:
            self.param(name + ".lora_A.default").copy_(torch.randn(init_r, fin, generator=g) * 0.02)
            self.param(name + ".lora_B.default").copy_(torch.randn(fout, init_r, generator=g) * 0.02)
        if state is not None:
            self.load_state_dict(state)
        zero_lora = {}
        for name, fin, fout in self.modules:
            zero_lora[name + ".lora_A.default.weight"] = torch.zeros(PAD_RANK, fin)
            zero_lora[name + ".lora_B.default.weight"] = torch.zeros(fout, PAD_RANK)
        self.engine = WhisperEEGEngine(self.dims, params, zero_lora, dtype=dtype, device=dev, lora_dropout=lora_dropout,
                                       dropout_seed=dropout_seed)
        self.opt_step = 0
        self.last_reg = None

    # ---- master parameters -------------------------------------------------------------------------------------------
    def param(self, key: str) -> torch.Tensor:
        off, shape = self.entries[key]
        return self.flat[off: off + int(math.prod(shape))].view(shape)

    def param_grad(self, key: str) -> torch.Tensor:
        off, shape = self.entries[key]
        return self.grad[off: off + int(math.prod(shape))].view(shape)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {k: self.param(k).detach().clone() for k in self.entries}
        for name, _, _ in self.modules:
            sd[name + ".ranknum.default"] = torch.tensor([float(self.r)])
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        for k in self.entries:
            self.param(k).copy_(sd[k].to(self.flat.device, torch.float32).view(self.entries[k][1]))

    # ---- effective operands <-> master --------------------------------------------------------------------------------
    def sync_effective(self):
        eng, r = self.engine, self.r
        for name, _, _ in self.modules:
            A, E, B = self.param(name + ".lora_A.default"), self.param(name + ".lora_E.default"), self.param(name + ".lora_B.default")
            a_eff = eng.trainable(name + ".lora_A.default.weight")
            b_eff = eng.trainable(name + ".lora_B.default.weight")
            torch.mul(A, E, out=a_eff[:r])
            b_eff[:, :r].copy_(B)
        eng._packed = False

    def _map_grads(self):
        """engine.grad (effective operands) -> self.grad (master), + the orthogonality regulariser; the engine's LoRA
        gradient region is cleared so that its optimizer step only moves the stem."""
        eng, r = self.engine, self.r
        self.grad.zero_()
        for name, _, _ in self.modules:
            A, E = self.param(name + ".lora_A.default"), self.param(name + ".lora_E.default")
            da_eff = eng.trainable_grad(name + ".lora_A.default.weight")[:r]
            db_eff = eng.trainable_grad(name + ".lora_B.default.weight")[:, :r]
            torch.mul(da_eff, E, out=self.param_grad(name + ".lora_A.default"))
            self.param_grad(name + ".lora_E.default").copy_((da_eff * A).sum(dim=1, keepdim=True))
            self.param_grad(name + ".lora_B.default").copy_(db_eff)
        reg = torch.zeros((), dtype=torch.float32, device=self.flat.device)
        n_mats = 2 * len(self.modules)
        groups: Dict[Tuple[int, int], List[str]] = {}
        for name, fin, fout in self.modules:
            groups.setdefault((fin, fout), []).append(name)
        for (_fin, _fout), names in groups.items():                 # modules of one shape as one batched product
            A = torch.stack([self.param(n + ".lora_A.default") for n in names])
            B = torch.stack([self.param(n + ".lora_B.default") for n in names])
            s, dA, dB = orth_regulariser(A, B)
            reg = reg + s
            for j, n in enumerate(names):
                self.param_grad(n + ".lora_A.default").add_(dA[j], alpha=self.w / n_mats)
                self.param_grad(n + ".lora_B.default").add_(dB[j], alpha=self.w / n_mats)
        lay = eng.layout
        eng.grad[lay.lora_begin: lay.lora_end].zero_()
        return reg / n_mats

    # ---- the training step --------------------------------------------------------------------------------------------
    def loss_and_grads(self, x: torch.Tensor, labels: torch.Tensor, aug: Optional[dict] = None) -> torch.Tensor:
        """CE + orth_reg_weight * regulariser; gradients of the master parameters in self.grad, of the stem in engine.grad."""
        from . import ops  # noqa: F401  (fails loudly when the library is missing)
        eng = self.engine
        self.sync_effective()
        eng._advance_seed()                      # a new dropout mask per step (oracle.next_dropout_seed)
        eng.pack_trainable()
        ce, _, _ = eng.forward_loss(x, labels, aug=aug, save=True, ce_grad_scale=1.0)
        eng.backward()
        reg = self._map_grads()
        self.last_reg = reg
        return ce + self.w * reg

    def train_step(self, x: torch.Tensor, labels: torch.Tensor, lr: float, aug: Optional[dict] = None, all_reduce=None,
                   max_grad_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> torch.Tensor:
        from . import ops
        eng = self.engine
        loss = self.loss_and_grads(x, labels, aug=aug)
        if all_reduce is not None:
            all_reduce(self.grad)
            all_reduce(eng.grad)
        # one global norm over master + stem gradients, then the two fused clip + AdamW launches share it
        self.sumsq.zero_()
        ops.sumsq(self.grad, self.sumsq)
        ops.sumsq(eng.grad, self.sumsq)
        self.opt_step += 1
        eng.opt_step = self.opt_step
        ops.adamw_clip(self.flat, self.grad, self.adam_m, self.adam_v, self.sumsq, 1.0, max_grad_norm, lr, betas[0], betas[1], eps,
                       weight_decay, self.opt_step)
        ops.adamw_clip(eng.flat, eng.grad, eng.adam_m, eng.adam_v, self.sumsq, 1.0, max_grad_norm, lr, betas[0], betas[1], eps,
                       weight_decay, self.opt_step)
        eng._packed = False
        return loss

    def merged_lora(self) -> Dict[str, torch.Tensor]:
        """Plain-LoRA view of the adapters (rank r, scale alpha / (r + 1e-5) folded into B) for merge / evaluation."""
        out = {}
        scale = self.alpha / (self.r + 1e-5)
        for name, _, _ in self.modules:
            A, E, B = self.param(name + ".lora_A.default"), self.param(name + ".lora_E.default"), self.param(name + ".lora_B.default")
            out[name + ".lora_A.default.weight"] = (A * E).detach().clone()
            out[name + ".lora_B.default.weight"] = (B * scale).detach().clone()
        return out

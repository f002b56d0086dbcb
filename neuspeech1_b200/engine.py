"""Execution engine of the EEG-conditioned Whisper hot path on B200.

Host-side orchestration (which kernel, on which buffer, in which order) of the path the reference runs through
`utils/load_model.py::WhisperForConditionalGeneration.forward/generate` (utils/load_model.py:371-476, :534-767,
:976-1070, :1072-1351) with PEFT LoRA on the encoder linears (finetune.py:194-212).  All arithmetic happens in
libneuspeech_b200.so (neuspeech1_b200/ops.py); torch only owns the HBM buffers and the stream.

Data layout in HBM (row-major, activations in `dtype` = bf16 or fp32, statistics / master weights / grads in fp32):
  x_cl   (B, T, Cp)        channels-last EEG after the augmentation pass, Cp = eeg_ch rounded up to 16
  a*,z*  (B, T|T/2|S, d)   stem activations / pre-activations (z kept for the GELU backward)
  h      (B*S, d)          residual stream;  qkv (B*S, 3d) packed [q|k|v];  m,z1 (B*S, F)
  t_*    (B*S, r|3r)       LoRA bottlenecks t = s * x A^T (s = alpha/r folded in)
  kv_all (B*S, Nd*2d)      cross-attention K|V of all decoder layers, one GEMM
  trainable parameters live in ONE flat fp32 buffer (LoRA A/B per layer, then the three stem convs) with a flat fp32
  gradient buffer of the same layout: one fused clip+AdamW launch and one all-reduce cover everything.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._abi import ACT_DGELU, ACT_GELU, ACT_NONE, NS_BF16, NS_F32

_NO_TRAIN_GRAPH = bool(os.environ.get("NS_NO_TRAIN_GRAPH"))     # developer switch: launch every kernel of train_step one by one
_NO_PLANE_OVERLAP = not os.environ.get("NS_PLANE_OVERLAP")      # NS_PLANE_OVERLAP=1: draw the next layer's dropout planes on a side stream
                                                                # (measured: 31.35 / 31.74 ms against 31.28 / 31.22 ms on the main stream --
                                                                # the step runs under the power cap, overlap saves no energy)
_TRAIN_PDL = bool(os.environ.get("NS_TRAIN_PDL"))               # experiment: programmatic dependent launch for the training step's GEMM / LN launches
_NO_PDL = bool(os.environ.get("NS_NO_PDL"))                     # developer A/B switch: plain stream-ordered launches in the decode loop
_NO_AR_OVERLAP = bool(os.environ.get("NS_NO_AR_OVERLAP"))       # developer A/B switch: one all-reduce after the whole backward
_NO_MASK_STAGE = bool(os.environ.get("NS_NO_MASK_STAGE"))       # developer A/B switch: mma.sync ns_lora_down / ns_lora_da instead of the mask stages
_NO_GELU_DERIV = bool(os.environ.get("NS_NO_GELU_DERIV"))       # developer A/B switch: save the GELU pre-activation, not the derivative
_ABSORB = os.environ.get("NS_ABSORB")                           # "0" / "1": force the absorbed decode cross-attention off / on (default: by batch size)
_NO_PLANE_FUSE = bool(os.environ.get("NS_NO_PLANE_FUSE"))       # developer A/B switch: dropout planes from ns_dropout_bits launches, not from the mask stage
_NO_FUSED_BWD_B = bool(os.environ.get("NS_NO_FUSED_BWD_B"))     # developer A/B switch: dt = dy B and dB = dy^T t as two passes over dy
_NO_GEMM_MASK = bool(os.environ.get("NS_NO_GEMM_MASK"))         # developer A/B switch: dropout correction pass instead of the masked GEMM product
ENC_LORA_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")


@dataclass
class ModelDims:
    """The WhisperConfig fields the path reads (+ the EEG stem's channel count and the LoRA hyper-parameters)."""
    d_model: int = 512
    enc_layers: int = 6
    dec_layers: int = 6
    enc_heads: int = 8
    dec_heads: int = 8
    enc_ffn: int = 2048
    dec_ffn: int = 2048
    vocab: int = 51865
    max_source_positions: int = 1500
    max_target_positions: int = 448
    eeg_ch: int = 208
    pad_token_id: int = 50257
    eos_token_id: int = 50257
    decoder_start_token_id: int = 50258
    begin_suppress_tokens: Tuple[int, ...] = (220, 50256)
    lora_r: int = 32
    lora_alpha: int = 64

    @property
    def T(self) -> int:
        return self.max_source_positions * 4

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r

    @property
    def Cp(self) -> int:
        return (self.eeg_ch + 15) // 16 * 16

    @property
    def Vp(self) -> int:
        return (self.vocab + 15) // 16 * 16

    @classmethod
    def from_any(cls, o) -> "ModelDims":
        """Build from any object with the same attribute names (e.g. the oracle's Dims or a dict)."""
        get = (lambda k: o[k]) if isinstance(o, dict) else (lambda k: getattr(o, k))
        return cls(**{f: get(f) for f in cls.__dataclass_fields__})


def absorb_query_key(wq: torch.Tensor, bq: torch.Tensor, wk: torch.Tensor, heads: int):
    """Query AND key projection of a cross-attention as one weight for the absorbed decode form (csrc/ns_attention_absorbed.cu):
    with q = (x Wq^T + bq) Dh^-0.5 and K = enc Wk^T (no bias, HF modeling_whisper.py:284-310), the scores of head h are
    q_h . K_h[j] = Q'_h . enc[j] with Q' = x wq_abs^T + bq_abs,
        wq_abs[h*d + n, m] = Dh^-0.5 sum_c Wk[h*Dh + c, n] Wq[h*Dh + c, m],   bq_abs[h*d + n] = Dh^-0.5 sum_c bq[h*Dh + c] Wk[h*Dh + c, n].
    fp32 in, fp32 out ((heads*d, d), (heads*d,)): the caller rounds to the storage type once."""
    d = wq.shape[1]
    dh = wq.shape[0] // heads
    wkh, wqh = wk.view(heads, dh, d), wq.view(heads, dh, d)
    w = torch.einsum("hcn,hcm->hnm", wkh, wqh) * dh ** -0.5
    b = torch.einsum("hc,hcn->hn", bq.view(heads, dh), wkh) * dh ** -0.5
    return w.reshape(heads * d, d).contiguous(), b.reshape(heads * d).contiguous()


def lora_module_name(layer: int, target: str) -> str:
    return f"model.encoder.layers.{layer}." + (target if target.startswith("fc") else f"self_attn.{target}")


class TrainableLayout:
    """Flat layout of the trainable set (finetune.py:176-212): LoRA A/B of 6 linears per encoder layer + 3 stem convs."""

    def __init__(self, dims: ModelDims, with_lora: bool = True):
        self.entries: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        d, r, F = dims.d_model, dims.lora_r, dims.enc_ffn

        def add(name, shape):
            nonlocal off
            n = int(math.prod(shape))
            self.entries[name] = (off, tuple(shape))
            off += (n + 63) // 64 * 64

        self.lora_begin = off
        if with_lora:
            for i in range(dims.enc_layers):
                for t in ("q_proj", "k_proj", "v_proj"):          # A_q, A_k, A_v contiguous -> stacked (3r, d) view
                    add(lora_module_name(i, t) + ".lora_A.default.weight", (r, d))
                for t in ("q_proj", "k_proj", "v_proj"):          # B_q, B_k, B_v contiguous -> stacked (3d, r) view
                    add(lora_module_name(i, t) + ".lora_B.default.weight", (d, r))
                add(lora_module_name(i, "out_proj") + ".lora_A.default.weight", (r, d))
                add(lora_module_name(i, "out_proj") + ".lora_B.default.weight", (d, r))
                add(lora_module_name(i, "fc1") + ".lora_A.default.weight", (r, d))
                add(lora_module_name(i, "fc1") + ".lora_B.default.weight", (F, r))
                add(lora_module_name(i, "fc2") + ".lora_A.default.weight", (r, F))
                add(lora_module_name(i, "fc2") + ".lora_B.default.weight", (d, r))
        self.lora_end = off
        add("model.encoder.conv1.0.weight", (d, dims.eeg_ch, 3)); add("model.encoder.conv1.0.bias", (d,))
        add("model.encoder.conv1.2.weight", (d, d, 3)); add("model.encoder.conv1.2.bias", (d,))
        add("model.encoder.conv2.weight", (d, d, 3)); add("model.encoder.conv2.bias", (d,))
        self.size = off
        if with_lora:
            assert (r * d) % 64 == 0, "stacked q/k/v views need unpadded entries"

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.entries[name]
        return flat[off: off + int(math.prod(shape))].view(shape)


class Workspace:
    """Named, lazily allocated, reused HBM buffers."""

    def __init__(self, device):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        self.gen = 0          # bumped on every (re)allocation: captured CUDA graphs are only replayed while it stands still

    def get(self, name: str, shape, dtype, zero: bool = False) -> torch.Tensor:
        shape = tuple(int(s) for s in shape)
        t = self.bufs.get(name)
        if t is None or t.shape != shape or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self.bufs[name] = t
            self.gen += 1
        return t

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


def _on_device(fn):
    """Run an engine entry point with the engine's GPU current: every ns_* call launches on the CURRENT device's current stream
    (ops._stream), so an engine built for cuda:N in a process whose current device is cuda:0 would otherwise launch on GPU 0
    with GPU N pointers."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped


class WhisperEEGEngine:
    def __init__(self, dims: ModelDims, params: Dict[str, torch.Tensor], lora: Optional[Dict[str, torch.Tensor]] = None,
                 dtype: torch.dtype = torch.bfloat16, device="cuda", lora_dropout: float = 0.0, dropout_seed: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("neuspeech1_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        ops.lib()
        self.dims = dims
        self.dtype = dtype
        self.ns = ops.ns_dtype(dtype)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.ws = Workspace(self.device)
        self.has_lora = lora is not None
        self.fuse_cross_bwd = False      # decoder cross-attention backward through the fused single-pass kernel
        self.layout = TrainableLayout(dims, with_lora=self.has_lora)
        self.flat = torch.zeros(self.layout.size, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros_like(self.flat)
        self.adam_m = torch.zeros_like(self.flat)
        self.adam_v = torch.zeros_like(self.flat)
        self.opt_step = 0
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        # LoRA-branch dropout (finetune.py:210 lora_dropout=0.05): applied by training forwards (save=True) while `training`;
        # the step seed is a device word that train_step advances on the device (a replayed CUDA graph draws a new mask)
        if not 0.0 <= lora_dropout < 1.0:
            raise ValueError(f"lora_dropout must be in [0, 1), got {lora_dropout}")
        self.lora_dropout = float(lora_dropout)
        self.training = True
        self.drop_seed = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._drop_p = 0.0
        self.use_lora_kernels = dtype == torch.bfloat16      # ns_lora_down / ns_lora_da (bf16 storage); else the GEMM kernels
        self.set_dropout_seed(dropout_seed)
        self.P: Dict[str, torch.Tensor] = {}
        with torch.cuda.device(self.device):
            self.load_params(params, lora)

    def set_dropout_seed(self, seed: int):
        """Step seed of the LoRA-branch dropout mask (the next train_step advances it first: oracle.next_dropout_seed)."""
        seed &= 0xFFFFFFFF
        self.drop_seed.fill_(seed - (1 << 32) if seed >= (1 << 31) else seed)

    def _salts(self, layer: int, targets) -> List[int]:
        import zlib
        return [zlib.crc32(lora_module_name(layer, t).encode()) & 0xFFFFFFFF for t in targets]

    def _fast_lora(self, K: int, G: int) -> bool:
        r = self.dims.lora_r
        return self.use_lora_kernels and K % 64 == 0 and r in (8, 16, 32) and G * r * (2 * K + 64) <= 220 * 1024

    _PLANE_ORDER = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1")       # the adapters whose input is d_model wide

    def _bits(self, layer: int, targets, M: int, K: int) -> torch.Tensor:
        """Bit planes (len(targets), M, K/32) of the dropped elements of `targets` in encoder layer `layer`: drawn before the layer's
        forward (`_draw_planes`), kept for the backward consumers.  q, k, v, out_proj and fc1 share one (5, M, d/32) buffer and
        one generator launch; fc2 (K = ffn) has its own."""
        dm = self.dims
        if targets[0] == "fc2":
            return self.ws.get(f"dropbits.{layer}.F", (1, M, (K + 31) // 32), torch.int32)
        buf = self.ws.get(f"dropbits.{layer}.d", (len(self._PLANE_ORDER), M, (dm.d_model + 31) // 32), torch.int32)
        i0 = self._PLANE_ORDER.index(targets[0])
        return buf[i0: i0 + len(targets)]

    def _mask_stage_draws(self, K: int) -> bool:
        """The rank-32 down product of a K-wide input goes through the tcgen05 mask stage, which then also DRAWS the dropout
        plane (ns_epilogue.drop_mode 2) and leaves it in `_bits` for the backward: no generator launch for that width."""
        return (self.use_lora_kernels and self.dims.lora_r == 32 and K % 64 == 0 and self.dtype == torch.bfloat16
                and not _NO_MASK_STAGE and not _NO_PLANE_FUSE)

    def _draw_planes(self, layer: int, M: int):
        """ns_dropout_bits for the adapters of one encoder layer whose planes the forward does not draw on the fly."""
        dm, p = self.dims, self._drop_p
        if not self._mask_stage_draws(dm.d_model):
            ops.dropout_bits(M, dm.d_model, self.drop_seed, self._salts(layer, self._PLANE_ORDER), p, self._bits(layer, ("q_proj", "k_proj", "v_proj", "out_proj", "fc1"), M, dm.d_model))
        if not self._mask_stage_draws(dm.enc_ffn):
            ops.dropout_bits(M, dm.enc_ffn, self.drop_seed, self._salts(layer, ("fc2",)), p, self._bits(layer, ("fc2",), M, dm.enc_ffn))

    def _planes_ahead(self, layer: int, M: int):
        """Draw layer `layer`'s planes on a side stream, forked here: the generator is pure integer arithmetic (no memory reads,
        issue-bound) and shares the SMs with whatever the main stream runs meanwhile -- the previous layer's attention leaves a
        third of the register file and most issue slots free.  `_planes_join` makes the main stream wait for them."""
        if self._drop_p == 0.0 or layer >= self.dims.enc_layers:
            return
        if _NO_PLANE_OVERLAP:
            self._draw_planes(layer, M)
            return
        side = self.__dict__.get("_plane_stream")
        if side is None:
            side = self._plane_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._draw_planes(layer, M)

    def _planes_join(self):
        if self._drop_p > 0.0 and not _NO_PLANE_OVERLAP and self.__dict__.get("_plane_stream") is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._plane_stream)

    def _lora_down(self, x: torch.Tensor, A: torch.Tensor, t: torch.Tensor, layer: int, targets):
        """t[:, g*r:(g+1)*r] = alpha' * dropout_g(x) A_g^T for the adapters `targets` stacked in A (PEFT lora.Linear:
        lora_A(lora_dropout(x)) * scaling); alpha' = (alpha/r) / (1 - p).  Without dropout this is the thin tcgen05 GEMM; with
        it the mask is drawn once into a bit plane and (bf16) ns_lora_down applies it on the way from HBM to the MMA."""
        p, r, G = self._drop_p, self.dims.lora_r, len(targets)
        a = self.dims.lora_scale / (1.0 - p)
        M, K = x.shape
        if p == 0.0:
            ops.gemm_nt(x, A, t, self._ep(alpha=a, alpha_cols=G * r))
            return
        bits = self._bits(layer, targets, M, K)           # drawn by _draw_planes before the layer started
        if self._mask_stage_draws(K):
            # the thin tcgen05 GEMM whose mask stage (between TMA and MMA, one 32-column tile per adapter) draws the planes
            ops.gemm_nt(x, A, t, self._ep(alpha=a, alpha_cols=G * r, drop_a=bits, drop_gen=(self.drop_seed, self._salts(layer, targets), p)))
        elif self.use_lora_kernels and r == 32 and K % 64 == 0 and not _NO_MASK_STAGE:
            ops.gemm_nt(x, A, t, self._ep(alpha=a, alpha_cols=G * r, drop_a=bits))
        elif self._fast_lora(K, G):
            ops.lora_down(x, A, t, a, G, bits)
        else:
            xm = self.ws.get(f"xm.{K}", x.shape, self.dtype)
            for g in range(G):
                ops.dropout_apply(x, xm, bits[g])
                ops.gemm_nt(xm, A[g * r:(g + 1) * r], t[:, g * r:(g + 1) * r], self._ep(alpha=a, alpha_cols=r))

    def _gelu_deriv(self, with_fc2_lora: bool = False) -> int:
        """1: the GELU forward saves gelu'(z) and the backward multiplies by it (ns_epilogue.aux_deriv) -- bf16 storage only; the
        parity mode keeps the pre-activation.  The encoder MLP under LoRA-branch dropout needs the masked second product for it
        (the correction pass reads z as a pre-activation)."""
        if self.dtype != torch.bfloat16 or _NO_GELU_DERIV:
            return 0
        if with_fc2_lora and self.has_lora and self._drop_p > 0.0:
            dm = self.dims
            ok = (self.use_lora_kernels and dm.enc_ffn % 64 == 0 and not _NO_GEMM_MASK and dm.lora_r % 16 == 0 and dm.d_model % 16 == 0)
            return 1 if ok else 0
        return 1

    def _drop_plane(self, layer: int, target: str, M: int, K: int, N: Optional[int] = None) -> Optional[torch.Tensor]:
        """The (M, K/32) dropout plane of ONE adapter for the input-gradient GEMM's masked second product (ns_epilogue.drop_bits:
        dx = g W + keep . (dt' A), the LoRA product masked in the epilogue), or None when the step runs without dropout or the
        shape / storage takes the correction pass instead (`_lora_da_fix(dx=...)`)."""
        if self._drop_p == 0.0 or not self.use_lora_kernels or K % 64 != 0 or _NO_GEMM_MASK:
            return None
        if self.dims.lora_r % 16 != 0 or (N is not None and N % 16 != 0):     # tcgen05 operand rules of both products
            return None
        return self._bits(layer, (target,), M, K)[0]

    def _fused_bwd_b(self, N: int, M: int, groups: int = 1):
        """Workspace of ns_lora_bwd_b for this shape (None: no workspace needed), or False when the one-pass kernel does not take it
        (fp32 parity storage, rank != 32, N not a multiple of 128)."""
        if _NO_FUSED_BWD_B or self.dtype != torch.bfloat16:
            return False
        n = ops.lora_bwd_b_workspace_bytes(M, N, self.dims.lora_r, groups)
        if n < 0:
            return False
        if n == 0:
            return None
        return self.ws.get(f"lora_bwd_b_ws.{n}", (n,), torch.uint8, zero=True)       # tickets zeroed once; the kernel leaves them zero

    def _lora_bwd_b(self, dy: torch.Tensor, Bt: torch.Tensor, t: torch.Tensor, dt_out: torch.Tensor, dB: torch.Tensor, s: float):
        """dt = alpha' dy B and dB += dy^T t of one adapter: one pass over dy when the shape qualifies, else two products."""
        r = self.dims.lora_r
        M, N = dy.shape
        wsb = self._fused_bwd_b(N, M)
        if wsb is not False:
            ops.lora_bwd_b(dy, Bt, t, dt_out, dB, N, r, [s], [1.0], workspace=wsb)
        else:
            ops.gemm_nt(dy, Bt, dt_out, self._ep(alpha=s, alpha_cols=r))
            ops.gemm_tn(dy, t, dB, r, 1)

    def _lora_da_fix(self, x: torch.Tensor, dt: torch.Tensor, dx: Optional[torch.Tensor], At: torch.Tensor, layer: int, targets,
                     z: Optional[torch.Tensor] = None):
        """dA_g += dt_g^T dropout_g(x) into the flat gradient buffer (the A gradients of `targets` are contiguous) and, under
        dropout with `dx` given, the correction of the input gradient: the GEMM that produced `dx` added dt' A for EVERY element
        (K-segment); the dropped ones are taken out again (times gelu'(z) where dx went through the GELU backward).  dx = None:
        the GEMM already masked its LoRA product (`_drop_plane`)."""
        p, r, G = self._drop_p, self.dims.lora_r, len(targets)
        M, K = x.shape
        off, _ = self.layout.entries[lora_module_name(layer, targets[0]) + ".lora_A.default.weight"]
        gA = self.grad[off: off + G * r * K].view(G * r, K)
        if p == 0.0:
            ops.gemm_tn(x, dt, gA, 1, K)
            return
        bits = self._bits(layer, targets, M, K)
        if dx is None and G == 1 and self.use_lora_kernels and K % 64 == 0 and not _NO_MASK_STAGE:
            ops.gemm_tn_masked(x, dt, gA, 1, K, bits[0])          # split-K tcgen05 wgrad, x masked in shared memory
        elif self.use_lora_kernels and K % 64 == 0 and r in (8, 16, 32):     # (ns_lora_da keeps A^T in registers: no K limit)
            ops.lora_da(x, dt, gA, G, bits, dx=dx, At=At if dx is not None else None, z=z if dx is not None else None)
        else:
            xm = self.ws.get(f"xm.{K}", x.shape, self.dtype)
            for g in range(G):
                ops.dropout_apply(x, xm, bits[g])
                ops.gemm_tn(xm, dt[:, g * r:(g + 1) * r], gA[g * r:(g + 1) * r], 1, K)
            if dx is not None:
                ops.lora_dx_fix(dx, dt, At, bits, G, z)

    # ------------------------------------------------------------------ parameters
    def _c(self, t: torch.Tensor) -> torch.Tensor:
        """fp32 device tensor -> compute-dtype copy (one ns_cast launch)."""
        t = t.contiguous()
        if self.dtype == torch.float32:
            return t.clone()
        out = torch.empty(t.shape, dtype=self.dtype, device=self.device)
        return ops.cast(t, out)

    def _ct(self, t: torch.Tensor, scale: float = 1.0, pad_to: Optional[int] = None) -> torch.Tensor:
        """fp32 (rows, cols) -> compute-dtype transposed copy (cols, rows[, padded])."""
        t = t.contiguous()
        rows, cols = t.shape
        out = torch.empty((cols, pad_to or rows), dtype=self.dtype, device=self.device)
        return ops.transpose(t, out, scale)

    @_on_device
    def load_params(self, params: Dict[str, torch.Tensor], lora: Optional[Dict[str, torch.Tensor]] = None):
        """(Re)load weights.  `params`: HF-named fp32 tensors (see oracle.init_params / state_dict of the HF model)."""
        dm = self.dims
        d = dm.d_model
        dev = self.device
        f32 = lambda k: params[k].detach().to(dev, torch.float32).contiguous()
        W: Dict[str, torch.Tensor] = {}
        for name in self.layout.entries:
            src = lora if (lora is not None and name in lora) else params
            self.layout.view(self.flat, name).copy_(src[name].detach().to(dev, torch.float32))
        W["enc_pos"] = self._c(f32("model.encoder.embed_positions.weight"))
        eh_scale = (d // dm.enc_heads) ** -0.5
        dh_scale = (d // dm.dec_heads) ** -0.5
        zeros_d = torch.zeros(d, dtype=torch.float32, device=dev)

        def ln(pre, key):
            W[key + ".g"] = f32(pre + ".weight"); W[key + ".b"] = f32(pre + ".bias")

        def qscaled_t(wq, wk, wv, scale):
            """transposed [q;k;v] weight with the q rows pre-multiplied by `scale` (backward through q * Dh^-0.5)."""
            return self._ct(torch.cat([wq * scale, wk, wv], dim=0))

        for i in range(dm.enc_layers):
            pre = f"model.encoder.layers.{i}"
            k = f"enc{i}"
            wq, wk, wv = (f32(f"{pre}.self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj"))
            W[k + ".wqkv"] = self._c(torch.cat([wq, wk, wv], dim=0))
            W[k + ".bqkv"] = torch.cat([f32(f"{pre}.self_attn.q_proj.bias"), zeros_d, f32(f"{pre}.self_attn.v_proj.bias")])
            W[k + ".wqkv_t"] = qscaled_t(wq, wk, wv, eh_scale)
            wo = f32(f"{pre}.self_attn.out_proj.weight")
            W[k + ".wo"] = self._c(wo); W[k + ".wo_t"] = self._ct(wo); W[k + ".bo"] = f32(f"{pre}.self_attn.out_proj.bias")
            w1 = f32(f"{pre}.fc1.weight"); w2 = f32(f"{pre}.fc2.weight")
            W[k + ".w1"] = self._c(w1); W[k + ".w1_t"] = self._ct(w1); W[k + ".b1"] = f32(f"{pre}.fc1.bias")
            W[k + ".w2"] = self._c(w2); W[k + ".w2_t"] = self._ct(w2); W[k + ".b2"] = f32(f"{pre}.fc2.bias")
            ln(pre + ".self_attn_layer_norm", k + ".ln1"); ln(pre + ".final_layer_norm", k + ".ln2")
        ln("model.encoder.layer_norm", "enc.lnf")

        E = f32("model.decoder.embed_tokens.weight")
        W["dec.E"] = self._c(E)
        W["dec.E_t"] = self._ct(E, pad_to=dm.Vp)                        # (d, Vp) for dy = dlogits @ E
        W["dec.pos"] = self._c(f32("model.decoder.embed_positions.weight"))
        kv_w, kv_b = [], []
        for i in range(dm.dec_layers):
            pre = f"model.decoder.layers.{i}"
            k = f"dec{i}"
            wq, wk, wv = (f32(f"{pre}.self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj"))
            W[k + ".wqkv"] = self._c(torch.cat([wq, wk, wv], dim=0))
            W[k + ".bqkv"] = torch.cat([f32(f"{pre}.self_attn.q_proj.bias"), zeros_d, f32(f"{pre}.self_attn.v_proj.bias")])
            W[k + ".wqkv_t"] = qscaled_t(wq, wk, wv, dh_scale)
            wo = f32(f"{pre}.self_attn.out_proj.weight")
            W[k + ".wo"] = self._c(wo); W[k + ".wo_t"] = self._ct(wo); W[k + ".bo"] = f32(f"{pre}.self_attn.out_proj.bias")
            wqc = f32(f"{pre}.encoder_attn.q_proj.weight")
            W[k + ".wqc"] = self._c(wqc); W[k + ".wqc_t"] = self._ct(wqc, scale=dh_scale)
            W[k + ".bqc"] = f32(f"{pre}.encoder_attn.q_proj.bias")
            kv_w += [f32(f"{pre}.encoder_attn.k_proj.weight"), f32(f"{pre}.encoder_attn.v_proj.weight")]
            kv_b += [zeros_d, f32(f"{pre}.encoder_attn.v_proj.bias")]
            woc = f32(f"{pre}.encoder_attn.out_proj.weight")
            W[k + ".woc"] = self._c(woc); W[k + ".woc_t"] = self._ct(woc); W[k + ".boc"] = f32(f"{pre}.encoder_attn.out_proj.bias")
            w1 = f32(f"{pre}.fc1.weight"); w2 = f32(f"{pre}.fc2.weight")
            W[k + ".w1"] = self._c(w1); W[k + ".w1_t"] = self._ct(w1); W[k + ".b1"] = f32(f"{pre}.fc1.bias")
            W[k + ".w2"] = self._c(w2); W[k + ".w2_t"] = self._ct(w2); W[k + ".b2"] = f32(f"{pre}.fc2.bias")
            ln(pre + ".self_attn_layer_norm", k + ".ln1"); ln(pre + ".encoder_attn_layer_norm", k + ".ln2")
            ln(pre + ".final_layer_norm", k + ".ln3")
        ln("model.decoder.layer_norm", "dec.lnf")
        # absorbed cross-attention of the decode step (csrc/ns_attention_absorbed.cu): query and key projections in one weight,
        # wq_abs[h*d + n, m] = Dh^-0.5 sum_c Wk[h*Dh + c, n] Wq[h*Dh + c, m], bq_abs[h*d + n] = Dh^-0.5 sum_c bq[h*Dh + c] Wk[h*Dh + c, n]
        # (products in fp32, rounded to the storage type once)
        Hd = dm.dec_heads
        for i in range(dm.dec_layers if (self.dtype == torch.bfloat16 and d == 512 and Hd <= 8) else 0):   # the shapes the kernel takes
            pre = f"model.decoder.layers.{i}"
            w_abs, b_abs = absorb_query_key(f32(f"{pre}.encoder_attn.q_proj.weight"), f32(f"{pre}.encoder_attn.q_proj.bias"), kv_w[2 * i], Hd)
            W[f"dec{i}.wq_abs"] = self._c(w_abs)
            W[f"dec{i}.bq_abs"] = b_abs
        wkv = torch.cat(kv_w, dim=0)                                     # (Nd*2d, d): [k0; v0; k1; v1; ...]
        W["dec.wkv"] = self._c(wkv); W["dec.wkv_t"] = self._ct(wkv); W["dec.bkv"] = torch.cat(kv_b)
        self.P = W
        self._decode_graphs = {}
        self._train_graphs = {}
        self.graph_launches = getattr(self, "graph_launches", 0)   # ns_* launches executed through CUDA-graph replays
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        self._lora_tb = None             # LoRA operand views / transpose job table are rebuilt by pack_trainable
        self.suppress = torch.tensor(list(dm.begin_suppress_tokens), dtype=torch.int32, device=dev)
        self._packed = False

    def trainable(self, name: str) -> torch.Tensor:
        return self.layout.view(self.flat, name)

    def trainable_grad(self, name: str) -> torch.Tensor:
        return self.layout.view(self.grad, name)

    @_on_device
    def pack_trainable(self):
        """Refresh the compute-dtype copies of the trainable weights (LoRA A/B + transposes, stem tap layouts)."""
        dm, lay, W = self.dims, self.layout, self.P
        d, r, F = dm.d_model, dm.lora_r, dm.enc_ffn
        if self.has_lora:
            nl = lay.lora_end - lay.lora_begin
            if self.dtype == torch.float32:
                lc = self.flat[lay.lora_begin: lay.lora_end]
            else:
                lc = ops.cast(self.flat[lay.lora_begin: lay.lora_end], self.ws.get("lora_c", (nl,), self.dtype))
            if getattr(self, "_lora_tb", None) is None or self._lora_tb_ptr != lc.data_ptr():
                # operand views + the (src, dst) list of the 10 transposes per layer; the buffers never move, so the job table
                # is built once and every later refresh is ONE ns_transpose_batched launch
                def cv(name):
                    off, shape = lay.entries[name]
                    return lc[off - lay.lora_begin: off - lay.lora_begin + int(math.prod(shape))].view(shape)

                pairs = []
                for i in range(dm.enc_layers):
                    k = f"enc{i}"
                    off_a = lay.entries[lora_module_name(i, "q_proj") + ".lora_A.default.weight"][0] - lay.lora_begin
                    off_b = lay.entries[lora_module_name(i, "q_proj") + ".lora_B.default.weight"][0] - lay.lora_begin
                    W[k + ".A_qkv"] = lc[off_a: off_a + 3 * r * d].view(3 * r, d)        # stacked [Aq;Ak;Av]
                    W[k + ".B_qkv"] = lc[off_b: off_b + 3 * d * r].view(3 * d, r)        # stacked [Bq;Bk;Bv]
                    W[k + ".A_qkv_t"] = self.ws.get(k + ".A_qkv_t", (d, 3 * r), self.dtype)
                    pairs.append((W[k + ".A_qkv"], W[k + ".A_qkv_t"]))
                    # [B_q^T * Dh^-0.5; B_k^T; B_v^T] stacked (3r, d): the block-diagonal dt = [dq|dk|dv] . B of the backward is one
                    # launch (q's gradient carries the Dh^-0.5 of the forward: folded into this copy)
                    W[k + ".B_qkv_t"] = self.ws.get(k + ".B_qkv_t", (3 * r, d), self.dtype)
                    for g in range(3):
                        pairs.append((W[k + ".B_qkv"][g * d:(g + 1) * d], W[k + ".B_qkv_t"][g * r:(g + 1) * r],
                                      (d // dm.enc_heads) ** -0.5 if g == 0 else 1.0))
                    for t, fin, fout in (("out_proj", d, d), ("fc1", d, F), ("fc2", F, d)):
                        a = cv(lora_module_name(i, t) + ".lora_A.default.weight")
                        b = cv(lora_module_name(i, t) + ".lora_B.default.weight")
                        W[f"{k}.A_{t}"] = a; W[f"{k}.B_{t}"] = b
                        W[f"{k}.A_{t}_t"] = self.ws.get(f"{k}.A_{t}_t", (fin, r), self.dtype)
                        W[f"{k}.B_{t}_t"] = self.ws.get(f"{k}.B_{t}_t", (r, fout), self.dtype)
                        pairs += [(a, W[f"{k}.A_{t}_t"]), (b, W[f"{k}.B_{t}_t"])]
                self._lora_tb = ops.TransposeBatch(pairs, self.device)
                self._lora_tb_ptr = lc.data_ptr()
            self._lora_tb.run()
        Cp = dm.Cp
        for key, name, cin, cp in (("stemA", "model.encoder.conv1.0", dm.eeg_ch, Cp), ("stemB", "model.encoder.conv1.2", d, d),
                                   ("stemC", "model.encoder.conv2", d, d)):
            w = self.trainable(name + ".weight")
            wt = self.ws.get(key + ".w", (3, d, cp), self.dtype)
            wtt = self.ws.get(key + ".wt", (3, cp, d), self.dtype)
            ops.conv_weight_pack(w, wt, wtt)
            W[key + ".w"] = wt; W[key + ".wt"] = wtt; W[key + ".b"] = self.trainable(name + ".bias")
        self._packed = True

    # ------------------------------------------------------------------ encoder forward
    def _ep(self, **kw):
        kw.setdefault("out_dtype", self.ns)
        return ops.epilogue(**kw)

    def input_to_channels_last(self, x: torch.Tensor, aug: Optional[dict] = None) -> torch.Tensor:
        """(B,C,T) fp32 -> (B,T,Cp) compute dtype through the augmentation/pad/cast pass (identity when aug is None).
        With aug["src_off"] / aug["src_ld"], x is the flat ragged sample store (reader.SampleStore.flat, fp32 or bf16) and the
        batch is gathered from it by the same pass."""
        dm = self.dims
        if aug is not None and "src_off" in aug:
            B = aug["src_off"].shape[0]
            y = self.ws.get("x_cl", (B, dm.T, dm.Cp), self.dtype)
            ops.aug_pass(x, y, layout=1, C_in=dm.eeg_ch, Tin=dm.T, **aug)
            return y
        B = x.shape[0]
        if x.shape[1] != dm.eeg_ch:
            raise ValueError(f"expected {dm.eeg_ch} EEG channels, got {x.shape[1]}")
        y = self.ws.get("x_cl", (B, dm.T, dm.Cp), self.dtype)
        x = x.to(self.device, torch.float32).contiguous()
        ops.aug_pass(x, y, layout=1, **(aug or {}))
        return y

    @_on_device
    def encode(self, x: torch.Tensor, aug: Optional[dict] = None, save: bool = False) -> torch.Tensor:
        """input_features (B, eeg_ch, T) -> encoder_last_hidden_state (B, S, d).  utils/load_model.py:371-476."""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        if x.shape[-1] != dm.T and (aug is None or "n" not in aug):
            raise ValueError(f"Whisper expects the input features to be of length {dm.T}, but found {x.shape[-1]}")  # HF:613
        if not self._packed:
            self.pack_trainable()
        B = aug["src_off"].shape[0] if (aug is not None and "src_off" in aug) else x.shape[0]
        d, S, T, F, r, H = dm.d_model, dm.max_source_positions, dm.T, dm.enc_ffn, dm.lora_r, dm.enc_heads
        M = B * S
        self._drop_p = self.lora_dropout if (save and self.training and self.has_lora) else 0.0
        self._planes_ahead(0, M)                      # layer 0's dropout planes: drawn beside the augmentation pass and the stem
        xcl = self.input_to_channels_last(x, aug)
        zA = ws.get("zA", (B, T, d), dt); aA = ws.get("aA", (B, T, d), dt)
        ops.conv3_fwd(xcl, W["stemA.w"], aA, 1, self._ep(bias=W["stemA.b"], act=ACT_GELU, aux_out=zA if save else None, ldaux=d, aux_deriv=self._gelu_deriv()))
        zB = ws.get("zB", (B, T // 2, d), dt); aB = ws.get("aB", (B, T // 2, d), dt)
        ops.conv3_fwd(aA, W["stemB.w"], aB, 2, self._ep(bias=W["stemB.b"], act=ACT_GELU, aux_out=zB if save else None, ldaux=d, aux_deriv=self._gelu_deriv()))
        zC = ws.get("zC", (B, S, d), dt)
        h = ws.get("h0", (M, d), dt)
        ops.conv3_fwd(aB, W["stemC.w"], h.view(B, S, d), 2,
                      self._ep(bias=W["stemC.b"], act=ACT_GELU, aux_out=zC if save else None, ldaux=d, residual=W["enc_pos"], ldr=d, res_mod=S))
        Dh = d // H
        shp = ops.attn_shape(B, H, S, S, Dh, False, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * d, d)
        for i in range(dm.enc_layers):
            k = f"enc{i}"
            self._planes_join()                   # this layer's dropout planes are ready ...
            self._planes_ahead(i + 1, M)          # ... and the next layer's are drawn on the side stream while this one runs
            sfx = f".{i}" if save else ""        # per-layer buffers only when the backward needs them
            u1 = ws.get("u1" + sfx, (M, d), dt)
            mean1 = ws.get("mean1" + sfx, (M,), torch.float32); rstd1 = ws.get("rstd1" + sfx, (M,), torch.float32)
            ops.layernorm_fwd(h, W[k + ".ln1.g"], W[k + ".ln1.b"], u1, mean1, rstd1)
            qkv = ws.get("qkv" + sfx, (M, 3 * d), dt)
            if self.has_lora:
                t_qkv = ws.get("t_qkv" + sfx, (M, 3 * r), dt)
                self._lora_down(u1, W[k + ".A_qkv"], t_qkv, i, ("q_proj", "k_proj", "v_proj"))
                ops.gemm_nt(u1, W[k + ".wqkv"], qkv, self._ep(bias=W[k + ".bqkv"], alpha=Dh ** -0.5, alpha_cols=d, a2_group_cols=d),
                            a2=t_qkv, w2=W[k + ".B_qkv"], k2=r)
            else:
                ops.gemm_nt(u1, W[k + ".wqkv"], qkv, self._ep(bias=W[k + ".bqkv"], alpha=Dh ** -0.5, alpha_cols=d))
            o = ws.get("o" + sfx, (M, d), dt)
            lse = ws.get("lse" + sfx, (B, H, S), torch.float32)
            ops.attention_fwd(shp, qkv, qkv[:, d:], qkv[:, 2 * d:], o, lse)
            hm = ws.get("hm" + sfx, (M, d), dt)
            if self.has_lora:
                t_o = ws.get("t_o" + sfx, (M, r), dt)
                self._lora_down(o, W[k + ".A_out_proj"], t_o, i, ("out_proj",))
                ops.gemm_nt(o, W[k + ".wo"], hm, self._ep(bias=W[k + ".bo"], residual=h, ldr=d), a2=t_o, w2=W[k + ".B_out_proj"], k2=r)
            else:
                ops.gemm_nt(o, W[k + ".wo"], hm, self._ep(bias=W[k + ".bo"], residual=h, ldr=d))
            u2 = ws.get("u2" + sfx, (M, d), dt)
            mean2 = ws.get("mean2" + sfx, (M,), torch.float32); rstd2 = ws.get("rstd2" + sfx, (M,), torch.float32)
            ops.layernorm_fwd(hm, W[k + ".ln2.g"], W[k + ".ln2.b"], u2, mean2, rstd2)
            z1 = ws.get("z1" + sfx, (M, F), dt) if save else None
            m = ws.get("m" + sfx, (M, F), dt)
            if self.has_lora:
                t_1 = ws.get("t_1" + sfx, (M, r), dt)
                self._lora_down(u2, W[k + ".A_fc1"], t_1, i, ("fc1",))
                ops.gemm_nt(u2, W[k + ".w1"], m, self._ep(bias=W[k + ".b1"], act=ACT_GELU, aux_out=z1, ldaux=F, aux_deriv=self._gelu_deriv(True)), a2=t_1,
                            w2=W[k + ".B_fc1"], k2=r)
            else:
                ops.gemm_nt(u2, W[k + ".w1"], m, self._ep(bias=W[k + ".b1"], act=ACT_GELU, aux_out=z1, ldaux=F, aux_deriv=self._gelu_deriv(True)))
            hn = ws.get(f"h{i + 1}" if save else f"h{(i + 1) % 2 + 1}", (M, d), dt)
            if self.has_lora:
                t_2 = ws.get("t_2" + sfx, (M, r), dt)
                self._lora_down(m, W[k + ".A_fc2"], t_2, i, ("fc2",))
                ops.gemm_nt(m, W[k + ".w2"], hn, self._ep(bias=W[k + ".b2"], residual=hm, ldr=d), a2=t_2, w2=W[k + ".B_fc2"], k2=r)
            else:
                ops.gemm_nt(m, W[k + ".w2"], hn, self._ep(bias=W[k + ".b2"], residual=hm, ldr=d))
            if save:
                self._saved_h = getattr(self, "_saved_h", {})
                self._saved_h[i] = h
            h = hn
        enc = ws.get("enc", (M, d), dt)
        meanf = ws.get("meanf", (M,), torch.float32); rstdf = ws.get("rstdf", (M,), torch.float32)
        ops.layernorm_fwd(h, W["enc.lnf.g"], W["enc.lnf.b"], enc, meanf, rstdf)
        self._h_last = h
        self._B = B
        return enc.view(B, S, d)

    # ------------------------------------------------------------------ decoder (teacher forced) + loss
    @staticmethod
    def shift_tokens_right(labels: torch.Tensor, pad: int, start: int) -> torch.Tensor:
        out = torch.empty_like(labels)
        out[:, 1:] = labels[:, :-1]
        out[:, 0] = start
        return out.masked_fill_(out == -100, pad)

    def _decoder_layers_fwd(self, hd: torch.Tensor, B: int, L: int, kv_all: torch.Tensor, save: bool):
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S, F, H = dm.d_model, dm.max_source_positions, dm.dec_ffn, dm.dec_heads
        Dh = d // H
        ML = B * L
        nkv = dm.dec_layers * 2 * d
        shp_s = ops.attn_shape(B, H, L, L, Dh, True, L * 3 * d, 3 * d, L * 3 * d, 3 * d, L * 3 * d, 3 * d, L * d, d)
        shp_c = ops.attn_shape(B, H, L, S, Dh, False, L * d, d, S * nkv, nkv, S * nkv, nkv, L * d, d)
        for i in range(dm.dec_layers):
            k = f"dec{i}"
            sfx = f".{i}" if save else ""
            u = ws.get("d_u", (ML, d), dt)
            m1 = ws.get("d_mean1" + sfx, (ML,), torch.float32); r1 = ws.get("d_rstd1" + sfx, (ML,), torch.float32)
            ops.layernorm_fwd(hd, W[k + ".ln1.g"], W[k + ".ln1.b"], u, m1, r1)
            qkv = ws.get("d_qkv" + sfx, (ML, 3 * d), dt)
            ops.gemm_nt(u, W[k + ".wqkv"], qkv, self._ep(bias=W[k + ".bqkv"], alpha=Dh ** -0.5, alpha_cols=d))
            o = ws.get("d_o" + sfx, (ML, d), dt); lse = ws.get("d_lse" + sfx, (B, H, L), torch.float32)
            ops.attention_fwd(shp_s, qkv, qkv[:, d:], qkv[:, 2 * d:], o, lse)
            h1 = ws.get("d_h1" + sfx, (ML, d), dt)
            ops.gemm_nt(o, W[k + ".wo"], h1, self._ep(bias=W[k + ".bo"], residual=hd, ldr=d))
            m2 = ws.get("d_mean2" + sfx, (ML,), torch.float32); r2 = ws.get("d_rstd2" + sfx, (ML,), torch.float32)
            ops.layernorm_fwd(h1, W[k + ".ln2.g"], W[k + ".ln2.b"], u, m2, r2)
            qc = ws.get("d_qc" + sfx, (ML, d), dt)
            ops.gemm_nt(u, W[k + ".wqc"], qc, self._ep(bias=W[k + ".bqc"], alpha=Dh ** -0.5, alpha_cols=d))
            oc = ws.get("d_oc" + sfx, (ML, d), dt); lsec = ws.get("d_lsec" + sfx, (B, H, L), torch.float32)
            ops.attention_fwd(shp_c, qc, kv_all[:, i * 2 * d:], kv_all[:, i * 2 * d + d:], oc, lsec)
            h2 = ws.get("d_h2" + sfx, (ML, d), dt)
            ops.gemm_nt(oc, W[k + ".woc"], h2, self._ep(bias=W[k + ".boc"], residual=h1, ldr=d))
            m3 = ws.get("d_mean3" + sfx, (ML,), torch.float32); r3 = ws.get("d_rstd3" + sfx, (ML,), torch.float32)
            ops.layernorm_fwd(h2, W[k + ".ln3.g"], W[k + ".ln3.b"], u, m3, r3)
            z = ws.get("d_z" + sfx, (ML, F), dt) if save else None
            mm = ws.get("d_m", (ML, F), dt)
            ops.gemm_nt(u, W[k + ".w1"], mm, self._ep(bias=W[k + ".b1"], act=ACT_GELU, aux_out=z, ldaux=F, aux_deriv=self._gelu_deriv()))
            h3 = ws.get(f"d_h{i + 1}" if save else f"d_h{(i + 1) % 2 + 1}x", (ML, d), dt)
            ops.gemm_nt(mm, W[k + ".w2"], h3, self._ep(bias=W[k + ".b2"], residual=h2, ldr=d))
            if save:
                self._saved_hd = getattr(self, "_saved_hd", {})
                self._saved_hd[i] = hd
            hd = h3
        return hd

    @_on_device
    def forward_loss(self, x: torch.Tensor, labels: Optional[torch.Tensor] = None, decoder_input_ids: Optional[torch.Tensor] = None,
                     aug: Optional[dict] = None, save: bool = True, logits_dtype: Optional[torch.dtype] = None,
                     ce_grad_scale: Optional[float] = None):
        """model(input_features, labels) -> (loss (0-d fp32 tensor or None), logits (B,L,V) view, enc (B,S,d)).
        utils/load_model.py:976-1070."""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S = dm.d_model, dm.max_source_positions
        if decoder_input_ids is not None and labels is not None:
            pass  # HF allows both; labels only drive the loss then
        enc = self.encode(x, aug=aug, save=save)
        B = enc.shape[0]
        if decoder_input_ids is None:
            if labels is None:
                raise ValueError("You have to specify either decoder_input_ids or labels")
            decoder_input_ids = self.shift_tokens_right(labels.to(self.device), dm.pad_token_id, dm.decoder_start_token_id)
        ids = decoder_input_ids.to(self.device, torch.long).contiguous()
        L = ids.shape[1]
        ML = B * L
        nkv = dm.dec_layers * 2 * d
        kv_all = ws.get("kv_all", (B * S, nkv), dt)
        ops.gemm_nt(enc.view(B * S, d), W["dec.wkv"], kv_all, self._ep(bias=W["dec.bkv"]))
        hd = ws.get("d_h0", (ML, d), dt)
        ops.embed(ids, W["dec.E"], W["dec.pos"], 0, hd)
        hd = self._decoder_layers_fwd(hd, B, L, kv_all, save)
        y = ws.get("d_y", (ML, d), dt)
        mf = ws.get("d_meanf", (ML,), torch.float32); rf = ws.get("d_rstdf", (ML,), torch.float32)
        ops.layernorm_fwd(hd, W["dec.lnf.g"], W["dec.lnf.b"], y, mf, rf)
        self._hd_last = hd
        ldt = logits_dtype or dt
        logits = ws.get("logits", (ML, dm.Vp), ldt)
        ops.gemm_nt(y, W["dec.E"], logits, self._ep(out_dtype=ops.ns_dtype(ldt)), N=dm.vocab)
        loss = None
        self._L = L
        if labels is not None:
            lab = labels.to(self.device, torch.long).contiguous().view(-1)
            self._labels = lab
            row_loss = ws.get("row_loss", (ML,), torch.float32)
            loss_sum = ws.get("loss_sum", (1,), torch.float32)
            n_valid = ws.get("n_valid", (1,), torch.int32)
            # ce_grad_scale: the training step never looks at the logits again, so the same launch also turns them into
            # dlogits (softmax - onehot) * scale / n_valid in place and backward() skips its own cross-entropy pass
            fuse = ce_grad_scale is not None and save
            ops.cross_entropy(logits, dm.vocab, lab, row_loss, loss_sum, n_valid, write_grad=fuse,
                              grad_scale=ce_grad_scale if fuse else 1.0)
            self._dlogits_ready = fuse
            loss = (loss_sum / n_valid.to(torch.float32).clamp_min(1.0)).squeeze(0)
        return loss, logits.view(B, L, dm.Vp)[:, :, :dm.vocab], enc

    # ------------------------------------------------------------------ backward
    @_on_device
    def backward(self, grad_scale: float = 1.0, part: Optional[int] = None):
        """Gradients of the mean token cross-entropy w.r.t. the trainable set into self.grad (flat fp32).
        Must follow forward_loss(..., labels=..., save=True).  Mirrors autograd through utils/load_model.py:976-1070 with
        every non-LoRA / non-stem weight frozen (finetune.py:176-177).
        part: None = everything; 0 = loss .. encoder layers (every LoRA gradient is final after it: `layout.lora_end`), 1 = the
        stem.  The data-parallel step runs them apart so that the all-reduce of the LoRA gradients overlaps the stem backward."""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S, T, F, r, H = dm.d_model, dm.max_source_positions, dm.T, dm.enc_ffn, dm.lora_r, dm.enc_heads
        B, L = self._B, self._L
        M, ML = B * S, B * L
        s = dm.lora_scale / (1.0 - self._drop_p)         # dt' = alpha' g B (the 1/(1-p) of the dropped branch input lives here)
        if part == 1:
            return self._backward_stem(self._dh_stem)
        self.grad.zero_()
        # ---- loss -> logits (in place) -> decoder output
        logits = ws.bufs["logits"]
        if getattr(self, "_dlogits_ready", False):
            self._dlogits_ready = False          # forward_loss(ce_grad_scale=...) already left dlogits in the buffer
        else:
            ops.cross_entropy(logits, dm.vocab, self._labels, ws.bufs["row_loss"], None, ws.bufs["n_valid"], write_grad=True, grad_scale=grad_scale)
        if logits.dtype != dt:
            dl = ops.cast(logits, ws.get("dlogits_c", logits.shape, dt))
        else:
            dl = logits
        dy = ws.get("d_dy", (ML, d), dt)
        ops.gemm_nt(dl, W["dec.E_t"], dy, self._ep())
        dh = ws.get("d_dh_a", (ML, d), dt)
        ops.layernorm_bwd(dy, self._hd_last, W["dec.lnf.g"], ws.bufs["d_meanf"], ws.bufs["d_rstdf"], dh)
        # ---- decoder layers (frozen: input gradients only)
        Hd = dm.dec_heads
        Dhd = d // Hd
        nkv = dm.dec_layers * 2 * d
        kv_all = ws.bufs["kv_all"]
        dkv_all = ws.get("dkv_all", (M, nkv), dt)
        delta = ws.get("delta", (B * max(H, Hd) * max(S, L),), torch.float32)
        shp_s = ops.attn_shape(B, Hd, L, L, Dhd, True, L * 3 * d, 3 * d, L * 3 * d, 3 * d, L * 3 * d, 3 * d, L * d, d)
        shp_c = ops.attn_shape(B, Hd, L, S, Dhd, False, L * d, d, S * nkv, nkv, S * nkv, nkv, L * d, d)
        Fd = dm.dec_ffn
        for i in reversed(range(dm.dec_layers)):
            k = f"dec{i}"
            sfx = f".{i}"
            g = lambda n: ws.bufs[n + sfx]
            dm_ = ws.get("d_dm", (ML, Fd), dt)
            ops.gemm_nt(dh, W[k + ".w2_t"], dm_, self._ep(act=ACT_DGELU, aux_in=g("d_z"), ldaux=Fd, aux_deriv=self._gelu_deriv()))
            du = ws.get("d_du", (ML, d), dt)
            ops.gemm_nt(dm_, W[k + ".w1_t"], du, self._ep())
            dh2 = ws.get("d_dh_b", (ML, d), dt)
            ops.layernorm_bwd(du, g("d_h2"), W[k + ".ln3.g"], g("d_mean3"), g("d_rstd3"), dh2, dres=dh)
            doc = ws.get("d_doc", (ML, d), dt)
            ops.gemm_nt(dh2, W[k + ".woc_t"], doc, self._ep())
            dqc = ws.get("d_dqc", (ML, d), dt)
            ops.attention_bwd_ws(shp_c, g("d_qc"), kv_all[:, i * 2 * d:], kv_all[:, i * 2 * d + d:], g("d_oc"), doc, g("d_lsec"), delta,
                                 dqc, dkv_all[:, i * 2 * d:], dkv_all[:, i * 2 * d + d:], self._attn_ws(shp_c) if self.fuse_cross_bwd else None)
            ops.gemm_nt(dqc, W[k + ".wqc_t"], du, self._ep())
            dh1 = ws.get("d_dh_c", (ML, d), dt)
            ops.layernorm_bwd(du, g("d_h1"), W[k + ".ln2.g"], g("d_mean2"), g("d_rstd2"), dh1, dres=dh2)
            dos = ws.get("d_dos", (ML, d), dt)
            ops.gemm_nt(dh1, W[k + ".wo_t"], dos, self._ep())
            dqkv = ws.get("d_dqkv", (ML, 3 * d), dt)
            qkv = g("d_qkv")
            ops.attention_bwd(shp_s, qkv, qkv[:, d:], qkv[:, 2 * d:], g("d_o"), dos, g("d_lse"), delta,
                              dqkv, dqkv[:, d:], dqkv[:, 2 * d:])
            ops.gemm_nt(dqkv, W[k + ".wqkv_t"], du, self._ep())
            ops.layernorm_bwd(du, self._saved_hd[i], W[k + ".ln1.g"], g("d_mean1"), g("d_rstd1"), dh, dres=dh1)
        # ---- cross K/V projections of all layers back to the encoder output
        denc = ws.get("denc", (M, d), dt)
        ops.gemm_nt(dkv_all, W["dec.wkv_t"], denc, self._ep())
        dh = ws.get("dh_a", (M, d), dt)
        ops.layernorm_bwd(denc, self._h_last, W["enc.lnf.g"], ws.bufs["meanf"], ws.bufs["rstdf"], dh)
        # ---- encoder layers
        Dh = d // H
        qs = Dh ** -0.5
        shp = ops.attn_shape(B, H, S, S, Dh, False, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * 3 * d, 3 * d, S * d, d)
        lay = self.layout
        for i in reversed(range(dm.enc_layers)):
            k = f"enc{i}"
            sfx = f".{i}"
            g = lambda n: ws.bufs[n + sfx]
            G = lambda t, ab: self.trainable_grad(lora_module_name(i, t) + f".lora_{ab}.default.weight")
            # fc2
            dz1 = ws.get("dz1", (M, F), dt)
            if self.has_lora:
                dt2 = ws.get("dt_r", (M, r), dt)
                self._lora_bwd_b(dh, W[k + ".B_fc2_t"], g("t_2"), dt2, G("fc2", "B"), s)
                db = self._drop_plane(i, "fc2", M, F, d)
                ops.gemm_nt(dh, W[k + ".w2_t"], dz1, self._ep(act=ACT_DGELU, aux_in=g("z1"), ldaux=F, drop_bits=db, aux_deriv=self._gelu_deriv(True)), a2=dt2,
                            w2=W[k + ".A_fc2_t"], k2=r)
                self._lora_da_fix(g("m"), dt2, dz1 if db is None else None, W[k + ".A_fc2_t"], i, ("fc2",), z=g("z1"))
            else:
                ops.gemm_nt(dh, W[k + ".w2_t"], dz1, self._ep(act=ACT_DGELU, aux_in=g("z1"), ldaux=F, aux_deriv=self._gelu_deriv(True)))
            # fc1
            du2 = ws.get("du", (M, d), dt)
            if self.has_lora:
                dt1 = ws.get("dt_r", (M, r), dt)
                self._lora_bwd_b(dz1, W[k + ".B_fc1_t"], g("t_1"), dt1, G("fc1", "B"), s)
                # (long contraction, narrow output: the 128-wide tiles of the masked product re-read dz1 from L2 twice as often
                # and cost more than the correction pass saves -- measured 234 + 43 us against 154 + 83 us)
                db = self._drop_plane(i, "fc1", M, d, F) if F <= d else None
                ops.gemm_nt(dz1, W[k + ".w1_t"], du2, self._ep(drop_bits=db), a2=dt1, w2=W[k + ".A_fc1_t"], k2=r)
                self._lora_da_fix(g("u2"), dt1, du2 if db is None else None, W[k + ".A_fc1_t"], i, ("fc1",))
            else:
                ops.gemm_nt(dz1, W[k + ".w1_t"], du2, self._ep())
            dhm = ws.get("dh_b", (M, d), dt)
            ops.layernorm_bwd(du2, g("hm"), W[k + ".ln2.g"], g("mean2"), g("rstd2"), dhm, dres=dh)
            # out_proj
            do = ws.get("do", (M, d), dt)
            if self.has_lora:
                dto = ws.get("dt_r", (M, r), dt)
                self._lora_bwd_b(dhm, W[k + ".B_out_proj_t"], g("t_o"), dto, G("out_proj", "B"), s)
                db = self._drop_plane(i, "out_proj", M, d, d)
                ops.gemm_nt(dhm, W[k + ".wo_t"], do, self._ep(drop_bits=db), a2=dto, w2=W[k + ".A_out_proj_t"], k2=r)
                self._lora_da_fix(g("o"), dto, do if db is None else None, W[k + ".A_out_proj_t"], i, ("out_proj",))
            else:
                ops.gemm_nt(dhm, W[k + ".wo_t"], do, self._ep())
            # attention
            dqkv = ws.get("dqkv", (M, 3 * d), dt)
            qkv = g("qkv")
            ops.attention_bwd_ws(shp, qkv, qkv[:, d:], qkv[:, 2 * d:], g("o"), do, g("lse"), delta, dqkv, dqkv[:, d:], dqkv[:, 2 * d:],
                                 self._attn_ws(shp))
            # q/k/v projections (dq carries the Dh^-0.5 of the forward: folded into alpha / pre-scaled transposed weight)
            du1 = du2
            if self.has_lora:
                dtq = ws.get("dt_qkv", (M, 3 * r), dt)
                t_qkv = g("t_qkv")
                # dt_g = alpha' dy_g B_g and dB_g = dy_g^T t_g for q, k, v: block-diagonal products, one launch each
                off_b, _ = lay.entries[lora_module_name(i, "q_proj") + ".lora_B.default.weight"]
                dB_qkv = self.grad[off_b: off_b + 3 * d * r].view(3 * d, r)
                wsb = self._fused_bwd_b(d, M, 3)
                if wsb is not False:
                    ops.lora_bwd_b(dqkv, W[k + ".B_qkv_t"], t_qkv, dtq, dB_qkv, d, r, [s, s, s], [qs, 1.0, 1.0], workspace=wsb)
                else:
                    ops.gemm_nt(dqkv, W[k + ".B_qkv_t"], dtq, self._ep(alpha=s, alpha_cols=3 * r, a_group_cols=r), K=d)
                    ops.gemm_tn_grouped(dqkv, t_qkv, dB_qkv, d, r, r, 1, [qs, 1.0, 1.0])
                # dA for q,k,v in one launch: the three (r,d) gradients are contiguous = one (3r, d) matrix
                ops.gemm_nt(dqkv, W[k + ".wqkv_t"], du1, self._ep(), a2=dtq, w2=W[k + ".A_qkv_t"], k2=3 * r)
                self._lora_da_fix(g("u1"), dtq, du1, W[k + ".A_qkv_t"], i, ("q_proj", "k_proj", "v_proj"))
            else:
                ops.gemm_nt(dqkv, W[k + ".wqkv_t"], du1, self._ep())
            ops.layernorm_bwd(du1, self._saved_h[i], W[k + ".ln1.g"], g("mean1"), g("rstd1"), dh, dres=dhm)
        self._dh_stem = dh
        if part == 0:
            return self.grad
        return self._backward_stem(dh)

    def _backward_stem(self, dh: torch.Tensor):
        """Stem part of `backward` (all three convs trainable; conv A's input needs no gradient)."""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S, T = dm.d_model, dm.max_source_positions, dm.T
        B = self._B
        Cp = dm.Cp
        dzC = ws.get("dzC", (B, S, d), dt)
        ops.dgelu_mul(dh, ws.bufs["zC"], dzC)
        aA, aB, xcl = ws.bufs["aA"], ws.bufs["aB"], ws.bufs["x_cl"]
        self._conv_wgrad("model.encoder.conv2", dzC, aB, 2, d, d)
        dzB = ws.get("dzB", (B, T // 2, d), dt)
        ops.conv3_dgrad(dzC, W["stemC.wt"], dzB, 2, self._ep(act=ACT_DGELU, aux_in=ws.bufs["zB"], ldaux=d, aux_deriv=self._gelu_deriv()))
        self._conv_wgrad("model.encoder.conv1.2", dzB, aA, 2, d, d)
        dzA = ws.get("dzA", (B, T, d), dt)
        ops.conv3_dgrad(dzB, W["stemB.wt"], dzA, 2, self._ep(act=ACT_DGELU, aux_in=ws.bufs["zA"], ldaux=d, aux_deriv=self._gelu_deriv()))
        self._conv_wgrad("model.encoder.conv1.0", dzA, xcl, 1, dm.eeg_ch, Cp)
        return self.grad

    def _attn_ws(self, shp) -> Optional[torch.Tensor]:
        """1024-byte aligned scratch for the fused attention backward (fp32 dQ accumulator + tile-major softmax statistics);
        None when the shape does not qualify (fp32 storage, causal, head_dim != 64) and the two-kernel path runs."""
        if self.dtype != torch.bfloat16:
            return None
        n = ops.attention_bwd_workspace_bytes(shp)
        if n <= 0:
            return None
        buf = self.ws.get(f"attn_bwd_ws.{n}", (n + 1024,), torch.uint8)
        off = (-buf.data_ptr()) % 1024
        return buf[off: off + n]

    def _conv_wgrad(self, name: str, dz: torch.Tensor, xin: torch.Tensor, stride: int, cin: int, cp: int):
        d = self.dims.d_model
        dwt = self.ws.get(f"dw_tap.{cp}", (3, d, cp), torch.float32)
        dwt.zero_()
        ops.conv3_wgrad(dz, xin, dwt, self.trainable_grad(name + ".bias"), stride)
        ops.conv_weight_unpack_grad(dwt, self.trainable_grad(name + ".weight"))

    # ------------------------------------------------------------------ optimizer
    @_on_device
    def optimizer_step(self, lr: float, max_grad_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8,
                       weight_decay: float = 0.0, grad_scale: float = 1.0):
        """clip_grad_norm_(max_grad_norm) + AdamW over the flat buffer (HF trainer.py:2493,1760; finetune.py:236-247).
        Returns the 0-d tensor holding sum(g^2) (pre-clip norm squared) without synchronising."""
        self.sumsq.zero_()
        ops.sumsq(self.grad, self.sumsq)
        self.opt_step += 1
        ops.adamw_clip(self.flat, self.grad, self.adam_m, self.adam_v, self.sumsq, grad_scale, max_grad_norm, lr, betas[0],
                       betas[1], eps, weight_decay, self.opt_step)
        self._packed = False
        return self.sumsq

    @_on_device
    def train_step(self, x, labels, lr: float, aug: Optional[dict] = None, all_reduce=None, use_graph: bool = True):
        """One Trainer.training_step + optimizer step.  `all_reduce(flat_grad)` is the data-parallel hook.

        use_graph: the ~380 launches of pack + forward + backward are replayed as ONE CUDA graph when the step is called
        again on the same input buffers (same addresses and shapes: a double-buffered loader, or device-resident data).
        Launched one by one the kernels leave ~1.3 ms of gaps in a 31 ms step.  The first call on a buffer pair runs
        eagerly (it also warms the workspace up), the second one captures, later ones replay; anything else -- new
        addresses every step, a workspace that grew, reloaded weights -- simply keeps running eagerly.  The gradient
        all-reduce and the three optimizer launches stay outside the graph (learning rate and step count are host values)."""
        if _TRAIN_PDL and not getattr(self, "_in_pdl", False):
            self._in_pdl = True
            prev = ops.set_pdl(True)
            try:
                return self.train_step(x, labels, lr, aug=aug, all_reduce=all_reduce, use_graph=use_graph)
            finally:
                ops.set_pdl(prev)
                self._in_pdl = False
        loss = None
        # data parallel: the LoRA gradients (first part of the flat buffer) are complete before the stem backward starts, so
        # their all-reduce runs on a side stream under it; the stem gradients follow on the main stream
        split = all_reduce is not None and self.has_lora and not _NO_AR_OVERLAP
        if (use_graph and not getattr(self, "_graphs_off", False) and not _NO_TRAIN_GRAPH and ops._prof is None and x.is_cuda and x.is_contiguous()
                and labels.is_cuda and labels.is_contiguous() and labels.dtype == torch.long):
            loss = self._fwd_bwd_graph(x, labels, aug, all_reduce if split else None)
            reduced = split
        if loss is None:
            self._advance_seed()
            self.pack_trainable()
            loss, _, _ = self.forward_loss(x, labels, aug=aug, save=True, ce_grad_scale=1.0)
            if split:
                self.backward(part=0)
                self._reduce_overlapped(all_reduce, lambda: self.backward(part=1))
            else:
                self.backward()
            reduced = split
        if all_reduce is not None and not reduced:
            all_reduce(self.grad)
        self.optimizer_step(lr)
        return loss

    def _reduce_overlapped(self, all_reduce, run_stem):
        """all_reduce(LoRA gradients) on a side stream while `run_stem` (the stem backward) runs on the current one, then
        all_reduce(stem gradients).  Both collectives are issued in the same order on every rank."""
        n = self.layout.lora_end
        cur = torch.cuda.current_stream(self.device)
        side = self.__dict__.get("_ar_stream")
        if side is None:
            side = self._ar_stream = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            all_reduce(self.grad[:n])
        run_stem()
        all_reduce(self.grad[n:])
        cur.wait_stream(side)

    def _advance_seed(self):
        if self.lora_dropout > 0.0 and self.training and self.has_lora:
            ops.seed_advance(self.drop_seed)

    def _fwd_bwd_graph(self, x, labels, aug, all_reduce=None):
        """Replay (or capture) the graph of pack_trainable + forward_loss + backward for these buffers; None = run eagerly.
        With `all_reduce` the step is TWO graphs (.. encoder backward | stem backward) and the gradient all-reduce is issued
        between / after them (`_reduce_overlapped`)."""
        from . import _abi
        akey = tuple(sorted((k, v.data_ptr() if torch.is_tensor(v) else v) for k, v in (aug or {}).items()))
        key = (x.data_ptr(), tuple(x.shape), x.dtype, labels.data_ptr(), tuple(labels.shape), akey, all_reduce is not None)
        stamp = (self.ws.gen, self._weights_version, self.lora_dropout, self.training)
        graphs = self.__dict__.setdefault("_train_graphs", {})
        ent = graphs.get(key)
        if ent is not None and ent[0] is not None and ent[3] == stamp:
            ent[0].replay()
            if all_reduce is not None:
                self._reduce_overlapped(all_reduce, ent[4].replay)
            self.graph_launches += ent[2]
            self._packed = False
            return ent[1].clone()                     # the graph's own output buffer is overwritten by the next replay
        if ent is None or ent[3] != stamp:
            if len(graphs) >= 4:                      # a loader that never reuses a buffer must not pile graphs up
                graphs.pop(next(iter(graphs)))
            graphs[key] = (None, None, 0, stamp)     # seen once: eager now, capture if it comes back unchanged
            return None
        self._packed = False
        torch.cuda.synchronize()
        before = sum(_abi.counters().values())
        g = torch.cuda.CUDAGraph()
        g2 = torch.cuda.CUDAGraph() if all_reduce is not None else None
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):   # loader threads may pin memory meanwhile
                self._advance_seed()
                self.pack_trainable()
                loss, _, _ = self.forward_loss(x, labels, aug=aug, save=True, ce_grad_scale=1.0)
                self.backward(part=None if g2 is None else 0)
            if g2 is not None:
                with torch.cuda.graph(g2, pool=g.pool(), capture_error_mode="thread_local"):
                    self.backward(part=1)
        except Exception as e:                          # something on this path is not capturable: stay eager from now on
            import warnings
            warnings.warn(f"neuspeech1_b200: CUDA-graph capture of the training step failed ({e}); running eagerly")
            self._graphs_off = True
            graphs.pop(key, None)
            torch.cuda.synchronize()
            self._packed = False
            return None
        n = sum(_abi.counters().values()) - before
        if (self.ws.gen, self._weights_version, self.lora_dropout, self.training) != stamp:   # the capture itself allocated: do not trust it
            graphs.pop(key, None)
            return None
        graphs[key] = (g, loss, n, stamp, g2)
        g.replay()
        if g2 is not None:
            self._reduce_overlapped(all_reduce, g2.replay)
        self.graph_launches += n
        self._packed = False
        return loss.clone()

    def _absorbed_decode(self, B: int) -> bool:
        """The decode step attends over the encoder output itself (key / value projections moved to the query / output side,
        csrc/ns_attention_absorbed.cu): half the bytes per position and no cross K|V buffer.  One CTA per sample streams its
        encoder rows, so the batch has to cover the SMs; smaller batches keep one CTA per (sample, head) over cached K|V."""
        dm = self.dims
        ok = self.dtype == torch.bfloat16 and dm.d_model == 512 and dm.dec_heads <= 8
        if _ABSORB is not None:
            return ok and _ABSORB != "0"
        return ok and B >= 96

    def _native_decoder(self, B: int, Tmax: int, cache, kv_layers: Optional[torch.Tensor], logits: torch.Tensor, enc: Optional[torch.Tensor] = None):
        """The `ns_decoder` argument block of ns_decode_prefill / ns_decode_step for this batch shape: pointers into the weight
        table, the self-attention cache, the per-layer cross K|V and persistent scratch.  Cached per (B, Tmax) while the
        workspace and the weights stay in place.  enc (B*S, d) instead of kv_layers: the absorbed form (`_absorbed_decode`)."""
        from . import _abi
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, F = dm.d_model, dm.dec_ffn
        key = (B, Tmax, enc is not None)
        p = lambda t: t.data_ptr()
        tag = f"nd.{B}"
        buf = lambda n, cols=d: ws.get(f"{tag}.{n}", (B, cols), dt)
        sc = [buf("h0"), buf("u"), buf("o"), buf("h1"), buf("qc"), buf("h2"), buf("mm", F), buf("h3a"), buf("h3b"), buf("y")]
        ab = [buf("qp", dm.dec_heads * d), buf("cp", dm.dec_heads * d)] if enc is not None else []
        stamp = (ws.gen, self._weights_version, tuple(p(c) for c in cache), p(kv_layers) if kv_layers is not None else p(enc), p(logits))
        ent = self.__dict__.setdefault("_native_dec", {}).get(key)
        if ent is not None and ent[0] == stamp:
            return ent[1]
        layers = (_abi.DecoderLayer * dm.dec_layers)()
        keep = []
        for i in range(dm.dec_layers):
            k = f"dec{i}"
            wkv = W["dec.wkv"][i * 2 * d:(i + 1) * 2 * d]; bkv = W["dec.bkv"][i * 2 * d:(i + 1) * 2 * d]
            keep += [wkv, bkv]
            layers[i] = _abi.DecoderLayer(p(W[k + ".ln1.g"]), p(W[k + ".ln1.b"]), p(W[k + ".wqkv"]), p(W[k + ".bqkv"]), p(W[k + ".wo"]),
                                          p(W[k + ".bo"]), p(W[k + ".ln2.g"]), p(W[k + ".ln2.b"]), p(W[k + ".wqc"]), p(W[k + ".bqc"]),
                                          p(W[k + ".woc"]), p(W[k + ".boc"]), p(W[k + ".ln3.g"]), p(W[k + ".ln3.b"]), p(W[k + ".w1"]),
                                          p(W[k + ".b1"]), p(W[k + ".w2"]), p(W[k + ".b2"]), p(wkv), p(bkv), p(cache[i]),
                                          p(kv_layers[i]) if kv_layers is not None else None,
                                          p(W[k + ".wq_abs"]) if enc is not None else None, p(W[k + ".bq_abs"]) if enc is not None else None)
        dec = _abi.Decoder(self.ns, dm.dec_layers, d, dm.dec_heads, F, dm.vocab, dm.max_source_positions, Tmax, B, ops.ns_dtype(logits),
                           kv_layers.stride(1) if kv_layers is not None else 0, logits.stride(0), p(W["dec.E"]), p(W["dec.pos"]),
                           p(W["dec.lnf.g"]), p(W["dec.lnf.b"]), layers, *[p(t) for t in sc], p(logits),
                           p(enc) if enc is not None else None, *([p(t) for t in ab] if ab else [None, None]))
        self._native_dec[key] = (stamp, dec, layers, keep, sc, ab)
        return dec

    def _cross_kv_per_layer(self, enc: torch.Tensor, B: int) -> torch.Tensor:
        """Cross-attention K|V for the decode loops as (N_dec, B*S, 2d): layer i's keys and values of one source position are
        2 KB contiguous and consecutive positions follow each other, so the single-query attention of a (sample, layer)
        streams one 3 MB run.  (The training step keeps all layers in one (B*S, N_dec*2d) matrix from one GEMM: there the
        1500-query tiles reuse K/V from L2 and a key's row stride does not matter; in the decode step it is the whole cost --
        2.36 GB per position at B = 128 -- and 128-byte pieces 12 KB apart reached 70 % of the copy bandwidth.)"""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S = dm.d_model, dm.max_source_positions
        kv = ws.get("kv_layers", (dm.dec_layers, B * S, 2 * d), dt)
        x = enc.view(B * S, d)
        for i in range(dm.dec_layers):
            ops.gemm_nt(x, W["dec.wkv"][i * 2 * d:(i + 1) * 2 * d], kv[i], self._ep(bias=W["dec.bkv"][i * 2 * d:(i + 1) * 2 * d]))
        return kv

    # ------------------------------------------------------------------ one decoder pass with KV cache
    def _decode_logits(self, ids: torch.Tensor, pos: int, cache, kv_all: torch.Tensor, logits: torch.Tensor, Tmax: int,
                       beams: int = 1, kv_rows: Optional[torch.Tensor] = None):
        """Decoder pass over `ids` (B, Lq) at cache position `pos` (utils/load_model.py:624,704,740-741, HF
        modeling_whisper.py:314-336): self-attention K/V appended to `cache[i]` (B, Tmax, 3d), cross-attention over the
        precomputed `kv_all` ((B / beams)*S, N_dec*2d); logits of the LAST position -> `logits` (B, Vp).
        beams > 1: rows b*beams + k are the beams of sample b.  They ride in the batch dimension of the self-attention and in
        the QUERY dimension of the cross-attention (beams * Lq independent queries per sample), so a sample's encoder K/V are
        read once per step for all its beams and never copied."""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S, F, H = dm.d_model, dm.max_source_positions, dm.dec_ffn, dm.dec_heads
        Dh = d // H
        nkv = dm.dec_layers * 2 * d
        B, Lq = ids.shape
        MLq = B * Lq
        tag = f"{B}.{Lq}"
        hd = ws.get(f"g_h0.{tag}", (MLq, d), dt)
        ops.embed(ids, W["dec.E"], W["dec.pos"], pos, hd)
        for i in range(dm.dec_layers):
            k = f"dec{i}"
            u = ws.get(f"g_u.{tag}", (MLq, d), dt)
            ops.layernorm_fwd(hd, W[k + ".ln1.g"], W[k + ".ln1.b"], u)
            qkv_new = cache[i][:, pos:pos + Lq]                     # rows (b, pos..pos+Lq) of the cache, written in place
            if Lq == 1:
                ops.gemm_nt(u, W[k + ".wqkv"], qkv_new.view(B, 3 * d) if Tmax == 1 else qkv_new.squeeze(1),
                            self._ep(bias=W[k + ".bqkv"], alpha=Dh ** -0.5, alpha_cols=d))
            else:
                tmp = ws.get(f"g_qkvtmp.{tag}", (MLq, 3 * d), dt)
                ops.gemm_nt(u, W[k + ".wqkv"], tmp, self._ep(bias=W[k + ".bqkv"], alpha=Dh ** -0.5, alpha_cols=d))
                qkv_new.copy_(tmp.view(B, Lq, 3 * d))
            Lk = pos + Lq
            shp_s = ops.attn_shape(B, H, Lq, Lk, Dh, True, Tmax * 3 * d, 3 * d, Tmax * 3 * d, 3 * d, Tmax * 3 * d, 3 * d, Lq * d, d)
            o = ws.get(f"g_o.{tag}", (MLq, d), dt)
            c = cache[i]
            if kv_rows is not None and Lq == 1:
                # beam search: key/value j of row b sits in cache row kv_rows[b, j] (the reorder permutes the table, not the cache)
                ops.attention_decode_rows(shp_s, c[:, pos:], c[:, :, d:], c[:, :, 2 * d:], o, kv_rows)
            else:
                ops.attention_fwd(shp_s, c[:, pos:], c[:, :, d:], c[:, :, 2 * d:], o)
            h1 = ws.get(f"g_h1.{tag}", (MLq, d), dt)
            ops.gemm_nt(o, W[k + ".wo"], h1, self._ep(bias=W[k + ".bo"], residual=hd, ldr=d))
            ops.layernorm_fwd(h1, W[k + ".ln2.g"], W[k + ".ln2.b"], u)
            qc = ws.get(f"g_qc.{tag}", (MLq, d), dt)
            ops.gemm_nt(u, W[k + ".wqc"], qc, self._ep(bias=W[k + ".bqc"], alpha=Dh ** -0.5, alpha_cols=d))
            Lc = beams * Lq
            if kv_all.dim() == 3:                                    # per-layer (N_dec, B*S, 2d) from _cross_kv_per_layer
                shp_c = ops.attn_shape(B // beams, H, Lc, S, Dh, False, Lc * d, d, S * 2 * d, 2 * d, S * 2 * d, 2 * d, Lc * d, d)
                ops.attention_fwd(shp_c, qc, kv_all[i], kv_all[i][:, d:], o)
            else:
                shp_c = ops.attn_shape(B // beams, H, Lc, S, Dh, False, Lc * d, d, S * nkv, nkv, S * nkv, nkv, Lc * d, d)
                ops.attention_fwd(shp_c, qc, kv_all[:, i * 2 * d:], kv_all[:, i * 2 * d + d:], o)
            h2 = ws.get(f"g_h2.{tag}", (MLq, d), dt)
            ops.gemm_nt(o, W[k + ".woc"], h2, self._ep(bias=W[k + ".boc"], residual=h1, ldr=d))
            ops.layernorm_fwd(h2, W[k + ".ln3.g"], W[k + ".ln3.b"], u)
            mm = ws.get(f"g_m.{tag}", (MLq, F), dt)
            ops.gemm_nt(u, W[k + ".w1"], mm, self._ep(bias=W[k + ".b1"], act=ACT_GELU))
            hd2 = ws.get(f"g_h3.{tag}.{i % 2}", (MLq, d), dt)
            ops.gemm_nt(mm, W[k + ".w2"], hd2, self._ep(bias=W[k + ".b2"], residual=h2, ldr=d))
            hd = hd2
        y = ws.get(f"g_y.{tag}", (MLq, d), dt)
        ops.layernorm_fwd(hd, W["dec.lnf.g"], W["dec.lnf.b"], y)
        y_last = y.view(B, Lq, d)[:, Lq - 1]                          # (B, d) view, row stride Lq*d
        ops.gemm_nt(y_last, W["dec.E"], logits, self._ep(out_dtype=ops.ns_dtype(logits)), N=dm.vocab, M=B, K=d)

    # ------------------------------------------------------------------ beam search (evaluation.py:370-385)
    @_on_device
    @torch.no_grad()
    def beam_search(self, x: torch.Tensor, max_length: int, num_beams: int = 5, repetition_penalty: float = 1.0,
                    no_repeat_ngram_size: int = 0, prompt: Optional[torch.Tensor] = None, aug: Optional[dict] = None,
                    length_penalty: float = 1.0, use_graphs: bool = True, sequence_bias=None) -> torch.Tensor:
        """`generate(num_beams=K, repetition_penalty, no_repeat_ngram_size)`: the beams ride in the batch dimension of the
        decoder pass (rows b*K + k) and in the query dimension of its cross-attention.  `_reorder_cache` (utils/load_model.py:
        1353-1360) becomes a permutation of a (rows, positions) table of cache rows that the self-attention reads through
        (ns_attention_decode_rows): no cache bytes move.  The vocabulary-sized scoring (log-softmax, repetition penalty, n-gram
        ban, begin-suppress, top-2K per row) is ONE kernel (ns_beam_row_topk); the host loop of neuspeech1_b200/generation.py only
        handles (B, 2K)-sized tensors.  use_graphs: the decoder pass of every position is captured once into a CUDA graph and
        replayed by later calls of the same shape (as in `greedy`).  sequence_bias (evaluation.py:339-343) takes the tensor-op
        scorer.  Returns the generated suffix (B, <= max_length - prompt_len), pad after EOS."""
        from .generation import beam_search as run_beams
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S = dm.d_model, dm.max_source_positions
        B, K = x.shape[0], num_beams
        if max_length > dm.max_target_positions:
            raise ValueError(f"max_length {max_length} exceeds max_target_positions {dm.max_target_positions}")
        enc = self.encode(x, aug=aug, save=False)
        nkv = dm.dec_layers * 2 * d
        kv_all = self._cross_kv_per_layer(enc, B)
        if prompt is None:
            prompt = torch.full((B, 1), dm.decoder_start_token_id, dtype=torch.long, device=self.device)
        prompt = prompt.to(self.device, torch.long).contiguous()
        L0 = prompt.shape[1]
        N = B * K
        cache = [ws.get(f"bs_qkv.{K}.{i}", (N, max_length, 3 * d), dt) for i in range(dm.dec_layers)]
        logits = ws.get(f"bs_logits.{K}", (N, dm.Vp), torch.float32 if dt == torch.float32 else dt)
        # cache-row table: position j of logical row r was written by (and still sits in) physical row rows[r, j]
        rows = ws.get(f"bs_rows.{K}", (N, max_length), torch.int32)
        rows.copy_(torch.arange(N, dtype=torch.int32, device=self.device)[:, None].expand(N, max_length))
        ident = ws.get(f"bs_ident.{K}", (N,), torch.int32)
        ident.copy_(torch.arange(N, dtype=torch.int32, device=self.device))
        state = {"pos": 0}

        graphs = self._decode_graphs if use_graphs else None

        def step_fn(tokens: torch.Tensor, pos: int) -> torch.Tensor:
            Lq = tokens.shape[1]
            ids = ws.get(f"bs_ids.{K}.{Lq}", (N, Lq), torch.long)       # static buffer: the pass below may be a graph replay
            ids.copy_(tokens)
            state["pos"] = pos + Lq
            if graphs is None:
                self._decode_logits(ids, pos, cache, kv_all, logits, max_length, beams=K, kv_rows=rows)
                return logits
            key = ("beam", B, K, max_length, Lq, pos, self._weights_version)
            ent = graphs.get(key)
            if ent is None or ent[1] != ws.gen:
                self._decode_logits(ids, pos, cache, kv_all, logits, max_length, beams=K, kv_rows=rows)   # eager once: sizes the workspace
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._decode_logits(ids, pos, cache, kv_all, logits, max_length, beams=K, kv_rows=rows)
                graphs[key] = (g, ws.gen)
            else:
                ent[0].replay()
            return logits

        def reorder_fn(beam_idx: torch.Tensor):
            # positions < pos: inherit the parent's table; the next token's K/V will be written by the row itself
            p = state["pos"]
            rows[:, :p] = rows[:, :p].index_select(0, beam_idx)

        scorer = None
        if not sequence_bias and 2 * K <= 16:
            C2 = 2 * K
            rs = ws.get(f"bs_rs.{K}", (N, C2), torch.float32)
            rt = ws.get(f"bs_rt.{K}", (N, C2), torch.int32)

            def scorer(lg, flat, run_score, first):
                ops.beam_row_topk(lg, dm.vocab, flat.contiguous(), run_score.reshape(-1).contiguous(), repetition_penalty,
                                  no_repeat_ngram_size, self.suppress if (first and self.suppress.numel()) else None, C2, rs, rt)
                top_score, idx = torch.topk(rs.view(B, K * C2), C2, dim=1)
                return top_score, idx // C2, rt.view(B, K * C2).gather(1, idx).long()

        out = run_beams(step_fn, reorder_fn, prompt, K, max_length, dm.vocab, dm.eos_token_id, dm.pad_token_id,
                        dm.begin_suppress_tokens, repetition_penalty, no_repeat_ngram_size, length_penalty,
                        sequence_bias=sequence_bias, scorer=scorer)
        return out[:, L0:].contiguous()

    # ------------------------------------------------------------------ greedy decode with KV cache
    @_on_device
    @torch.no_grad()
    def greedy(self, x: torch.Tensor, max_length: int, prompt: Optional[torch.Tensor] = None, aug: Optional[dict] = None,
               use_graphs: bool = True, eos_check_every: int = 16) -> torch.Tensor:
        """Batched greedy generate (utils/load_model.py:1072-1351 -> GenerationMixin greedy): encoder once, cross-K/V once,
        then one-token decoder steps against the self-attention cache.  Returns the generated suffix (B, n_new) int64;
        rows that hit EOS emit pad afterwards; begin_suppress_tokens are masked at the first generated position.  Every
        `eos_check_every` positions the host looks at the finished flags and stops once every row has emitted EOS (the
        remaining positions are pad, as HF pads finished rows)."""
        dm, W, ws, dt = self.dims, self.P, self.ws, self.dtype
        d, S, F, H = dm.d_model, dm.max_source_positions, dm.dec_ffn, dm.dec_heads
        Dh = d // H
        B = x.shape[0]
        enc = self.encode(x, aug=aug, save=False)
        if prompt is None:
            prompt = torch.full((B, 1), dm.decoder_start_token_id, dtype=torch.long, device=self.device)
        prompt = prompt.to(self.device, torch.long).contiguous()
        L0 = prompt.shape[1]
        Tmax = max_length
        if Tmax > dm.max_target_positions:
            raise ValueError(f"max_length {Tmax} exceeds max_target_positions {dm.max_target_positions}")
        n_new = Tmax - L0
        if n_new <= 0:
            return torch.empty((B, 0), dtype=torch.long, device=self.device)
        cache = [ws.get(f"g_qkv.{i}", (B, Tmax, 3 * d), dt) for i in range(dm.dec_layers)]
        absorb = self._absorbed_decode(B)
        kv_all = None if absorb else ws.get("kv_layers", (dm.dec_layers, B * S, 2 * d), dt)
        finished = ws.get("g_fin", (B,), torch.uint8); finished.zero_()
        nxt = ws.get("g_next", (B,), torch.long)
        logits = ws.get("g_logits", (B, dm.Vp), torch.float32 if dt == torch.float32 else dt)
        out = ws.get(f"g_out.{n_new}", (B, n_new), torch.long)
        out.fill_(dm.pad_token_id)
        ids0 = ws.get(f"g_ids0.{L0}", (B, L0), torch.long)
        ids0.copy_(prompt)
        # the loop body is native: ns_decode_prefill (cross K|V of all layers) and one ns_decode_step per position
        dec = self._native_decoder(B, Tmax, cache, kv_all, logits, enc=enc.view(B * S, d) if absorb else None)
        ops.decode_prefill(dec, enc)

        def decode_step(step: int, ids: torch.Tensor, pos: int):
            """One decoder pass over `ids` (B, Lq) at cache position `pos` -> next token in `nxt`, appended to out[:, step]."""
            sup = self.suppress if step == 0 else None
            if ids.shape[1] == 1:
                ops.decode_step(dec, ids, pos, sup, dm.eos_token_id, dm.pad_token_id, finished, nxt, out_col=out[:, step])
            else:                                                     # a multi-token prompt: the general pass, once
                self._decode_logits(ids, pos, cache, kv_all, logits, Tmax)
                ops.greedy_pick(logits, dm.vocab, sup, dm.eos_token_id, dm.pad_token_id, finished, nxt, out_col=out[:, step])

        # A position is one native call (~70 launches issued from C++ in ~0.2 ms of host time against ~0.9 ms of device time), so
        # the loop is device-bound without CUDA graphs; use_graphs=True still replays one captured graph per position (keyed by
        # batch / prompt length / position; all buffers live in the persistent workspace).
        graphs = self._decode_graphs if use_graphs else None
        # programmatic dependent launch: every kernel of a step starts its prologue under the tail of the one before it
        pdl_prev = ops.set_pdl(not _NO_PDL)
        try:
            first, pos0 = ids0, 0
            if absorb and L0 > 1:
                # The absorbed form has no cached K|V for the general multi-token pass: the prompt (evaluation.py:357-359 hands
                # generate() four decoder tokens) goes through the native step one position at a time -- the same causal
                # computation, its picks discarded (scratch flags / ids) -- and the loop starts at the last prompt token.
                fin_d = ws.get("g_fin_prompt", (B,), torch.uint8); nxt_d = ws.get("g_next_prompt", (B,), torch.long)
                cols = ws.get(f"g_prompt_cols.{L0}", (L0, B), torch.long)
                cols.copy_(ids0.t())
                for j in range(L0 - 1):
                    fin_d.zero_()
                    ops.decode_step(dec, cols[j], j, None, dm.eos_token_id, dm.pad_token_id, fin_d, nxt_d)
                first, pos0 = cols[L0 - 1].view(B, 1), L0 - 1
            return self._greedy_loop(decode_step, graphs, n_new, B, Tmax, L0, first, nxt, finished, out, eos_check_every, pos0)
        finally:
            ops.set_pdl(pdl_prev)

    def _greedy_loop(self, decode_step, graphs, n_new, B, Tmax, L0, ids0, nxt, finished, out, eos_check_every, pos0: int = 0):
        ws = self.ws
        pos = pos0
        for step in range(n_new):
            ids = ids0 if step == 0 else nxt.view(B, 1)
            key = (B, Tmax, L0, step, self._weights_version)
            if graphs is None:
                decode_step(step, ids, pos)
            else:
                ent = graphs.get(key)
                if ent is None or ent[1] != ws.gen:
                    decode_step(step, ids, pos)                       # eager once: allocates this step's workspace buffers
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):                         # capture only records; state stays as after the eager run
                        decode_step(step, ids, pos)
                    graphs[key] = (g, ws.gen)
                else:
                    ent[0].replay()
            pos += ids.shape[1]
            if eos_check_every and step % eos_check_every == eos_check_every - 1 and step + 1 < n_new and bool(finished.all()):
                break                                                 # every row is past its EOS: the rest stays pad
        return out.clone()

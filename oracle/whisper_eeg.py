"""CPU oracle for the NeuSpeech hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-fp32 (CPU) restatement of the arithmetic of the
reference's EEG-conditioned Whisper path.  Nothing in the product package
(`neuspeech1_b200/`) may import it: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs do, and only as the checker
or the timed CPU baseline.

What it restates (reference file:line  /  HF = transformers 5.5.0 models/whisper/modeling_whisper.py):
  * stem            utils/model_utils.py:9-17 (projection_module 'base'),
                    utils/load_model.py:410-417 (gelu(conv1), gelu(conv2), permute, + embed_positions)
  * encoder layers  utils/load_model.py:428-468 ; HF:380-414 (pre-LN block) ; HF:284-358 (attention:
                    q = (Wq h + bq) * Dh**-0.5, bias-free k, softmax(q k^T) v, out_proj)
  * LoRA            finetune.py:194-212 (r=32, alpha=64 -> scale 2.0, bias none, targets q/k/v/out/fc1/fc2 of
                    the encoder); PEFT lora.Linear forward  y = base(x) + scale * B(A(dropout(x)))
                    (PEFT itself is absent from this image: restated from the call site, dropout=0 for parity)
  * decoder         utils/load_model.py:534-767 ; HF:449-507 (self-attn causal, cross-attn, MLP), final LN
  * loss            utils/load_model.py:1027 (shift_tokens_right), :1047 (tied proj_out), :1051-1054 (CE, ignore -100)
  * greedy decode   utils/load_model.py:1072-1351 -> GenerationMixin greedy with KV cache; begin-suppress
                    tokens at the first generated position (HF generation_whisper.py:1774-1813)
  * training step   HF trainer.py:1867-1934 + finetune.py:231-253: loss.backward, clip_grad_norm_(1.0),
                    AdamW(betas=(0.9,0.999), eps=1e-8, weight_decay=0)

Pinning: `tests/test_oracle.py` checks this restatement against stock
`transformers.WhisperForConditionalGeneration` + the reference's own `projection_module`
(golden fixtures in tests/golden/, made by oracle/make_golden.py which imports /root/reference),
and live against stock HF (transformers is part of the image on both boxes).
The reference itself ships no tests/golden vectors for this path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
ENC_LORA_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")  # finetune.py:194


@dataclass
class Dims:
    """Model dimensions (mirrors the WhisperConfig fields the path reads)."""
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
    def T(self) -> int:  # input samples: S * stride(conv1 'B') * stride(conv2)
        return self.max_source_positions * 4

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r


WHISPER_BASE = Dims()
TINY = Dims(d_model=128, enc_layers=2, dec_layers=2, enc_heads=2, dec_heads=2, enc_ffn=256, dec_ffn=256,
            vocab=1000, max_source_positions=64, max_target_positions=32, eeg_ch=16, pad_token_id=997,
            eos_token_id=997, decoder_start_token_id=998, begin_suppress_tokens=(220, 996), lora_r=8, lora_alpha=16)


# --------------------------------------------------------------------------- parameters

def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> Tensor:
    """Whisper encoder position table: [sin | cos] halves (HF:55-65)."""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def init_params(dims: Dims, seed: int = 0, std: float = 0.02) -> Dict[str, Tensor]:
    """Random-init parameter dict with the reference's state-dict names (HF names + Sequential stem)."""
    g = torch.Generator().manual_seed(seed)
    d = dims.d_model
    P: Dict[str, Tensor] = {}

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    def conv(name, cout, cin):
        bound = 1.0 / math.sqrt(cin * 3)
        P[name + ".weight"] = (torch.rand(cout, cin, 3, generator=g) * 2 - 1) * bound
        P[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    conv("model.encoder.conv1.0", d, dims.eeg_ch)
    conv("model.encoder.conv1.2", d, d)
    conv("model.encoder.conv2", d, d)
    P["model.encoder.embed_positions.weight"] = sinusoids(dims.max_source_positions, d)

    def ln(name):
        P[name + ".weight"] = 1.0 + n(d, s=0.1)
        P[name + ".bias"] = n(d, s=0.1)

    def attn(prefix):
        for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
            P[f"{prefix}.{p}.weight"] = n(d, d)
            if p != "k_proj":
                P[f"{prefix}.{p}.bias"] = n(d)

    for i in range(dims.enc_layers):
        pre = f"model.encoder.layers.{i}"
        attn(pre + ".self_attn")
        ln(pre + ".self_attn_layer_norm")
        P[pre + ".fc1.weight"] = n(dims.enc_ffn, d); P[pre + ".fc1.bias"] = n(dims.enc_ffn)
        P[pre + ".fc2.weight"] = n(d, dims.enc_ffn); P[pre + ".fc2.bias"] = n(d)
        ln(pre + ".final_layer_norm")
    ln("model.encoder.layer_norm")

    P["model.decoder.embed_tokens.weight"] = n(dims.vocab, d)
    P["model.decoder.embed_positions.weight"] = n(dims.max_target_positions, d)
    for i in range(dims.dec_layers):
        pre = f"model.decoder.layers.{i}"
        attn(pre + ".self_attn"); ln(pre + ".self_attn_layer_norm")
        attn(pre + ".encoder_attn"); ln(pre + ".encoder_attn_layer_norm")
        P[pre + ".fc1.weight"] = n(dims.dec_ffn, d); P[pre + ".fc1.bias"] = n(dims.dec_ffn)
        P[pre + ".fc2.weight"] = n(d, dims.dec_ffn); P[pre + ".fc2.bias"] = n(d)
        ln(pre + ".final_layer_norm")
    ln("model.decoder.layer_norm")
    return P


def init_lora(dims: Dims, seed: int = 1, b_std: float = 0.0) -> Dict[str, Tensor]:
    """LoRA A/B for the 6 target linears of every encoder layer (finetune.py:194-212).

    PEFT init: A ~ kaiming_uniform(a=sqrt(5)) i.e. U(-1/sqrt(in), 1/sqrt(in)), B = 0.  `b_std > 0` gives a non-zero B so
    that parity tests exercise the adapter path (a trained adapter has B != 0).
    Names follow PEFT: <module>.lora_A.default.weight (r,in), <module>.lora_B.default.weight (out,r).
    """
    g = torch.Generator().manual_seed(seed)
    L: Dict[str, Tensor] = {}
    d, r = dims.d_model, dims.lora_r
    for i in range(dims.enc_layers):
        for t in ENC_LORA_TARGETS:
            mod = f"model.encoder.layers.{i}." + (t if t.startswith("fc") else f"self_attn.{t}")
            fin = dims.enc_ffn if t == "fc2" else d
            fout = dims.enc_ffn if t == "fc1" else d
            bound = 1.0 / math.sqrt(fin)
            L[mod + ".lora_A.default.weight"] = (torch.rand(r, fin, generator=g) * 2 - 1) * bound
            L[mod + ".lora_B.default.weight"] = torch.randn(fout, r, generator=g) * b_std
    return L


def trainable_names(P: Dict[str, Tensor], lora: Dict[str, Tensor]) -> List[str]:
    """Trainable set of finetune.py:176-212: LoRA A/B + modules_to_save conv1 (stem A,B) and conv2 (stem C)."""
    stem = [k for k in P if k.startswith("model.encoder.conv1.") or k.startswith("model.encoder.conv2.")]
    return sorted(k for k in lora if not k.startswith("__")) + sorted(stem)


# --------------------------------------------------------------------------- forward pieces

def lowbias32(x):
    """The 32-bit mixer of the dropout mask / seed sequence (numpy uint64 arithmetic, masked to 32 bits)."""
    import numpy as np
    m = np.uint64(0xFFFFFFFF)
    x = x & m
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def next_dropout_seed(seed: int) -> int:
    """The step-seed sequence of the training step: seed <- lowbias32(seed + 0x9E3779B9)  (ns_seed_advance)."""
    import numpy as np
    return int(lowbias32(np.uint64((seed + 0x9E3779B9) & 0xFFFFFFFF)))


def module_salt(name: str) -> int:
    import zlib
    return zlib.crc32(name.encode()) & 0xFFFFFFFF


def lora_dropout_plane(seed: int, name: str, rows: int, cols: int, p: float):
    """The dropped-element bit plane of module `name` as the device draws it (include/neuspeech_b200.h, csrc/ns_lora.cu):
    uint32 array (rows, (cols+31)//32); bit (col % 32) of word (row, col // 32) set <=> dropped.
    The 32 flags of a word come from 16 hashed words R_i = mix1(km + (i + 1) * 0xC2B2AE35), km = lowbias32((row * 0x9E3779B1) ^
    (w * 0x85EBCA77) ^ seed ^ crc32(name)), mix1(x): x ^= x >> 16; x *= 0x7FEB352D; x ^= x >> 15 (all mod 2^32), combined along
    the binary expansion of thr = round(p * 65536), least significant bit first:
    D = bit_i(thr) ? (D | R_i) : (D & R_i), i.e. flag_b = [U_b < thr] for the 16-bit number U_b made of bit b of R_15..R_0."""
    import numpy as np
    m = np.uint64(0xFFFFFFFF)
    ms = np.uint64((seed ^ module_salt(name)) & 0xFFFFFFFF)
    row = np.arange(rows, dtype=np.uint64)[:, None]
    w = np.arange((cols + 31) // 32, dtype=np.uint64)[None, :]
    km = lowbias32(((row * np.uint64(0x9E3779B1)) ^ (w * np.uint64(0x85EBCA77)) ^ ms) & m)
    thr = min(65535, int(p * 65536.0 + 0.5))
    d = np.zeros(km.shape, dtype=np.uint64)
    for i in range(16):
        x = (km + np.uint64(((i + 1) * 0xC2B2AE35) & 0xFFFFFFFF)) & m
        x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & m
        x ^= x >> np.uint64(15)
        d = (d | x) if (thr >> i) & 1 else (d & x)
    valid = np.clip(cols - 32 * np.arange(d.shape[1]), 0, 32)                # columns that exist in each 32-column block
    d &= ((np.uint64(1) << valid.astype(np.uint64)) - np.uint64(1))[None, :]
    return d.astype(np.uint32)


def lora_dropout_keep(seed: int, name: str, rows: int, cols: int, p: float) -> Tensor:
    """Keep mask (rows, cols) of the LoRA-branch dropout of module `name` (finetune.py:210 lora_dropout=0.05; PEFT lora.Linear
    applies nn.Dropout to the LoRA branch input only), read off the counter-based bit plane above: P(dropped) = round(p * 65536)
    / 65536.  PEFT itself draws from torch's generator, so with dropout on, parity with the reference is statistical by
    construction; between this oracle and the CUDA path it is exact."""
    import numpy as np
    plane = lora_dropout_plane(seed, name, rows, cols, p).astype(np.uint64)                 # (rows, words)
    sh = np.arange(32, dtype=np.uint64)[None, None, :]
    dropped = ((plane[:, :, None] >> sh) & np.uint64(1)).reshape(rows, -1)[:, :cols]
    return torch.from_numpy(dropped == 0)


def linear(x: Tensor, P, name: str, lora=None, scale: float = 0.0) -> Tensor:
    y = F.linear(x, P[name + ".weight"], P.get(name + ".bias"))
    if lora is not None and (name + ".lora_A.default.weight") in lora:
        a = lora[name + ".lora_A.default.weight"]; b = lora[name + ".lora_B.default.weight"]
        xin = x
        if "__dropout__" in lora:                       # (p, seed): training-mode dropout on the LoRA branch input only
            pdrop, seed = lora["__dropout__"]
            if pdrop > 0:
                keep = lora_dropout_keep(seed, name, x.numel() // x.shape[-1], x.shape[-1], pdrop).view(x.shape)
                xin = x * keep.to(device=x.device, dtype=x.dtype) / (1.0 - pdrop)
        y = y + scale * F.linear(F.linear(xin, a), b)
    return y


def stem(x: Tensor, P) -> Tensor:
    """(B,C,T) -> (B,S,d): three convs, exact-erf GELU."""
    a = F.conv1d(x, P["model.encoder.conv1.0.weight"], P["model.encoder.conv1.0.bias"], padding=1)
    a = F.gelu(a)
    b = F.conv1d(a, P["model.encoder.conv1.2.weight"], P["model.encoder.conv1.2.bias"], stride=2, padding=1)
    b = F.gelu(b)                                                   # utils/load_model.py:410
    c = F.gelu(F.conv1d(b, P["model.encoder.conv2.weight"], P["model.encoder.conv2.bias"], stride=2, padding=1))
    return c.permute(0, 2, 1) + P["model.encoder.embed_positions.weight"]


def mha(q: Tensor, k: Tensor, v: Tensor, heads: int, causal: bool = False) -> Tensor:
    """q already carries the Dh**-0.5 factor (HF:310); plain softmax(q k^T) v (HF:215-238)."""
    B, Lq, d = q.shape
    Lk = k.shape[1]
    dh = d // heads
    q = q.view(B, Lq, heads, dh).transpose(1, 2)
    k = k.view(B, Lk, heads, dh).transpose(1, 2)
    v = v.view(B, Lk, heads, dh).transpose(1, 2)
    w = q @ k.transpose(2, 3)
    if causal:
        off = Lk - Lq
        m = torch.ones(Lq, Lk, dtype=torch.bool, device=q.device).tril(off)
        w = w.masked_fill(~m, float("-inf"))
    w = w.softmax(dim=-1)
    return (w @ v).transpose(1, 2).reshape(B, Lq, d)


def encoder_layer(h: Tensor, P, pre: str, dims: Dims, lora=None) -> Tensor:
    s = dims.lora_scale
    u = F.layer_norm(h, (dims.d_model,), P[pre + ".self_attn_layer_norm.weight"], P[pre + ".self_attn_layer_norm.bias"], 1e-5)
    dh = dims.d_model // dims.enc_heads
    q = linear(u, P, pre + ".self_attn.q_proj", lora, s) * dh ** -0.5
    k = linear(u, P, pre + ".self_attn.k_proj", lora, s)
    v = linear(u, P, pre + ".self_attn.v_proj", lora, s)
    o = mha(q, k, v, dims.enc_heads)
    h = h + linear(o, P, pre + ".self_attn.out_proj", lora, s)
    u = F.layer_norm(h, (dims.d_model,), P[pre + ".final_layer_norm.weight"], P[pre + ".final_layer_norm.bias"], 1e-5)
    m = F.gelu(linear(u, P, pre + ".fc1", lora, s))
    return h + linear(m, P, pre + ".fc2", lora, s)


def encoder(x: Tensor, P, dims: Dims, lora=None) -> Tensor:
    """input_features (B,C,T) -> encoder_last_hidden_state (B,S,d)   [parity tensor 1]"""
    if x.shape[-1] != dims.T:
        raise ValueError(f"expected input length {dims.T}, got {x.shape[-1]}")  # HF:613-617
    h = stem(x, P)
    for i in range(dims.enc_layers):
        h = encoder_layer(h, P, f"model.encoder.layers.{i}", dims, lora)
    return F.layer_norm(h, (dims.d_model,), P["model.encoder.layer_norm.weight"], P["model.encoder.layer_norm.bias"], 1e-5)


def shift_tokens_right(labels: Tensor, pad: int, start: int) -> Tensor:
    out = labels.new_zeros(labels.shape)
    out[:, 1:] = labels[:, :-1]
    out[:, 0] = start
    return out.masked_fill(out == -100, pad)


def decoder(ids: Tensor, enc: Tensor, P, dims: Dims, past: Optional[list] = None) -> Tuple[Tensor, list]:
    """Decoder hidden states for `ids` (B,L).  `past` = per-layer [self_k, self_v, cross_k, cross_v] (KV cache)."""
    d, H = dims.d_model, dims.dec_heads
    dh = d // H
    t0 = 0 if past is None else past[0][0].shape[1]
    L = ids.shape[1]
    h = P["model.decoder.embed_tokens.weight"][ids] + P["model.decoder.embed_positions.weight"][t0:t0 + L]
    new_past = []
    for i in range(dims.dec_layers):
        pre = f"model.decoder.layers.{i}"
        u = F.layer_norm(h, (d,), P[pre + ".self_attn_layer_norm.weight"], P[pre + ".self_attn_layer_norm.bias"], 1e-5)
        q = linear(u, P, pre + ".self_attn.q_proj") * dh ** -0.5
        k = linear(u, P, pre + ".self_attn.k_proj")
        v = linear(u, P, pre + ".self_attn.v_proj")
        if past is not None:
            k = torch.cat([past[i][0], k], dim=1); v = torch.cat([past[i][1], v], dim=1)
        h = h + linear(mha(q, k, v, H, causal=True), P, pre + ".self_attn.out_proj")
        u = F.layer_norm(h, (d,), P[pre + ".encoder_attn_layer_norm.weight"], P[pre + ".encoder_attn_layer_norm.bias"], 1e-5)
        q = linear(u, P, pre + ".encoder_attn.q_proj") * dh ** -0.5
        if past is not None:
            ck, cv = past[i][2], past[i][3]
        else:
            ck = linear(enc, P, pre + ".encoder_attn.k_proj"); cv = linear(enc, P, pre + ".encoder_attn.v_proj")
        h = h + linear(mha(q, ck, cv, H), P, pre + ".encoder_attn.out_proj")
        u = F.layer_norm(h, (d,), P[pre + ".final_layer_norm.weight"], P[pre + ".final_layer_norm.bias"], 1e-5)
        h = h + linear(F.gelu(linear(u, P, pre + ".fc1")), P, pre + ".fc2")
        new_past.append([k, v, ck, cv])
    h = F.layer_norm(h, (d,), P["model.decoder.layer_norm.weight"], P["model.decoder.layer_norm.bias"], 1e-5)
    return h, new_past


def forward_loss(x: Tensor, labels: Tensor, P, dims: Dims, lora=None):
    """model(input_features, labels) -> (loss, logits (B,L,V), encoder_last_hidden_state)   [parity tensor 2 = loss]"""
    enc = encoder(x, P, dims, lora)
    dec_in = shift_tokens_right(labels, dims.pad_token_id, dims.decoder_start_token_id)
    y, _ = decoder(dec_in, enc, P, dims)
    logits = y @ P["model.decoder.embed_tokens.weight"].t()            # tied proj_out, no bias
    loss = F.cross_entropy(logits.view(-1, dims.vocab), labels.reshape(-1), ignore_index=-100)
    return loss, logits, enc


@torch.no_grad()
def greedy_decode(x: Tensor, P, dims: Dims, max_length: int, lora=None, prompt: Optional[Tensor] = None) -> Tensor:
    """Greedy generate with KV cache.  Returns the generated suffix (B, max_length - prompt_len) like HF 5.5's plain-tensor
    return.  Finished rows (EOS seen) emit pad.  begin-suppress tokens get -inf at the first generated position."""
    enc = encoder(x, P, dims, lora)
    B = x.shape[0]
    if prompt is None:
        prompt = torch.full((B, 1), dims.decoder_start_token_id, dtype=torch.long)
    E = P["model.decoder.embed_tokens.weight"]
    out = []
    done = torch.zeros(B, dtype=torch.bool)
    ids, past = prompt, None
    cur = prompt.shape[1]
    while cur < max_length:
        y, past = decoder(ids, enc, P, dims, past)
        logits = y[:, -1] @ E.t()
        if cur == prompt.shape[1] and len(dims.begin_suppress_tokens):
            logits[:, list(dims.begin_suppress_tokens)] = float("-inf")
        nxt = logits.argmax(dim=-1)
        nxt = torch.where(done, torch.full_like(nxt, dims.pad_token_id), nxt)
        out.append(nxt)
        done = done | (nxt == dims.eos_token_id)
        ids = nxt[:, None]
        cur += 1
        if bool(done.all()):
            break
    return torch.stack(out, dim=1)


# --------------------------------------------------------------------------- AdaLoRA (finetune.py:205-208)  -- PARITY UNPINNED

def adalora_loss(x: Tensor, labels: Tensor, P, dims: Dims, master: Dict[str, Tensor], modules, init_r: int = 12,
                 lora_alpha: float = 32.0, orth_reg_weight: float = 0.5, dropout=None) -> Tensor:
    """Loss of the reference's AdaLoRA configuration (`AdaLoraConfig(init_r=12, target_r=4, lora_alpha=32, lora_dropout=0.1,
    orth_reg_weight=0.5)`, finetune.py:205-208; `update_and_allocate` is never called, so the rank stays init_r):
        y    = base(x) + (dropout(x) @ (A * E).T @ B.T) * lora_alpha / (ranknum + 1e-5)
        loss = CE + orth_reg_weight * mean over all A, B of ||A A^T - I||_F resp. ||B^T B - I||_F
    PARITY UNPINNED: PEFT is not on this box and the reference ships no vectors for it; this restates PEFT's SVDLinear /
    AdaLoraModel.forward from the call site and from memory of PEFT ~0.5.  `master`: {<module>.lora_A.default (r, in),
    .lora_E.default (r, 1), .lora_B.default (out, r)}; `modules`: module names; dropout = (p, seed) or None."""
    eff = {}
    for name in modules:
        eff[name + ".lora_A.default.weight"] = master[name + ".lora_A.default"] * master[name + ".lora_E.default"]
        eff[name + ".lora_B.default.weight"] = master[name + ".lora_B.default"]
    if dropout is not None:
        eff["__dropout__"] = dropout
    od = Dims(**{**dims.__dict__, "lora_r": init_r, "lora_alpha": lora_alpha * init_r / (init_r + 1e-5)})
    ce, _, _ = forward_loss(x, labels, P, od, eff)
    eye = torch.eye(init_r)
    reg = sum(torch.norm(master[n + ".lora_A.default"] @ master[n + ".lora_A.default"].T - eye, p="fro")
              + torch.norm(master[n + ".lora_B.default"].T @ master[n + ".lora_B.default"] - eye, p="fro") for n in modules)
    return ce + orth_reg_weight * reg / (2 * len(modules))


# --------------------------------------------------------------------------- training step

@dataclass
class AdamWState:
    step: int = 0
    m: Dict[str, Tensor] = field(default_factory=dict)
    v: Dict[str, Tensor] = field(default_factory=dict)


def grads(x: Tensor, labels: Tensor, P, dims: Dims, lora) -> Tuple[Tensor, Dict[str, Tensor], Tensor]:
    """loss + gradients of the trainable set (LoRA A/B + 3 stem convs) via autograd on the restated forward."""
    names = trainable_names(P, lora)
    Pg = dict(P); Lg = dict(lora)
    leaves = []
    for nme in names:
        src = Lg if nme in Lg else Pg
        t = src[nme].detach().clone().requires_grad_(True)
        src[nme] = t
        leaves.append(t)
    loss, _, enc = forward_loss(x, labels, Pg, dims, Lg)
    gs = torch.autograd.grad(loss, leaves)
    return loss.detach(), {nme: g for nme, g in zip(names, gs)}, enc.detach()


def clip_and_adamw(params: Dict[str, Tensor], g: Dict[str, Tensor], st: AdamWState, lr: float,
                   max_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, wd: float = 0.0) -> float:
    """clip_grad_norm_(max_norm) then torch.optim.AdamW semantics, in place on `params`.  Returns the pre-clip norm."""
    total = math.sqrt(sum(float((v.double() ** 2).sum()) for v in g.values()))
    coef = min(1.0, max_norm / (total + 1e-6))
    st.step += 1
    b1, b2 = betas
    for nme, gr in g.items():
        gr = gr * coef
        if nme not in st.m:
            st.m[nme] = torch.zeros_like(gr); st.v[nme] = torch.zeros_like(gr)
        p = params[nme]
        p.mul_(1 - lr * wd)
        st.m[nme].mul_(b1).add_(gr, alpha=1 - b1)
        st.v[nme].mul_(b2).addcmul_(gr, gr, value=1 - b2)
        bc1 = 1 - b1 ** st.step; bc2 = 1 - b2 ** st.step
        denom = (st.v[nme].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(st.m[nme], denom, value=-lr / bc1)
    return total


def train_step(x, labels, P, dims, lora, st: AdamWState, lr: float = 1e-3):
    loss, g, _ = grads(x, labels, P, dims, lora)
    both = {**{k: lora[k] for k in lora if not k.startswith("__")}, **{k: P[k] for k in g if k in P}}
    norm = clip_and_adamw(both, g, st, lr)
    return float(loss), norm


def synthetic_batch(dims: Dims, B: int, L: int = 32, seed: int = 1, ragged: bool = True):
    """Config #1 inputs: (0.3*randn).clamp(-1,1) signal of random length, zero tail to T; labels with -100 tail."""
    g = torch.Generator().manual_seed(seed)
    x = (0.3 * torch.randn(B, dims.eeg_ch, dims.T, generator=g)).clamp_(-1, 1)
    if ragged:
        lo, hi = max(1, dims.T // 15), max(2, dims.T * 5 // 6)
        for b in range(B):
            n = int(torch.randint(lo, hi, (1,), generator=g))
            x[b, :, n:] = 0
    labels = torch.randint(0, dims.vocab - 10, (B, L), generator=g)
    labels[:, -min(4, L // 2):] = -100
    return x, labels

"""Drop-in module for the object `utils/load_model.py` builds in the reference.

Mirrors the reference's Python contract -- same module tree and parameter names (`model.encoder.conv1.0`, `...layers.N.
self_attn.q_proj`, `proj_out`, ...), `forward(input_features, labels=...) -> .loss/.logits`, `generate(...)`,
`encoder.set_input_embeddings(...)`, `get_encoder()` -- while every FLOP runs in the hand-written CUDA kernels behind the
C-ABI (neuspeech1_b200/engine.py).  The nn.Parameters here are the fp32 masters; the engine keeps bf16 copies of the frozen
ones and owns the trainable ones (LoRA A/B + stem convs) in one flat buffer that the Parameters alias.

Reference call sites: finetune.py:127-177 (construction, stem swap, freezing), finetune.py:194-212 (LoRA by module name),
utils/load_model.py:976-1070 (forward), :1072-1351 (generate), evaluation.py:370-395.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn

from .engine import ENC_LORA_TARGETS, ModelDims, WhisperEEGEngine, lora_module_name
from .lora import LoraLinear, ModulesToSaveWrapper
from .model_utils import projection_module
from . import weights as W


class Seq2SeqLMOutput(dict):
    """Minimal stand-in for transformers' ModelOutput: attribute and key access (`out.loss`, `out["loss"]`, out[0])."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __getitem__(self, k):
        if isinstance(k, int):
            return [v for v in self.values() if v is not None][k]
        return super().__getitem__(k)


class WhisperAttention(nn.Module):
    def __init__(self, d: int, heads: int):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = d, heads, d // heads
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)


class WhisperEncoderLayer(nn.Module):
    def __init__(self, d: int, heads: int, ffn: int):
        super().__init__()
        self.self_attn = WhisperAttention(d, heads)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class WhisperDecoderLayer(nn.Module):
    def __init__(self, d: int, heads: int, ffn: int):
        super().__init__()
        self.self_attn = WhisperAttention(d, heads)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = WhisperAttention(d, heads)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class WhisperEncoder(nn.Module):
    def __init__(self, dims: ModelDims, owner):
        super().__init__()
        d = dims.d_model
        self._owner = [owner]                   # list: keep the back-reference out of the module tree
        self.conv1 = projection_module("base", meg_ch=dims.eeg_ch, d_model=d)
        self.conv2 = nn.Conv1d(d, d, kernel_size=3, stride=2, padding=1)
        self.embed_positions = nn.Embedding(dims.max_source_positions, d)
        self.embed_positions.requires_grad_(False)
        self.layers = nn.ModuleList([WhisperEncoderLayer(d, dims.enc_heads, dims.enc_ffn) for _ in range(dims.enc_layers)])
        self.layer_norm = nn.LayerNorm(d)

    def get_input_embeddings(self) -> nn.Module:
        return self.conv1

    def set_input_embeddings(self, value: nn.Module):
        """utils/load_model.py:368-369; finetune.py:148,163 (also the cross-dataset 208 -> 273 channel stem swap)."""
        self.conv1 = value
        self._owner[0]._stem_swapped()

    def forward(self, input_features, **kw):
        owner = self._owner[0]
        enc = owner._engine().encode(input_features)
        return SimpleNamespace(last_hidden_state=enc.clone())


class WhisperDecoder(nn.Module):
    def __init__(self, dims: ModelDims):
        super().__init__()
        d = dims.d_model
        self.embed_tokens = nn.Embedding(dims.vocab, d, padding_idx=dims.pad_token_id if dims.pad_token_id < dims.vocab else None)
        self.embed_positions = nn.Embedding(dims.max_target_positions, d)
        self.layers = nn.ModuleList([WhisperDecoderLayer(d, dims.dec_heads, dims.dec_ffn) for _ in range(dims.dec_layers)])
        self.layer_norm = nn.LayerNorm(d)


class WhisperModel(nn.Module):
    def __init__(self, dims: ModelDims, owner):
        super().__init__()
        self.encoder = WhisperEncoder(dims, owner)
        self.decoder = WhisperDecoder(dims)

    def get_encoder(self):
        return self.encoder

    def get_decoder(self):
        return self.decoder


class _HotPath(torch.autograd.Function):
    """One autograd node for the whole model: forward = engine.forward_loss, backward = engine.backward.  Gradients are
    returned for the trainable Parameters (LoRA A/B, stem convs) so torch optimizers / DDP / HF Trainer see them."""

    @staticmethod
    def forward(ctx, owner, x, labels, dec_ids, *trainables):
        eng = owner._engine()
        owner._sync_trainables_to_engine()
        loss, logits, enc = eng.forward_loss(x, labels, decoder_input_ids=dec_ids, save=True)   # only reached when grads are needed
        ctx.owner = owner
        ctx.n = len(trainables)
        out_loss = loss.clone() if loss is not None else torch.zeros((), device=eng.device)
        ctx.mark_non_differentiable(logits, enc)
        return out_loss, logits, enc

    @staticmethod
    def backward(ctx, gloss, glogits, genc):
        owner = ctx.owner
        eng = owner._engine()
        eng.backward()
        eng.grad.mul_(gloss.to(torch.float32))
        grads = [eng.trainable_grad(name).clone() for name in owner._trainable_names]
        return (None, None, None, None, *grads)


class WhisperEEGForConditionalGeneration(nn.Module):
    """`WhisperForConditionalGeneration` of utils/load_model.py with the EEG stem, on the B200 engine."""

    main_input_name = "input_features"

    def __init__(self, dims: ModelDims, params: Optional[Dict[str, torch.Tensor]] = None, lora: Optional[Dict[str, torch.Tensor]] = None,
                 dtype: torch.dtype = torch.bfloat16, device="cuda", lora_dropout: float = 0.0):
        super().__init__()
        self.dims = dims
        self.compute_dtype = dtype
        self.device_ = torch.device(device)
        self.config = SimpleNamespace(
            d_model=dims.d_model, vocab_size=dims.vocab, encoder_layers=dims.enc_layers, decoder_layers=dims.dec_layers,
            encoder_attention_heads=dims.enc_heads, decoder_attention_heads=dims.dec_heads, encoder_ffn_dim=dims.enc_ffn,
            decoder_ffn_dim=dims.dec_ffn, max_source_positions=dims.max_source_positions, max_target_positions=dims.max_target_positions,
            pad_token_id=dims.pad_token_id, eos_token_id=dims.eos_token_id, decoder_start_token_id=dims.decoder_start_token_id,
            begin_suppress_tokens=list(dims.begin_suppress_tokens), forced_decoder_ids=None, suppress_tokens=[], use_cache=True,
            is_encoder_decoder=True)
        self.model = WhisperModel(dims, self)
        self.proj_out = nn.Linear(dims.d_model, dims.vocab, bias=False)
        self.proj_out.weight = self.model.decoder.embed_tokens.weight           # tied (utils/load_model.py:1047)
        with torch.no_grad():
            self.model.encoder.embed_positions.weight.copy_(W.sinusoids(dims.max_source_positions, dims.d_model))
        self._eng: Optional[WhisperEEGEngine] = None
        self._trainable_names = []
        self._trainable_params = []
        self._lora_cfg = None
        if params is not None:
            missing = self.load_state_dict({k: v for k, v in params.items()}, strict=False)
            assert all(k == "proj_out.weight" for k in missing.missing_keys), missing
        self.requires_grad_(False)
        for p in list(self.model.encoder.conv1.parameters()) + list(self.model.encoder.conv2.parameters()):
            p.requires_grad_(True)                                              # modules_to_save (finetune.py:202)
        if lora is not None:
            from .lora import lora_inject
            lora_inject(self, r=dims.lora_r, lora_alpha=dims.lora_alpha, lora_dropout=lora_dropout, state=lora)

    # ---- construction helpers -------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path: str, eeg_ch: int = 208, dtype=torch.bfloat16, device="cuda", local_files_only: bool = True, **_):
        """Load a HF Whisper checkpoint directory (config.json + model.safetensors | pytorch_model.bin).  The mel stem of the
        checkpoint is dropped: the EEG stem is created fresh, like finetune.py:138-148 does right after loading."""
        cfg = json.load(open(os.path.join(path, "config.json")))
        dims = ModelDims(d_model=cfg["d_model"], enc_layers=cfg["encoder_layers"], dec_layers=cfg["decoder_layers"],
                         enc_heads=cfg["encoder_attention_heads"], dec_heads=cfg["decoder_attention_heads"],
                         enc_ffn=cfg["encoder_ffn_dim"], dec_ffn=cfg["decoder_ffn_dim"], vocab=cfg["vocab_size"],
                         max_source_positions=cfg["max_source_positions"], max_target_positions=cfg["max_target_positions"],
                         eeg_ch=eeg_ch, pad_token_id=cfg.get("pad_token_id", 50257), eos_token_id=cfg.get("eos_token_id", 50257),
                         decoder_start_token_id=cfg.get("decoder_start_token_id", 50258),
                         begin_suppress_tokens=tuple(cfg.get("begin_suppress_tokens") or ()))
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        sd = W.params_from_state_dict(sd)
        own_stem = sd.get("model.encoder.conv1.0.weight") is not None and sd["model.encoder.conv1.0.weight"].shape[1] == eeg_ch
        if not own_stem:                                    # mel stem (80/128 bins): keep conv2, re-create conv1 for EEG
            sd = {k: v for k, v in sd.items() if not k.startswith("model.encoder.conv1.")}
        model = cls(dims, None, None, dtype=dtype, device=device)
        res = model.load_state_dict(sd, strict=False)
        bad = [k for k in res.missing_keys if not k.startswith("model.encoder.conv1.") and k != "proj_out.weight"]
        assert not bad, f"checkpoint is missing {bad[:5]}..."
        return model

    @property
    def device(self):
        return self.device_

    @property
    def engine(self) -> WhisperEEGEngine:
        return self._engine()

    def post_init(self):
        return None

    def get_encoder(self):
        return self.model.encoder

    def get_decoder(self):
        return self.model.decoder

    # ---- engine plumbing ------------------------------------------------------------------------------------------
    def _collect(self):
        """Flatten the module tree to (frozen params, lora dict) with plain HF names."""
        sd = {k: v for k, v in self.state_dict().items()}
        return W.params_from_state_dict(sd), (W.lora_from_state_dict(sd) or None)

    def _engine(self) -> WhisperEEGEngine:
        if self._eng is None:
            params, lora = self._collect()
            stem_ch = params["model.encoder.conv1.0.weight"].shape[1]
            if stem_ch != self.dims.eeg_ch:
                self.dims = ModelDims(**{**self.dims.__dict__, "eeg_ch": stem_ch})
            if lora is not None and self._lora_cfg is not None:
                self.dims = ModelDims(**{**self.dims.__dict__, "lora_r": self._lora_cfg["r"], "lora_alpha": self._lora_cfg["lora_alpha"]})
            drop = float((self._lora_cfg or {}).get("lora_dropout", 0.0)) if lora is not None else 0.0
            self._eng = WhisperEEGEngine(self.dims, params, lora, dtype=self.compute_dtype, device=self.device_, lora_dropout=drop,
                                         dropout_seed=int(torch.initial_seed()) & 0xFFFFFFFF)
            self._alias_trainables()
        return self._eng

    def _named_trainable_modules(self):
        """(engine name, nn.Parameter) for every trainable parameter, PEFT wrappers resolved."""
        out = []
        for k, p in self.named_parameters():
            if ".original_module." in k:
                continue
            name = k.replace(".base_layer.", ".").replace(".modules_to_save.default.", ".")
            if name in self._eng.layout.entries:
                out.append((name, p))
        return out

    def _alias_trainables(self):
        """Point the trainable nn.Parameters at the engine's flat fp32 buffer (zero-copy: any optimizer updates it in place)."""
        eng = self._eng
        self._trainable_names, self._trainable_params = [], []
        for name, p in self._named_trainable_modules():
            view = eng.trainable(name)
            view.copy_(p.data.to(view.device))
            p.data = view
            self._trainable_names.append(name)
            self._trainable_params.append(p)

    def _sync_trainables_to_engine(self):
        """An external optimizer (HF Trainer's AdamW) updates the aliased flat master buffer in place: the engine's bf16 LoRA
        copies / packed stem taps are stale then.  Detected through the Parameters' version counters, so evaluation and
        generate() right after optimizer.step() never run one step behind."""
        ver = tuple(p._version for p in self._trainable_params)
        if ver != getattr(self, "_trainable_versions", None):
            self._trainable_versions = ver
            self._eng._packed = False

    def _stem_swapped(self):
        self._eng = None

    def invalidate_engine(self):
        """Call after changing frozen weights in place (load_state_dict, merge) so the bf16 copies are rebuilt."""
        self._eng = None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._eng = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def to(self, *a, **kw):          # device placement is fixed at construction; keep HF-style chaining working
        return self

    # ---- the reference's forward / generate contract -------------------------------------------------------------
    def forward(self, input_features=None, labels=None, decoder_input_ids=None, return_dict=True, **unused):
        if decoder_input_ids is None and labels is None:
            raise ValueError("You have to specify either decoder_input_ids or decoder_inputs_embeds")   # utils/load_model.py:613-614
        self._engine().training = self.training          # nn.Dropout semantics: the LoRA-branch dropout is off under model.eval()
        x = input_features.to(self.device_)
        train_params = self._trainable_params if torch.is_grad_enabled() else []
        need_grad = labels is not None and torch.is_grad_enabled() and any(p.requires_grad for p in train_params)
        if need_grad:
            loss, logits, enc = _HotPath.apply(self, x, labels, decoder_input_ids, *train_params)
        else:
            self._sync_trainables_to_engine()
            with torch.no_grad():
                loss, logits, enc = self._eng.forward_loss(x, labels, decoder_input_ids=decoder_input_ids, save=False)
                loss = loss.clone() if loss is not None else None
        return Seq2SeqLMOutput(loss=loss, logits=logits, encoder_last_hidden_state=enc)

    def training_step(self, input_features, labels, lr: float, all_reduce=None, aug: Optional[dict] = None, use_graph: bool = True):
        """Fused Trainer.training_step + clip + AdamW (HF trainer.py:1867-1934, finetune.py:231-253) without autograd.
        use_graph: see WhisperEEGEngine.train_step (CUDA-graph replay when the same input buffers come back)."""
        eng = self._engine()
        eng.training = True
        loss = eng.train_step(input_features, labels, lr=lr, aug=aug, all_reduce=all_reduce, use_graph=use_graph)
        return Seq2SeqLMOutput(loss=loss)

    @torch.no_grad()
    def generate(self, input_features=None, do_sample: bool = False, num_beams: int = 1, max_length: Optional[int] = None,
                 max_new_tokens: Optional[int] = None, decoder_input_ids=None, repetition_penalty: float = 1.0,
                 no_repeat_ngram_size: int = 0, sequence_bias=None, **unused):
        """`generate` as the reference calls it: greedy with KV cache (utils/process_str.py:54-55) and the beam search of
        evaluation.py:370-385 (`num_beams=5, repetition_penalty=5.0, no_repeat_ngram_size=2`, optionally `sequence_bias`,
        :339-343).  Sampling and beam groups are refused loudly rather than approximated.  The decoder passes replay CUDA
        graphs (captured on first use per shape).  Returns the generated suffix like HF's Whisper wrapper."""
        if do_sample or unused.get("num_beam_groups", 1) != 1:
            raise NotImplementedError("neuspeech1_b200.generate implements greedy and plain beam search (no sampling, no beam groups)")
        L0 = 1 if decoder_input_ids is None else decoder_input_ids.shape[1]
        if max_new_tokens is not None:
            max_length = L0 + max_new_tokens
        if max_length is None:
            max_length = self.dims.max_target_positions
        eng = self._engine()
        self._sync_trainables_to_engine()
        x = input_features.to(self.device_)
        if num_beams == 1 and repetition_penalty == 1.0 and not no_repeat_ngram_size and not sequence_bias:
            return eng.greedy(x, max_length=max_length, prompt=decoder_input_ids)
        if sequence_bias:
            sequence_bias = {tuple(int(t) for t in k): float(v) for k, v in dict(sequence_bias).items()}
        return eng.beam_search(x, max_length=max_length, num_beams=num_beams, repetition_penalty=repetition_penalty,
                               no_repeat_ngram_size=no_repeat_ngram_size, prompt=decoder_input_ids,
                               length_penalty=float(unused.get("length_penalty", 1.0)), sequence_bias=sequence_bias)

    @staticmethod
    def _reorder_cache(past_key_values, beam_idx):
        """utils/load_model.py:1353-1360: reorder a tuple-of-tuples KV cache along the batch axis.  The engine's own beam search
        never calls it (it permutes a cache-row table instead, engine.beam_search); kept for callers that drive the decoder
        with an external generation loop."""
        return tuple(tuple(t.index_select(0, beam_idx.to(t.device)) for t in layer) for layer in past_key_values)

    def prepare_inputs_for_generation(self, decoder_input_ids, past_key_values=None, use_cache=None, encoder_outputs=None, **kw):
        if past_key_values is not None:                      # utils/load_model.py:1332-1351: feed only the last token
            decoder_input_ids = decoder_input_ids[:, -1:]
        return {"encoder_outputs": encoder_outputs, "past_key_values": past_key_values, "decoder_input_ids": decoder_input_ids,
                "use_cache": use_cache}


WhisperForConditionalGeneration = WhisperEEGForConditionalGeneration   # the reference's class name

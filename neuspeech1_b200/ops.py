"""Thin torch-tensor wrappers over the C-ABI (include/neuspeech_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every computation below is one `ns_*` call into
libneuspeech_b200.so on the current CUDA stream.  Tensors must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi
from ._abi import ACT_DGELU, ACT_GELU, ACT_NONE, AttnShape, AugArgs, Epilogue, NS_BF16, NS_F32, check

_lib = None
_fn_cache = {}
_prof = None          # when a list: (name, flops, bytes, start_event, end_event) per launch


def profile_begin():
    """Record a CUDA-event pair around every ns_* call (on the current stream) until profile_end()."""
    global _prof
    _prof = []


def profile_end():
    """-> list of dicts {name, ms, flops, bytes} in launch order (synchronises)."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    return [dict(name=n, flops=f, bytes=b, ms=s.elapsed_time(e)) for (n, f, b, s, e) in rec]


def _call(name, work, *args, tag=None):
    """tag: name the launch is filed under by the profiler (default: the entry point)."""
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(lib(), name)
    if _prof is None:
        st = fn(*args)
    else:
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        st = fn(*args)
        e.record()
        _prof.append((tag or name, float(work[0]), float(work[1]), s, e))
    if st != 0:
        check(st, name)


def lib():
    global _lib
    if _lib is None:
        _lib = _abi.load()
    return _lib


def ns_dtype(t) -> int:
    dt = t if isinstance(t, torch.dtype) else t.dtype
    if dt == torch.float32:
        return NS_F32
    if dt == torch.bfloat16:
        return NS_BF16
    raise TypeError(f"neuspeech1_b200 supports float32 and bfloat16 storage, got {dt}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _abi.NeuSpeechB200Error("neuspeech1_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def epilogue(bias=None, alpha=1.0, alpha_cols=0, act=ACT_NONE, aux_in=None, aux_out=None, ldaux=0, residual=None, ldr=0,
             res_mod=0, out_dtype=NS_BF16, a2_group_cols=0, drop_bits=None, drop_a=None, a_group_cols=0, aux_deriv=0, drop_gen=None) -> Epilogue:
    """drop_bits: ONE adapter's (rows, words) plane of dropout_bits -- the second product is masked with it (input gradient of a
    LoRA branch under dropout).  drop_a: the (G, rows, words) planes of G stacked rank-32 adapters -- the A operand of the single
    product is masked per 32-column output tile (the LoRA down product); with drop_gen = (seed tensor, salts, p) the kernel
    draws those planes itself and stores them into drop_a (drop_mode 2).  See include/neuspeech_b200.h ns_epilogue."""
    if drop_a is not None:
        ep = Epilogue(_p(bias), alpha, alpha_cols, act, _p(aux_in), _p(aux_out), ldaux, _p(residual), ldr, res_mod,
                      out_dtype, a2_group_cols, _p(drop_a), drop_a.stride(1), 1, drop_a.stride(0), a_group_cols, aux_deriv, None, None, 0.0)
        if drop_gen is not None:
            seed, salts, p = drop_gen
            arr = (C.c_uint * 4)(*([int(x) & 0xFFFFFFFF for x in salts] + [0] * (4 - len(salts))))
            ep.drop_mode = 2
            ep.drop_seed = _p(seed)
            ep.drop_salts = C.cast(arr, C.c_void_p)
            ep.drop_p = float(p)
            ep._salts_keepalive = arr                    # the library copies the salts during the call
        return ep
    return Epilogue(_p(bias), alpha, alpha_cols, act, _p(aux_in), _p(aux_out), ldaux, _p(residual), ldr, res_mod,
                    out_dtype, a2_group_cols, _p(drop_bits), drop_bits.stride(0) if drop_bits is not None else 0, 0, 0, a_group_cols, aux_deriv,
                    None, None, 0.0)


def gemm_nt(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, ep: Optional[Epilogue] = None, a2=None, w2=None,
            k2: int = 0, M: Optional[int] = None, N: Optional[int] = None, K: Optional[int] = None):
    """out[M,N] = epi(a[M,K] @ w[N,K]^T (+ a2[M,k2] @ w2[N,k2]^T)).  Leading dims are taken from .stride(0)."""
    M = a.shape[0] if M is None else M
    K = a.shape[1] if K is None else K
    N = w.shape[0] if N is None else N
    if ep is None:
        ep = epilogue(out_dtype=ns_dtype(out))
    # the profiler files the rank-r products (N <= 96: t = x A^T, dt = g B; 32-wide kernel, bound by the read of the activation)
    # apart from the dense products (256 / 128-wide kernels, tensor bound): different kernels, different rooflines
    thin = N <= 96
    _call("ns_gemm_nt", (2.0 * M * N * (K + k2), (M * K + M * N) * float(a.element_size()) if thin else 0), ns_dtype(a), M, N, K, _p(a),
          a.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), C.byref(ep), _p(a2), a2.stride(0) if a2 is not None else 0, _p(w2),
          w2.stride(0) if w2 is not None else 0, k2, _stream(), tag="ns_gemm_nt.rank_r" if thin else None)
    return out


def ln_gemm_nt(x: torch.Tensor, gamma, beta, w: torch.Tensor, out: torch.Tensor, ep: Optional[Epilogue] = None, eps: float = 1e-5):
    """out[M,N] = epi(LN(x)[M,K] @ w[N,K]^T) for M <= 128 rows in one launch (gamma = beta = None: no LayerNorm)."""
    M, K = x.shape
    N = w.shape[0]
    if ep is None:
        ep = epilogue(out_dtype=ns_dtype(out))
    _call("ns_ln_gemm_nt", (2.0 * M * N * K, 0), ns_dtype(x), M, N, K, _p(x), x.stride(0), _p(gamma), _p(beta), float(eps), _p(w), w.stride(0),
          _p(out), out.stride(0), C.byref(ep), _stream())
    return out


def gemm_tn(x: torch.Tensor, y: torch.Tensor, g: torch.Tensor, si: int, sj: int, alpha: float = 1.0,
            I: Optional[int] = None, J: Optional[int] = None):
    """g[i*si + j*sj] += alpha * sum_m x[m,i] * y[m,j]   (g fp32)."""
    I = x.shape[1] if I is None else I
    J = y.shape[1] if J is None else J
    _call("ns_gemm_tn", (2.0 * x.shape[0] * I * J, 0), ns_dtype(x), x.shape[0], I, J, _p(x), x.stride(0), _p(y), y.stride(0),
          _p(g), si, sj, alpha, _stream())
    return g


def conv3_fwd(x, w_tap, y, stride: int, ep: Epilogue):
    B, Tin, Cp = x.shape
    N = w_tap.shape[1]
    _call("ns_conv3_fwd", (6.0 * B * (Tin // stride) * Cp * N, 0), ns_dtype(x), B, Tin, Cp, N, stride, _p(x), _p(w_tap), _p(y), C.byref(ep), _stream())
    return y


def conv3_dgrad(dz, w_tap_t, dx, stride: int, ep: Epilogue):
    B, Tin, Cp = dx.shape
    N = dz.shape[2]
    _call("ns_conv3_dgrad", (6.0 * B * (Tin // stride) * Cp * N, 0), ns_dtype(dz), B, Tin, Cp, N, stride, _p(dz), _p(w_tap_t), _p(dx), C.byref(ep), _stream())
    return dx


def conv3_wgrad(dz, x, dw_tap, db, stride: int):
    B, Tin, Cp = x.shape
    N = dz.shape[2]
    _call("ns_conv3_wgrad", (6.0 * B * (Tin // stride) * Cp * N, 0), ns_dtype(dz), B, Tin, Cp, N, stride, _p(dz), _p(x), _p(dw_tap), _p(db), _stream())


def layernorm_fwd(x, gamma, beta, y, mean=None, rstd=None, eps: float = 1e-5):
    d = x.shape[-1]
    _call("ns_layernorm_fwd", (0, 2.0 * x.numel() * x.element_size()), ns_dtype(x), x.numel() // d, d, _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), eps, _stream())
    return y


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dres=None):
    d = x.shape[-1]
    _call("ns_layernorm_bwd", (0, (4.0 if dres is not None else 3.0) * x.numel() * x.element_size()), ns_dtype(x), x.numel() // d, d, _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx), _stream())
    return dx


def attn_shape(B, H, Lq, Lk, Dh, causal, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs) -> AttnShape:
    return AttnShape(B, H, Lq, Lk, Dh, int(causal), q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs)


def attention_fwd(shape: AttnShape, q, k, v, o, lse=None):
    _call("ns_attention_fwd", (4.0 * shape.B * shape.H * shape.Lq * shape.Lk * shape.Dh * (0.5 if shape.causal else 1.0), 0), ns_dtype(q), C.byref(shape), _p(q), _p(k), _p(v), _p(o), _p(lse), _stream())
    return o


def attention_bwd(shape: AttnShape, q, k, v, o, d_o, lse, delta, dq, dk, dv):
    _call("ns_attention_bwd", (10.0 * shape.B * shape.H * shape.Lq * shape.Lk * shape.Dh * (0.5 if shape.causal else 1.0), 0), ns_dtype(q), C.byref(shape), _p(q), _p(k), _p(v), _p(o), _p(d_o), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv), _stream())


def attention_bwd_workspace_bytes(shape: AttnShape) -> int:
    return int(lib().ns_attention_bwd_workspace_bytes(C.byref(shape)))


def attention_bwd_ws(shape: AttnShape, q, k, v, o, d_o, lse, delta, dq, dk, dv, ws: Optional[torch.Tensor]):
    """attention_bwd through the fused single-pass kernel when `ws` (uint8 device buffer, >= attention_bwd_workspace_bytes)
    is given and the shape qualifies; identical results contract."""
    nbytes = 0 if ws is None else ws.numel() * ws.element_size()
    _call("ns_attention_bwd_ws", (10.0 * shape.B * shape.H * shape.Lq * shape.Lk * shape.Dh * (0.5 if shape.causal else 1.0), 0), ns_dtype(q), C.byref(shape), _p(q), _p(k), _p(v), _p(o), _p(d_o), _p(lse), _p(delta), _p(dq), _p(dk), _p(dv), _p(ws), nbytes, _stream())


def attention_decode_rows(shape: AttnShape, q, k, v, o, kv_row):
    """Single-query attention with a per-(row, position) cache-row table (beam search: no cache copies)."""
    _call("ns_attention_decode_rows", (4.0 * shape.B * shape.H * shape.Lk * shape.Dh, 0), ns_dtype(q), C.byref(shape), _p(q), _p(k), _p(v), _p(o),
          _p(kv_row), kv_row.stride(0), _stream())
    return o


def beam_row_topk(logits, V: int, seqs, run_score, penalty: float, ngram: int, suppress, C2: int, out_score, out_tok):
    """Per beam row the C2 best continuations after log-softmax, repetition penalty, n-gram ban, begin-suppress, + running score."""
    n_sup = 0 if suppress is None else suppress.numel()
    _call("ns_beam_row_topk", (0, float(logits.shape[0]) * V * logits.element_size()), ns_dtype(logits), logits.shape[0], V, logits.stride(0),
          _p(logits), _p(seqs), seqs.stride(0), seqs.shape[1], _p(run_score), float(penalty), int(ngram), _p(suppress), n_sup, C2,
          _p(out_score), _p(out_tok), _stream())
    return out_score, out_tok


def embed(ids, E, P, pos0: int, h):
    B, L = ids.shape
    _call("ns_embed", (0, 0), ns_dtype(E), B, L, E.shape[1], _p(ids), _p(E), _p(P), pos0, _p(h), _stream())
    return h


def cross_entropy(logits, V: int, labels, row_loss, loss_sum, n_valid, write_grad: bool, grad_scale: float = 1.0):
    rows = logits.shape[0]
    _call("ns_cross_entropy", (0, (3.0 if write_grad else 2.0) * logits.numel() * logits.element_size()), ns_dtype(logits), rows, V, logits.stride(0), _p(logits), _p(labels), _p(row_loss), _p(loss_sum), _p(n_valid), int(write_grad), grad_scale, _stream())


def greedy_pick(logits, V: int, suppress, eos: int, pad: int, finished, next_ids, out_col=None):
    """out_col: a column view sequences[:, step] (int64) that receives the picked token as well."""
    n_sup = 0 if suppress is None else suppress.numel()
    _call("ns_greedy_pick", (0, 0), ns_dtype(logits), logits.shape[0], V, logits.stride(0), _p(logits), _p(suppress), n_sup, eos, pad, _p(finished),
          _p(next_ids), _p(out_col), out_col.stride(0) if out_col is not None else 0, _stream())
    return next_ids


def cross_attention_absorbed(qp: torch.Tensor, enc: torch.Tensor, ctx: torch.Tensor):
    """ctx[b, h] = softmax_j(qp[b, h] . enc[b, j]) enc[b]  (qp, ctx (B, H, d); enc (B, S, d); bf16, d = 512): the decode-step
    cross-attention with the key / value projections absorbed into the query / output side (ns_cross_attention_absorbed)."""
    B, H, d = qp.shape
    S = enc.shape[1]
    _call("ns_cross_attention_absorbed", (4.0 * B * H * S * d, 2.0 * B * S * d), ns_dtype(qp), B, S, H, d, _p(qp), qp.stride(0), _p(enc),
          enc.stride(0), _p(ctx), ctx.stride(0), _stream())
    return ctx


def decode_prefill(dec, enc):
    """Cross-attention K|V of every decoder layer (ns_decode_prefill); dec = _abi.Decoder built by the engine."""
    _call("ns_decode_prefill", (0, 0), C.byref(dec), _p(enc), _stream())


def decode_step(dec, ids, pos: int, suppress, eos: int, pad: int, finished, next_ids, out_col=None):
    """One decoder position + greedy pick in one native call (ns_decode_step); ids (B,) int64 on the device."""
    n_sup = 0 if suppress is None else suppress.numel()
    _call("ns_decode_step", (0, 0), C.byref(dec), _p(ids), pos, _p(suppress), n_sup, eos, pad, _p(finished), _p(next_ids), _p(out_col),
          out_col.stride(0) if out_col is not None else 0, _stream())
    return next_ids


def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch for this thread's decoder-step kernels (include/neuspeech_b200.h ns_set_pdl)."""
    return bool(lib().ns_set_pdl(1 if on else 0))


def cast(src, dst):
    _call("ns_cast", (0, 0), ns_dtype(src), ns_dtype(dst), src.numel(), _p(src), _p(dst), _stream())
    return dst


def transpose(src, dst, scale: float = 1.0):
    """dst (cols, ldd>=rows) = scale * src(rows, cols)^T ; columns [rows, ldd) of dst are zero-filled."""
    rows, cols = src.shape
    _call("ns_transpose", (0, 0), ns_dtype(src), ns_dtype(dst), rows, cols, _p(src), src.stride(0), _p(dst), dst.stride(0), scale, _stream())
    return dst


class TransposeBatch:
    """A fixed set of (src, dst) transposes refreshed with one launch (ns_transpose_batched).  The job table lives on the
    device; the tensors it points at must stay alive and in place (the engine's workspace guarantees that)."""

    def __init__(self, pairs, device):
        import numpy as np
        assert pairs
        self.sdt, self.ddt = ns_dtype(pairs[0][0]), ns_dtype(pairs[0][1])
        self.keep = pairs
        jobs = (_abi.TransposeJob * len(pairs))()
        self.max_ldd = self.max_cols = 0
        for i, pr in enumerate(pairs):                   # (src, dst) or (src, dst, scale)
            src, dst = pr[0], pr[1]
            scale = float(pr[2]) if len(pr) > 2 else 1.0
            assert ns_dtype(src) == self.sdt and ns_dtype(dst) == self.ddt and src.dim() == 2 and dst.dim() == 2
            rows, cols = src.shape
            assert dst.shape[0] == cols and dst.stride(0) >= rows
            jobs[i] = _abi.TransposeJob(_p(src), _p(dst), rows, cols, src.stride(0), dst.stride(0), scale, 0)
            self.max_ldd = max(self.max_ldd, dst.stride(0)); self.max_cols = max(self.max_cols, cols)
        raw = np.frombuffer(bytes(jobs), dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(device)
        self.n = len(pairs)

    def run(self):
        _call("ns_transpose_batched", (0, 0), self.sdt, self.ddt, self.n, self.max_ldd, self.max_cols, _p(self.table), _stream())


def conv_weight_pack(w, w_tap, w_tap_t):
    N, Cin, _ = w.shape
    ref = w_tap if w_tap is not None else w_tap_t
    Cp = w_tap.shape[2] if w_tap is not None else w_tap_t.shape[1]
    _call("ns_conv_weight_pack", (0, 0), ns_dtype(ref), N, Cin, Cp, _p(w), _p(w_tap), _p(w_tap_t), _stream())


def conv_weight_unpack_grad(dw_tap, dw):
    N, Cin, _ = dw.shape
    _call("ns_conv_weight_unpack_grad", (0, 0), N, Cin, dw_tap.shape[2], _p(dw_tap), _p(dw), _stream())


def add(a, b, y):
    _call("ns_add", (0, 0), ns_dtype(a), a.numel(), _p(a), _p(b), _p(y), _stream())
    return y


def dgelu_mul(dy, z, dz):
    _call("ns_dgelu_mul", (0, 0), ns_dtype(dy), dy.numel(), _p(dy), _p(z), _p(dz), _stream())
    return dz


def seed_advance(seed: torch.Tensor):
    """seed (1,) int32 device word <- lowbias32(seed + 0x9E3779B9): one new dropout mask per (replayed) training step."""
    _call("ns_seed_advance", (0, 0), _p(seed), _stream())


def _salts(salts):
    return (C.c_uint * 8)(*([int(s) & 0xFFFFFFFF for s in salts] + [0] * (8 - len(salts))))


def gemm_tn_grouped(x: torch.Tensor, y: torch.Tensor, g: torch.Tensor, I: int, J: int, si: int, sj: int, alphas):
    """Block-diagonal weight gradients in one launch: g[(k*I + i)*si + j*sj] += alphas[k] * sum_m x[m, k*I + i] * y[m, k*J + j]."""
    groups = len(alphas)
    arr = (C.c_float * groups)(*[float(a) for a in alphas])
    _call("ns_gemm_tn_grouped", (2.0 * x.shape[0] * I * J * groups, 0), ns_dtype(x), x.shape[0], I, J, groups, _p(x), x.stride(0), _p(y),
          y.stride(0), _p(g), si, sj, arr, _stream())
    return g


def lora_bwd_b_workspace_bytes(M: int, N: int, r: int, groups: int) -> int:
    """Bytes of caller-owned workspace ns_lora_bwd_b needs for this shape: 0 = none, -1 = the shape does not qualify."""
    return int(lib().ns_lora_bwd_b_workspace_bytes(M, N, r, groups))


def lora_bwd_b(dy: torch.Tensor, bt: torch.Tensor, t: torch.Tensor, dt: torch.Tensor, db: torch.Tensor, N: int, r: int, alpha_dt, alpha_db,
               workspace: Optional[torch.Tensor] = None):
    """One pass over dy (M, groups*N): dt[:, g*r:(g+1)*r] = alpha_dt[g] * dy_g @ bt_g^T and db[g*N:(g+1)*N] += alpha_db[g] * dy_g^T @ t_g
    (ns_lora_bwd_b; bt (groups*r, N) = B^T, db (groups*N, r) fp32 contiguous).  workspace: uint8 tensor of
    lora_bwd_b_workspace_bytes() bytes whose ticket words were zeroed once (wide N only)."""
    groups = len(alpha_dt)
    a1 = (C.c_float * groups)(*[float(a) for a in alpha_dt])
    a2 = (C.c_float * groups)(*[float(a) for a in alpha_db])
    M = dy.shape[0]
    assert db.is_contiguous() and db.dtype == torch.float32 and db.shape == (groups * N, r)
    _call("ns_lora_bwd_b", (4.0 * M * N * r * groups, 2.0 * M * N * groups), ns_dtype(dy), M, N, r, groups, _p(dy), dy.stride(0), _p(bt),
          bt.stride(0), _p(t), t.stride(0), _p(dt), dt.stride(0), _p(db), a1, a2, _p(workspace),
          workspace.numel() if workspace is not None else 0, _stream())
    return dt


def gemm_tn_masked(x: torch.Tensor, y: torch.Tensor, g: torch.Tensor, si: int, sj: int, xbits: torch.Tensor, alpha: float = 1.0):
    """g[i*si + j*sj] += alpha * sum_m (x . keep)[m,i] * y[m,j]; xbits = ONE adapter's (rows, words) plane of dropout_bits."""
    I, J = x.shape[1], y.shape[1]
    _call("ns_gemm_tn_masked", (2.0 * x.shape[0] * I * J, 0), ns_dtype(x), x.shape[0], I, J, _p(x), x.stride(0), _p(y), y.stride(0),
          _p(g), si, sj, float(alpha), _p(xbits), xbits.stride(0), _stream())
    return g


def dropout_apply(x, y, bits):
    """y = x with the dropped elements of one adapter's bit plane zeroed (unscaled); x, y 2-D views (rows, cols)."""
    _call("ns_dropout_apply", (0, 2.0 * x.numel() * x.element_size()), ns_dtype(x), x.shape[0], x.shape[1], _p(x), x.stride(0), _p(y),
          y.stride(0), _p(bits), _stream())
    return y


def dropout_bits_words(rows: int, cols: int) -> int:
    return int(lib().ns_dropout_bits_words(rows, cols))


def dropout_bits(rows: int, cols: int, seed, salts, p: float, bits):
    """bits (G, rows, (cols+31)//32) int32 <- the dropped-element bit plane of the G modules `salts` (hashed once per step)."""
    _call("ns_dropout_bits", (0, float(bits.numel() * 4)), rows, cols, len(salts), _p(seed), _salts(salts), float(p), _p(bits), _stream())
    return bits


def lora_down(x, A, t, alpha: float, G: int = 1, bits=None):
    """t[M, G*r] = alpha * (x . keep_g) A_g^T, A = stacked (G*r, K) bf16; bits = dropout_bits(...) or None."""
    M, K = x.shape
    r = A.shape[0] // G
    _call("ns_lora_down", (2.0 * M * K * G * r, float(x.numel() * 2)), M, K, G, r, _p(x), x.stride(0), _p(A), A.stride(0), _p(t), t.stride(0),
          float(alpha), _p(bits), _stream())
    return t


def lora_da(x, dt, dA, G: int = 1, bits=None, dx=None, At=None, z=None):
    """dA[G*r, K] (fp32) += dt_g^T (x . keep_g); with dx (and At = A^T, optionally z) the same pass takes the dropped terms of
    the LoRA product out of the input gradient (see include/neuspeech_b200.h)."""
    M, K = x.shape
    r = dA.shape[0] // G
    _call("ns_lora_da", (2.0 * M * K * G * r * (2 if dx is not None else 1), float(x.numel() * 2)), M, K, G, r, _p(x), x.stride(0), _p(dt),
          dt.stride(0), _p(dA), dA.stride(0), _p(bits), _p(dx), dx.stride(0) if dx is not None else 0, _p(At),
          At.stride(0) if At is not None else 0, _p(z), z.stride(0) if z is not None else 0, _stream())
    return dA


def lora_dx_fix(dx, dt, At, bits, G: int = 1, z=None):
    """dx -= dropped_g * (dt_g . At[k, g]) (* gelu'(z)), in place, any storage dtype."""
    M, K = dx.shape
    r = At.shape[1] // G
    _call("ns_lora_dx_fix", (0, 2.0 * dx.numel() * dx.element_size()), ns_dtype(dx), M, K, G, r, _p(dx), dx.stride(0), _p(dt), dt.stride(0),
          _p(At), At.stride(0), _p(bits), _p(z), z.stride(0) if z is not None else 0, _stream())
    return dx


def sumsq(g, out):
    _call("ns_sumsq", (0, 0), g.numel(), _p(g), _p(out), _stream())


def adamw_clip(p, g, m, v, sumsq_t, gscale, max_norm, lr, beta1, beta2, eps, wd, step):
    _call("ns_adamw_clip", (0, 0), p.numel(), _p(p), _p(g), _p(m), _p(v), _p(sumsq_t), gscale, max_norm, lr, beta1, beta2, eps, wd, step, _stream())


def aug_pass(x, y, layout: int, n=None, shift=None, e0=None, e1=None, flags=None, grid=None, grid_stride=0, gl=None,
             rep_c=None, rep_t=None, sigma=None, seed: int = 0, src_off=None, src_ld=None, C_in: Optional[int] = None,
             Tin: Optional[int] = None):
    """x: the dense (B, C, Tin) batch, or (src_off given) the flat ragged sample store with row c of slot b at
    x[src_off[b] + c * src_ld[b] :][: n[b]]; fp32 or bf16."""
    if layout == 0:
        T, Cp = y.shape[2], y.shape[1]
    else:
        T, Cp = y.shape[1], y.shape[2]
    if src_off is None:
        B, Cc, Tin = x.shape
    else:
        B, Cc, Tin = y.shape[0], int(C_in), int(Tin or T)
    if layout == 0:
        Cp = Cc
    a = AugArgs(B, Cc, Tin, T, Cp, layout, ns_dtype(y), _p(n), _p(shift), _p(e0), _p(e1), _p(flags), _p(grid), grid_stride,
                _p(gl), _p(rep_c), _p(rep_t), _p(sigma), seed, ns_dtype(x), _p(src_off), _p(src_ld))
    _call("ns_aug_pass", (0, x.numel() * float(x.element_size()) + y.numel() * y.element_size()), C.byref(a), _p(x), _p(y), _stream())
    return y


def channel_meansq(x, n, ms, src_off=None, src_ld=None):
    if src_off is None:
        B, Cc, Tin = x.shape
    else:
        B, Cc = ms.shape
        Tin = 1
    _call("ns_channel_meansq", (0, 0), ns_dtype(x), B, Cc, Tin, _p(n), _p(x), _p(ms), _p(src_off), _p(src_ld), _stream())
    return ms

"""ctypes binding of libneuspeech_b200.so (the C-ABI declared in include/neuspeech_b200.h).

The library is built in-tree by `make` / `__graft_entry__.build()`.  There is no CPU fallback: importing this module
without the shared library raises, and every entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libneuspeech_b200.so")

NS_F32, NS_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_DGELU = 0, 1, 2
PATH_AUTO, PATH_SIMT, PATH_FAST = 0, 1, 2

c_vp = C.c_void_p
c_ll = C.c_longlong
c_i = C.c_int
c_f = C.c_float


class Epilogue(C.Structure):
    _fields_ = [("bias", c_vp), ("alpha", c_f), ("alpha_cols", c_i), ("act", c_i), ("aux_in", c_vp), ("aux_out", c_vp),
                ("ldaux", c_ll), ("residual", c_vp), ("ldr", c_ll), ("res_mod", c_i), ("out_dtype", c_i),
                ("a2_group_cols", c_i), ("drop_bits", c_vp), ("drop_ld", c_ll), ("drop_mode", c_i),
                ("drop_gstride", c_ll), ("a_group_cols", c_i), ("aux_deriv", c_i), ("drop_seed", c_vp), ("drop_salts", c_vp),
                ("drop_p", c_f)]


class DecoderLayer(C.Structure):
    _fields_ = [(n, c_vp) for n in ("ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "wqc", "bqc", "woc", "boc",
                                    "ln3_g", "ln3_b", "w1", "b1", "w2", "b2", "wkv", "bkv", "self_cache", "cross_kv", "wq_abs", "bq_abs")]


class Decoder(C.Structure):
    _fields_ = [(n, c_i) for n in ("dtype", "n_layers", "d", "heads", "ffn", "vocab", "S", "Tmax", "B", "logits_dtype")] + \
               [("cross_ld", c_ll), ("logits_ld", c_ll), ("E", c_vp), ("pos_table", c_vp), ("lnf_g", c_vp), ("lnf_b", c_vp),
                ("layers", C.POINTER(DecoderLayer))] + \
               [(n, c_vp) for n in ("h0", "u", "o", "h1", "qc", "h2", "mm", "h3a", "h3b", "y", "logits", "enc", "qp", "cp")]


class AttnShape(C.Structure):
    _fields_ = [("B", c_i), ("H", c_i), ("Lq", c_i), ("Lk", c_i), ("Dh", c_i), ("causal", c_i),
                ("q_bs", c_ll), ("q_rs", c_ll), ("k_bs", c_ll), ("k_rs", c_ll), ("v_bs", c_ll), ("v_rs", c_ll),
                ("o_bs", c_ll), ("o_rs", c_ll)]


class TransposeJob(C.Structure):
    _fields_ = [("src", c_vp), ("dst", c_vp), ("rows", c_i), ("cols", c_i), ("lds", c_ll), ("ldd", c_ll), ("scale", c_f), ("pad_", c_i)]


class AugArgs(C.Structure):
    _fields_ = [("B", c_i), ("C", c_i), ("Tin", c_i), ("T", c_i), ("Cp", c_i), ("layout", c_i), ("out_dtype", c_i),
                ("n", c_vp), ("shift", c_vp), ("e0", c_vp), ("e1", c_vp), ("flags", c_vp),
                ("grid", c_vp), ("grid_stride", c_ll), ("gl", c_vp), ("rep_c", c_vp), ("rep_t", c_vp),
                ("sigma", c_vp), ("seed", C.c_ulonglong), ("in_dtype", c_i), ("src_off", c_vp), ("src_ld", c_vp)]


# name -> argtypes (restype is int unless noted).  Kept in one table so tests can check every symbol of the header.
SIGNATURES = {
    "ns_version": [],
    "ns_set_path": [c_i],
    "ns_device_info": [C.POINTER(c_i), C.POINTER(c_i), C.POINTER(c_i)],
    "ns_get_counters": [C.POINTER(c_ll), c_i],
    "ns_reset_counters": [],
    "ns_gemm_nt": [c_i, c_ll, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, C.POINTER(Epilogue), c_vp, c_ll, c_vp, c_ll, c_i, c_vp],
    "ns_ln_gemm_nt": [c_i, c_ll, c_i, c_i, c_vp, c_ll, c_vp, c_vp, c_f, c_vp, c_ll, c_vp, c_ll, C.POINTER(Epilogue), c_vp],
    "ns_gemm_tn": [c_i, c_ll, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_ll, c_f, c_vp],
    "ns_gemm_tn_grouped": [c_i, c_ll, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_ll, c_vp, c_vp],
    "ns_lora_bwd_b": [c_i, c_ll, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp],
    "ns_lora_bwd_b_workspace_bytes": [c_ll, c_i, c_i, c_i],
    "ns_gemm_tn_masked": [c_i, c_ll, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_ll, c_f, c_vp, c_ll, c_vp],
    "ns_conv3_fwd": [c_i, c_i, c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, C.POINTER(Epilogue), c_vp],
    "ns_conv3_dgrad": [c_i, c_i, c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, C.POINTER(Epilogue), c_vp],
    "ns_conv3_wgrad": [c_i, c_i, c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_vp, c_vp],
    "ns_layernorm_fwd": [c_i, c_ll, c_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_vp],
    "ns_layernorm_bwd": [c_i, c_ll, c_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "ns_attention_fwd": [c_i, C.POINTER(AttnShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "ns_attention_bwd": [c_i, C.POINTER(AttnShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "ns_attention_bwd_workspace_bytes": [C.POINTER(AttnShape)],
    "ns_attention_bwd_ws": [c_i, C.POINTER(AttnShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp],
    "ns_debug_attn_trace": [c_vp],
    "ns_embed": [c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_i, c_vp, c_vp],
    "ns_cross_entropy": [c_i, c_ll, c_i, c_ll, c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_f, c_vp],
    "ns_greedy_pick": [c_i, c_i, c_i, c_ll, c_vp, c_vp, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_ll, c_vp],
    "ns_set_pdl": [c_i],
    "ns_cross_attention_absorbed": [c_i, c_i, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp],
    "ns_decode_prefill": [C.POINTER(Decoder), c_vp, c_vp],
    "ns_decode_step": [C.POINTER(Decoder), c_vp, c_i, c_vp, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_ll, c_vp],
    "ns_aug_pass": [C.POINTER(AugArgs), c_vp, c_vp, c_vp],
    "ns_channel_meansq": [c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "ns_cast": [c_i, c_i, c_ll, c_vp, c_vp, c_vp],
    "ns_transpose": [c_i, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_f, c_vp],
    "ns_transpose_batched": [c_i, c_i, c_i, c_i, c_i, c_vp, c_vp],
    "ns_conv_weight_pack": [c_i, c_i, c_i, c_i, c_vp, c_vp, c_vp, c_vp],
    "ns_conv_weight_unpack_grad": [c_i, c_i, c_i, c_vp, c_vp, c_vp],
    "ns_add": [c_i, c_ll, c_vp, c_vp, c_vp, c_vp],
    "ns_dgelu_mul": [c_i, c_ll, c_vp, c_vp, c_vp, c_vp],
    "ns_beam_row_topk": [c_i, c_i, c_i, c_ll, c_vp, c_vp, c_ll, c_i, c_vp, c_f, c_i, c_vp, c_i, c_i, c_vp, c_vp, c_vp],
    "ns_attention_decode_rows": [c_i, C.POINTER(AttnShape), c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp],
    "ns_seed_advance": [c_vp, c_vp],
    "ns_dropout_apply": [c_i, c_ll, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_vp],
    "ns_lora_down": [c_ll, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_f, c_vp, c_vp],
    "ns_lora_da": [c_ll, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp],
    "ns_lora_dx_fix": [c_i, c_ll, c_i, c_i, c_i, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_ll, c_vp],
    "ns_dropout_bits_words": [c_ll, c_i],
    "ns_dropout_bits": [c_ll, c_i, c_i, c_vp, c_vp, c_f, c_vp, c_vp],
    "ns_sumsq": [c_ll, c_vp, c_vp, c_vp],
    "ns_adamw_clip": [c_ll, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_vp],
}

RESTYPES = {"ns_attention_bwd_workspace_bytes": c_ll, "ns_dropout_bits_words": c_ll, "ns_lora_bwd_b_workspace_bytes": c_ll}      # everything else returns an int status

_lib = None


class NeuSpeechB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NeuSpeechB200Error(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "neuspeech1_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = RESTYPES.get(name, c_i)
    lib.ns_last_error_string.argtypes = []
    lib.ns_last_error_string.restype = C.c_char_p
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().ns_last_error_string().decode("utf-8", "replace")
        raise NeuSpeechB200Error(f"{what} failed with status {status}: {msg}")


def counters() -> dict:
    buf = (c_ll * 8)()
    check(load().ns_get_counters(buf, 8), "ns_get_counters")
    names = ["gemm_tcgen05", "gemm_simt", "attn_tc", "attn_simt", "other", "wgrad_tcgen05"]
    return {n: int(buf[i]) for i, n in enumerate(names)}


def reset_counters():
    check(load().ns_reset_counters(), "ns_reset_counters")


def set_path(path: int) -> int:
    return load().ns_set_path(path)

"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol include/*.h declares, and the
ctypes table in neuspeech1_b200/_abi.py covers them with the right arity.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "neuspeech_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:long long|int|const char\*)\s+(ns_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib():
    from neuspeech1_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _abi.load()


def test_header_symbols_exported(lib):
    fns = declared_functions()
    assert len(fns) >= 31
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in include/neuspeech_b200.h but not exported"


def test_ctypes_table_matches_header(lib):
    from neuspeech1_b200 import _abi
    fns = declared_functions()
    for name, nargs in fns.items():
        if name == "ns_last_error_string":
            continue
        assert name in _abi.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_abi.SIGNATURES[name]) == nargs, (name, len(_abi.SIGNATURES[name]), nargs)
    assert set(_abi.SIGNATURES) <= set(fns)


def test_struct_sizes_match_c_layout():
    from neuspeech1_b200 import _abi
    # ns_epilogue: ptr,float,int,int,(pad),ptr,ptr,ll,ptr,ll,int,int,int,(pad),ptr,ll,int,(pad),ll,int,int,ptr,ptr,float,(pad)
    assert ctypes.sizeof(_abi.Epilogue) == 144
    assert ctypes.sizeof(_abi.DecoderLayer) == 24 * 8 and ctypes.sizeof(_abi.Decoder) == 10 * 4 + 2 * 8 + 19 * 8
    assert ctypes.sizeof(_abi.AttnShape) == 6 * 4 + 8 * 8
    # ... seed, then in_dtype (int, padded to 8), src_off, src_ld
    assert ctypes.sizeof(_abi.AugArgs) == 7 * 4 + 4 + 5 * 8 + 8 + 8 + 3 * 8 + 8 + 8 + 8 + 8 + 8


def test_argument_errors_without_gpu(lib):
    """Entry points validate arguments before touching the device and explain the failure."""
    from neuspeech1_b200 import _abi
    assert lib.ns_version() >= 100
    st = lib.ns_gemm_nt(7, 1, 1, 1, None, 1, None, 1, None, 1, None, None, 0, None, 0, 0, None)
    assert st == -1
    assert b"dtype" in lib.ns_last_error_string()
    st = lib.ns_layernorm_fwd(0, 4, 0, None, None, None, None, None, None, 1e-5, None)
    assert st == -1
    prev = lib.ns_set_path(_abi.PATH_SIMT)
    assert lib.ns_set_path(prev) == _abi.PATH_SIMT


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under neuspeech1_b200/ may reference it."""
    pkg = os.path.join(ROOT, "neuspeech1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f

"""CPU: the C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every symbol that
include/jammy_b200.h declares; the ctypes mirrors agree with the compiled struct sizes.  No compute calls here."""
import ctypes
import os
import re

from jammy_flows_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "jammy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jf_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_declare_the_same_symbols():
    assert _declared_symbols() == sorted(_cabi.SYMBOLS.keys())


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_binding_loads_and_struct_sizes_match(lib_built):
    lib = _cabi.load()          # asserts ABI version and sizeof() of every struct
    assert lib.jf_abi_version() == _cabi.JF_ABI_VERSION
    assert lib.jf_launch_count() == 0


def test_constants_match_header():
    text = open(os.path.join(ROOT, "include", "jammy_b200.h")).read()
    for name, val in re.findall(r"#define\s+(JF_[A-Z0-9_]+)\s+(-?\d+)\b", text):
        if hasattr(_cabi, name):
            assert getattr(_cabi, name) == int(val), name


def test_bad_descriptors_are_rejected_without_a_gpu(lib_built):
    lib = _cabi.load()
    d = _cabi.JfPdfDesc()
    assert lib.jf_pdf_workspace_bytes(ctypes.byref(d), 1024) == -1          # abi_version 0
    sd = _cabi.JfSubPdfDesc()
    rc = lib.jf_subpdf_apply(ctypes.byref(sd), _cabi.JF_F64, 0, None, 0, None, 0, 0, None, None, None, None, None, 0,
                             None, 0, 16, None, None)
    assert rc == -3                                                          # JF_ERR_BAD_ARG (null buffers)

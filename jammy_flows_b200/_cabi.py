"""ctypes binding of libjammy_b200.so (C-ABI declared in include/jammy_b200.h).

The library is the ONLY compute path: if it cannot be loaded the import of this module raises -- there is no CPU or
eager-PyTorch fallback anywhere in the package.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# JF_LIB_PATH: experiment variants of the same library built by tools/build_variants.py (never another implementation)
LIB_PATH = os.environ.get("JF_LIB_PATH") or os.path.join(PKG_DIR, "libjammy_b200.so")

# ---- constants (keep in sync with include/jammy_b200.h; checked by tests/test_cabi_symbols.py) -----------------------
JF_ABI_VERSION = 5
JF_MAX_LAYERS = 16
JF_MAX_SUBPDFS = 8
JF_MAX_MLP_LINEAR = 6
JF_MAX_MLP_SEGMENTS = 10
JF_MAX_DIM = 16
JF_MAX_KDE = 32
JF_STATUS_WORDS = 4
JF_MAX_NESTED = 4
JF_MAX_BINS = 32
JF_F32, JF_F64 = 0, 1
JF_DIR_LOGPDF, JF_DIR_SAMPLE = 0, 1
JF_LAYER_GF, JF_LAYER_FVM, JF_LAYER_RQS, JF_LAYER_S1SPLINE, JF_LAYER_MOEBIUS, JF_LAYER_EXPMAP, JF_LAYER_MVN = 1, 2, 3, 4, 5, 6, 7
JF_COV_IDENTITY, JF_COV_DIAGONAL_SYMMETRIC, JF_COV_DIAGONAL, JF_COV_FULL = 0, 1, 2, 3
JF_SPLINE_PLAIN, JF_SPLINE_SMOOTH, JF_SPLINE_CIRCULAR = 0, 1, 2
JF_BD_PARAMS, JF_BD_FIXED, JF_BD_PERIODIC = 0, 1, 2
JF_NORM_NONE, JF_NORM_RAW, JF_NORM_REGULATED = 0, 1, 2
JF_ROT_HOUSEHOLDER, JF_ROT_NONE, JF_ROT_ANGLES, JF_ROT_CAYLEY, JF_ROT_TRIANGULAR = 0, 1, 2, 3, 4
JF_ROT_XYZ, JF_ROT_QUATERNION = 5, 6
(JF_KAPPA_DIRECT_LOG, JF_KAPPA_SOFTPLUS, JF_KAPPA_LOG_BOUNDED, JF_KAPPA_MU, JF_KAPPA_MU_SQUARED, JF_KAPPA_QUATVEC,
 JF_KAPPA_QUATVEC_SQUARED) = 0, 1, 2, 3, 4, 5, 6
JF_WIDTH_SMOOTH, JF_WIDTH_EXP, JF_WIDTH_SOFTPLUS = 0, 1, 2
JF_STRETCH_CLASSIC, JF_STRETCH_RQS = 0, 1
JF_POT_EXPONENTIAL, JF_POT_LINEAR, JF_POT_QUADRATIC = 0, 1, 2
JF_STATUS_NONFINITE, JF_STATUS_UNCONVERGED, JF_STATUS_OUT_OF_RANGE, JF_STATUS_ITERATIONS = 0, 1, 2, 3
ERRORS = {-1: "JF_ERR_BAD_DESC (invalid descriptor)", -2: "JF_ERR_UNSUPPORTED (no kernel for this configuration)",
          -3: "JF_ERR_BAD_ARG (invalid argument)", -4: "JF_ERR_WORKSPACE (workspace missing or too small)"}


class JfSplineDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_bins", C.c_int32), ("n_w", C.c_int32), ("n_h", C.c_int32), ("n_d", C.c_int32),
                ("fix_first", C.c_int32), ("fix_second", C.c_int32), ("indep", C.c_int32), ("bd_mode", C.c_int32),
                ("natural_direction", C.c_int32), ("param_offset", C.c_int32), ("reserved", C.c_int32),
                ("lo", C.c_double), ("hi", C.c_double), ("min_w", C.c_double), ("min_h", C.c_double),
                ("min_d", C.c_double), ("bd_fixed", C.c_double), ("max_ratio", C.c_double), ("reserved1", C.c_double)]


class JfLayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("n_params", C.c_int32), ("param_offset", C.c_int32),
                ("K", C.c_int32), ("hh_iter", C.c_int32), ("inv_type", C.c_int32), ("norm_mode", C.c_int32),
                ("has_offset", C.c_int32), ("first", C.c_int32), ("natural_direction", C.c_int32),
                ("max_iter", C.c_int32), ("n_vertical", C.c_int32), ("n_circular", C.c_int32),
                ("rotation_mode", C.c_int32), ("width_mode", C.c_int32), ("width_clamp", C.c_int32),
                ("skew", C.c_int32), ("center_mean", C.c_int32), ("stretch", C.c_int32),
                ("w_min", C.c_double), ("w_max", C.c_double), ("n_min", C.c_double), ("n_max", C.c_double),
                ("z_sign", C.c_double), ("min_kappa", C.c_double), ("lo", C.c_double), ("hi", C.c_double),
                ("clamp_lo", C.c_double), ("clamp_hi", C.c_double),
                ("spline", JfSplineDesc * JF_MAX_NESTED)]


class JfSubPdfDesc(C.Structure):
    _fields_ = [("manifold", C.c_int32), ("dim", C.c_int32), ("n_layers", C.c_int32), ("n_params", C.c_int32),
                ("layers", JfLayerDesc * JF_MAX_LAYERS)]


class JfMlpDesc(C.Structure):
    _fields_ = [("n_linear", C.c_int32), ("dims", C.c_int32 * (JF_MAX_MLP_LINEAR + 1)), ("n_segments", C.c_int32),
                ("seg_cols", C.c_int32 * JF_MAX_MLP_SEGMENTS)]


class JfPdfDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dtype", C.c_int32), ("n_sub", C.c_int32), ("cond_dim", C.c_int32),
                ("total_target_dim", C.c_int32), ("total_base_dim", C.c_int32), ("reserved0", C.c_int32),
                ("reserved1", C.c_int32),
                ("sub", JfSubPdfDesc * JF_MAX_SUBPDFS),
                ("target_col", C.c_int32 * JF_MAX_SUBPDFS), ("base_col", C.c_int32 * JF_MAX_SUBPDFS),
                ("emb_dim", C.c_int32 * JF_MAX_SUBPDFS), ("has_mlp", C.c_int32 * JF_MAX_SUBPDFS),
                ("mlp", JfMlpDesc * JF_MAX_SUBPDFS)]


class JfPdfParams(C.Structure):
    _fields_ = [("shared", C.c_void_p * JF_MAX_SUBPDFS),
                ("weights", (C.c_void_p * JF_MAX_MLP_LINEAR) * JF_MAX_SUBPDFS),
                ("biases", (C.c_void_p * JF_MAX_MLP_LINEAR) * JF_MAX_SUBPDFS)]


# every symbol include/jammy_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _i32p = C.c_void_p, C.c_int64, C.POINTER(C.c_int32)
SYMBOLS = {
    "jf_subpdf_apply": (C.c_int, [C.POINTER(JfSubPdfDesc), C.c_int, C.c_int, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp,
                                  _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    "jf_subpdf_backward": (C.c_int, [C.POINTER(JfSubPdfDesc), C.c_int, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    "jf_subpdf_sample_backward": (C.c_int, [C.POINTER(JfSubPdfDesc), C.c_int, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp,
                                            _vp, _i64, _i64, _vp, _vp]),
    "jf_subpdf_jacobian": (C.c_int, [C.POINTER(JfSubPdfDesc), C.c_int, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp,
                                     _i64, _vp, _vp]),
    "jf_subpdf_forward_backward": (C.c_int, [C.POINTER(JfSubPdfDesc), C.c_int, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64,
                                             _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "jf_mlp_forward": (C.c_int, [C.POINTER(JfMlpDesc), C.c_int, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp),
                                 C.POINTER(_vp), _vp, _i64, _i64, _i64, _vp]),
    "jf_mlp_forward_acc": (C.c_int, [C.POINTER(JfMlpDesc), C.c_int, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp),
                                     C.POINTER(_vp), _vp, _i64, _i64, _i64, C.c_int, _vp]),
    "jf_mlp_workspace_bytes": (_i64, [C.POINTER(JfMlpDesc), C.c_int]),
    "jf_mlp_forward_ws": (C.c_int, [C.POINTER(JfMlpDesc), C.c_int, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp),
                                    C.POINTER(_vp), _vp, _i64, _i64, _i64, _vp, _i64, C.c_int, _vp]),
    "jf_mlp_backward_workspace_bytes": (_i64, [C.POINTER(JfMlpDesc), C.c_int, _i64]),
    "jf_mlp_backward": (C.c_int, [C.POINTER(JfMlpDesc), C.c_int, _vp, _i64, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _i64,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "jf_subpdf_generated_workspace_bytes": (_i64, [C.POINTER(JfSubPdfDesc), C.POINTER(JfMlpDesc), C.c_int]),
    "jf_subpdf_apply_generated": (C.c_int, [C.POINTER(JfSubPdfDesc), C.POINTER(JfMlpDesc), C.c_int, C.c_int,
                                            C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_vp), C.POINTER(_vp), _vp, _i64,
                                            _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, C.c_int, _vp, _vp]),
    "jf_pdf_workspace_bytes": (_i64, [C.POINTER(JfPdfDesc), _i64]),
    "jf_pdf_logpdf": (C.c_int, [C.POINTER(JfPdfDesc), C.POINTER(JfPdfParams), _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64,
                                _i64, _vp, _i64, _i64, _vp, _vp]),
    "jf_pdf_sample": (C.c_int, [C.POINTER(JfPdfDesc), C.POINTER(JfPdfParams), _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp,
                                _i64, _vp, _i64, _i64, _vp, _vp]),
    "jf_pdf_host_workspace_bytes": (_i64, [C.POINTER(JfPdfDesc), _i64]),
    "jf_pdf_logpdf_host": (C.c_int, [C.POINTER(JfPdfDesc), C.POINTER(JfPdfParams), _vp, _i64, _vp, _i64, _vp, _vp, _vp,
                                     _i64, _i64, _vp, _i64, _i64, _vp]),
    "jf_pdf_sample_host": (C.c_int, [C.POINTER(JfPdfDesc), C.POINTER(JfPdfParams), _vp, _i64, _vp, _i64, _vp, _i64, _vp,
                                     _vp, _i64, _vp, _i64, _i64, _vp]),
    "jf_pdf_transform_target": (C.c_int, [C.POINTER(JfPdfDesc), C.c_int, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp]),
    "jf_row_logmeanexp": (C.c_int, [C.c_int, _vp, _i64, _i64, _vp, _vp]),
    "jf_normal_rows": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, _i64, C.c_int32, _vp, _i64, _vp]),
    "jf_rowwise_linear": (C.c_int, [C.c_int, _vp, _i64, _i64, _i64, _vp, _i64, C.c_int32, C.c_int32, C.c_int, C.c_int,
                                    _vp, _i64, _i64, _i64, _vp]),
    "jf_abi_version": (C.c_int, []),
    "jf_probe_fma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), _vp, _vp]),
    "jf_launch_count": (_i64, []),
    "jf_struct_size": (_i64, [C.c_int]),
}

_lib = None


def load():
    """dlopen the library once; raise loudly when it is missing or its ABI does not match this binding."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "jammy_flows_b200: %s is missing. Build it with `python -m jammy_flows_b200.build` (needs nvcc). "
            "There is no CPU fallback: the sm_100a kernels are the only compute path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.jf_abi_version() != JF_ABI_VERSION:
        raise RuntimeError("libjammy_b200.so ABI version %d != binding %d" % (lib.jf_abi_version(), JF_ABI_VERSION))
    for which, st in enumerate((JfLayerDesc, JfSubPdfDesc, JfMlpDesc, JfPdfDesc, JfPdfParams, JfSplineDesc)):
        if lib.jf_struct_size(which) != C.sizeof(st):
            raise RuntimeError("ABI struct %s: library sizeof %d != binding %d"
                               % (st.__name__, lib.jf_struct_size(which), C.sizeof(st)))
    _lib = lib
    return lib


def check(rc, what):
    """Translate a C-ABI return code into a Python exception (the reference raises assert/Exception)."""
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError("%s failed: %s" % (what, ERRORS.get(rc, "error %d" % rc)))
    raise RuntimeError("%s failed: CUDA runtime error %d" % (what, rc))

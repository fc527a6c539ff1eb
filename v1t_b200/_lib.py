"""ctypes binding of libv1t_b200.so (the C-ABI declared in include/v1t_b200.h).

There is no CPU path and no fallback: if the library cannot be loaded this module raises, and every
wrapper raises ``RuntimeError`` with ``v1t_last_error()`` when a call fails (so callers such as the
reference's OOM probe, utils/utils.py:435-464, see an ordinary Python exception).
"""
from __future__ import annotations

import ctypes as C
import os

V1T_MAX_BLOCKS = 16
IMPL_FP32, IMPL_BF16X3, IMPL_BF16 = 0, 1, 2
PHASES = ["patch", "ln_qkv", "attn_fwd", "proj", "mlp", "attn_bwd", "linear_bwd", "readout_fwd", "readout_bwd",
          "attn_fwd_kernel", "attn_bwd_kernel", "attn_bwd_pair", "attn_bwd_dq"]  # the last four are nested scopes
NESTED_PHASES = ("attn_fwd_kernel", "attn_bwd_kernel", "attn_bwd_pair", "attn_bwd_dq")
IMPL_NAMES = {"fp32": IMPL_FP32, "bf16x3": IMPL_BF16X3, "exact": IMPL_BF16X3, "bf16": IMPL_BF16, "fast": IMPL_BF16}

_f32p = C.POINTER(C.c_float)


class CoreShape(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("in_ch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("patch", C.c_int32), ("stride", C.c_int32), ("emb", C.c_int32), ("heads", C.c_int32),
        ("mlp", C.c_int32), ("blocks", C.c_int32), ("bdim", C.c_int32), ("impl", C.c_int32),
        ("p_drop_tokens", C.c_float), ("p_drop_block", C.c_float), ("seed", C.c_uint64),
    ]


class CoreDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("gh", "gw", "tokens", "emb_ld", "inner", "mlp_ld", "patch_dim", "hid",
                                         "attn_path")]


ATTN_MATERIALISED, ATTN_FUSED = 0, 1


BLOCK_FIELDS = ("ln1_w", "ln1_b", "wqkv", "wproj", "bproj", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2",
                "bw0", "bb0", "bw3", "bb3")


class BlockPtrs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BLOCK_FIELDS]


class CorePtrs(C.Structure):
    _fields_ = [("cls", C.c_void_p), ("pos", C.c_void_p), ("wpe", C.c_void_p), ("bpe", C.c_void_p),
                ("blk", BlockPtrs * V1T_MAX_BLOCKS)]


class ReadoutShape(C.Structure):
    _fields_ = [("batch", C.c_int32), ("neurons", C.c_int32), ("channels", C.c_int32), ("gh", C.c_int32),
                ("gw", C.c_int32), ("fs_b", C.c_int64), ("fs_y", C.c_int64), ("fs_x", C.c_int64)]


class GemmDesc(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32), ("batch1", C.c_int32), ("batch2", C.c_int32),
                ("a_m", C.c_int64), ("a_k", C.c_int64), ("a_b1", C.c_int64), ("a_b2", C.c_int64),
                ("b_k", C.c_int64), ("b_n", C.c_int64), ("b_b1", C.c_int64), ("b_b2", C.c_int64),
                ("c_m", C.c_int64), ("c_b1", C.c_int64), ("c_b2", C.c_int64),
                ("r_m", C.c_int64), ("r_b1", C.c_int64), ("r_b2", C.c_int64),
                ("alpha", C.c_float), ("accumulate", C.c_int32)]


class OptTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("lr", C.c_float), ("l1", C.c_float), ("weight_decay", C.c_float),
                ("group", C.c_int32)]


MLP_MAX_LAYERS, MLP_MAX_WIDTH = 3, 32
ACT_NONE, ACT_TANH, ACT_ELU = 0, 1, 2


class MlpSpec(C.Structure):
    _fields_ = [("rows", C.c_int32), ("layers", C.c_int32), ("width", C.c_int32 * (MLP_MAX_LAYERS + 1)),
                ("act", C.c_int32 * MLP_MAX_LAYERS), ("x_ld", C.c_int64)]


class MlpPtrs(C.Structure):
    _fields_ = [("w", C.c_void_p * MLP_MAX_LAYERS), ("b", C.c_void_p * MLP_MAX_LAYERS)]


class CropShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "channels", "in_h", "in_w", "crop_h", "crop_w", "out_h", "out_w",
                                         "behavior_planes")]


ENSEMBLE_MAX = 16


class EnsembleMembers(C.Structure):
    _fields_ = [("x", C.c_void_p * ENSEMBLE_MAX), ("count", C.c_int32)]


LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libv1t_b200.so")

# every symbol include/v1t_b200.h declares: (restype, argtypes)
_vp, _i64, _f = C.c_void_p, C.c_int64, C.c_float
SYMBOLS = {
    "v1t_last_error": (C.c_char_p, []),
    "v1t_version": (C.c_int, []),
    "v1t_launch_count": (C.c_uint64, []),
    "v1t_prof_enable": (C.c_int, [C.c_int]),
    "v1t_prof_reset": (C.c_int, []),
    "v1t_prof_read": (C.c_int, [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "v1t_core_dims_of": (C.c_int, [C.POINTER(CoreShape), C.POINTER(CoreDims)]),
    "v1t_core_saved_bytes": (C.c_size_t, [C.POINTER(CoreShape)]),
    "v1t_core_scratch_bytes": (C.c_size_t, [C.POINTER(CoreShape)]),
    "v1t_core_forward": (C.c_int, [C.POINTER(CoreShape), C.POINTER(CorePtrs), _vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "v1t_core_backward": (C.c_int, [C.POINTER(CoreShape), C.POINTER(CorePtrs), _vp, _vp, _vp, _vp, _vp,
                                    C.POINTER(CorePtrs), _vp, _vp]),
    "v1t_attention_probs": (C.c_int, [C.POINTER(CoreShape), _vp, C.c_int, _vp, _vp]),
    "v1t_readout_scratch_bytes": (C.c_size_t, [C.POINTER(ReadoutShape)]),
    "v1t_readout_forward": (C.c_int, [C.POINTER(ReadoutShape), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f,
                                      _vp, _vp, _vp, _vp, _vp]),
    "v1t_readout_backward": (C.c_int, [C.POINTER(ReadoutShape), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "v1t_elu1_forward": (C.c_int, [_vp, _vp, _i64, _vp]),
    "v1t_elu1_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "v1t_poisson_scratch_bytes": (C.c_size_t, [_i64]),
    "v1t_poisson_forward": (C.c_int, [_vp, _vp, _i64, _f, _f, _vp, _vp, _vp]),
    "v1t_poisson_backward": (C.c_int, [_vp, _vp, _i64, _f, _f, _vp, _vp, _vp]),
    "v1t_gemm_fp32": (C.c_int, [C.POINTER(GemmDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "v1t_gemm_tc_set_mn_major": (C.c_int, [C.c_int]),
    "v1t_gemm_tc": (C.c_int, [C.POINTER(GemmDesc), _vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "v1t_matrix_plane_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "v1t_matrix_planes": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "v1t_gemm_tc_planes": (C.c_int, [C.POINTER(GemmDesc), _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int64,
                                     C.c_int64, _vp, _vp, C.c_int64, C.c_int64, _vp]),
    "v1t_attn_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "v1t_attn_forward": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f, C.c_uint64, C.c_uint32,
                                   _vp, _vp, _vp, _vp]),
    "v1t_attn_backward": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f, C.c_uint64,
                                    C.c_uint32, _vp, _vp, _vp]),
    "v1t_dropout_mask": (C.c_int, [_vp, _i64, C.c_uint64, C.c_uint32, _f, _vp]),
    "v1t_opt_chunk_elems": (C.c_int, []),
    "v1t_adamw_l1_scratch_bytes": (C.c_size_t, [C.c_int]),
    "v1t_adamw_l1_step": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_int, _vp, C.c_int, _vp, _vp]),
    "v1t_small_mlp_scratch_bytes": (C.c_size_t, [C.POINTER(MlpSpec)]),
    "v1t_small_mlp_forward": (C.c_int, [C.POINTER(MlpSpec), C.POINTER(MlpPtrs), _vp, _vp, _vp]),
    "v1t_small_mlp_backward": (C.c_int, [C.POINTER(MlpSpec), C.POINTER(MlpPtrs), _vp, _vp, C.POINTER(MlpPtrs), _vp,
                                         _vp]),
    "v1t_crop_resize": (C.c_int, [C.POINTER(CropShape), _vp, _vp, _vp, _vp, _vp, _vp]),
    "v1t_ensemble_scratch_bytes": (C.c_size_t, [_i64, C.c_int]),
    "v1t_ensemble_forward": (C.c_int, [C.POINTER(EnsembleMembers), _vp, _vp, _i64, _vp, _vp]),
    "v1t_ensemble_backward": (C.c_int, [C.POINTER(EnsembleMembers), _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "v1t_rollout_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "v1t_attention_rollout": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        _vp, _vp, _vp]),
}

# include/v1t_b200_diag.h: micro-benchmarks / self-test, in their own library (not the product ABI)
DIAG_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libv1t_b200_diag.so")
DIAG_SYMBOLS = {
    "v1t_diag_last_error": (C.c_char_p, []),
    "v1t_mma_microbench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "v1t_bulk_microbench": (C.c_int, [_vp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "v1t_ts_selftest": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "v1t_diag_attn_pair_trace": (C.c_int, [_vp]),
}
PRODUCT_HOSTED_DIAG = ("v1t_diag_attn_pair_trace",)  # diagnostics hooks that live inside libv1t_b200.so itself

_lib = None
_diag = None


def load_diag():
    """The diagnostics library (scripts/*_microbench.py, one GPU self-test)."""
    global _diag
    if _diag is None:
        load()  # builds both libraries when needed
        lib = C.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in DIAG_SYMBOLS.items():
            fn = getattr(_lib if name in PRODUCT_HOSTED_DIAG else lib, name)
            fn.restype = res
            fn.argtypes = args
            if name in PRODUCT_HOSTED_DIAG:
                setattr(lib, name, fn)
        _diag = lib
    return _diag


def load(build_if_missing: bool = True):
    """Load (building in-tree first if the .so is absent and nvcc exists).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import _build

    if not _build.is_current():
        # absent, or stale relative to csrc/ + include/v1t_b200.h (a struct-layout change behind ctypes would corrupt
        # memory silently): rebuild when nvcc is here, otherwise refuse to load
        import shutil

        have_nvcc = shutil.which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")
        if not build_if_missing or not have_nvcc:
            state = "stale (sources changed since it was built)" if os.path.exists(LIB_PATH) else "missing"
            raise RuntimeError(f"{LIB_PATH} is {state}; run `python -m v1t_b200._build`")
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().v1t_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        msg = last_error()
        # the reference's micro-batch probe expects RuntimeError on out-of-memory (utils/utils.py:435-464)
        raise RuntimeError(f"v1t_b200.{what} failed (code {rc}): {msg}")

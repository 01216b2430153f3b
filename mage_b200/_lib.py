"""ctypes binding of libmage_sm100.so (the C ABI in include/mage_b200.h).

There is no Python or CPU fallback: if the shared library is missing the import of any
product module fails with instructions to build it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.environ.get("MAGE_LIB") or os.path.join(CSRC, "libmage_sm100.so")   # MAGE_LIB: experiment builds (tools/experiments)

ABI_VERSION = 5   # include/mage_b200.h: mage_abi_version()

_c_f = ctypes.c_void_p  # device pointers travel as integers
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f32 = ctypes.c_float

# name -> argtypes (restype is int for all but the two bookkeeping calls)
SIGNATURES = {
    "mage_abi_version": [],
    "mage_ctx_create": [_i, ctypes.POINTER(ctypes.c_void_p)],   # no ctx argument (special-cased below)
    "mage_ctx_destroy": [_c_f],
    "mage_ctx_device": [_c_f],
    "mage_launch_count": [_c_f],
    "mage_pdl": [_c_f] + [_i],
    "mage_sm_share": [_c_f] + [_i],
    "mage_temporal_attn_ring": [_c_f] + [_i],
    "mage_gemm_f32": [_c_f, _c_f, _i64, _c_f, _i64, _c_f, _c_f, _i64, _i, _c_f, _i64, _i, _i, _i, _i, _i, _c_f],
    "mage_conv2d_nhwc_f32": [_c_f] + [_c_f] * 5 + [_i] * 22 + [_i64, _c_f],
    "mage_tc_tuning": [_c_f] + [_i, _i],
    "mage_tc_conv_halo": [_c_f] + [_i],
    "mage_tc_nsplit": [_c_f] + [_i],
    "mage_split_f32": [_c_f, _c_f, _i64, _c_f, _i64, _i, _i, _i, _c_f, _c_f],
    "mage_patch_rows_split_f32": [_c_f, _c_f, _c_f, _i64, _i, _i, _i, _i, _i, _i, _c_f],
    "mage_s2d_pad_split_f32": [_c_f, _c_f, _c_f, _i64, _i, _i, _i, _i, _i, _c_f, _c_f],
    "mage_embedding_split": [_c_f, _c_f, _c_f, _i64, _c_f, _i64, _i, _i, _c_f],
    "mage_gemm_tc": [_c_f, _c_f, _i64, _i64, _c_f, _i64, _i64, _c_f, _c_f, _i64, _i, _c_f, _c_f, _c_f, _i64, _i64, _i, _i, _i, _i,
                     _c_f, _c_f],
    "mage_gemm_tc_ln": [_c_f, _c_f, _i64, _i64, _c_f, _i64, _i64, _c_f, _c_f, _i64, _c_f, _i, _i, _c_f, _c_f, _f32, _c_f, _i64, _c_f, _c_f, _c_f],
    "mage_token_taps_ln_f32": [_c_f] + [_c_f] * 5 + [_i] * 6 + [_c_f, _c_f, _f32, _c_f, _i64, _c_f, _c_f],
    "mage_qkv_axial_attn_tc": [_c_f, _c_f, _i64, _c_f, _i64, _c_f, _c_f, _i64, _i, _i, _i, _i, _i, _f32, _c_f, _c_f],
    "mage_conv2d_tc": [_c_f, _c_f, _i64, _c_f, _i64, _c_f, _c_f, _c_f, _c_f, _c_f, _i64] + [_i] * 19 + [_i64, _i, _c_f, _c_f],
    "mage_conv2d_tc_pixel_head": [_c_f, _c_f, _i64, _c_f, _i64, _c_f, _c_f] + [_i] * 12 + [_c_f, _c_f, _i, _c_f, _i64, _i, _c_f, _c_f],
    "mage_conv2d_first_f32": [_c_f] + [_c_f] * 4 + [_i] * 12 + [_c_f],
    "mage_conv1x1_tanh_nchw_f32": [_c_f] + [_c_f] * 4 + [_i] * 4 + [_i64, _c_f],
    "mage_maxpool2x2_nhwc_f32": [_c_f, _c_f, _c_f, _i, _i, _i, _i, _c_f],
    "mage_layernorm_f32": [_c_f] + [_c_f] * 5 + [_i64, _c_f, _i, _i, _f32, _c_f],
    "mage_mha_f32": [_c_f] + [_c_f] * 4 + [_i] * 5 + [_i64] * 12 + [_c_f, _f32, _c_f, _i64, _c_f, _c_f],
    "mage_axial_attn_f32": [_c_f, _c_f, _c_f, _c_f, _i64, _c_f, _i, _i, _i, _i, _f32, _c_f],
    "mage_temporal_attn_step_f32": [_c_f] + [_c_f] * 5 + [_i64, _c_f, _i, _i, _i, _f32, _c_f],
    "mage_temporal_attn_seq_f32": [_c_f] + [_c_f] * 4 + [_i64, _c_f, _i, _i, _i, _i, _f32, _c_f],
    "mage_kv_append_f32": [_c_f] + [_c_f] * 3 + [_i] * 4 + [_c_f],
    "mage_vq_argmin_f32": [_c_f] + [_c_f] * 4 + [_i] * 3 + [_c_f],
    "mage_argmax_rows_f32": [_c_f, _c_f, _i64, _c_f, _i, _i, _c_f],
    "mage_embedding_f32": [_c_f] + [_c_f] * 3 + [_i, _i, _c_f],
    "mage_token_taps_f32": [_c_f] + [_c_f] * 5 + [_i] * 6 + [_c_f],
    "mage_text_embed_f32": [_c_f] + [_c_f] * 7 + [_i] * 4 + [_f32, _i, _c_f, _c_f],
    "mage_adain_nhwc_f32": [_c_f] + [_c_f] * 4 + [_i] * 3 + [_f32, _c_f],
    "mage_add_scaled_vec_f32": [_c_f] + [_c_f] * 3 + [_i] * 3 + [_c_f],
    "mage_nchw_to_nhwc_f32": [_c_f, _c_f, _c_f, _i, _i, _i, _c_f],
    "mage_gn_partial_f32": [_c_f, _c_f, _c_f, _i, _i, _i, _i, _i, _c_f],
    "mage_gn_silu_head_f32": [_c_f] + [_c_f] * 7 + [_i] * 7 + [_f32, _c_f],
    "mage_gn_apply_f32": [_c_f] + [_c_f] * 8 + [_i64, _c_f] + [_i] * 6 + [_f32, _c_f],
    "mage_cross_entropy_rows_f32": [_c_f, _c_f, _i64, _c_f, _c_f, _i, _i, _c_f, _c_f],
    "mage_reparam_kl_f32": [_c_f] + [_c_f] * 4 + [_i] * 3 + [_c_f],
    "mage_scaled_sum_f32": [_c_f, _c_f, _c_f, _i64, ctypes.c_double, _c_f],
    "mage_scaled_sqdiff_sum_f32": [_c_f, _c_f, _c_f, _c_f, _i64, ctypes.c_double, _c_f],
}


def build(verbose: bool = False) -> str:
    """Compile libmage_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libmage_sm100.so failed (see output above)")
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the MAGE sampling path has no CPU/PyTorch fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C mage_b200/csrc`.")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.argtypes = args
            fn.restype = _i64 if name == "mage_launch_count" else _i
        _lib = L
    return _lib


class MageCudaError(RuntimeError):
    pass


class MageSplitRangeError(MageCudaError):
    """A tensor-core operand left the range of the fp16 hi/lo split format (|x| > 65504, or NaN) during the call."""


_ctx = {}


def ctx(device_index: int) -> int:
    """The library handle (`mage_ctx*`, include/mage_b200.h) of a CUDA device: created on first use, one per (process, device).
    All state the library keeps between calls -- per-device kernel configuration, tuning switches, launch counter -- lives in it."""
    h = _ctx.get(device_index)
    if h is None:
        out = ctypes.c_void_p()
        check(lib().mage_ctx_create(int(device_index), ctypes.byref(out)), f"mage_ctx_create(device {device_index})")
        h = _ctx[device_index] = out.value
    return h


def check(code: int, what: str) -> None:
    if code != 0:
        kind = {-1: "MAGE_EINVAL (unsupported shape/alignment)", -2: "MAGE_ENOTSUP"}.get(code, f"cudaError {code}")
        raise MageCudaError(f"{what} failed: {kind}")

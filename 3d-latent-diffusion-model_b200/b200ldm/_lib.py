"""ctypes binding of libb200ldm.so (the C ABI declared in include/b200ldm.h).

The library is built in-tree by `python 3d-latent-diffusion-model_b200/build.py`.  There is NO
fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200ldm.so")

F32, BF16 = 0, 1
ENGINE_AUTO, ENGINE_TCGEN05, ENGINE_SIMT = 0, 1, 2


class ConvParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x2", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("temb", C.c_void_p),
        ("residual", C.c_void_p), ("y", C.c_void_p), ("gn_stats", C.c_void_p),
        ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("C1", C.c_int32), ("C2", C.c_int32), ("Cout", C.c_int32),
        ("Do", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad_lo", C.c_int32),
        ("temb_stride", C.c_int32), ("gn_groups", C.c_int32), ("dtype", C.c_int32),
        ("engine", C.c_int32), ("out_f32", C.c_int32), ("upsample2x", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("chunk_stats", C.c_void_p),
        ("skip_x", C.c_void_p), ("skip_x2", C.c_void_p), ("skip_w", C.c_void_p),
        ("skip_C1", C.c_int32), ("skip_C2", C.c_int32), ("chunk_width", C.c_int32), ("defer_finalize", C.c_int32),
    ]


class SchedParams(C.Structure):
    _fields_ = [
        ("eps", C.c_void_p), ("x_t", C.c_void_p), ("noise", C.c_void_p), ("x0_out", C.c_void_p),
        ("x_model", C.c_void_p), ("coef", C.c_void_p), ("counter", C.c_void_p),
        ("n_vox", C.c_int64), ("noise_step_stride", C.c_int64),
        ("C", C.c_int32), ("C_eps", C.c_int32), ("C_model", C.c_int32), ("dtype", C.c_int32),
        ("seed", C.c_uint64), ("seed_dev", C.c_void_p), ("var_pred", C.c_void_p), ("var_coef", C.c_void_p),
        ("philox_offset", C.c_int64),
    ]


# name -> (restype, argtypes); mirrors include/b200ldm.h one to one (tests check the export list)
_vp, _i32, _i64, _u64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
SIGNATURES = {
    "b200ldm_version": (C.c_int, []),
    "b200ldm_last_error": (C.c_int, [C.c_char_p, _sz]),
    "b200ldm_check_device": (C.c_int, []),
    "b200ldm_launch_count": (C.c_uint64, []),
    "b200ldm_sizeof_params": (_sz, [_i32]),
    "b200ldm_conv3d": (C.c_int, [C.POINTER(ConvParams), _vp]),
    "b200ldm_conv3d_tune": (C.c_int, [C.POINTER(ConvParams), _vp]),
    "b200ldm_conv3d_tune_pending": (C.c_int, [C.POINTER(ConvParams)]),
    "b200ldm_conv3d_workspace_bytes": (C.c_int64, [C.POINTER(ConvParams)]),
    "b200ldm_conv3d_emits_chunk_stats": (C.c_int, [C.POINTER(ConvParams)]),
    "b200ldm_conv3d_splits": (C.c_int, [C.POINTER(ConvParams)]),
    "b200ldm_conv3d_finalize": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "b200ldm_groupnorm_small_slabs": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32,
                                                _i32, _i32, _f32, _i32, _vp]),
    "b200ldm_repack_conv_weight": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200ldm_repack_upconv_weight": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200ldm_groupnorm_stats": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp]),
    "b200ldm_groupnorm_apply": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _f32, _i32,
                                          _i32, _vp]),
    "b200ldm_groupnorm_apply_chunked": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _f32,
                                                  _i32, _i32, _vp]),
    "b200ldm_groupnorm_apply_chunked_w": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32,
                                                    _f32, _i32, _i32, _i32, _vp]),
    "b200ldm_groupnorm_small_fits": (C.c_int, [_i64, _i32, _i32, _i32]),
    "b200ldm_groupnorm_small": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _f32, _i32, _i32, _vp]),
    "b200ldm_fill_zero": (C.c_int, [_vp, _sz, _vp]),
    "b200ldm_attention": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _f32, _i32, _i32, _vp]),
    "b200ldm_attention_workspace_bytes": (C.c_int64, [_i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "b200ldm_attention_ws": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _f32, _i32, _i32, _vp, _i64, _vp]),
    "b200ldm_cross_attention_ws": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _f32, _i32, _i32,
                                             _vp, _i64, _vp]),
    "b200ldm_cross_attention": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _f32, _i32, _i32,
                                          _vp]),
    "b200ldm_layernorm": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _vp]),
    "b200ldm_geglu": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "b200ldm_embedding_add": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "b200ldm_tap_gather": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                     C.POINTER(SchedParams), _vp]),
    "b200ldm_upsample_nearest2x": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200ldm_pack_taps": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200ldm_ncdhw_to_ndhwc": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _i32, _i32, _i32, _vp]),
    "b200ldm_ndhwc_to_ncdhw": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _i32, _i32, _vp]),
    "b200ldm_silu_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "b200ldm_timestep_embedding": (C.c_int, [_vp, _vp, _i32, _i32, _f32, _vp]),
    "b200ldm_select_row": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "b200ldm_counter_add": (C.c_int, [_vp, _i32, _vp]),
    "b200ldm_scheduler_step": (C.c_int, [C.POINTER(SchedParams), _vp]),
    "b200ldm_philox_normal": (C.c_int, [_vp, _i64, _u64, _i32, _vp]),
    "b200ldm_ae_sampling": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "b200ldm_mul_add_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "b200ldm_add_noise": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built - never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python 3d-latent-diffusion-model_b200/build.py` "
            "(b200ldm has no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype, fn.argtypes = res, args
    for which, st in ((0, ConvParams), (1, SchedParams)):
        if lib.b200ldm_sizeof_params(which) != C.sizeof(st):
            raise RuntimeError(f"{st.__name__}: ctypes layout ({C.sizeof(st)} B) != compiled struct "
                               f"({lib.b200ldm_sizeof_params(which)} B); rebuild libb200ldm.so")
    _lib = lib
    return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().b200ldm_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        names = {1: "BAD_SHAPE", 2: "BAD_ALIGN", 3: "UNSUPPORTED_ARCH", 4: "CUDA_ERROR", 5: "BAD_ARG"}
        raise RuntimeError(f"b200ldm {what} failed [{names.get(rc, rc)}]: {last_error()}")

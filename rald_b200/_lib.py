"""ctypes binding of librald_b200.so (the C ABI declared in include/rald_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_LIB = None
# RALD_B200_LIB: an alternative build of the same sources (kernel-variant experiments)
LIB_PATH = Path(os.environ.get("RALD_B200_LIB") or Path(__file__).resolve().parent / "_C" / "librald_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_f64 = ctypes.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "rald_abi_version": [],
    "rald_last_error": [],
    "rald_launch_count": [],
    "rald_launch_count_add": [ctypes.c_uint64],
    "rald_tmap_cache_stats": [c_void_p, c_void_p],
    "rald_prof_enable": [ctypes.c_uint],
    "rald_prof_collect": [c_int, c_void_p, c_void_p, c_void_p],
    "rald_prof_dump": [c_int, c_void_p, c_void_p, c_i64],
    "rald_gemm_bf16": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_i64,
                       c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rald_gemm_bf16_accum": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_void_p],
    "rald_gemm_bf16_accum_shift": [c_void_p, c_i64, c_void_p, c_i64, c_int, c_void_p, c_i64, c_int, c_int, c_int, c_void_p],
    "rald_gn_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_f32, c_int, c_void_p,
                    c_void_p, c_void_p, c_void_p],
    "rald_enc_pad_transpose": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_void_p, c_i64, c_int, c_int, c_int, c_void_p, c_void_p],
    "rald_enc_stuff": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "rald_enc_attn_bwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "rald_gemm_bf16_accum_taps": [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_i64,
                                  c_int, c_int, c_void_p],
    "rald_gemm_bf16_f16cols": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_void_p],
    "rald_gemm_bf16_wsplit": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_i64,
                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rald_gemm_debug_buffer": [c_void_p],
    "rald_attn_debug_buffer": [c_void_p],
    "rald_attn_streams_debug_buffer": [c_void_p],
    "rald_xattn_debug_buffer": [c_void_p],
    "rald_ae_query_debug_buffer": [c_void_p],
    "rald_attn_d64": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_int,
                      c_f32, c_void_p],
    "rald_attn_d64_long": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_int,
                           c_f32, c_void_p, c_void_p, c_void_p],
    "rald_ln_rows": [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_i64, c_int, c_i64, c_int,
                     c_f32, c_void_p],
    "rald_dit_mod_table": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "rald_dit_boundary": [c_void_p] * 10 + [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int, c_i64, c_int, c_f32,
                                            c_void_p, c_void_p],
    "rald_dit_boundary_pack": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "rald_dit_boundary_pack_bytes": [],
    "rald_radar_tokens": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "rald_dit_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_int,
                         c_void_p],
    "rald_dit_sample": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                        c_void_p],
    "rald_xattn_fold": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "rald_xattn_fused": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rald_xattn_split": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rald_ae_stack": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "rald_linear_smallk": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p],
    "rald_ln_dot_rows": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_f32, c_void_p],
    "rald_ae_query": [c_void_p, c_int, c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                      c_void_p, c_void_p, c_int, c_int, c_void_p],
    "rald_conv3d_cl": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                       c_int, c_void_p],
    "rald_enc_conv_in": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rald_gn_stats": [c_void_p, c_int, c_i64, c_int, c_int, c_void_p, c_void_p],
    "rald_gn_apply": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_f32, c_int,
                      c_void_p],
    "rald_enc_attn": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "rald_radar_encoder": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rald_point_features": [c_void_p, c_void_p, c_int, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_void_p],
    "rald_fps": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "rald_ae_posterior": [c_void_p, c_i64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p],
    "rald_ae_encode_stats": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "rald_occupancy_compact": [c_void_p, c_void_p, c_int, c_i64, c_f32, c_void_p, c_int, c_i64, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p],
    "rald_occupancy_ws_elems": [c_int, c_i64],
    "rald_refine_queries": [c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, ctypes.c_uint64, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "rald_chamfer": [c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "rald_chamfer_ws_elems": [c_int, c_i64, c_i64],
    "rald_radar_cube_prep": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_int,
                             c_f32, c_void_p, c_void_p],
    "rald_ema_update": [c_void_p, c_int, c_i64, c_i64, c_f32, c_f32, c_void_p],
    "rald_ema_chunk_elems": [],
    "rald_occupancy_iou": [c_void_p, c_void_p, c_int, c_i64, c_f32, c_void_p, c_void_p, c_void_p],
    "rald_attn_d64_stats": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_int,
                            c_int, c_f32, c_void_p, c_void_p],
    "rald_attn_d64_bwd": [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64,
                          c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_int, c_int,
                          c_int, c_int, c_f32, c_void_p],
    "rald_cast_transpose": [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p],
    "rald_colsum_finish": [c_void_p, c_int, c_i64, c_void_p, c_int, c_void_p],
    "rald_center_cast_f16_bf16": [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_i64, c_void_p],
    "rald_colsum": [c_void_p, c_int, c_i64, c_i64, c_i64, c_void_p, c_i64, c_void_p, c_int, c_void_p],
    "rald_ln_bwd": [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_int, c_void_p, c_i64, c_void_p, c_i64,
                    c_i64, c_int, c_i64, c_int, c_f32, c_void_p],
    "rald_geglu_fwd": [c_void_p, c_i64, c_int, c_void_p, c_void_p],
    "rald_geglu_bwd": [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p],
    "rald_sgemm_f32": [c_int, c_int, c_int, c_int, c_int, c_f32, c_void_p, c_i64, c_void_p, c_i64, c_f32, c_void_p,
                       c_i64, c_void_p],
    "rald_radar_tokens_bwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p],
}
_RESTYPES = {"rald_last_error": ctypes.c_char_p, "rald_launch_count_add": None, "rald_launch_count": ctypes.c_uint64,
             "rald_occupancy_ws_elems": c_i64, "rald_prof_dump": c_i64, "rald_chamfer_ws_elems": c_i64,
             "rald_dit_boundary_pack_bytes": c_i64}


class RaldError(RuntimeError):
    pass


class RuntimeNotCopied:
    """Mixin of the per-module runtimes (packed device weights, workspaces, ctypes structs, CUDA graphs): they are
    derived state, so ``copy.deepcopy(module)`` (ema_model = deepcopy(model)) and pickling (torch.save(module)) drop
    them — the copy rebuilds its own runtime lazily on first use."""

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (_no_runtime, ())


def _no_runtime():
    return None


def lib() -> ctypes.CDLL:
    """Loads the shared library once. Raises if it has not been built (python -m rald_b200.build)."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            if os.environ.get("RALD_B200_AUTOBUILD", "1") == "1":
                from . import build as _build
                _build.build()
            if not LIB_PATH.exists():
                raise RaldError(f"{LIB_PATH} is missing: build it with `python -m rald_b200.build` "
                                "(there is no fallback path)")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _LIB = handle
    return _LIB


def exported_symbols():
    return list(_SIGNATURES.keys())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rald_last_error()
        raise RaldError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


FAMILIES = {"gemm": 0, "attn": 1, "ln": 2, "boundary": 3, "conv3d": 4, "gn": 5, "ae_query": 6, "other": 7, "fps": 8,
            "xattn": 9}


def boundary_pack_bytes() -> int:
    return int(lib().rald_dit_boundary_pack_bytes())


def launch_count() -> int:
    return int(lib().rald_launch_count())


def prof_enable(*families: str) -> None:
    mask = 0
    for f in families:
        mask |= 1 << FAMILIES[f]
    check(lib().rald_prof_enable(mask), "rald_prof_enable")


def prof_collect(family: str):
    """(total ms, total algorithmic work, launches) of one kernel family since prof_enable()."""
    ms, work, n = c_f64(), c_f64(), c_i64()
    check(lib().rald_prof_collect(FAMILIES[family], ctypes.addressof(ms), ctypes.addressof(work), ctypes.addressof(n)),
          "rald_prof_collect")
    return ms.value, work.value, n.value


def prof_dump(family: str, cap: int = 1 << 16):
    """[(ms, work)] of every recorded launch of one kernel family since prof_enable()."""
    import numpy as np
    ms = np.zeros(cap, dtype=np.float32)
    work = np.zeros(cap, dtype=np.float64)
    n = int(lib().rald_prof_dump(FAMILIES[family], ms.ctypes.data, work.ctypes.data, cap))
    if n < 0:
        raise RaldError("rald_prof_dump failed")
    return ms[:n], work[:n]

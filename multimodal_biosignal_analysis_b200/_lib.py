"""ctypes binding of libcmc_b200.so (the C ABI declared in include/cmc.h).

No fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used only as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libcmc_b200.so")
# developer switch for A/B runs of instrumented or alternative builds (same ABI)
LIB_PATH = os.environ.get("CMC_B200_LIB", LIB_PATH)

_lib = None

_i32, _i64, _u64, _f32, _f64, _vp = C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_void_p

_SIGNATURES = {
    "cmc_abi_version": (C.c_int, []),
    "cmc_last_error": (C.c_char_p, []),
    "cmc_launch_count": (_i64, []),
    "cmc_fft_segments": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32,
                                   _vp, _i64, _vp]),
    "cmc_fft_prepare": (C.c_int, [_i32]),
    "cmc_fft_segments_pair": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _i32, _i64, _vp, _i64, _vp, _i32, _vp, _i32, _i32,
                                        _i32, _i32, _i32, _i64, _vp]),
    "cmc_welch_hann_plan_create": (C.c_int, [_vp, _i32, _i32, _i32, _i32, C.POINTER(C.c_void_p)]),
    "cmc_welch_hann_plan_destroy": (C.c_int, [_vp]),
    "cmc_welch_hann_plan_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cmc_welch_hann_spectra": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _i32, _i64, _vp, _i64, _i32, _i64, _vp]),
    "cmc_psd_from_spectra": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i64, _f32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "cmc_msc_windows": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _i32, _f32, _f32,
                                  _vp, _vp, _vp, _vp, _vp]),
    "cmc_msc_windows_maxemg": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _i32, _f32,
                                         _f32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cmc_csd_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "cmc_csd_workspace_bytes_min": (_i64, [_i32, _i32, _i32]),
    "cmc_csd_coherence": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cmc_csd_operands": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _i64, _vp]),
    "cmc_csd_msc": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cmc_phase_table": (C.c_int, [_vp]),
    "cmc_surrogate_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i64]),
    "cmc_surrogate_null": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _u64, _i64, _i64, _vp, _vp,
                                     _vp, _vp, _i64, _vp]),
    "cmc_surrogate_null_range": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _u64, _i64, _i64, _i32, _i32,
                                           _vp, _vp, _vp, _vp, _i64, _vp]),
    "cmc_surrogate_null_hist": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _u64, _i64, _i64, _i32, _i32, _i32,
                                          _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "cmc_hist_select": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "cmc_cbpa_workspace_bytes": (_i64, [_i32, _i32]),
    "cmc_cbpa_permute": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _i64, _f64, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cmc_cbpa_observed": (C.c_int, [_vp, _i32, _i32, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                    _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class CmcError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CmcError(
                f"{LIB_PATH} not found - build it with `python -m multimodal_biosignal_analysis_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        if lib.cmc_abi_version() != 1:
            raise CmcError("libcmc_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cmc_last_error().decode("utf-8", "replace")
        raise CmcError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None stays NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().cmc_launch_count())

"""ctypes loader for ``libosr_sm100a.so`` (the C ABI declared in ``include/osr.h``).

There is deliberately no fallback: if the library is missing or a call fails, the op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSR_LIB_PATH", os.path.join(_HERE, "libosr_sm100a.so"))  # override = A/B builds only
OSR_MAX_LEVELS = 8

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)


class RpnLevel(C.Structure):
    _fields_ = [
        ("deltas", C.c_void_p), ("scores", C.c_void_p), ("anchors", C.c_void_p),
        ("num_anchors", C.c_int64),
        ("delta_stride_n", C.c_int64), ("delta_stride_a", C.c_int64), ("delta_stride_c", C.c_int64),
        ("score_stride_n", C.c_int64), ("score_stride_a", C.c_int64),
    ]


class FeatLevel(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("sN", C.c_int64), ("sC", C.c_int64), ("sH", C.c_int64), ("sW", C.c_int64),
        ("H", C.c_int32), ("W", C.c_int32), ("scale", C.c_float),
    ]


# symbol -> (restype, argtypes); also the list tests/test_abi.py checks against include/osr.h
SIGNATURES = {
    "osr_version": (C.c_int, []),
    "osr_last_error": (C.c_char_p, []),
    "osr_launch_count": (C.c_longlong, []),
    "osr_reset_launch_count": (None, []),
    "osr_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "osr_get_tuning": (C.c_int, [C.c_int]),
    "osr_rpn_kmax": (C.c_int64, [C.POINTER(RpnLevel), C.c_int, C.c_int]),
    "osr_rpn_select_decode_workspace": (C.c_size_t, [C.POINTER(RpnLevel), C.c_int, C.c_int, C.c_int]),
    "osr_rpn_select_decode": (C.c_int, [
        C.POINTER(RpnLevel), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_nms_workspace": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "osr_nms_segmented": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_roi_align_fwd_workspace": (C.c_size_t, [C.c_int]),
    "osr_roi_align_fwd": (C.c_int, [
        C.POINTER(FeatLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_roi_align_fwd_bf16": (C.c_int, [
        C.POINTER(FeatLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_linear_bf16_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p]),
    "osr_cast_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "osr_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "osr_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "osr_roi_align_bwd_workspace": (C.c_size_t, [C.POINTER(FeatLevel), C.c_int, C.c_int, C.c_int, C.c_int]),
    "osr_roi_align_bwd": (C.c_int, [
        C.POINTER(FeatLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_roi_align_bwd_prepare": (C.c_int, [
        C.POINTER(FeatLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_roi_align_bwd_prepared": (C.c_int, [
        C.POINTER(FeatLevel), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "osr_pln_loss_fwd": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_loss_bwd": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_loss_fwd_bwd": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_loss_fwd_bwd_phase": (C.c_int, [
        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_encode_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "osr_pln_encode_fwd": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_encode_gather_fwd": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64,
        C.c_void_p, C.c_size_t, C.c_void_p]),
    "osr_pln_nearest": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int64, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p]),
    "osr_rcnn_decode_score": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
        C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "osr_match_label": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int64,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "osr_sample_rois": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


class OsrError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the library; raises if it is not built - there is no CPU fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise OsrError(
                        f"{LIB_PATH} not found: build it with `make -C openset-rcnn_b200` "
                        "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
                h = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    try:
                        fn = getattr(h, name)
                    except AttributeError:  # reported by missing_symbols(); calling it raises AttributeError
                        continue
                    fn.restype = res
                    fn.argtypes = args
                _lib = h
    return _lib


def missing_symbols():
    """Symbols declared in include/osr.h (SIGNATURES) that the built library does not export."""
    h = lib()
    return [name for name in SIGNATURES if not hasattr(h, name)]


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().osr_last_error().decode("utf-8", "replace")
        raise OsrError(f"{what} failed (status {rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise OsrError("osr_b200 ops run only on CUDA tensors (B200); there is no CPU fallback")


def launch_count() -> int:
    return int(lib().osr_launch_count())


def reset_launch_count() -> None:
    lib().osr_reset_launch_count()


TUNE_KEYS = {"bwd": 0, "fwd": 1, "pln": 2, "rpn": 3, "nms": 4, "bwd_split": 5}


def set_tuning(key: str, value: int) -> int:
    """Kernel-variant switch for A/B measurements (include/osr.h: osr_set_tuning); returns the previous value."""
    return int(lib().osr_set_tuning(TUNE_KEYS[key], int(value)))


def get_tuning(key: str) -> int:
    return int(lib().osr_get_tuning(TUNE_KEYS[key]))

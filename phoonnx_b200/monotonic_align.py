"""Monotonic alignment search on the B200 -- the host-side mirror of the reference's `monotonic_align` package.

Reference interface (same names, same argument meaning):
  * `maximum_path(neg_cent, mask)`                      /root/reference/phoonnx_train/vits/monotonic_align/__init__.py:7-21
  * `maximum_path_c(paths, values, t_ys, t_xs)`         /root/reference/phoonnx_train/vits/monotonic_align/core.pyx:38-42
called once per training step from `SynthesizerTrn.forward` (phoonnx_train/vits/models.py:646-650).  The reference copies neg_cent
to the host, runs an OpenMP loop over the batch, and copies the path back; here the tensors never leave the device: both functions
call `mas_maximum_path` (include/mas_b200.h, csrc/mas.cu) in libvits_b200.so on the caller's current CUDA stream.  There is no CPU
fallback: a CPU tensor or a missing library raises.

`install()` registers this module under the reference's package name, so an unmodified `models.py` picks it up.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import engine

MAS_DEVICE_PTRS = 0x1
MAS_PATH_F32 = 0x2
MAS_TIMED = 0x4
MAS_ROW_KERNEL = 0x8
MAS_SYMBOLS = ("mas_maximum_path", "mas_last_error", "mas_last_ms", "mas_last_forward_ms")

_bound = None


def _lib():
    global _bound
    if _bound is None:
        lib = engine.load_library()
        lib.mas_maximum_path.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.mas_maximum_path.restype = C.c_int
        lib.mas_last_error.restype = C.c_char_p
        lib.mas_last_ms.restype = C.c_float
        lib.mas_last_forward_ms.restype = C.c_float
        _bound = lib
    return _bound


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"mas_maximum_path failed ({rc}): {_lib().mas_last_error().decode()}")


def maximum_path_c(paths: np.ndarray, values: np.ndarray, t_ys: np.ndarray, t_xs: np.ndarray) -> None:
    """core.pyx:38-42 on host arrays: int32 `paths` [b, t_y, t_x] is filled in place.  Unlike the Cython loop, `values` is left
    untouched (the reference accumulates into it; its only caller passes a private copy)."""
    if paths.dtype != np.int32 or values.dtype != np.float32 or t_ys.dtype != np.int32 or t_xs.dtype != np.int32:
        raise TypeError("maximum_path_c: paths int32, values float32, t_ys / t_xs int32 (the Cython signature)")
    if paths.ndim != 3 or paths.shape != values.shape or t_ys.shape != (paths.shape[0],) or t_xs.shape != (paths.shape[0],):
        raise ValueError("maximum_path_c: paths and values [b, t_y, t_x], t_ys and t_xs [b]")
    for a in (paths, values, t_ys, t_xs):
        if not a.flags.c_contiguous:
            raise ValueError("maximum_path_c: arrays must be C-contiguous (`[:,:,::1]` memoryviews in the reference)")
    b, ty, tx = paths.shape
    _check(_lib().mas_maximum_path(paths.ctypes.data, values.ctypes.data, t_ys.ctypes.data, t_xs.ctypes.data, b, ty, tx, 0, None))


def maximum_path(neg_cent, mask):
    """monotonic_align/__init__.py:7-21.  neg_cent, mask: [b, t_t, t_s] CUDA tensors; returns the 0/1 path, same shape, device and
    dtype as neg_cent.  Stream-ordered on torch's current stream; no host synchronisation."""
    import torch
    if not neg_cent.is_cuda:
        raise RuntimeError("phoonnx_b200.monotonic_align.maximum_path needs CUDA tensors (there is no CPU fallback)")
    dtype = neg_cent.dtype
    values = neg_cent.detach().to(torch.float32).contiguous()
    t_t_max = mask.sum(1)[:, 0].to(torch.int32).contiguous()          # rows: spectrogram frames   (__init__.py:18)
    t_s_max = mask.sum(2)[:, 0].to(torch.int32).contiguous()          # columns: text positions    (__init__.py:19)
    b, ty, tx = values.shape
    path = torch.empty((b, ty, tx), dtype=torch.float32, device=values.device)
    with torch.cuda.device(values.device):
        stream = torch.cuda.current_stream().cuda_stream
        _check(_lib().mas_maximum_path(path.data_ptr(), values.data_ptr(), t_t_max.data_ptr(), t_s_max.data_ptr(), b, ty, tx,
                                       MAS_DEVICE_PTRS | MAS_PATH_F32, C.c_void_p(stream)))
    return path if dtype == torch.float32 else path.to(dtype)


def maximum_path_timed(values, t_ys, t_xs, row_kernel: bool = False):
    """Device tensors in (float32 [b, t_y, t_x], int32 [b], int32 [b]); returns (int32 path tensor, kernel milliseconds by CUDA
    events).  For tools/bench_mas.py and the tests."""
    import torch
    b, ty, tx = values.shape
    path = torch.empty((b, ty, tx), dtype=torch.int32, device=values.device)
    with torch.cuda.device(values.device):
        stream = torch.cuda.current_stream().cuda_stream
        _check(_lib().mas_maximum_path(path.data_ptr(), values.data_ptr(), t_ys.data_ptr(), t_xs.data_ptr(), b, ty, tx,
                                       MAS_DEVICE_PTRS | MAS_TIMED | (MAS_ROW_KERNEL if row_kernel else 0), C.c_void_p(stream)))
    return path, float(_lib().mas_last_ms())


def install(package: str = "phoonnx_train.vits.monotonic_align") -> None:
    """Make `from . import monotonic_align` in the reference's models.py resolve to this module."""
    sys.modules[package] = sys.modules[__name__]

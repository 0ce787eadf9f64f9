"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's monotonic alignment search.

Follows /root/reference/phoonnx_train/vits/monotonic_align/core.pyx:7-42 (`maximum_path_each`, `maximum_path_c`) and the
wrapper /root/reference/phoonnx_train/vits/monotonic_align/__init__.py:7-21 (`maximum_path`).  Only `tests/`, `tools/bench_mas.py`'s
CPU leg and `__graft_entry__.smoke()` may import this module; the product (`phoonnx_b200/monotonic_align.py` -> `mas_maximum_path` in
libvits_b200.so) never does.

Pinned (tests/test_oracle_mas.py): against `tests/golden/mas_cases.npz`, minted by `oracle/make_golden_mas.py` from the reference's
own `core.pyx` compiled here (`oracle/build_ref_mas.py` -> `oracle/_ref/monotonic_align/core*.so`), and against that compiled
reference directly on random cases whenever `oracle/_ref` is present.

Two restatements: `maximum_path_each_loops` is the Cython loop nest line for line (pure Python, small cases only);
`maximum_path_each` does a row at a time in numpy float32 (same additions in the same order per cell -- a row only reads the row
above), for the sizes the GPU tests use.
"""
from __future__ import annotations

import numpy as np

MAX_NEG_VAL = np.float32(-1e9)          # core.pyx:7


def maximum_path_each_loops(path: np.ndarray, value: np.ndarray, t_y: int, t_x: int) -> None:
    """core.pyx:7-34 as written (in place on `path` int32 [t_y_max, t_x_max] and `value` float32)."""
    index = t_x - 1
    for y in range(t_y):                                                      # core.pyx:16
        for x in range(max(0, t_x + y - t_y), min(t_x, y + 1)):              # core.pyx:17
            v_cur = MAX_NEG_VAL if x == y else value[y - 1, x]               # core.pyx:18-21
            if x == 0:                                                       # core.pyx:22-28
                v_prev = np.float32(0.0) if y == 0 else MAX_NEG_VAL
            else:
                v_prev = value[y - 1, x - 1]
            # Cython's max(v_prev, v_cur): `v_cur > v_prev ? v_cur : v_prev`
            value[y, x] = np.float32(value[y, x] + (v_cur if v_cur > v_prev else v_prev))     # core.pyx:29
    for y in range(t_y - 1, -1, -1):                                          # core.pyx:31
        if index < 0:
            break                                                             # t_x == 0: the reference would write out of bounds
        path[y, index] = 1                                                    # core.pyx:32
        # y == 0 with index != 0 (only when t_x > t_y) reads row -1 in the reference: undefined there, "no move" here
        if index != 0 and (index == y or (y > 0 and value[y - 1, index] < value[y - 1, index - 1])):    # core.pyx:33
            index -= 1


def maximum_path_each(path: np.ndarray, value: np.ndarray, t_y: int, t_x: int) -> None:
    """Same result as `maximum_path_each_loops`, one numpy row operation per y."""
    xs = np.arange(t_x)
    for y in range(t_y):
        lo, hi = max(0, t_x + y - t_y), min(t_x, y + 1)
        if hi <= lo:
            continue
        x = xs[lo:hi]
        if y == 0:
            v_cur = np.full(x.shape, MAX_NEG_VAL, np.float32)               # only x == 0 == y is in the band
            v_prev = np.zeros(x.shape, np.float32)
        else:
            v_cur = np.where(x == y, MAX_NEG_VAL, value[y - 1, lo:hi]).astype(np.float32)
            v_prev = np.where(x == 0, MAX_NEG_VAL, value[y - 1, np.maximum(x - 1, 0)]).astype(np.float32)
        value[y, lo:hi] += np.where(v_cur > v_prev, v_cur, v_prev)
    index = t_x - 1
    for y in range(t_y - 1, -1, -1):
        if index < 0:
            break
        path[y, index] = 1
        if index != 0 and (index == y or (y > 0 and value[y - 1, index] < value[y - 1, index - 1])):
            index -= 1


def maximum_path_c(paths: np.ndarray, values: np.ndarray, t_ys: np.ndarray, t_xs: np.ndarray, loops: bool = False) -> None:
    """core.pyx:38-42: every batch item independently (the reference uses an OpenMP prange).  In place, like the reference."""
    assert paths.dtype == np.int32 and values.dtype == np.float32
    each = maximum_path_each_loops if loops else maximum_path_each
    for i in range(paths.shape[0]):
        each(paths[i], values[i], int(t_ys[i]), int(t_xs[i]))


def maximum_path(neg_cent: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """monotonic_align/__init__.py:7-21 on numpy arrays: neg_cent, mask [b, t_t, t_s] -> path [b, t_t, t_s] in neg_cent's dtype."""
    dtype = neg_cent.dtype
    values = np.ascontiguousarray(neg_cent, dtype=np.float32).copy()
    path = np.zeros(values.shape, np.int32)
    t_t_max = mask.sum(1)[:, 0].astype(np.int32)
    t_s_max = mask.sum(2)[:, 0].astype(np.int32)
    maximum_path_c(path, values, t_t_max, t_s_max)
    return path.astype(dtype)

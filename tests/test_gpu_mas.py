"""Monotonic alignment search on the GPU (include/mas_b200.h -> csrc/mas.cu) against the oracle and the reference-minted goldens.
Integer / index work: the bar is bit-exact paths."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import mas_oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden", "mas_cases.npz")


def _oracle(values, t_ys, t_xs):
    paths = np.zeros(values.shape, np.int32)
    mas_oracle.maximum_path_c(paths, values.copy(), t_ys, t_xs)
    return paths


def _random_case(rs, b, ty, tx, scale=3.0, ragged=True):
    values = (rs.randn(b, ty, tx) * scale).astype(np.float32)
    if ragged:
        t_xs = rs.randint(1, min(tx, ty) + 1, size=b).astype(np.int32)
        t_ys = np.array([rs.randint(t_xs[i], ty + 1) for i in range(b)], np.int32)
        t_xs[0], t_ys[0] = min(tx, ty), ty
    else:
        t_xs = np.full(b, tx, np.int32); t_ys = np.full(b, ty, np.int32)
    return values, t_ys, t_xs


def test_golden_cases_host_pointers():
    """The reference's own outputs, through the Cython-shaped entry point on host arrays (library stages the copies)."""
    from phoonnx_b200 import monotonic_align as ma
    z = np.load(GOLDEN)
    for name in sorted({k.rsplit(".", 1)[0] for k in z.files}):
        values, t_ys, t_xs = z[name + ".values"], z[name + ".t_ys"], z[name + ".t_xs"]
        paths = np.full(values.shape, 7, np.int32)                 # every cell must be overwritten, zeros included
        before = values.copy()
        ma.maximum_path_c(paths, values, t_ys, t_xs)
        assert np.array_equal(paths, z[name + ".paths"].astype(np.int32)), name
        assert np.array_equal(values, before), "values must stay read-only"


def test_golden_cases_device_tensors_and_wrapper():
    import torch
    from phoonnx_b200 import monotonic_align as ma
    z = np.load(GOLDEN)
    for name in sorted({k.rsplit(".", 1)[0] for k in z.files}):
        values, t_ys, t_xs = z[name + ".values"], z[name + ".t_ys"], z[name + ".t_xs"]
        if (t_ys == 0).any() or (t_xs == 0).any():
            continue                                               # a mask cannot express an empty item next to a non-empty column 0
        b, ty, tx = values.shape
        mask = np.zeros((b, ty, tx), np.float32)
        for i in range(b):
            mask[i, :t_ys[i], :t_xs[i]] = 1
        for dt in (torch.float32, torch.float16):
            nc = torch.from_numpy(values).cuda().to(dt)
            got = ma.maximum_path(nc, torch.from_numpy(mask).cuda())
            assert got.dtype == dt and got.device == nc.device and got.shape == nc.shape
            want = _oracle(nc.float().cpu().numpy(), t_ys, t_xs)    # fp16 inputs: the reference also widens first (__init__.py:14)
            assert np.array_equal(got.float().cpu().numpy().astype(np.int32), want), (name, dt)


@pytest.mark.parametrize("b,ty,tx,what", [
    (8, 700, 180, "training-sized items, wavefront kernel, bit matrix in shared memory"),
    (3, 900, 1024, "wavefront kernel at its widest: 32 warps, 32-row FIFO"),
    (4, 130, 600, "wavefront kernel, 19 warps, short items (the boundary ring never wraps)"),
    (2, 5, 70, "fewer rows than one hand-over block"),
    (3, 1700, 1500, "two columns per thread"),
    (2, 3300, 3000, "four columns per thread, bit matrix in the global scratch"),
    (2, 6000, 1024, "bit matrix in global memory, 1024 threads"),
    (1, 60000, 40, "row index table in global memory"),
    (5, 257, 33, "odd pitch: scalar output stores"),
    (4, 64, 64, "square"),
])
def test_random_cases_match_oracle(b, ty, tx, what):
    import torch
    from phoonnx_b200 import monotonic_align as ma
    rs = np.random.RandomState(b * 1000 + tx)
    values, t_ys, t_xs = _random_case(rs, b, ty, tx)
    want = _oracle(values, t_ys, t_xs)
    dev = [torch.from_numpy(a).cuda() for a in (values, t_ys, t_xs)]
    path, ms = ma.maximum_path_timed(*dev)
    assert ms > 0
    assert np.array_equal(path.cpu().numpy(), want), what
    # the general (barrier-per-row) kernel on the same case; above 1024 columns it is what ran already
    path, _ = ma.maximum_path_timed(*dev, row_kernel=True)
    assert np.array_equal(path.cpu().numpy(), want), what + " [row kernel]"


def test_ties_and_items_without_a_monotonic_path():
    """Rounded values (many equal neighbours: `<` must keep the column) and items with t_x > t_y, where the reference reads one row
    above the array; the library defines that read as "no move" like the oracle."""
    import torch
    from phoonnx_b200 import monotonic_align as ma
    rs = np.random.RandomState(8)
    values = np.round(rs.randn(6, 90, 50) * 2).astype(np.float32)
    t_ys = np.array([90, 60, 50, 10, 1, 33], np.int32)
    t_xs = np.array([50, 50, 50, 30, 5, 34], np.int32)
    want = _oracle(values, t_ys, t_xs)
    path, _ = ma.maximum_path_timed(torch.from_numpy(values).cuda(), torch.from_numpy(t_ys).cuda(), torch.from_numpy(t_xs).cuda())
    assert np.array_equal(path.cpu().numpy(), want)


def test_full_training_batch_properties_and_stream_order():
    """b = 64 items of up to 1000 frames x 300 text positions on a side stream: equal to the oracle, and the properties every
    alignment has (one column per row, starts at 0, ends at t_x - 1, non-decreasing, steps of at most one)."""
    import torch
    from phoonnx_b200 import monotonic_align as ma
    rs = np.random.RandomState(21)
    values, t_ys, t_xs = _random_case(rs, 64, 1000, 300, scale=40.0)
    mask = np.zeros(values.shape, np.float32)
    for i in range(64):
        mask[i, :t_ys[i], :t_xs[i]] = 1
    side = torch.cuda.Stream()
    nc_host = torch.from_numpy(values).pin_memory(); mask_host = torch.from_numpy(mask).pin_memory()
    with torch.cuda.stream(side):
        nc = nc_host.cuda(non_blocking=True)                         # the kernel must queue behind this copy on the same stream
        got = ma.maximum_path(nc, mask_host.cuda(non_blocking=True))
    side.synchronize()
    p = got.cpu().numpy().astype(np.int32)
    assert np.array_equal(p, _oracle(values, t_ys, t_xs))
    for i in range(64):
        q = p[i, :t_ys[i], :t_xs[i]]
        assert p[i].sum() == t_ys[i] and np.all(q.sum(1) == 1)
        col = q.argmax(1)
        assert col[0] == 0 and col[-1] == t_xs[i] - 1 and np.diff(col).min(initial=0) >= 0 and np.diff(col).max(initial=0) <= 1


def test_argument_errors():
    from phoonnx_b200 import monotonic_align as ma
    paths = np.zeros((1, 4, 2), np.int32); values = np.zeros((1, 4, 2), np.float32); one = np.ones(1, np.int32)
    with pytest.raises(TypeError):
        ma.maximum_path_c(paths.astype(np.int64), values, one, one)
    with pytest.raises(ValueError):
        ma.maximum_path_c(paths, values[:, :3], one, one)
    lib = ma._lib()
    assert lib.mas_maximum_path(None, None, None, None, 1, 4, 2, 0, None) == -1 and b"null" in lib.mas_last_error()
    assert lib.mas_maximum_path(None, None, None, None, 0, 4, 2, 0, None) == 0          # empty batch: nothing to do
    big = 24577
    assert lib.mas_maximum_path(paths.ctypes.data, values.ctypes.data, one.ctypes.data, one.ctypes.data, 1, 1, big, 0, None) == -1


def test_install_registers_under_the_reference_package_name():
    import sys
    from phoonnx_b200 import monotonic_align as ma
    name = "phoonnx_train.vits.monotonic_align"
    saved = sys.modules.get(name)
    try:
        ma.install()
        assert sys.modules[name].maximum_path is ma.maximum_path
    finally:
        if saved is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = saved

"""The monotonic-alignment oracle (oracle/mas_oracle.py) against the reference: golden cases minted from the reference's own
core.pyx (tests/golden/mas_cases.npz, oracle/make_golden_mas.py) and, where oracle/_ref is present, the compiled reference itself."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import mas_oracle

GOLDEN = os.path.join(ROOT, "tests", "golden", "mas_cases.npz")


def golden_cases():
    z = np.load(GOLDEN)
    names = sorted({k.rsplit(".", 1)[0] for k in z.files})
    return [(n, z[n + ".values"], z[n + ".t_ys"], z[n + ".t_xs"], z[n + ".paths"].astype(np.int32)) for n in names]


def test_golden_file_has_the_edge_cases():
    names = {c[0] for c in golden_cases()}
    assert {"tx1", "square_forced_diagonal", "ragged_small", "ties", "garbage_padding", "empty_items", "crosses_a_warp"} <= names


@pytest.mark.parametrize("loops", [True, False])
def test_oracle_matches_reference_goldens(loops):
    for name, values, t_ys, t_xs, want in golden_cases():
        paths = np.zeros(values.shape, np.int32)
        mas_oracle.maximum_path_c(paths, values.copy(), t_ys, t_xs, loops=loops)
        assert np.array_equal(paths, want), name


def test_path_properties():
    """Size-independent properties of a valid alignment: one column per row, starts at 0, ends at t_x - 1, never moves left, moves by
    at most one, visits every column."""
    rs = np.random.RandomState(3)
    b, ty, tx = 6, 300, 90
    values = rs.randn(b, ty, tx).astype(np.float32) * 5
    t_ys = rs.randint(tx, ty + 1, size=b).astype(np.int32)
    t_xs = rs.randint(1, tx + 1, size=b).astype(np.int32)
    paths = np.zeros(values.shape, np.int32)
    mas_oracle.maximum_path_c(paths, values.copy(), t_ys, t_xs)
    for i in range(b):
        p = paths[i, :t_ys[i], :t_xs[i]]
        assert paths[i].sum() == t_ys[i] and np.all(p.sum(1) == 1)
        col = p.argmax(1)
        assert col[0] == 0 and col[-1] == t_xs[i] - 1
        d = np.diff(col)
        assert d.min() >= 0 and d.max() <= 1


def test_oracle_matches_compiled_reference_on_random_cases():
    from oracle import build_ref_mas
    ref = build_ref_mas.load()
    if ref is None:
        pytest.skip("oracle/_ref/monotonic_align not built (no /root/reference on this machine)")
    rs = np.random.RandomState(11)
    for trial in range(12):
        b = int(rs.randint(1, 5)); ty = int(rs.randint(1, 260)); tx = int(rs.randint(1, min(ty, 70) + 1))
        values = (rs.randn(b, ty, tx) * rs.choice([0.5, 3.0, 200.0])).astype(np.float32)
        if trial % 3 == 0:
            values = np.round(values)                      # ties
        t_xs = rs.randint(1, tx + 1, size=b).astype(np.int32)
        t_ys = np.array([rs.randint(t_xs[i], ty + 1) for i in range(b)], np.int32)
        want = np.zeros(values.shape, np.int32); work = values.copy()
        ref.maximum_path_c(want, work, t_ys, t_xs)
        got = np.zeros(values.shape, np.int32); mine = values.copy()
        mas_oracle.maximum_path_c(got, mine, t_ys, t_xs)
        assert np.array_equal(got, want), trial
        # the in-place running sums too, bit for bit, inside every item's corner
        for i in range(b):
            assert np.array_equal(mine[i, :t_ys[i], :t_xs[i]], work[i, :t_ys[i], :t_xs[i]]), (trial, i)


def test_wrapper_semantics():
    """monotonic_align/__init__.py:7-21: lengths come from the mask's first column / first row; the result has neg_cent's dtype."""
    rs = np.random.RandomState(5)
    nc = rs.randn(2, 10, 6).astype(np.float64)
    mask = np.zeros((2, 10, 6))
    mask[0, :10, :6] = 1; mask[1, :7, :3] = 1
    p = mas_oracle.maximum_path(nc, mask)
    assert p.dtype == np.float64 and p.shape == nc.shape
    assert p[0].sum() == 10 and p[1].sum() == 7 and p[1, 7:].sum() == 0 and p[1, :, 3:].sum() == 0

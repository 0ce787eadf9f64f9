"""Mints tests/golden/mas_cases.npz from the reference's own compiled `maximum_path_c` (oracle/build_ref_mas.py).

    python oracle/make_golden_mas.py        (needs /root/reference; the fixture is committed, this script documents its origin)

Cases cover what the reference's caller produces (models.py:628-650: neg_cent is a sum of four Gaussian log-likelihood terms, every
item has t_x <= t_y) and the edges of core.pyx's band: t_x == 1, t_x == t_y (the diagonal is forced), t_y == t_x + 1, ragged
batches inside one padded array, ties (equal neighbours: `<` keeps the column), values large enough to lose float32 bits, and
cells outside an item's corner filled with garbage (must not be read into the result).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def cases():
    rs = np.random.RandomState(20261017)
    out = []

    def add(name, values, t_ys, t_xs):
        out.append((name, np.ascontiguousarray(values, np.float32), np.asarray(t_ys, np.int32), np.asarray(t_xs, np.int32)))

    add("single_1x1", rs.randn(1, 1, 1), [1], [1])
    add("tx1", rs.randn(2, 9, 1), [9, 4], [1, 1])
    add("square_forced_diagonal", rs.randn(2, 6, 6), [6, 5], [6, 5])
    add("ty_is_tx_plus_1", rs.randn(3, 8, 7), [8, 7, 4], [7, 6, 3])
    add("ragged_small", rs.randn(4, 12, 7) * 3, [12, 9, 7, 3], [7, 4, 7, 2])
    v = rs.randn(3, 40, 33)
    add("crosses_a_warp", v, [40, 37, 33], [33, 32, 31])
    add("ties", np.round(rs.randn(3, 30, 11)), [30, 22, 11], [11, 9, 11])
    add("all_equal", np.zeros((2, 17, 5)), [17, 9], [5, 5])
    add("large_magnitude", rs.randn(2, 50, 20) * 1e4 - 3e5, [50, 31], [20, 13])
    # shaped like the caller's neg_cent: -0.5 * sum_d (z - m)^2 / s^2 - log terms, 192 channels -> magnitudes of a few hundred
    z = rs.randn(2, 120, 16); m = rs.randn(2, 37, 16)
    nc = -0.5 * ((z[:, :, None, :] - m[:, None, :, :]) ** 2).sum(-1) - 0.5 * 16 * np.log(2 * np.pi)
    add("gaussian_like", nc, [120, 97], [37, 29])
    g = rs.randn(2, 20, 10)
    g[1, 13:, :] = 1e30; g[1, :, 6:] = -1e30          # padding of item 1 holds garbage
    add("garbage_padding", g, [20, 13], [10, 6])
    add("empty_items", rs.randn(3, 6, 4), [6, 0, 3], [4, 2, 0 + 1])
    return out


def main():
    from oracle import build_ref_mas
    ref = build_ref_mas.load()
    if ref is None:
        raise SystemExit("the reference source is not available here")
    blob = {}
    for name, values, t_ys, t_xs in cases():
        paths = np.zeros(values.shape, np.int32)
        work = values.copy()
        ref.maximum_path_c(paths, work, t_ys, t_xs)
        blob[name + ".values"] = values
        blob[name + ".t_ys"] = t_ys
        blob[name + ".t_xs"] = t_xs
        blob[name + ".paths"] = paths.astype(np.int8)
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "mas_cases.npz")
    np.savez_compressed(dst, **blob)
    print(dst, os.path.getsize(dst), "bytes,", len(blob) // 4, "cases")


if __name__ == "__main__":
    main()

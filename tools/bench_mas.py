"""Measurement of the monotonic alignment search: the GPU kernel, device-resident, against the reference's own
Cython/OpenMP `maximum_path_c` (oracle/_ref, else the oracle port) on the box's host cores, on a training-shaped batch.

    python tools/bench_mas.py [--b 64] [--ty 1000] [--tx 300] [--iters 50]

Prints one JSON line: alignments per second both ways, the kernel's HBM roofline (algorithmic bytes = the valid cells of neg_cent read once + the padded path
array written once) against MEASURED_PEAKS.json, and what the reference's call costs end to end
including the two copies it needs (D2H neg_cent, H2D path: monotonic_align/__init__.py:14-21).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--ty", type=int, default=1000)
    ap.add_argument("--tx", type=int, default=300)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    import torch
    from phoonnx_b200 import monotonic_align as ma
    from oracle import build_ref_mas, mas_oracle

    rs = np.random.RandomState(0)
    b, ty, tx = args.b, args.ty, args.tx
    values = (rs.randn(b, ty, tx) * 40).astype(np.float32)
    t_xs = rs.randint(tx // 2, tx + 1, size=b).astype(np.int32)
    t_ys = np.array([rs.randint(max(t_xs[i], ty // 2), ty + 1) for i in range(b)], np.int32)
    dv, dy, dx = torch.from_numpy(values).cuda(), torch.from_numpy(t_ys).cuda(), torch.from_numpy(t_xs).cuda()

    for _ in range(5):
        path, _ = ma.maximum_path_timed(dv, dy, dx)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms, ms_fwd = [], []
    for _ in range(args.iters):
        flush.zero_()                                             # L2 flush between timed iterations
        path, t = ma.maximum_path_timed(dv, dy, dx)
        ms.append(t); ms_fwd.append(float(ma._lib().mas_last_forward_ms()))
    ms = float(np.median(ms)); ms_fwd = float(np.median(ms_fwd))
    got = path.cpu().numpy()

    # the reference's way on this box: device tensor -> host, compiled core (OpenMP over the batch), host -> device
    ref = build_ref_mas.load()
    kind = "reference" if ref is not None else "port"
    def cpu_call():
        host = dv.cpu().numpy().astype(np.float32)
        p = np.zeros(host.shape, np.int32)
        t0 = time.perf_counter()
        if ref is not None:
            ref.maximum_path_c(p, host, t_ys, t_xs)
        else:
            mas_oracle.maximum_path_c(p, host, t_ys, t_xs)
        t1 = time.perf_counter()
        torch.from_numpy(p).cuda().float()
        torch.cuda.synchronize()
        return p, t1 - t0
    cpu_call()
    core_s, e2e_s = [], []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p, core = cpu_call()
        e2e_s.append(time.perf_counter() - t0); core_s.append(core)
    assert np.array_equal(p, got), "GPU path differs from the CPU reference"

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbps_burst", peaks.get("hbm_gbps", 6557.0))) if isinstance(peaks, dict) else 6557.0
    # algorithmic bytes: every VALID cell of neg_cent read once (padding is never touched) + the whole padded path array written once
    bytes_algo = 4.0 * float((t_ys.astype(np.int64) * t_xs).sum()) + 4.0 * b * ty * tx
    line = {
        "metric": "alignments_per_s", "unit": "alignments/s", "value": b / (ms * 1e-3), "ms_per_batch": ms, "ms_forward_and_backtrack": ms_fwd, "ms_output_pass": ms - ms_fwd,
        "config": {"workload": f"maximum_path on b={b} items of up to {ty} frames x {tx} text positions (lengths drawn in the upper half), "
                               "float32 neg_cent resident in HBM, int32 path out", "l2_policy": "256 MiB write between timed calls"},
        "roofline": {"bound": "hbm", "achieved": bytes_algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_algo / (ms * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes": bytes_algo,
                     "note": "the forward pass is a chain of t_y dependent row steps per item (~100 cycles each: shuffle, select, max, add): latency-bound by construction; the output pass alone runs at ~4 TB/s; DRAM traffic = algorithmic bytes (profiles/r02v_ncu_mas.txt)"},
        "cpu_baseline": {"kind": kind, "cores": os.cpu_count(), "value": b / float(np.median(core_s)), "unit": "alignments/s",
                         "ms_core": 1e3 * float(np.median(core_s)), "ms_with_copies": 1e3 * float(np.median(e2e_s)),
                         "sample": "the same batch, 5 calls, median; `ms_with_copies` adds the D2H of neg_cent and the H2D of the path the reference's wrapper performs"},
        "parity": "path identical to the CPU reference on this batch",
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()

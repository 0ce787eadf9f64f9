"""Table of the memory-bound kernels from an `ncu --csv --metrics ...` log of tools/membound_workload.py:
achieved DRAM GB/s (dram bytes / duration) against MEASURED_PEAKS.json hbm_gbs, and DRAM bytes against the ALGORITHMIC bytes of
SURVEY.md 8(d).  usage: python tools/ncu_membound.py <csv> <sizes.json>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, sizes_path):
    sz = json.load(open(sizes_path))
    R, Fr, B, H, C, Fd, hop = sz["ids"], sz["frames"], sz["utterances"], sz["hidden"], sz["inter"], sz["dp_filter"], sz["hop"]
    peak = 6557.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["ID"], row["Kernel Name"].split("(")[0], row["Grid Size"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # us
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per.setdefault(key, {})[m] = v
    # algorithmic bytes per launch: per id R / per frame (chunk) / per sample
    def alg(name, grid):
        nchunks = max(1, -(-Fr // sz["chunk_frames"]))
        fr = Fr / nchunks
        if name.startswith("k_layernorm"):
            width = H if ("<0>" in name or name.endswith("<(int)0>")) else Fd
            return None, f"2-3 x rows x C x 4 B (rows = {R} ids)"
        return None, ""
    agg = collections.OrderedDict()
    for (_id, name, grid), m in per.items():
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1; a["us"] += m.get("gpu__time_duration.sum", 0.0)
        a["rd"] += m.get("dram__bytes_read.sum", 0.0); a["wr"] += m.get("dram__bytes_write.sum", 0.0)
    # algorithmic bytes of ALL launches of a kernel in one pass of the workload
    npass = int(os.environ.get("PASSES", "2"))
    samples = Fr * hop
    algo = {
        "k_expand_sample": Fr * C * 4 + R * 2 * C * 4 + Fr * 12,            # write z_p, read stats once (L2-resident across frames), frame tables
        "k_frame_index": Fr * 12 + R * 4,
        "k_durations": R * 4 + R * 8 + B * 4,
        "k_spline_inverse": 3 * R * (32 + 2 + 1) * 4,                        # three ConvFlows: 32 padded params + x0/x1 in, x1 out
        "k_absmax": samples * 4,
        "k_to_int16": samples * (4 + 2),
        "k_noise_dp": R * 8,
        "k_cf_pre": 3 * R * (Fd * 4 * 2 + 4),                                # read g, write h per ConvFlow
        "k_ea_logw": R * 8,
        "k_embed": R * (H * 4 + 4),
        "k_row_pos": R * 8,
    }
    print(f"sizes: {R} ids, {Fr} frames, {B} utterances; HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json); {npass} passes profiled")
    print("kernel | launches | avg us | DRAM read MB/launch | DRAM write MB/launch | achieved GB/s | frac of HBM peak | DRAM bytes / algorithmic bytes (all launches of one pass)")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        bytes_ = a["rd"] + a["wr"]
        gbs = bytes_ / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
        base = name.split("<")[0]
        al = algo.get(base)
        if base == "k_layernorm":
            # modes: <0> LN (2 x C), <1> depthwise+LN+GELU (2 x C), <2> LN+GELU+residual (3 x C) -- widths differ per launch; quote per byte moved
            ratio = "rows x (2..3) x C x 4 B per launch: see README"
        else:
            ratio = f"{bytes_ / npass / al:.2f}" if al else "-"
        print(f"{name[:40]:40s} | {n:4d} | {a['us'] / n:9.1f} | {a['rd'] / n / 1e6:9.2f} | {a['wr'] / n / 1e6:9.2f} | {gbs:8.0f} | {gbs / peak:5.2f} | {ratio}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

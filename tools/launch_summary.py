"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys


def main(path, tail=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
        seq.append((name, row["Grid Size"], v))
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {len(seq)} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:48]:48s} n={v[0]:5d} total={v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  avg={v[1] / v[0]:8.1f} us")
    if tail:
        for s in seq[-tail:]:
            print(s)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)

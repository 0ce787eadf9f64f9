"""Hottest SASS instructions (by warp-stall samples) of one kernel in an .ncu-rep, with stall reasons.
usage: python tools/ncu_hot.py <rep> <launch-skip> [top]"""
import csv, io, subprocess, sys

rep, skip = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
view = sys.argv[4] if len(sys.argv) > 4 else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view, "--launch-skip", str(skip), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0])
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[1:]))) if (r.get("# Samples") or "").isdigit()]
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"]) for r in rows)
print("total samples", tot, "instructions", len(rows))
agg = {c: sum(int(r[c]) for r in rows) for c in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, r in enumerate(rows):
    r["_i"] = i
for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:top]:
    st = {c[6:]: int(r[c]) for c in stall_cols if int(r[c])}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{r["_i"]:5d} {int(r["# Samples"]):7d} {100*int(r["# Samples"])/tot:5.1f}% {r["Source"].strip()[:90]:90s} {st}')

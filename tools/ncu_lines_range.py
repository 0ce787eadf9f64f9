"""Stall-reason breakdown of the source lines in [lo, hi] of one file for one kernel launch in an .ncu-rep
usage: python tools/ncu_lines_range.py <rep> <launch-skip> <file substring> <lo> <hi>"""
import csv, io, subprocess, sys
from collections import Counter

rep, skip, fsub, lo, hi = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows, hdr, fpath, kern = [], None, "", ""
for rec in csv.reader(io.StringIO(out)):
    if not rec:
        continue
    if rec[0] == "File Path":
        fpath = rec[1].split("/")[-1]; continue
    if rec[0] == "Function Name":
        kern = rec[1]; continue
    if rec[0] == "Line No":
        hdr = rec; continue
    if hdr and len(rec) == len(hdr) and rec[0].isdigit():
        d = dict(zip(hdr, rec)); d["_file"] = fpath; d["_src"] = rec[1]
        rows.append(d)
print(kern)
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"]) for r in rows)
sel = [r for r in rows if fsub in r["_file"] and lo <= int(r["Line No"]) <= hi]
agg = Counter()
for r in sel:
    for c in stall_cols:
        if r[c].isdigit():
            agg[c[6:]] += int(r[c])
n = sum(int(r["# Samples"]) for r in sel)
print(f"lines {lo}-{hi} of {fsub}: {n} samples of {tot} ({100 * n / max(tot, 1):.1f} %)")
for k, v in agg.most_common(12):
    print(f"   {k:24s} {v:8d} {100 * v / max(n, 1):5.1f} %")
for r in sorted(sel, key=lambda r: -int(r["# Samples"]))[:25]:
    st = {c[6:]: int(r[c]) for c in stall_cols if r[c].isdigit() and int(r[c])}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f'{r["_file"]}:{r["Line No"]:>4s} {int(r["# Samples"]):7d}  {r["_src"].strip()[:90]:90s} {st}')

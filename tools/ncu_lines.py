"""Hottest CUDA source lines (by warp-stall samples) of one kernel in an .ncu-rep.
usage: python tools/ncu_lines.py <rep> <launch-skip> [top]"""
import csv, io, subprocess, sys

rep, skip = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
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
print("total samples", tot)
for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:top]:
    st = {c[6:]: int(r[c]) for c in stall_cols if r[c].isdigit() and int(r[c])}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{r["_file"]}:{r["Line No"]:>4s} {int(r["# Samples"]):7d} {100*int(r["# Samples"])/tot:5.1f}%  {r["_src"].strip()[:100]:100s} {st}')

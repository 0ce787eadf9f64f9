"""Key metrics of every launch in an .ncu-rep (raw page) as one table.  usage: python tools/ncu_metrics.py <rep>"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("lts__t_bytes.sum", "l2_bytes"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem"), ("sm__cycles_elapsed.max", "cycles")]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
cols = [(hdr.index(m), n) for m, n in WANT if m in hdr]
print("kernel | grid | " + " | ".join(f"{n} [{units[i]}]" for i, n in cols))
for r in rows[2:]:
    print(r[ki].split("(")[0][:28], "|", r[gi], "|", " | ".join(r[i][:12] for i, _ in cols))

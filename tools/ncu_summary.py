"""Summarise an ncu report (.ncu-rep) as a markdown table: python tools/ncu_summary.py rep.ncu-rep > out.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
cols = [(c, n) for c, n in cols if c in ix]
print("| kernel | " + " | ".join(f"{n} ({units[ix[c]]})" if units[ix[c]] else n for c, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0][-60:]
    vals = []
    for c, _ in cols:
        v = r[ix[c]]
        try:
            v = f"{float(v.replace(',', '')):.4g}"
        except ValueError:
            pass
        vals.append(v)
    print(f"| `{name}` | " + " | ".join(vals) + " |")

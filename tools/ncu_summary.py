"""Summarise an `ncu --set full` report (read here, no GPU needed): python tools/ncu_summary.py <rep> > profiles/x.md
Prints one row per captured launch with the metrics B200_PROFILING.md names."""
import csv
import io
import subprocess
import sys

WANT = [("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print(f"# {rep.split('/')[-1]}\n")
print("| # | kernel | " + " | ".join(f"{n} ({units[col[m]]})" if units[col[m]] else n for m, n in WANT) + " |")
print("|---|---|" + "---|" * len(WANT))
for r in rows[2:]:
    if not r or not r[0].isdigit():
        continue
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = []
    for m, _ in WANT:
        v = r[col[m]]
        try:
            vals.append(f"{float(v):.4g}")
        except ValueError:
            vals.append(v)
    print(f"| {r[0]} | `{name}` | " + " | ".join(vals) + " |")

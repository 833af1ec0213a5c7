"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/ (the reports are too big for git).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" [label1,label2,...] > profiles/xyz.md"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
labels = sys.argv[3].split(",") if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"),
        ("dram__bytes_write.sum", "dram wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu(mufu) %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
cols = [(c, n) for c, n in cols if c in hdr]
print(f"# {title}\n\nSource: `{rep}` (`ncu --set full --clock-control none`), one row per captured launch.\n")
print("| # | " + " | ".join(n + (f" ({units[hdr.index(c)]})" if units[hdr.index(c)] else "") for c, n in cols) + " |")
print("|---|" + "---|" * len(cols))
for i, r in enumerate(data):
    vals = []
    for c, n in cols:
        v = r[hdr.index(c)]
        if n == "kernel":
            v = v.split("(")[0][-60:]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        vals.append(v)
    lab = labels[i] if i < len(labels) else str(i)
    print(f"| {lab} | " + " | ".join(vals) + " |")

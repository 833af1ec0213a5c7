"""Per-layer table from an ncu launch list (csv, one row per metric) of one forward: python tools/launch_table.py file.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, data = rows[0], rows[1:]
iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
L = OrderedDict()
for r in data:
    L.setdefault(r[iID], {"k": r[iK]})[r[iM]] = float(r[iV].replace(",", ""))
ids = list(L)
labs = ("e0g e1c e1g e2c e2g e3c e3g e4c e4g e5c e5g e6c e6g e7c e7g t1 m0in m0xp m0dt m0out m1in m1xp m1dt m1out m2in m2xp m2dt "
        "m2out t2 d0g d0c d1g d1c d2g d2c d3g d3c d4g d4c d5g d5c d6g d6c d7g").split()
g = [i for i in ids if "gemm_tc" in L[i]["k"]]
last = g[-44:]
tot = 0.0
print("| layer | kernel | time (us) | dram rd (MB) | dram wr (MB) | tensor pipe active (%) |\n|---|---|---|---|---|---|")
for lab, i in zip(labs, last):
    d = L[i]
    t = d["gpu__time_duration.sum"] / 1e3
    kn = d["k"].split("gemm_tc_kernel")[1].split("(CUt")[0].replace("(int)", "").replace("(bool)", "")
    print(f"| {lab} | gemm_tc{kn} | {t:.1f} | {d['dram__bytes_read.sum'] / 1e6:.1f} | {d['dram__bytes_write.sum'] / 1e6:.1f} | "
          f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} |")
    tot += t
print(f"\nGEMM total {tot / 1e3:.2f} ms (cold-cache, serialised launches)\n")
others = {}
for i in ids[ids.index(last[0]) - 2:]:
    if "gemm_tc" not in L[i]["k"]:
        k = L[i]["k"].split("(")[0][-40:]
        others[k] = others.get(k, 0) + L[i]["gpu__time_duration.sum"] / 1e3
print("other kernels of the same forward (us):", {k: round(v, 1) for k, v in others.items()})

"""Per-launch table of ONE steady-state streaming call from an ncu launch list (csv, one row per metric):
    python tools/stream_launch_table.py gpurun_out/launches_stream_h1.csv > profiles/r02_launches_stream_h1.md
The list is taken with `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...
--clock-control none -k regex:_kernel` around `bench.py --mode stream --model e6 --streams-total 4096 --hops 1`; the last call of the
run (from its input staging copy to its stream_shift launch) is printed with the layer each launch belongs to."""
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
start = ids.index([i for i in ids if "conv_in" in L[i]["k"]][-1]) - 2      # input staging copy, stream_std, conv_in, ...
call = ids[start:]
enc = ["e0g"] + [f"e{i}{s}" for i in range(1, 6) for s in "cg"]
mam = [x for l in range(3) for x in (f"m{l} ln", f"m{l} in_proj", f"m{l} dwconv+state", f"m{l} x_proj", f"m{l} dt_proj", f"m{l} state update", f"m{l} out_proj")]
dec = [f"d{j}{s}" for j in range(5) for s in "gc"] + ["d5g"]
names = ["input staging (torch copy_)", "running std", "conv_in"] + enc + ["t1"] + mam + ["norm_f", "t2"] + dec + ["convt_out", "FIFO maintenance"]
print("| # | layer | kernel | time (us) | dram rd (MB) | dram wr (MB) | tensor pipe active (%) |\n|---|---|---|---|---|---|---|")
tot = rd = wr = 0.0
kinds = {}
for n, i in enumerate(call):
    d = L[i]
    t = d["gpu__time_duration.sum"] / 1e3
    k = d["k"].split("(")[0].replace("void ", "").replace("cum::", "").replace("at::", "torch ")
    if "gemm_tc_kernel" in d["k"]:
        k = "gemm_tc" + d["k"].split("gemm_tc_kernel")[1].split("(CUt")[0].replace("(int)", "").replace("(bool)", "")
    lab = names[n] if n < len(names) else "?"
    print(f"| {n} | {lab} | {k[:48]} | {t:.1f} | {d['dram__bytes_read.sum'] / 1e6:.1f} | {d['dram__bytes_write.sum'] / 1e6:.1f} | "
          f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} |")
    tot += t; rd += d["dram__bytes_read.sum"]; wr += d["dram__bytes_write.sum"]
    kind = "gemm" if "gemm_tc" in d["k"] else k.split("<")[0]
    kinds[kind] = kinds.get(kind, 0.0) + t
print(f"\n{len(call)} launches, {tot / 1e3:.3f} ms serialised (ncu: un-capped clocks, one kernel at a time), DRAM {rd / 1e9:.2f} GB read + {wr / 1e9:.2f} GB written.\n")
print("| kind | us per call | share |\n|---|---|---|")
for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]):
    print(f"| {k} | {v:.0f} | {100 * v / tot:.1f} % |")

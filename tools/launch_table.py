"""Per-layer table from an ncu launch list (csv, one row per metric) of one forward:
    python tools/launch_table.py file.csv [traffic.json]
Round 2: the first / last U-Net blocks are single fused kernels (fused_end_kernel<0> / <1>), so one forward has 42 tap-GEMM launches
+ 2 fused blocks.  With a second argument the DRAM bytes of those 44 tensor-core launches are written as roofline.traffic input."""
import csv
import json
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, data = rows[0], rows[1:]
iK, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
L = OrderedDict()
for r in data:
    L.setdefault(r[iID], {"k": r[iK]})[r[iM]] = float(r[iV].replace(",", ""))
ids = list(L)
labs = ("e1c e1g e2c e2g e3c e3g e4c e4g e5c e5g e6c e6g e7c e7g t1 m0in m0xp m0dt m0out m1in m1xp m1dt m1out m2in m2xp m2dt "
        "m2out t2 d0g d0c d1g d1c d2g d2c d3g d3c d4g d4c d5g d5c d6g d6c").split()
fused = [i for i in ids if "fused_end_kernel" in L[i]["k"]]
g = [i for i in ids if "gemm_tc" in L[i]["k"]]
if fused:
    last = g[-42:]
    first_id = fused[-2]
    tc = [fused[-2]] + last + [fused[-1]]
    names = ["e0 (fused block)"] + labs + ["d7 (fused block)"]
else:
    labs = ["e0g"] + labs + ["d7g"]
    last = g[-44:]
    first_id = last[0]
    tc, names = last, labs
tot = dr = dw = 0.0
print("| layer | kernel | time (us) | dram rd (MB) | dram wr (MB) | tensor pipe active (%) |\n|---|---|---|---|---|---|")
for lab, i in zip(names, tc):
    d = L[i]
    t = d["gpu__time_duration.sum"] / 1e3
    if "gemm_tc" in d["k"]:
        kn = "gemm_tc" + d["k"].split("gemm_tc_kernel")[1].split("(CUt")[0].replace("(int)", "").replace("(bool)", "")
    else:
        kn = "fused_end" + d["k"].split("fused_end_kernel")[1].split("(CUt")[0].replace("(int)", "")
    print(f"| {lab} | {kn} | {t:.1f} | {d['dram__bytes_read.sum'] / 1e6:.1f} | {d['dram__bytes_write.sum'] / 1e6:.1f} | "
          f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} |")
    tot += t
    dr += d["dram__bytes_read.sum"]
    dw += d["dram__bytes_write.sum"]
print(f"\nTensor-core launches: {len(tc)}; total {tot / 1e3:.2f} ms (cold-cache, serialised launches); DRAM read {dr / 1e9:.2f} GB + "
      f"write {dw / 1e9:.2f} GB = {(dr + dw) / 1e9:.2f} GB per step\n")
others = {}
for i in ids[ids.index(first_id) - 1:]:
    if "gemm_tc" not in L[i]["k"] and "fused_end" not in L[i]["k"]:
        k = L[i]["k"].split("(")[0][-40:]
        others[k] = others.get(k, 0) + L[i]["gpu__time_duration.sum"] / 1e3
print("other kernels of the same forward (us):", {k: round(v, 1) for k, v in others.items()})
if len(sys.argv) > 2:
    json.dump({"source": sys.argv[1], "dram_bytes_per_step_gemm_launches": dr + dw, "ncu_time_ms": tot / 1e3, "batch": 64,
               "launches_per_step": len(tc),
               "note": "bytes per STEP over the 44 tensor-core launches of one forward (42 tap-GEMMs + the two fused end blocks; ncu "
                       "dram__bytes_read + dram__bytes_write, profiles/r02_launches_f16x3.md).  SURVEY 8(d) fused-minimum activation "
                       "traffic of the conv stacks is 0.44 GB per clip = 28.2 GB per step: every block fused end to end.  Blocks with "
                       "H >= 256 cannot chain their two GEMMs on one SM (accumulators of both exceed the 512 TMEM columns, DESIGN.md 4) "
                       "and are tensor-bound, so their H-wide intermediate still makes one HBM round trip."},
              open(sys.argv[2], "w"))

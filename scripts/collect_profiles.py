#!/usr/bin/env python
"""Copies one gpu_round2.sh run from gpurun_out/ (scratch) into profiles/r1/ (tracked) and prints the
numbers the docs quote.   python scripts/collect_profiles.py <tag>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
G, P = "gpurun_out", "profiles/r1"
os.makedirs(P, exist_ok=True)
for src, dst in ((f"{tag}_bench.json", "bench_n1.json"), (f"{tag}_bench_ref.json", "bench_reference_arm_n1.json"), (f"{tag}_box.txt", "box.txt"),
                 (f"{tag}_kernels.txt", "kernel_timings_100m.txt"), (f"{tag}_pytest.log", "pytest_gpu.log")):
    shutil.copy(os.path.join(G, src), os.path.join(P, dst))
for src, dst in ((f"{tag}_ncu_add", "ncu_full_binary_spans_kernel.csv"), (f"{tag}_ncu_kernels", "ncu_full_other_kernels.csv")):
    if os.path.exists(os.path.join(G, src + ".csv")):
        shutil.copy(os.path.join(G, src + ".csv"), os.path.join(P, dst))
    else:  # older runs brought the raw report back
        subprocess.check_call([sys.executable, "scripts/ncu_summary.py", f"{G}/{src.replace('_ncu_', '_prof_')}.ncu-rep", f"{P}/{dst}"])
rows = list(csv.reader(open(f"{G}/{tag}_launches.csv", errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv, mu, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
    a = agg.setdefault(r[kn].split("(")[0], [0, 0.0, r[gs]])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(f"{P}/launches_bench_n1.csv", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu\n"
            "# cold-cache, serialised launches: compare SHARES, not absolutes.  binary_spans_kernel launches with grid (1,1,1)..(123,1,1)\n"
            "# come from bench.py's 'one launch per span' context leg, not from the timed step.\n")
    f.write("kernel,launches,total_us,mean_us,share_of_gpu_time,last_grid\n")
    for k, (c, t, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{c},{t:.1f},{t / c:.2f},{t / tot:.4f},"{g}"\n')
d = json.load(open(f"{P}/bench_n1.json"))
print("ADD chunked ms", d["ms_per_step"], "Grows/s", d["value"] / 1e9, "frac", d["roofline"]["frac"], "contig ms", d["config"]["contiguous_ms_per_step"],
      "per-span ms", d["config"]["per_span_launch_ms_per_step"], "traffic", d["roofline"].get("traffic"))
print("e2e ms", d["e2e"]["ms_per_step"], "Grows/s", d["e2e"]["value"] / 1e9, "link", d["e2e"]["link_gbs"])
print("cpu 1t", d["cpu_baseline"]["value"] / 1e9, "all", d["cpu_baseline"]["all_cores"]["value"] / 1e9)
print("clocks", d["clocks"])
for k, v in d["others"].items():
    print(k, round(v["ms"] * 1e3, 1), "us", round(v["rows_per_s"] / 1e9, 1), "Grows/s", round(v.get("gbs_per_gpu", 0)), "GB/s", round(v.get("frac", 0), 3))
r = json.load(open(f"{P}/bench_reference_arm_n1.json"))
print("ref arm", r["value"] / 1e9, r["cpu_baseline"]["all_cores"]["value"] / 1e9)
for f in (f"{P}/ncu_full_other_kernels.csv", f"{P}/ncu_full_binary_spans_kernel.csv"):
    rr = list(csv.reader(open(f)))
    for x in rr[1:]:
        print(x[0][:60], "| t", x[3], "rd", x[4], "wr", x[5], "dram%", x[6])

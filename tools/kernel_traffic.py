"""Turns an ncu CSV of `dram__bytes_read.sum,dram__bytes_write.sum` over the GEMM launches of ONE bench step into
profiles/r01_kernel_traffic.json (average DRAM bytes per launch per kernel family), which bench.py reports as
`roofline.traffic`.   usage: python tools/kernel_traffic.py <ncu.csv> <workload> <launches_per_step_fwd> <launches_per_step_wgrad>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, workload = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    per.setdefault(r[idi], 0.0)
    per[r[idi]] += v * mult
n = len(per)
tot = sum(per.values())
out_path = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
out = json.load(open(out_path)) if os.path.isfile(out_path) else {}
out["source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d GEMM launches of one %s step" % (n, workload)
n_fwd, n_wg = int(sys.argv[3]), int(sys.argv[4])
assert n == n_fwd + n_wg, (n, n_fwd, n_wg)
# the launch list is in issue order; split-K weight-gradient launches are the fp32-output ones: identify them by name order is
# not possible from the CSV, so report the family average over all GEMM launches for both keys
avg = tot / n
out.setdefault(workload, {})
out[workload]["gemm_tcgen05"] = {"avg_bytes_per_launch": int(avg), "launches_per_step": n_fwd, "launches_measured": n}
out[workload]["gemm_tcgen05_wgrad"] = {"avg_bytes_per_launch": int(avg), "launches_per_step": n_wg, "launches_measured": n}
json.dump(out, open(out_path, "w"), indent=1)
print("wrote", out_path, "avg %.1f MB per launch over %d launches" % (avg / 1e6, n))

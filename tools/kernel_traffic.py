"""Per-kernel-family DRAM traffic and device time of ONE training step, from an ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch) and the issue-ordered C-ABI call list tools/traffic_step.py wrote in the same run.
Writes profiles/r02_kernel_traffic.json, which bench.py reports as `roofline.traffic` (per launch of the dominant family).
    python tools/kernel_traffic.py <step_metrics.csv> <step_kinds.json>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
kinds = json.load(open(sys.argv[2]))
hdr = rows[0]
ki, mi, vi, ui, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for r in rows[1:]:
    e = launches.setdefault(int(r[idi]), {"name": r[ki].split("(")[0], "bytes": 0.0, "us": 0.0})
    v = float(r[vi].replace(",", ""))
    if r[mi].startswith("dram__bytes"):
        e["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    else:
        e["us"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
# the C-ABI calls whose kernels carry a family-specific name map by name; the two GEMM families share one kernel and are told apart
# by issue order: the k-th gemm_tcgen05_kernel launch belongs to the k-th tensor-core GEMM call
gemm_calls = [c for c in kinds["calls"] if c["kind"] in ("gemm_tcgen05", "gemm_tcgen05_wgrad")]
NAME = [("attn_fwd", "attn_fwd"), ("attn_bwd", "attn_bwd"), ("ln_fwd", "layernorm_fwd"), ("ln_bwd", "layernorm_bwd"), ("bn_bwd", "bn_bwd"),
        ("bn_apply", "bn_apply"), ("colstats", "bn_colstats"), ("colsum", "bias_colsum"), ("adamw", "adamw"), ("ctc_", "ctc"),
        ("log_softmax", "ctc"), ("ce_sumexp", "ce_sumexp"), ("ce_finalize", "ce_sumexp"), ("gelu", "gelu")]
fam = collections.defaultdict(lambda: {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
gi = 0
for e in launches.values():
    n = e["name"]
    if "gemm_tcgen05_kernel" in n:
        k = gemm_calls[gi]["kind"] if gi < len(gemm_calls) else "gemm_tcgen05"
        gi += 1
    else:
        k = next((f for pat, f in NAME if pat in n), "other")
    fam[k]["launches"] += 1
    fam[k]["dram_bytes"] += e["bytes"]
    fam[k]["us"] += e["us"]
alg = collections.defaultdict(lambda: [0.0, 0])
for c in kinds["calls"]:
    alg[c["kind"]][0] += c["bytes"]
    alg[c["kind"]][1] += 1
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the %d kernel "
                 "launches of one %s training step (tools/traffic_step.py); per family: DRAM bytes summed over the step's launches"
                 % (len(launches), kinds["workload"]), kinds["workload"]: {}}
tot_us = sum(f["us"] for f in fam.values())
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    ent = {"launches_per_step": f["launches"], "dram_bytes_per_step": int(f["dram_bytes"]), "avg_bytes_per_launch": int(f["dram_bytes"] / f["launches"]),
           "serialized_us": round(f["us"], 1), "share_of_step": round(f["us"] / tot_us, 4)}
    if alg[k][1]:
        ent["algorithmic_bytes_per_step"] = int(alg[k][0])
        if alg[k][0] > 0:
            ent["traffic_over_algorithmic"] = round(f["dram_bytes"] / alg[k][0], 3)
    out[kinds["workload"]][k] = ent
assert gi == len(gemm_calls), "GEMM launch count (%d) does not match the recorded GEMM calls (%d)" % (gi, len(gemm_calls))
path = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
for k, e in out[kinds["workload"]].items():
    print("%-22s %4d launches %9.1f MB/launch  share %5.1f%%  traffic/algorithmic %s" % (
        k, e["launches_per_step"], e["avg_bytes_per_launch"] / 1e6, 100 * e["share_of_step"], e.get("traffic_over_algorithmic")))

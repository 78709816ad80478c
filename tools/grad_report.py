"""Debug (not a pytest file): per-tensor gradient error report of the fp32 engine step against a golden fixture."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import load_golden, golden_inputs
from test_engine_gpu import make_engine, run_step

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_enc_ctc"
dtype = torch.float32 if (len(sys.argv) < 3 or sys.argv[2] == "fp32") else torch.bfloat16
z, meta = load_golden(name)
cfg, sd, batch = golden_inputs(meta)
eng = make_engine(cfg, sd, dtype)
out_enc, out_dec, loss, loss_dec, loss_enc, G, ctx = run_step(eng, cfg, batch)
rows = []
tot_num = tot_den = 0.0
for n in meta["grad_names"]:
    g = G[n].double().reshape(-1)[torch.from_numpy(z["gidx/" + n])]
    ref = torch.from_numpy(z["gval/" + n]).double()
    truth = torch.from_numpy(z["gtruth/" + n]).double()
    num = float(((g - ref) ** 2).sum()); den = float((ref ** 2).sum())
    rows.append((num, den, n, float(((ref - truth) ** 2).sum())))
    tot_num += num; tot_den += den
print("global L2 rel err %.3e" % (tot_num / tot_den) ** 0.5)
for num, den, n, rt in sorted(rows, reverse=True)[:15]:
    print("%-60s contrib %.3e  own rel %.3e  (reference-vs-float64 contrib %.3e)" % (n, num / tot_den, (num / max(den, 1e-300)) ** 0.5, rt / tot_den))

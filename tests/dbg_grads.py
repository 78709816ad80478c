"""Diagnostic: per-parameter-group gradient errors of the CUDA path vs reference golden and vs fp64 truth."""
import sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from helpers import load_golden, golden_inputs, rel_err, check_grads_against_golden
from test_engine_gpu import make_engine, run_step
for name in ["short_hybrid", "ragged_hybrid", "cfg1_enc_ctc"]:
    for dtype in (torch.float32, torch.bfloat16):
        z, meta = load_golden(name)
        cfg, sd, batch = golden_inputs(meta)
        eng = make_engine(cfg, sd, dtype)
        out_enc, out_dec, loss, loss_dec, loss_enc, G, ctx = run_step(eng, cfg, batch)
        rep = []
        try:
            check_grads_against_golden(z, meta, G, 1e9, report=rep)
        except AssertionError as e:
            print("ASSERT", e)
        grp = collections.defaultdict(lambda: [0, 0, 0])
        for n, a, b, c in rep:
            parts = n.split(".")
            key = ".".join(parts[:2]) if parts[0].startswith("conv") else parts[0] + ("." + parts[-2] + "." + parts[-1] if parts[0].startswith("transformer") else "")
            g = grp[key]
            g[0] = max(g[0], a); g[1] = max(g[1], b); g[2] = max(g[2], c)
        print("==== %s %s  loss %.5f (ref %.5f)  out_enc err %.2e" % (name, dtype, loss, float(z["loss"]),
              max(rel_err(out_enc[b, :l], z["out_enc"][b, :l], floor=float(abs(z["out_enc"]).max())) for b, l in enumerate(batch["lengths"]))))
        for k in sorted(grp):
            print("   %-48s vs_ref %.2e  vs_truth %.2e  (ref_vs_truth %.2e)" % (k, *grp[k]))
# ---- bf16: tensor-core engine vs CUDA-core engine (same bf16 operands) per group
for name in ["short_hybrid"]:
    z, meta = load_golden(name)
    cfg, sd, batch = golden_inputs(meta)
    res = []
    for simt in (False, True):
        eng = make_engine(cfg, sd, torch.bfloat16)
        eng.force_simt = simt
        res.append(run_step(eng, cfg, batch))
    print("==== bf16 tcgen05 vs simt grads")
    gm = max(float(v.abs().max()) for v in res[1][5].values())
    for n in sorted(res[0][5]):
        e = rel_err(res[0][5][n], res[1][5][n], floor=1e-4 * gm)
        if e > 1e-2: print("   %-60s %.2e" % (n, e))

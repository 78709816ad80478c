"""Per-tensor relative-L2 gradient report of the engine step against the golden fixtures of the UNMODIFIED reference
(tests/golden/*.npz): every trainable tensor, fp32 and bf16 mode.  Writes the table tests/helpers.check_grads_l2 asserts on.
    python tools/grad_l2_report.py [out.txt] [case ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from helpers import load_golden, golden_inputs, grad_l2_table, l2_verdict, GOLDEN_CASES   # noqa: E402
from test_engine_gpu import make_engine, run_step                            # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else None
cases = sys.argv[2:] or GOLDEN_CASES
lines = []
for name in cases:
    z, meta = load_golden(name)
    cfg, sd, batch = golden_inputs(meta)
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 2e-2)):
        eng = make_engine(cfg, sd, dtype)
        G = run_step(eng, cfg, batch)[5]
        rows = grad_l2_table(z, meta, G)
        bf16 = dtype == torch.bfloat16
        verdicts = [l2_verdict(r, tol, bf16) for r in rows]
        lines.append("== %s %s, tol %.0e: %d tensors -- %d within tol of the reference, %d by the float64 clause, %d by the "
                     "autocast-bf16 clause, %d FAIL" % (name, str(dtype).replace("torch.", ""), tol, len(rows), verdicts.count("ref"),
                                                        verdicts.count("f64"), verdicts.count("bf16"), verdicts.count(None)))
        lines.append("%-64s %10s %10s %10s %10s" % ("tensor (relative L2 over the fixture's entries)", "vs ref", "vs f64", "ref vs f64", "ref-bf16"))
        for r, v in sorted(zip(rows, verdicts), key=lambda rv: -rv[0][1]):
            lines.append("%-64s %10.3e %10.3e %10.3e %10.3e  %s" % (r + ({"ref": "", "f64": "(f64 clause)", "bf16": "(bf16 clause)", None: "FAIL"}[v],)))
txt = "\n".join(lines)
print(txt)
if out_path:
    open(out_path, "w").write(txt + "\n")

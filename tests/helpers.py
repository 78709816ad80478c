"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np
import torch

import sst_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def golden_inputs(meta):
    cfg = meta["cfg"]
    sd = O.synthetic_state_dict(cfg, meta["wseed"])
    batch = O.synthetic_batch(**meta["batch"])
    return cfg, sd, batch


def rel_err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor)  -- the per-tensor relative error used throughout."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    denom = max(float(b.abs().max()), floor, 1e-30)
    return float((a - b).abs().max()) / denom


def check_grads_against_golden(z, meta, grads, tol, label=""):
    """grads: dict name -> tensor.  Tensors whose true gradient is rounding noise (conv biases in front
    of a training-mode BatchNorm: SURVEY.md Q7) are held to an absolute bound relative to the global
    gradient scale instead."""
    gmax = max(float(z["gstat/" + n][0]) for n in meta["grad_names"])
    worst = (0.0, None)
    for n in meta["grad_names"]:
        assert n in grads, "missing gradient for %s" % n
        g = grads[n].detach().double().cpu().reshape(-1).numpy()
        idx = z["gidx/" + n]
        ref = z["gval/" + n].astype(np.float64)
        ref_max = float(z["gstat/" + n][0])
        err = float(np.abs(g[idx] - ref).max())
        floor = 1e-4 * gmax
        e = err / max(ref_max, floor)
        if e > worst[0]:
            worst = (e, n)
        assert e <= tol, "%s grad %s: rel err %.3e > %.1e (abs %.3e, ref max %.3e)" % (label, n, e, tol, err, ref_max)
        nrm = float(np.sqrt((g ** 2).sum()))
        ref_nrm = float(z["gstat/" + n][1])
        assert abs(nrm - ref_nrm) <= 10 * tol * max(ref_nrm, floor * np.sqrt(g.size)), \
            "%s grad-norm %s: %.6e vs %.6e" % (label, n, nrm, ref_nrm)
    for n in meta["none_grad"]:
        assert n not in grads or grads[n] is None, "%s must not receive a gradient (Q2/Q14)" % n
    return worst

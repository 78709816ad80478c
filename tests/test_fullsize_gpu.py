"""BASELINE-sized checks on a B200 beyond the fixtures of tests/test_engine_gpu.py (full_2p1 / full_6p6 pin the benchmarked
utterance length and model against the unmodified reference): a RAGGED batch at full length -- padded query rows / keys,
skipped key tiles, decollate + pad 42 -- against the CPU oracle run live (2 + 1 layers, ~15 s of CPU time)."""
import pytest
import torch

import sst_oracle as O
from helpers import l2_rows, assert_l2_rows

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ragged_full_length_batch_matches_the_oracle(dtype):
    from sst_b200.synthetic import make_batch
    from test_engine_gpu import make_engine, run_step, oracle_autocast_bf16_grads, _valid_frames_err
    cfg = O.make_cfg(n_enc=2, n_dec=1, rel_dist=100, alpha=0.2)
    sd = O.synthetic_state_dict(cfg, 3)
    lengths = [1000, 640, 333, 128, 97, 1002]
    batch = make_batch(lengths=lengths, tgt_min=40, tgt_max=60, seed=5)
    res, grads, _ = O.loss_and_grads({k: v.clone() for k, v in sd.items()}, cfg, batch, True, 0)
    eng = make_engine(cfg, sd, dtype)
    out_enc, out_dec, loss, loss_dec, loss_enc, G, ctx = run_step(eng, cfg, batch)
    bf16 = dtype == torch.bfloat16
    tol = 2e-2 if bf16 else 1e-4
    assert _valid_frames_err(out_enc, res["out_enc"], lengths) < tol
    assert abs(loss_enc - float(res["loss_enc"])) < tol * abs(float(res["loss_enc"]))
    assert abs(loss_dec - float(res["loss_dec"])) < tol * abs(float(res["loss_dec"]))
    names = sorted(grads)
    flat = lambda d: {n: d[n].detach().double().cpu().reshape(-1).numpy() for n in names}      # noqa: E731
    gbf = oracle_autocast_bf16_grads(sd, cfg, batch) if bf16 else None
    truth = None
    if not bf16:                                       # float64 yardstick (the reference's own fp32 distance from exact arithmetic)
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        b64 = dict(batch)
        b64["raw_emg"] = [x.double() for x in batch["raw_emg"]]
        truth = flat(O.loss_and_grads(sd64, cfg, b64, True, 0)[1])
    gmax = max(float(g.abs().max()) for g in grads.values())
    rows = l2_rows(names, flat(G), flat(grads), truth, flat(gbf) if gbf is not None else None, gmax)
    assert_l2_rows(rows, tol, bf16, str(dtype))

"""Live pin of the oracle against the UNMODIFIED reference; only runs where /root/reference exists."""
import pytest
import torch
import torch.nn.functional as F

import ref_harness
import sst_oracle as O
from helpers import rel_err

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not present")]


@pytest.mark.parametrize("dropout", [0.0])
def test_hybrid_step_live(dropout):
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=60, alpha=0.3, dropout=dropout)
    sd = O.synthetic_state_dict(cfg, 11)
    model = ref_harness.build_model(cfg, sd)
    arch, tr, LS, du, FLAGS = ref_harness.load(ref_harness.cfg_to_argv(cfg))
    batch = O.synthetic_batch(seed=8, ragged=[120, 80], tgt_lens=[10, 6])
    model.train()
    X = du.combine_fixed_length(batch["raw_emg"], 1600).clone()
    tgt_in, tgt_out, ctc_tgt, ctc_lens = O.make_targets(batch)
    # Q13: the reference's in-place overlapping shift (architecture.py:107) raises on torch-2.11 CPU,
    # so the live pin runs with r forced to 0; the shift itself is pinned by test_shift_semantics.
    real = arch.random.randrange
    arch.random.randrange = lambda n: 0
    r = 0
    try:
        out_enc, out_dec = model(batch["lengths"], "cpu", x_raw=X, y=tgt_in)
    finally:
        arch.random.randrange = real
    lp = F.log_softmax(out_enc, 2).transpose(1, 0)
    loss_enc = F.ctc_loss(lp, ctc_tgt, batch["lengths"], ctc_lens, blank=43)
    loss_dec = LS.LabelSmoothingLoss(epsilon=0.1, num_classes=43)(out_dec.permute(0, 2, 1), tgt_out)
    loss = 0.7 * loss_dec + 0.3 * loss_enc
    loss.backward()
    res, grads, _ = O.loss_and_grads(sd, cfg, batch, True, shift_r=r)
    assert rel_err(res["out_enc"], out_enc.detach()) < 2e-5
    assert rel_err(res["out_dec"], out_dec.detach()) < 2e-5
    assert abs(float(res["loss"]) - float(loss)) < 1e-5 * abs(float(loss))
    gmax = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    for n, p in model.named_parameters():
        if p.grad is None:
            assert n not in grads
            continue
        assert rel_err(grads[n], p.grad, floor=1e-4 * gmax) < 2e-4, n


def test_shift_semantics():
    """architecture.py:104-108: shift left by r within each 1600-chunk, zero fill."""
    x = torch.arange(2 * 10 * 3.0).view(2, 10, 3)
    y = O.shift_left_(x.clone(), 3)
    assert torch.equal(y[:, :7], x[:, 3:]) and torch.all(y[:, 7:] == 0)
    assert torch.equal(O.shift_left_(x.clone(), 0), x)

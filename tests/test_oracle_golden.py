"""The oracle (CPU restatement) against the fixtures generated from the UNMODIFIED reference
(tests/golden/*.npz, oracle/make_golden.py).  Runs on any box; no GPU, no /root/reference."""
import numpy as np
import pytest
import torch

import sst_oracle as O
from helpers import load_golden, golden_inputs, rel_err, check_grads_l2

CASES = ["cfg1_enc_ctc", "ragged_hybrid", "short_hybrid"]


@pytest.mark.parametrize("name", CASES + ["full_2p1", "full_6p6"])
def test_forward_backward_matches_reference(name):
    z, meta = load_golden(name)
    cfg, sd, batch = golden_inputs(meta)
    res, grads, stats = O.loss_and_grads(sd, cfg, batch, training=True, shift_r=0)
    assert rel_err(res["out_enc"], z["out_enc"]) < 2e-5
    assert abs(float(res["loss_enc"]) - float(z["loss_enc"])) < 1e-5 * abs(float(z["loss_enc"]))
    if meta["mode"] == "hybrid":
        assert rel_err(res["out_dec"], z["out_dec"]) < 2e-5
        assert abs(float(res["loss_dec"]) - float(z["loss_dec"])) < 1e-5 * abs(float(z["loss_dec"]))
    assert abs(float(res["loss"]) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    assert sorted(grads) == sorted(meta["grad_names"])
    check_grads_l2(z, meta, grads, 1e-4, False, "oracle")
    for k, v in stats.items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == 1
        else:
            assert rel_err(v, z["bn/" + k]) < 1e-5


@pytest.mark.parametrize("name", CASES)
def test_adamw_step_matches_reference(name):
    z, meta = load_golden(name)
    cfg, sd, batch = golden_inputs(meta)
    _, grads, _ = O.loss_and_grads(sd, cfg, batch, training=True, shift_r=0)
    lr = O.lr_schedule(0)
    assert lr == pytest.approx(3e-4 / 1500)
    for n in meta["grad_names"]:
        p = sd[n].clone()
        O.adamw_step(p, grads[n], torch.zeros_like(p), torch.zeros_like(p), 1, lr)
        got = p.reshape(-1)[torch.from_numpy(z["gidx/" + n])]
        # AdamW's first step moves every weight by ~lr: compare the *update*, not the weight
        ref_upd = torch.from_numpy(z["pnew/" + n]) - sd[n].reshape(-1)[torch.from_numpy(z["gidx/" + n])]
        got_upd = got - sd[n].reshape(-1)[torch.from_numpy(z["gidx/" + n])]
        if "conv" in n and n.endswith("bias") or "residual_path.bias" in n:
            continue        # sign(noise) -- Q7, true gradient is zero
        assert float((got_upd - ref_upd).abs().max()) <= 2e-8 + 0.02 * float(ref_upd.abs().max()), n


@pytest.mark.parametrize("name", ["ragged_hybrid", "short_hybrid"])
def test_greedy_decode_bit_exact(name):
    z, meta = load_golden(name)
    cfg, sd, batch = golden_inputs(meta)
    import make_golden
    sd = make_golden.jitter_running_stats(sd, meta["wseed"])
    X = O.combine_fixed_length(batch["raw_emg"])
    max_len = z["greedy_ids"].shape[1]
    seqs, ids = O.greedy_decode(sd, cfg, X, batch["lengths"], max_len)
    assert np.array_equal(ids.numpy(), z["greedy_ids"])
    assert meta["greedy_min_margin"] > 1e-4


@pytest.mark.parametrize("name", CASES)
def test_ctc_best_path_decode(name):
    z, meta = load_golden(name)
    _, _, batch = golden_inputs(meta)
    dec = O.ctc_greedy_collapse(torch.from_numpy(z["out_enc"]), batch["lengths"])
    assert dec == meta["ctc_decode"]


def test_relpos_closed_form_equals_as_written():
    """SURVEY.md Q3 self-check: in-band q.E, exactly -1e8 out of band, for L>R and L<R."""
    g = torch.Generator().manual_seed(3)
    for L, R in ((130, 50), (40, 50), (50, 50)):
        H, dh, B = 2, 16, 2
        q = torch.randn(B, H, L, dh, generator=g)
        emb = torch.randn(H, 2 * R - 1, dh, 1, generator=g)
        a = O.relpos_logits_as_written(q.permute(2, 0, 1, 3).reshape(L, B * H, dh), emb, R, H).view(B, H, L, L)
        c = O.relpos_logits_closed_form(q, emb, R)
        i = torch.arange(L)[:, None]
        j = torch.arange(L)[None, :]
        inband = ((j - i).abs() < R)[None, None].expand_as(a)
        assert torch.all(a[~inband] == -1e8) and torch.all(c[~inband] == -1e8)
        assert float((a[inband] - c[inband]).abs().max()) < 1e-5


def test_ctc_alpha_beta_matches_torch():
    g = torch.Generator().manual_seed(7)
    for T, S in ((12, 3), (9, 4), (5, 0), (7, 3)):
        logits = torch.randn(T, 44, generator=g, dtype=torch.float64, requires_grad=True)
        tgt = torch.randint(0, 40, (S,), generator=g)
        if S >= 2:
            tgt[1] = tgt[0]                       # repeated label: the s-2 skip must be disabled
        lp = torch.log_softmax(logits, 1)
        if T < 2 * S - 1 + 1 and S > 0 and T < S + int((tgt[1:] == tgt[:-1]).sum()):
            continue
        loss = torch.nn.functional.ctc_loss(lp[:, None, :], tgt[None], [T], [S], blank=43, reduction="sum")
        loss.backward()
        nll, grad = O.ctc_alpha_beta(logits.detach().numpy(), tgt.numpy())
        assert abs(nll - float(loss)) < 1e-9 * max(1.0, abs(float(loss)))
        assert np.abs(grad - logits.grad.numpy()).max() < 1e-9
        assert np.abs(grad.sum(1)).max() < 1e-9       # d/dlogits sums to zero over classes


def test_combine_and_decollate_roundtrip():
    batch = O.synthetic_batch(seed=5, ragged=[30, 41, 7], tgt_lens=[3, 4, 2])
    X = O.combine_fixed_length(batch["raw_emg"])
    assert X.shape == (1, 1600, 8)
    assert torch.all(X.reshape(-1, 8)[(30 + 41 + 7) * 8:] == 42.0)
    parts = O.decollate_tensor(torch.arange(2 * 5 * 3.0).view(2, 5, 3), [4, 3, 2])
    assert [p.shape[0] for p in parts] == [4, 3, 2] and float(parts[1][0, 0]) == 12.0

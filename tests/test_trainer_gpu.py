"""sst_b200.train.Trainer on a B200 against the oracle's restatement of the reference training step
(recognition_model.py:57-64, :76-118, :293): loss values, the AdamW update of every parameter after one and two steps
(fp32 mode, dropout 0, shift 0), summed gradient accumulation over micro-batches (Q12), and the bf16 shadow / re-packing
of the GEMM operands after an optimizer step."""
import random

import pytest
import torch

import sst_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(cfg, sd, dtype):
    import sst_b200  # noqa: F401
    from sst_b200 import architecture as A
    A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=cfg["n_enc"], num_layers_decoder=cfg["n_dec"],
                n_heads_encoder=8, n_heads_decoder=8, relative_distance=cfg["rel_dist"], dropout_model=0.0, dropout_pos_emb=0.0,
                sst_dtype=dtype)
    model = A.Model(112, 44, 43, DEV).to(DEV)
    model.load_state_dict(sd)
    return model


def _oracle_steps(cfg, sd, batches, accumulate_every):
    """Reference loop: loss.backward() sums into .grad, AdamW when the accumulated chunk count reaches batch_size_grad."""
    sd = {k: v.clone() for k, v in sd.items()}
    names = O.trainable_names(sd, cfg)
    m = {n: torch.zeros_like(sd[n]) for n in names}
    v = {n: torch.zeros_like(sd[n]) for n in names}
    acc = {n: torch.zeros_like(sd[n]) for n in names}
    losses, step = [], 0
    for it, batch in enumerate(batches):
        res, grads, stats = O.loss_and_grads(sd, cfg, batch, True, 0)
        for k, val in stats.items():
            sd[k] = val
        losses.append(float(res["loss"]))
        for n in names:
            if n in grads:
                acc[n] += grads[n]
        if (it + 1) % accumulate_every == 0:
            step += 1
            lr = O.lr_schedule(it)
            for n in names:
                O.adamw_step(sd[n], acc[n], m[n], v[n], step, lr)
                acc[n].zero_()
    return sd, losses


@pytest.mark.parametrize("accumulate_every", [1, 2])
def test_trainer_matches_reference_loop(accumulate_every):
    from sst_b200.train import Trainer
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2)
    sd0 = O.synthetic_state_dict(cfg, 7)
    batches = [O.synthetic_batch(seed=40 + k, ragged=[70, 100, 30], tgt_lens=[9, 14, 5]) for k in range(2)]
    ref_sd, ref_losses = _oracle_steps(cfg, sd0, batches, accumulate_every)
    model = _model(cfg, sd0, "fp32")
    # one batch = 1 chunk of 1600 samples: batch_size_grad = accumulate_every chunks
    tr = Trainer(model, alpha_loss=cfg["alpha"], batch_size_grad=accumulate_every, learning_rate_warmup=1500)
    got_losses = []
    for batch in batches:
        dev = tr.to_device(tr.prepare(batch))
        losses = tr.step_device(dev, shift_r=0)
        tr.fetch_losses(losses)
        got_losses.append(tr.wait_losses()[0])
    for a, b in zip(got_losses, ref_losses):
        assert abs(a - b) < 1e-4 * abs(b), (got_losses, ref_losses)
    # parameters after the update(s): AdamW's first steps move every weight by ~lr regardless of the gradient scale, so
    # compare the UPDATE (new - old) against the reference's update
    new = dict(model.named_parameters())
    worst = 0.0
    for n in O.trainable_names(sd0, cfg):
        if n.startswith("conv_blocks.") and n.endswith(".bias") and ".bn" not in n and "res_norm" not in n:
            continue      # conv biases in front of BatchNorm: the true gradient is exactly 0, AdamW amplifies rounding noise
        upd_ref = (ref_sd[n] - sd0[n]).double()
        upd = (new[n].detach().cpu() - sd0[n]).double()
        denom = float(upd_ref.abs().max()) + 1e-12
        frac_bad = float(((upd - upd_ref).abs() > 0.05 * denom).double().mean())
        worst = max(worst, frac_bad)
        # sign(g)-like first updates flip where |g| ~ 0: allow a small fraction of elements to differ
        assert frac_bad < 0.02, "%s: %.3f of the elements moved differently from the reference" % (n, frac_bad)
    print("largest fraction of deviating update elements: %.4f" % worst)
    for k in ref_sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            buf = dict(model.named_buffers())[k].cpu()
            assert float((buf - ref_sd[k]).abs().max()) < 1e-4 * (float(ref_sd[k].abs().max()) + 1e-6), k


def test_bf16_shadow_and_repack_follow_the_optimizer():
    """After a step the packed bf16 operands must equal a fresh cast of the updated fp32 masters."""
    from sst_b200.train import Trainer
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2)
    sd0 = O.synthetic_state_dict(cfg, 8)
    model = _model(cfg, sd0, "bf16")
    tr = Trainer(model, alpha_loss=0.2, batch_size_grad=1)
    batch = O.synthetic_batch(seed=50, ragged=[70, 100, 30], tgt_lens=[9, 14, 5])
    random.seed(0)
    tr.step(batch)
    torch.cuda.synchronize()
    eng = tr.eng
    P = dict(model.named_parameters())
    w1 = P["transformerEncoder.layers.0.linear1.weight"].detach()
    assert torch.equal(eng.pk["transformerEncoder.layers.0.linear1"], w1.bfloat16())
    assert torch.equal(eng.pk["transformerEncoder.layers.0.linear1.T"], w1.bfloat16().t().contiguous())
    assert not torch.equal(w1.cpu(), sd0["transformerEncoder.layers.0.linear1.weight"])      # the step did move it
    wq = P["transformerEncoder.layers.0.self_attn.w_q"].detach()                                # (H, D, dh) -> rows (h, a), cols f
    packed = eng.pk["transformerEncoder.layers.0.self_attn.qkv"][:768]
    assert torch.equal(packed, wq.permute(0, 2, 1).reshape(768, 768).bfloat16())


def test_checkpoint_resume_continues_the_run():
    """Two steps straight through vs one step, save, fresh Trainer, load, one step (dropout on: the seeds derive from the
    step counter that the checkpoint carries).  fp32 atomics (split-K weight gradients, bias sums) make two runs agree
    to rounding, not bitwise, so the comparison is to 1e-3 of the moment / update scale."""
    from sst_b200.train import Trainer
    from sst_b200 import architecture as A
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2)
    sd0 = O.synthetic_state_dict(cfg, 9)
    batches = [O.synthetic_batch(seed=60 + k, ragged=[70, 100, 30], tgt_lens=[9, 14, 5]) for k in range(2)]

    def fresh():
        model = _model(cfg, sd0, "bf16")
        A.configure(dropout_model=0.2, dropout_pos_emb=0.2)
        model.cfg["dropout"], model.cfg["dropout_pos"] = 0.2, 0.2
        return Trainer(model, alpha_loss=0.2, batch_size_grad=1, seed=3)

    t1 = fresh()
    for b in batches:
        l1 = t1.step_device(t1.to_device(t1.prepare(b)), shift_r=2)
    t2 = fresh()
    t2.step_device(t2.to_device(t2.prepare(batches[0])), shift_r=2)
    ck = t2.state_dict(dataparallel_prefix=True)
    assert all(k.startswith("module.") for k in ck["model"])
    t3 = fresh()
    t3.load_state_dict(ck)
    assert (t3.flat.step_count, t3.batch_idx) == (1, 1)
    l3 = t3.step_device(t3.to_device(t3.prepare(batches[1])), shift_r=2)
    torch.cuda.synchronize()
    assert (t3.flat.step_count, t3.batch_idx) == (t1.flat.step_count, t1.batch_idx) == (2, 2)
    assert float((l1 - l3).abs().max()) < 1e-3 * float(l1.abs().max())              # same dropout masks, same weights
    assert float((t1.flat.m - t3.flat.m).abs().max()) < 1e-3 * float(t1.flat.m.abs().max())
    assert float((t1.flat.p - t3.flat.p).abs().max()) < 2.1 * t1.lr                   # at most one AdamW step apart anywhere
    assert float(((t1.flat.p - t3.flat.p).abs() > 0.05 * t1.lr).float().mean()) < 0.02
    b1, b3 = dict(t1.model.named_buffers()), dict(t3.model.named_buffers())
    for k in b1:
        assert float((b1[k].float() - b3[k].float()).abs().max()) <= 1e-4 * (float(b1[k].float().abs().max()) + 1e-6), k


def test_pipelined_run_gives_the_stepwise_losses():
    """Trainer.run (upload of batch i+1 on a copy stream, deferred loss read-back: the path bench.py's `e2e` times) against
    the plain step-by-step loop on the same batches: same losses per step, same parameters at the end (fp32 mode, dropout 0,
    to the rounding of the fp32 atomics), and every step's loss is reported exactly once and in order."""
    import random
    from sst_b200.train import Trainer
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2)
    sd0 = O.synthetic_state_dict(cfg, 11)
    batches = [O.synthetic_batch(seed=70 + k, ragged=[[70, 100, 30], [100, 64, 90], [55, 80, 100], [100, 100, 21]][k],
                                 tgt_lens=[9, 14, 5]) for k in range(4)]

    t1 = Trainer(_model(cfg, sd0, "fp32"), alpha_loss=0.2, batch_size_grad=1, seed=0)
    random.seed(5)
    want = []
    for b in batches:
        t1.step(b)
        want.append(t1.wait_losses())

    t2 = Trainer(_model(cfg, sd0, "fp32"), alpha_loss=0.2, batch_size_grad=1, seed=0)
    random.seed(5)                                   # the same random input shifts (architecture.py:105)
    got = list(t2.run(t2.prepare(b) for b in batches))
    torch.cuda.synchronize()
    assert len(got) == len(want) == 4
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            assert abs(float(x) - float(y)) < 1e-4 * max(1.0, abs(float(y)))
    assert t2.batch_idx == 4 and t2.flat.step_count == 4
    scale = float(t1.flat.p.abs().max())
    assert float((t1.flat.p - t2.flat.p).abs().max()) < 1e-4 * scale


def _graph_model(cfg, sd, dtype, dropout):
    import sst_b200  # noqa: F401
    from sst_b200 import architecture as A
    A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=cfg["n_enc"], num_layers_decoder=cfg["n_dec"],
                n_heads_encoder=8, n_heads_decoder=8, relative_distance=cfg["rel_dist"], dropout_model=dropout, dropout_pos_emb=dropout,
                sst_dtype=dtype)
    model = A.Model(112, 44, 43, DEV).to(DEV)
    model.load_state_dict(sd)
    return model


@pytest.mark.parametrize("accumulate_every", [1, 2])
def test_graphed_steps_equal_eager_steps(accumulate_every):
    """Trainer.step_graphed replays the micro-step as one CUDA graph (per batch signature, captured after two eager
    steps).  Without dropout the graphed loop must BE the eager loop: same losses, and the same parameters after 8 steps through the
    warm-up learning-rate schedule and Adam's bias corrections (those reach the replay through device memory) -- fp32 mode, compared
    at 2e-5 (the weight-gradient GEMMs accumulate their split-K partials with atomics, in a different order run to run)."""
    from sst_b200.train import Trainer
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2)
    sd0 = O.synthetic_state_dict(cfg, 7)
    batches = [O.synthetic_batch(seed=40 + (k % 2), ragged=[70, 100, 30], tgt_lens=[9, 14, 5]) for k in range(8)]
    out = {}
    for mode in ("eager", "graphed"):
        model = _graph_model(cfg, sd0, "fp32", 0.0)
        tr = Trainer(model, alpha_loss=cfg["alpha"], batch_size_grad=accumulate_every, learning_rate_warmup=1500)
        losses = []
        for batch in batches:
            dev = tr.to_device(tr.prepare(batch))
            ls = tr.step_device(dev, shift_r=3) if mode == "eager" else tr.step_graphed(dev, shift_r=3)
            tr.fetch_losses(ls)
            losses.append(tr.wait_losses())
        out[mode] = (losses, {n: p.detach().clone() for n, p in model.named_parameters()}, tr)
    tr = out["graphed"][2]
    live = [e for e in tr._graphs["graphs"].values() if e["graph"] is not None]
    assert len(live) == (1 if accumulate_every == 1 else 2), "one graph per signature (accumulate / step) expected, got %d" % len(live)
    assert tr.flat.step_count == out["eager"][2].flat.step_count == 8 // accumulate_every and tr.batch_idx == 8
    # Two runs of the SAME eager loop already differ a little: the weight-gradient GEMMs add their split-K partials with atomics, and
    # AdamW's first steps move a weight by ~lr * sign(g), so an entry whose gradient is ~0 can go the other way.  Yardsticks that
    # ignore those few entries but catch a replay that used a stale learning rate / bias correction / input: the loss of every step
    # to 1e-3 (a stale lr would be off by up to 8/3 at step 8), and each tensor's total update in relative L2.
    for a, b in zip(out["eager"][0], out["graphed"][0]):
        for x, y in zip(a, b):
            assert abs(x - y) <= 1e-3 * abs(x), (out["eager"][0], out["graphed"][0])
    worst = (0.0, None)
    for n, p in out["eager"][1].items():
        upd = (p - sd0[n].to(DEV)).double()
        if float(upd.norm()) == 0.0:
            assert torch.equal(out["graphed"][1][n], p)
            continue
        if n.startswith("conv_blocks.") and n.endswith(".bias") and ".bn" not in n and "res_norm" not in n:
            continue      # conv biases in front of BatchNorm: true gradient 0, AdamW amplifies pure rounding noise
        e = float((out["graphed"][1][n].double() - p.double()).norm() / upd.norm())
        worst = max(worst, (e, n))
    assert worst[0] < 0.1, worst


def test_graphed_replays_draw_fresh_dropout_masks():
    """Dropout seeds cross the C ABI by value, so a replayed graph would repeat its masks: every Philox-drawing kernel adds the
    registered device salt (sst_set_dropout_salt), rewritten in front of each replay.  The same batch replayed on frozen weights
    (accumulation only, no optimizer step) must give DIFFERENT losses each time, all within the spread eager steps show."""
    from sst_b200.train import Trainer
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2)
    sd0 = O.synthetic_state_dict(cfg, 9)
    batch = O.synthetic_batch(seed=77, ragged=[70, 100, 30], tgt_lens=[9, 14, 5])
    model = _graph_model(cfg, sd0, "bf16", 0.2)
    tr = Trainer(model, alpha_loss=0.2, batch_size_grad=10 ** 9)
    vals = []
    for k in range(8):
        dev = tr.to_device(tr.prepare(batch))
        tr.fetch_losses(tr.step_graphed(dev, shift_r=0))
        vals.append(tr.wait_losses()[0])
    eager, replayed = vals[:2], vals[2:]
    assert any(e["graph"] is not None for e in tr._graphs["graphs"].values())
    assert len(set(round(v, 5) for v in replayed)) == len(replayed), replayed
    mean = sum(vals) / len(vals)
    assert all(abs(v - mean) < 0.25 * abs(mean) for v in vals), vals

"""The drop-in seam itself (BASELINE.json north_star: "the architecture.py Model / recognition_model.py training-loop API stays
intact"): the reference's training-loop body, recognition_model.py:76-118, restated line for line against
`sst_b200.architecture.Model` + `sst_b200.LabelSmoothingLoss` -- torch autograd (`loss.backward()` -> `p.grad`),
`torch.optim.AdamW`, `F.ctc_loss` on the returned logits -- and checked against the fixtures of the UNMODIFIED reference
(losses, every parameter's gradient, the parameters after one optimizer step).  Plus the two other callers of the API:
BeamSearch.py:111-114 (`memory.repeat(n_hyp, 1, 1)`) and a decoder head count that differs from the encoder's."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

import sst_oracle as O
from helpers import load_golden, golden_inputs, rel_err, check_grads_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(cfg, sd, dtype):
    import sst_b200  # noqa: F401
    from sst_b200 import architecture as A
    A.configure(model_size=cfg["d_model"], feed_forward_layer_size=cfg["d_ff"], num_layers_encoder=cfg["n_enc"],
                num_layers_decoder=cfg["n_dec"], n_heads_encoder=cfg["n_heads"], n_heads_decoder=cfg.get("n_heads_dec", cfg["n_heads"]),
                relative_distance=cfg["rel_dist"], dropout_model=cfg["dropout"], dropout_pos_emb=cfg["dropout_pos"],
                sst_dtype="bf16" if dtype == torch.bfloat16 else "fp32")
    model = A.Model(112, 44, 43, DEV).to(DEV)
    model.load_state_dict(sd)
    return A, model


@pytest.mark.parametrize("name", ["short_hybrid", "ragged_hybrid"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reference_training_loop_body_verbatim(name, dtype, monkeypatch):
    from sst_b200.data_utils import combine_fixed_length
    from sst_b200.LabelSmoothingLoss import LabelSmoothingLoss
    z, meta = load_golden(name)
    cfg, sd, example = golden_inputs(meta)
    A, model = _model(cfg, sd, dtype)
    monkeypatch.setattr(A.random, "randrange", lambda n: 0)          # the fixtures force the time shift r = 0 (Q13)
    device, n_phones, pad, alpha_loss = DEV, 43, 42, cfg["alpha"]
    loss_fn = LabelSmoothingLoss(epsilon=0.1, num_classes=n_phones)  # recognition_model.py:285
    optim = torch.optim.AdamW(model.parameters(), lr=3e-4)           # :293
    for param_group in optim.param_groups:                            # schedule_lr(0), :57-64
        param_group['lr'] = 1 * 3e-4 / 1500
    optim.zero_grad()
    # ---- recognition_model.py:71-114 --------------------------------------------------------------------------------
    model.train()
    X = combine_fixed_length(example['raw_emg'], 200 * 8).to(device)
    y = example['phonemes_int']
    target = nn.utils.rnn.pad_sequence(example['phonemes_int'], batch_first=True, padding_value=pad).to(device)
    tgt = target[:, :-1]
    target = target[:, 1:]
    out_enc, out_dec = model(example['lengths'], device, x_raw=X, y=tgt)
    logits_enc, logits_dec = out_enc.detach().cpu(), out_dec.detach().cpu()
    out_enc = F.log_softmax(out_enc, 2)
    out_enc = out_enc.transpose(1, 0)
    phonemes_int_lengths = [item - 2 for item in example['phonemes_int_lengths']]
    y = [item[1:-1] for item in y]
    y = nn.utils.rnn.pad_sequence(y, batch_first=True, padding_value=pad).to(device)
    loss_enc = F.ctc_loss(out_enc, y, example['lengths'], phonemes_int_lengths, blank=n_phones)
    out_dec = out_dec.permute(0, 2, 1)
    loss_dec = loss_fn(out_dec, target)
    loss = (1 - alpha_loss) * loss_dec + alpha_loss * loss_enc
    loss.backward()
    # ------------------------------------------------------------------------------------------------------------------
    bf16 = dtype == torch.bfloat16
    tol = 2e-2 if bf16 else 1e-4
    assert logits_enc.shape == z["out_enc"].shape and logits_dec.shape == z["out_dec"].shape
    ref_enc = torch.from_numpy(z["out_enc"])
    for b, l in enumerate(example["lengths"]):
        assert rel_err(logits_enc[b, :l], ref_enc[b, :l], floor=float(ref_enc.abs().max())) < tol
    assert rel_err(logits_dec, z["out_dec"]) < tol
    assert abs(float(loss_enc) - float(z["loss_enc"])) < tol * abs(float(z["loss_enc"]))
    assert abs(float(loss_dec) - float(z["loss_dec"])) < tol * abs(float(z["loss_dec"]))
    assert abs(float(loss) - float(z["loss"])) < tol * abs(float(z["loss"]))
    params = dict(model.named_parameters())
    for n in meta["none_grad"]:                                        # Q2 / Q14: AdamW must skip them exactly as in the reference
        assert params[n].grad is None, n
    check_grads_l2(z, meta, {n: params[n].grad for n in meta["grad_names"]}, tol, bf16, "%s %s autograd" % (name, dtype))
    before = {n: params[n].detach().clone() for n in meta["grad_names"]}
    optim.step()                                                       # :115-118
    optim.zero_grad()
    if not bf16:
        for n in meta["grad_names"]:
            if ("conv" in n and n.endswith("bias")) or "residual_path.bias" in n:
                continue                                               # true gradient zero (Q7): the update is sign(noise) * lr
            idx = torch.from_numpy(z["gidx/" + n])
            ref_upd = torch.from_numpy(z["pnew/" + n]) - sd[n].reshape(-1)[idx]
            got_upd = (params[n].detach() - before[n]).reshape(-1).cpu()[idx]
            assert float((got_upd - ref_upd).abs().max()) <= 2e-8 + 0.02 * float(ref_upd.abs().max()), n
    # the next forward must see the stepped weights without any extra call (the reference loop makes none): take a large
    # step so that the difference is far above the tolerance, then compare with the oracle on the model's own state_dict
    for param_group in optim.param_groups:
        param_group['lr'] = 3e-3
    X2 = combine_fixed_length(example['raw_emg'], 200 * 8).to(device)
    out_enc2, out_dec2 = model(example['lengths'], device, x_raw=X2, y=tgt)
    (out_enc2.float().pow(2).mean() + out_dec2.float().pow(2).mean()).backward()
    optim.step()
    optim.zero_grad()
    sd_now = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    X3 = combine_fixed_length(example['raw_emg'], 200 * 8)
    with torch.no_grad():
        ref3 = O.forward_training(sd_now, cfg, X3.clone(), tgt.cpu(), example['lengths'], training=True, shift_r=0)
        got3 = model(example['lengths'], device, x_raw=X3.to(device), y=tgt)
    moved = rel_err(ref3[1], logits_dec)
    assert moved > 10 * tol, "the large step did not move the logits (%.2e): the check below would be vacuous" % moved
    assert rel_err(got3[1].cpu(), ref3[1]) < tol


def test_beam_search_call_pattern_memory_repeat():
    """BeamSearch.py:84,111-114: encoder once on ONE utterance, then every decoder step on memory.repeat(n_hyp, 1, 1) with the
    hypotheses as the batch; the cached src_key_padding_mask broadcasts over them.  fp32, 1e-4 against the oracle."""
    cfg = O.make_cfg(n_enc=1, n_dec=2, rel_dist=100)
    sd = O.synthetic_state_dict(cfg, 31)
    A, model = _model(cfg, sd, torch.float32)
    model.eval()
    batch = O.synthetic_batch(seed=9, ragged=[137], tgt_lens=[6])
    X = O.combine_fixed_length(batch["raw_emg"])
    histories = torch.tensor([[41, 3, 7, 1], [41, 3, 9, 0], [41, 12, 7, 5], [41, 30, 30, 2], [41, 0, 1, 2]], dtype=torch.int64)
    with torch.no_grad():
        mem_ref, kpm = O.encode(sd, cfg, X.clone(), batch["lengths"], False)
        n = histories.shape[0]
        ref = F.linear(O.decode(sd, cfg, histories, mem_ref.repeat(n, 1, 1), kpm.repeat(n, 1), False), sd["w_out.weight"], sd["w_out.bias"])
        memory, _ = model(batch["lengths"], DEV, mode='beam_search', part='encoder', x_raw=X.to(DEV))
        memory_stub = memory.repeat(n, 1, 1)
        step_logits = model(batch["lengths"], DEV, mode='beam_search', part='decoder', y=histories.to(DEV), memory=memory_stub)
    assert step_logits.shape == ref.shape
    assert rel_err(step_logits.cpu(), ref) < 1e-4
    with pytest.raises(Exception):
        model(batch["lengths"], DEV, mode='beam_search', part='decoder', y=histories[:3].to(DEV), memory=memory_stub)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_beam_decoder_shares_the_memory_across_hypotheses(dtype):
    """sst_b200.beam_search.BeamDecoder projects the encoder memory once per utterance and lets every hypothesis
    attend to the same key / value rows (SstAttnDesc.k_off = 0) instead of BeamSearch.py:111's memory.repeat -- the logits must be
    the memory.repeat call's BIT FOR BIT (same arithmetic on the same values), step after step of a growing, re-ordered beam, and
    (fp32) sit on the oracle at 1e-4.  Two utterances in the encoder batch: the search runs on the SHORTER one, whose memory is padded."""
    from sst_b200.beam_search import BeamDecoder
    cfg = O.make_cfg(n_enc=1, n_dec=2, rel_dist=100)
    sd = O.synthetic_state_dict(cfg, 33)
    A, model = _model(cfg, sd, dtype)
    model.eval()
    batch = O.synthetic_batch(seed=19, ragged=[200, 137], tgt_lens=[6, 6])
    X = O.combine_fixed_length(batch["raw_emg"])
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        memory, _ = model(batch["lengths"], DEV, mode='beam_search', part='encoder', x_raw=X.to(DEV))
        dec = BeamDecoder(model, memory, index=1)
        mem1 = memory[1:2]
        model._mem_lens_saved = model._mem_lens
        hist = torch.tensor([[41]], dtype=torch.int64)
        for step in range(6):
            n = hist.shape[0]
            got = dec(hist.to(DEV))
            # the reference call pattern on the same utterance: its cached mask row repeated with the memory
            model._mem_lens, model._mem_shape = model._mem_lens_saved[1:2], (1, memory.shape[1])
            want = model(batch["lengths"], DEV, mode='beam_search', part='decoder', y=hist.to(DEV), memory=mem1.repeat(n, 1, 1))
            model._mem_lens, model._mem_shape = model._mem_lens_saved, (2, memory.shape[1])
            assert got.shape == want.shape == (n, step + 1, 43)
            assert torch.equal(got, want), (step, float((got - want).abs().max()))
            assert torch.equal(dec.step_logits(hist.to(DEV)), want[:, -1, :-2])
            if dtype == torch.float32:
                mem_ref, kpm = O.encode(sd, cfg, X.clone(), batch["lengths"], False)
                ref = F.linear(O.decode(sd, cfg, hist, mem_ref[1:2].repeat(n, 1, 1), kpm[1:2].repeat(n, 1), False), sd["w_out.weight"], sd["w_out.bias"])
                assert rel_err(got.cpu(), ref) < 1e-4
            # grow and re-order the beam: top-3 continuations of every hypothesis, shuffled, at most 23 kept
            top = torch.topk(torch.log_softmax(got[:, -1, :-2].float().cpu(), 1), 3, dim=1).indices
            new = torch.cat([torch.cat([hist, top[:, k:k + 1]], 1) for k in range(3)], 0)
            hist = new[torch.randperm(new.shape[0], generator=g)[:23]]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decoder_head_count_differs_from_encoder(dtype):
    """FLAGS.n_heads_decoder != FLAGS.n_heads_encoder (architecture.py:16-17): 4 decoder heads of 192 dims next to 8 encoder
    heads of 96; one training step against the oracle."""
    from test_engine_gpu import make_engine, run_step, _valid_frames_err
    cfg = O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.3)
    cfg["n_heads_dec"] = 4
    sd = O.synthetic_state_dict(cfg, 41)
    assert sd["transformerDecoder.layers.0.self_attn.w_q"].shape == (4, 768, 192)
    batch = O.synthetic_batch(seed=10, ragged=[150, 90], tgt_lens=[11, 6])
    res, grads, _ = O.loss_and_grads({k: v.clone() for k, v in sd.items()}, cfg, batch, True, 0)
    eng = make_engine(cfg, sd, dtype)
    out_enc, out_dec, loss, loss_dec, loss_enc, G, ctx = run_step(eng, cfg, batch)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert _valid_frames_err(out_enc, res["out_enc"], batch["lengths"]) < tol
    assert rel_err(out_dec, res["out_dec"]) < tol
    assert abs(loss - float(res["loss"])) < tol * abs(float(res["loss"]))
    from test_engine_gpu import oracle_autocast_bf16_grads
    from helpers import l2_rows, assert_l2_rows
    bf16 = dtype == torch.bfloat16
    names = sorted(grads)
    flat = lambda d: {n: d[n].detach().double().cpu().reshape(-1).numpy() for n in names}      # noqa: E731
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    b64 = dict(batch)
    b64["raw_emg"] = [x.double() for x in batch["raw_emg"]]
    g64 = O.loss_and_grads(sd64, cfg, b64, True, 0)[1]
    gbf = oracle_autocast_bf16_grads(sd, cfg, batch) if bf16 else None
    gmax = max(float(g.abs().max()) for g in grads.values())
    rows = l2_rows(names, flat(G), flat(grads), flat(g64), flat(gbf) if gbf is not None else None, gmax)
    assert_l2_rows(rows, tol, bf16, str(dtype))

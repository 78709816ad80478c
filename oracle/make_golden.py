"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py) on the seeded synthetic inputs/weights of
oracle/sst_oracle.py.  Run in the build container:  python oracle/make_golden.py

Each fixture holds: the case description (cfg, seeds, lengths), the reference's logits and losses,
and for every trainable parameter the gradient's max-abs, L2 norm, a seeded sample of <=256 entries
(full tensor when <=4096 elements), so the files stay small while pinning every tensor.
"""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness            # noqa: E402
import sst_oracle as O        # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # BASELINE.json config 1: 6-layer encoder + CTC, B=4 x 1600 samples
    "cfg1_enc_ctc": dict(cfg=O.make_cfg(n_enc=6, n_dec=0, rel_dist=100), wseed=0,
                         batch=dict(n_utt=4, frames=200, tgt_len=30, seed=1234), mode="encoder"),
    # ragged hybrid CTC/attention step (Q8/Q9/Q11), alpha = 0.7 (config 3 weighting)
    "ragged_hybrid": dict(cfg=O.make_cfg(n_enc=2, n_dec=2, rel_dist=100, alpha=0.7), wseed=1,
                          batch=dict(seed=4321, ragged=[200, 180, 200, 150], tgt_lens=[30, 25, 30, 12]),
                          mode="hybrid"),
    # L < R branch of the rel-pos module (start_pos > 0, no -1e8 edges) and L not multiple of anything
    "short_hybrid": dict(cfg=O.make_cfg(n_enc=1, n_dec=1, rel_dist=100, alpha=0.2), wseed=2,
                         batch=dict(seed=99, ragged=[70, 100, 30], tgt_lens=[9, 14, 5]), mode="hybrid"),
    # the benchmarked shape (BASELINE.json configs[1]): utterances of 1000 frames -- banded attention over 8 query tiles with
    # skipped key tiles, R = 100 -- with one ragged utterance (padded query rows / keys), 2 + 1 layers
    "full_2p1": dict(cfg=O.make_cfg(n_enc=2, n_dec=1, rel_dist=100, alpha=0.2), wseed=3,
                     batch=dict(seed=77, ragged=[1000, 777, 1000], tgt_lens=[60, 45, 52]), mode="hybrid"),
    # the benchmarked MODEL (6 + 6 layers, d768 h8 ffn3072 R100) at the benchmarked utterance length, batch 2
    "full_6p6": dict(cfg=O.make_cfg(n_enc=6, n_dec=6, rel_dist=100, alpha=0.2), wseed=4,
                     batch=dict(n_utt=2, frames=1000, tgt_len=100, seed=78), mode="hybrid"),
}


def sample_indices(numel, name):
    if numel <= 4096:
        return np.arange(numel)
    g = np.random.default_rng(abs(hash_name(name)) % (2 ** 32))
    return np.sort(g.choice(numel, 256, replace=False))


def hash_name(name):
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) % (2 ** 64)
    return h


def jitter_running_stats(sd, seed):
    """Non-trivial BN running buffers so eval-mode (decode) parity exercises them."""
    for i, name in enumerate(sorted(sd)):
        g = torch.Generator().manual_seed(seed * 7 + i)
        if name.endswith("running_mean"):
            sd[name] = 0.05 * torch.randn(sd[name].shape, generator=g)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(sd[name].shape, generator=g)
    return sd


def run_case(name, case):
    cfg = case["cfg"]
    sd = O.synthetic_state_dict(cfg, case["wseed"])
    # the reference Model always owns a decoder; n_dec=0 cases build a 1-layer one that is never called
    ref_cfg = dict(cfg)
    model = ref_harness.build_model(ref_cfg, sd)
    arch, tr, LS, du, FLAGS = ref_harness.load(ref_harness.cfg_to_argv(ref_cfg))
    batch = O.synthetic_batch(**case["batch"])
    model.train()
    arch.random.seed(0)
    real_randrange = arch.random.randrange
    arch.random.randrange = lambda n: 0            # Q13: parity runs force r = 0
    X = du.combine_fixed_length(batch["raw_emg"], 1600).clone()
    tgt_in, tgt_out, ctc_tgt, ctc_lens = O.make_targets(batch)
    out = {}
    if case["mode"] == "encoder":
        _, out_enc = model(batch["lengths"], "cpu", x_raw=X, mode="greedy_search", part="encoder")
        out_dec = None
    else:
        out_enc, out_dec = model(batch["lengths"], "cpu", x_raw=X, y=tgt_in)
    lp = F.log_softmax(out_enc, 2).transpose(1, 0)
    loss_enc = F.ctc_loss(lp, ctc_tgt, batch["lengths"], ctc_lens, blank=43)
    if out_dec is not None:
        loss_dec = LS.LabelSmoothingLoss(epsilon=cfg["eps_ls"], num_classes=43)(out_dec.permute(0, 2, 1), tgt_out)
        loss = (1 - cfg["alpha"]) * loss_dec + cfg["alpha"] * loss_enc
        out["out_dec"] = out_dec.detach().numpy()
        out["loss_dec"] = np.float64(loss_dec.item())
    else:
        loss = loss_enc
    loss.backward()
    arch.random.randrange = real_randrange
    out["out_enc"] = out_enc.detach().numpy()
    out["loss_enc"] = np.float64(loss_enc.item())
    out["loss"] = np.float64(loss.item())
    names = []
    for pname, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().reshape(-1).numpy()
        idx = sample_indices(g.size, pname)
        names.append(pname)
        out["gidx/" + pname] = idx.astype(np.int64)
        out["gval/" + pname] = g[idx].astype(np.float32)
        out["gstat/" + pname] = np.array([np.abs(g).max(), np.sqrt((g.astype(np.float64) ** 2).sum())])
    # exact-arithmetic statement of the same step: the (reference-pinned) oracle evaluated in float64.  The reference's own
    # fp32 rounding error against it is what bounds any fp32 re-implementation on ill-conditioned tensors (BatchNorm /
    # LayerNorm backward cancellation), so the fixtures carry both.
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    b64 = dict(batch)
    b64["raw_emg"] = [x.double() for x in batch["raw_emg"]]
    r64, g64, _ = O.loss_and_grads(sd64, cfg, b64, True, 0)
    for pname in names:
        g = g64[pname].reshape(-1).numpy()
        out["gtruth/" + pname] = g[out["gidx/" + pname]]
        out["gtruthstat/" + pname] = np.array([np.abs(g).max()])
    out["out_enc_truth"] = r64["out_enc"].numpy().astype(np.float32)
    out["loss_truth"] = np.float64(r64["loss"].item())
    none_grad = [n for n, p in model.named_parameters() if p.grad is None]
    # The same UNMODIFIED reference step under torch.autocast(bfloat16) -- PyTorch's own bf16 mode of the reference code
    # (matmul / conv / linear operands rounded to bf16, normalisations and losses in fp32).  Its distance from the fp32
    # reference is the noise floor of ANY bf16 evaluation of this network: a pre-activation within bf16 rounding of zero flips
    # its ReLU mask bit and moves the gradient by a whole term.  The bf16-mode parity test uses it as the yardstick for the
    # tensors behind ReLU / BatchNorm kinks (tests/helpers.check_grads_l2).
    model_bf = ref_harness.build_model(ref_cfg, sd)
    model_bf.train()
    arch.random.randrange = lambda n: 0
    Xb = du.combine_fixed_length(batch["raw_emg"], 1600).clone()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        if case["mode"] == "encoder":
            _, oe_bf = model_bf(batch["lengths"], "cpu", x_raw=Xb, mode="greedy_search", part="encoder")
            od_bf = None
        else:
            oe_bf, od_bf = model_bf(batch["lengths"], "cpu", x_raw=Xb, y=tgt_in)
    oe_bf = oe_bf.float()
    l_bf = F.ctc_loss(F.log_softmax(oe_bf, 2).transpose(1, 0), ctc_tgt, batch["lengths"], ctc_lens, blank=43)
    out["loss_enc_bf16"] = np.float64(l_bf.item())
    if od_bf is not None:
        ld_bf = LS.LabelSmoothingLoss(epsilon=cfg["eps_ls"], num_classes=43)(od_bf.float().permute(0, 2, 1), tgt_out)
        out["loss_dec_bf16"] = np.float64(ld_bf.item())
        l_bf = (1 - cfg["alpha"]) * ld_bf + cfg["alpha"] * l_bf
    l_bf.backward()
    arch.random.randrange = real_randrange
    bfp = dict(model_bf.named_parameters())
    for pname in names:
        out["gbf16/" + pname] = bfp[pname].grad.detach().float().reshape(-1).numpy()[out["gidx/" + pname]].astype(np.float32)
    out["out_enc_bf16_err"] = np.float64(float((oe_bf.detach() - out_enc.detach()).abs().max() / out_enc.detach().abs().max()))
    del model_bf, bfp
    msd = model.state_dict()
    for k in msd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["bn/" + k] = msd[k].numpy().copy()
    # CTC best-path decode of the training-mode logits (config 5 decode rule on reference logits)
    ctc_dec = O.ctc_greedy_collapse(out_enc.detach(), batch["lengths"])
    top2 = out_enc.detach().topk(2, dim=-1).values
    margin_ctc = float((top2[..., 0] - top2[..., 1]).min())
    meta = dict(case=name, cfg=cfg, wseed=case["wseed"], batch=case["batch"], mode=case["mode"],
                grad_names=names, none_grad=none_grad, ctc_decode=ctc_dec, ctc_min_margin=margin_ctc,
                torch=torch.__version__)

    # one AdamW step with the warm-up lr of iteration 0 (recognition_model.py:57-64,115-118,293)
    optim = torch.optim.AdamW(model.parameters(), lr=3e-4)
    lr0 = O.lr_schedule(0)
    for gparam in optim.param_groups:
        gparam["lr"] = lr0
    optim.step()
    for pname, p in model.named_parameters():
        if p.grad is None:
            continue
        flat = p.detach().reshape(-1).numpy()
        out["pnew/" + pname] = flat[out["gidx/" + pname]].astype(np.float32)

    # greedy attention decode in eval mode with jittered running stats (greedy_search.py:7-53)
    if case["mode"] == "hybrid":
        sd2 = jitter_running_stats(O.synthetic_state_dict(cfg, case["wseed"]), case["wseed"])
        model2 = ref_harness.build_model(ref_cfg, sd2)
        model2.eval()
        import greedy_search
        Xe = du.combine_fixed_length(batch["raw_emg"], 1600).clone()
        phones, ids = greedy_search.run_greedy(model2, batch["lengths"], Xe, tgt_out, 43, "cpu")
        out["greedy_ids"] = ids.numpy()
        meta["greedy_phones"] = phones
        # margin of every argmax taken along the way (re-run the last full prefix)
        with torch.no_grad():
            mem, _ = model2(batch["lengths"], "cpu", mode="greedy_search", part="encoder", x_raw=Xe.clone())
            out["eval_memory_sample"] = mem.reshape(-1)[::997].numpy().copy()
            n_steps = int((ids != 42).sum(1).max())
            logits = model2(batch["lengths"], "cpu", mode="greedy_search", part="decoder",
                            y=ids[:, :max(n_steps - 1, 1)].long(), memory=mem)
            t2 = logits.topk(2, dim=-1).values
            meta["greedy_min_margin"] = float((t2[..., 0] - t2[..., 1]).min())
            out["eval_dec_logits"] = logits.numpy()
    out["meta"] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss", out["loss"], "loss_enc", out["loss_enc"], "ctc margin", margin_ctc,
          "greedy margin", meta.get("greedy_min_margin"), "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    sel = sys.argv[1:] or list(CASES)
    for n in sel:
        run_case(n, CASES[n])

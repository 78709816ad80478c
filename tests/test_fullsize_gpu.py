"""BASELINE-sized checks on a B200 (the oracle is too slow at this size): the bf16 tensor-core path against the fp32
CUDA-core parity path of the same engine on identical inputs and weights -- utterances of 1000 frames (banded attention
with skipped key tiles, 8 query tiles per head) and a ragged batch (padded query rows / keys, decollate + pad 42)."""
import pytest
import torch

import sst_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(cfg, sd, batch, dtype):
    from test_engine_gpu import make_engine, run_step
    eng = make_engine(cfg, sd, dtype)
    out_enc, out_dec, loss, loss_dec, loss_enc, G, ctx = run_step(eng, cfg, batch)
    return out_enc, out_dec, loss_dec, loss_enc, G


@pytest.mark.parametrize("lengths", [[1000] * 6, [1000, 640, 333, 128, 97, 1002]])
def test_bf16_tensor_core_path_tracks_fp32_path_at_full_length(lengths):
    from sst_b200.synthetic import make_batch
    cfg = O.make_cfg(n_enc=2, n_dec=1, rel_dist=100, alpha=0.2)
    sd = O.synthetic_state_dict(cfg, 3)
    total = sum(lengths)
    if total % 200:
        lengths = lengths[:-1] + [lengths[-1] + 200 - total % 200]      # fill the last 1600-sample chunk exactly
    batch = make_batch(lengths=lengths, tgt_min=40, tgt_max=60, seed=5)
    ref = _run(cfg, sd, batch, torch.float32)
    got = _run(cfg, sd, batch, torch.bfloat16)
    scale = float(ref[0].abs().max())
    for b, l in enumerate(lengths):
        e = float((got[0][b, :l] - ref[0][b, :l]).abs().max()) / scale
        assert e < 3e-2, "encoder logits of utterance %d: %.3e" % (b, e)
    assert abs(got[3] - ref[3]) < 2e-2 * abs(ref[3]), "CTC loss %f vs %f" % (got[3], ref[3])
    assert abs(got[2] - ref[2]) < 2e-2 * abs(ref[2]), "decoder loss %f vs %f" % (got[2], ref[2])
    # gradients: relative L2 over the attention / FFN / head tensors (conv + BN tensors sit behind ReLU/BN kinks)
    num = den = 0.0
    for n in ref[4]:
        if n.startswith("conv_blocks") or "relative_positional" in n or n.startswith("emg_projection"):
            continue
        a, r = got[4][n].double(), ref[4][n].double()
        assert bool(torch.isfinite(a).all()), n
        num += float(((a - r) ** 2).sum()); den += float((r ** 2).sum())
    assert (num / den) ** 0.5 < 3e-2, "gradient L2 error %.3e" % (num / den) ** 0.5

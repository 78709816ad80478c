"""Batching contract: the DynamicBatchSampler / collate_raw mirrors of sst_b200/read_emg.py against batches
emitted by the UNMODIFIED reference sampler (tests/golden/sampler_batches.json, oracle/make_golden_sampler.py), and the
properties of the data-parallel sharding the reference lacks.  CPU only."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import sst_b200  # noqa: E402,F401
from sst_b200 import read_emg as RE  # noqa: E402


def golden():
    with open(os.path.join(ROOT, "tests", "golden", "sampler_batches.json")) as f:
        return json.load(f)


def keep_fn(texts):
    import string
    return lambda i: any(ch in string.ascii_letters for ch in texts[i])


def make(g, s, **kw):
    return RE.DynamicBatchSampler(None, s["max_batch_length"], s["num_buckets"], shuffle=s["shuffle"], batch_ordering=s["ordering"],
                                  seed=s["seed"], epoch=s["epoch"], drop_last=s["drop_last"], max_batch_ex=s.get("max_batch_ex"),
                                  lengths_list=g["lengths"], keep=keep_fn(g["texts"]), **kw)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_same_batches_as_the_reference_sampler(case):
    g = golden()
    c = g["cases"][case]
    smp = make(g, c["setting"])
    assert np.allclose(smp._bucket_boundaries, c["boundaries"], rtol=1e-12)
    assert [list(b) for b in smp] == c["batches"]
    assert len(smp) == len(c["batches"])
    smp.set_epoch(c["setting"]["epoch"] + 1)
    assert [list(b) for b in smp] == c["batches_next_epoch"]


def test_live_reference_sampler_matches_the_fixture():
    import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not present")
    import make_golden_sampler as M
    g = golden()
    live = M.reference_batches(g["lengths"], g["texts"], [c["setting"] for c in g["cases"]])
    for a, b in zip(live, g["cases"]):
        assert a["batches"] == b["batches"] and a["batches_next_epoch"] == b["batches_next_epoch"]


def test_bucket_boundary_validation_errors_match():
    with pytest.raises(ValueError):
        RE.DynamicBatchSampler(None, 100, bucket_boundaries=[5, -1], lengths_list=[1, 2])
    with pytest.raises(ValueError):
        RE.DynamicBatchSampler(None, 100, bucket_boundaries=[5, 5], lengths_list=[1, 2])
    with pytest.raises(AssertionError):
        RE.DynamicBatchSampler(None, 100, bucket_boundaries=[9, 5], lengths_list=[1, 2])
    with pytest.raises(NotImplementedError):
        RE.DynamicBatchSampler(None, 100, 4, batch_ordering="sideways", lengths_list=[10, 20, 30])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_rank_sharding_partitions_the_batches_and_balances_the_steps(world):
    g = golden()
    s = g["cases"][0]["setting"]
    ref = [tuple(b) for b in make(g, s)]
    shards = [make(g, s, rank=r, world_size=world) for r in range(world)]
    n = len(shards[0])
    assert all(len(sh) == n for sh in shards) and n >= len(ref) // world - 1          # same step count on every rank
    seen = [tuple(b) for sh in shards for b in sh]
    assert len(set(seen)) == len(seen) and set(seen) <= set(ref)                      # disjoint batches of the reference list
    assert len(ref) - len(seen) < world                                              # at most one partial step dropped
    # a step's batches come from one bucket (or neighbouring ones at a bucket's tail): frames per rank within a few percent
    assert shards[0].step_imbalance() < 1.15
    per_rank = [sum(sh._frames(b) for b in sh) for sh in shards]
    assert max(per_rank) / (sum(per_rank) / world) < 1.25                            # heaviest-first dealing biases rank 0 only mildly
    # every rank sees the same permutation after set_epoch
    for sh in shards:
        sh.set_epoch(3)
    seen2 = [tuple(b) for sh in shards for b in sh]
    assert len(set(seen2)) == len(seen2)


def test_collate_raw_keys_and_lengths():
    ex = []
    for i, silent in enumerate([False, True]):
        ex.append({"silent": silent, "audio_features": np.zeros((5 + i, 26)), "parallel_voiced_audio_features": np.zeros((7, 26)),
                   "parallel_voiced_emg": np.zeros((9, 112)), "phonemes": np.arange(4), "phonemes_int": np.arange(3 + i),
                   "emg": np.zeros((10 + i, 112)), "raw_emg": np.zeros((80 + 8 * i, 8)), "session_ids": np.zeros(10 + i),
                   "text": "t%d" % i, "text_int": np.arange(6 + i)})
    out = RE.collate_raw(ex)
    assert sorted(out) == sorted(["audio_features", "audio_feature_lengths", "emg", "raw_emg", "parallel_voiced_emg", "phonemes",
                                  "phonemes_int", "phonemes_int_lengths", "session_ids", "lengths", "silent", "text", "text_int",
                                  "text_int_lengths"])
    assert out["lengths"] == [10, 11] and out["phonemes_int_lengths"] == [3, 4] and out["text_int_lengths"] == [6, 7]
    assert out["audio_feature_lengths"] == [5, 7] and out["parallel_voiced_emg"][0].shape == (1,) and out["parallel_voiced_emg"][1].shape == (9, 112)

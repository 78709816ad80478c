"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/sst.h declares, the host half of the
training step (batch preparation, warm-up schedule, flat parameter order, gradient buckets) and the data-parallel
gradient exchange on a world_size-2 gloo group."""
import os
import re
import sys

import pytest
import torch

import sst_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import sst_b200  # noqa: E402,F401
from sst_b200 import lib as L  # noqa: E402
from sst_b200 import train as T  # noqa: E402
from sst_b200.synthetic import make_batch, lognormal_lengths  # noqa: E402
from sst_b200.data_utils import combine_fixed_length, ChunkStager  # noqa: E402


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sst.h")).read()
    return re.findall(r"^(?:int|long long|size_t|const char\*) (sst_\w+)\(", src, re.M)


def test_cabi_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 24
    lib = L.lib()                       # raises SstError when libsst.so has not been built: there is no fallback
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.sst_version()
    assert lib.sst_launch_count() == 0  # nothing was launched by loading


def test_comm_entry_points_load_nccl_and_report_errors():
    """sst_comm_* (include/sst.h): NCCL is dlopen'ed at the first call, from the library PyTorch itself uses; no GPU is needed to
    load it and ask for its version; a path that does not exist must come back as SST_E_COMM with a message, not crash."""
    import ctypes as C
    v = L.comm_nccl_version()
    assert v >= 21000, v                                         # NCCL >= 2.10 (ncclAvg): 2.28.9 reports 22809
    assert L.lib().sst_comm_allreduce_bucket(None, None, C.c_int64(0), 0, 1, None) == -1        # SST_E_ARG: null communicator
    assert b"bad arguments" in L.lib().sst_last_error()
    assert L.lib().sst_comm_destroy(None) == 0


def test_missing_device_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.SstError):
        L.require_device()


def test_prepare_batch_matches_reference_step_arithmetic():
    """train.prepare_batch == recognition_model.py:77,85-87,95-97 as restated by the oracle."""
    batch = O.synthetic_batch(seed=7, ragged=[70, 100, 29], tgt_lens=[9, 14, 5])
    host = T.prepare_batch(batch)
    X = O.combine_fixed_length(batch["raw_emg"])
    tgt_in, tgt_out, ctc_tgt, ctc_lens = O.make_targets(batch)
    assert torch.equal(host["X"], X) and host["X"].shape == (1, 1600, 8)
    assert float(host["X"][0, -1, 0]) == 42.0               # tail padded with the VALUE 42 (data_utils.py:170)
    assert torch.equal(host["tgt_in"], tgt_in) and torch.equal(host["tgt_out"], tgt_out.reshape(-1))
    assert torch.equal(host["ctc_tgt"], ctc_tgt) and host["ctc_lens"].tolist() == ctc_lens
    assert host["n_valid"] == int((tgt_out != 42).sum())
    assert host["lengths"] == [70, 100, 29]
    # decoder key lengths: number of non-pad tokens in tgt_in
    assert host["tgt_lens"].tolist() == (tgt_in != 42).sum(1).tolist()


def test_chunk_stager_packs_like_the_reference_packer():
    """sst_b200.data_utils.combine_fixed_length / ChunkStager against the oracle's restatement of data_utils.py:165-174 (itself
    pinned to the unmodified reference): exact multiple of the chunk (no tail), ragged tail filled with the VALUE 42, a single
    short utterance, and buffer reuse by a second, smaller batch."""
    g = torch.Generator().manual_seed(0)
    stager = ChunkStager(pin=False)
    for sizes in ((800, 1234, 66), (1600, 1600), (5,), (3200, 1)):
        parts = [torch.randn(n, 8, generator=g) for n in sizes]
        want = O.combine_fixed_length(parts, 1600)
        for got in (combine_fixed_length(parts, 1600), combine_fixed_length(parts, 1600, stager=stager)):
            assert got.shape == want.shape and got.is_contiguous() and torch.equal(got, want)
    flat = combine_fixed_length([torch.zeros(2100, 8)], 1600).view(-1, 8)
    assert bool((flat[:2100] == 0).all()) and bool((flat[2100:] == 42).all())


def test_synthetic_batch_contract():
    b = make_batch(n_utt=3, frames=50, tgt_min=5, tgt_max=9, seed=1)
    assert set(b) == {"raw_emg", "lengths", "phonemes_int", "phonemes_int_lengths"}
    assert [t.shape for t in b["raw_emg"]] == [(400, 8)] * 3
    for p, n in zip(b["phonemes_int"], b["phonemes_int_lengths"]):
        assert int(p[0]) == 41 and int(p[-1]) == 40 and p.numel() == n and int(p[1:-1].max()) < 40
    lens = lognormal_lengths(1000, seed=3)
    assert min(lens) >= 100 and max(lens) <= 1500 and 380 < sum(lens) / 1000 < 620


def test_lr_warmup_schedule_matches_reference():
    class Dummy(T.Trainer):
        def __init__(self):
            self.lr_target, self.warmup, self.lr = 3e-4, 1500, 3e-4
    tr = Dummy()
    for it in (0, 1, 10, 1499, 1500, 5000):
        tr.schedule_lr(it)
        ref = O.lr_schedule(it)
        if ref is not None:
            assert tr.lr == pytest.approx(ref, rel=1e-12)
    assert tr.lr == pytest.approx(3e-4)


def _names(n_enc, n_dec):
    cfg = O.make_cfg(n_enc=n_enc, n_dec=n_dec)
    sd = O.synthetic_state_dict(dict(cfg, d_model=64, d_ff=128, n_heads=2), 0)
    return O.trainable_names(sd, cfg), sd


def test_backward_order_follows_gradient_completion():
    names, _ = _names(2, 2)
    order = T.backward_order(names, 2, 2)
    assert sorted(order) == sorted(names)
    stages = [T.stage_of(n, 2, 2) for n in order]
    seq = ["heads", "dec1", "dec0", "embed", "enc1", "enc0", "w_raw_in", "conv2", "conv1", "conv0"]
    idx = [seq.index(s) for s in stages]
    assert idx == sorted(idx), "flat buffer must be laid out in backward-completion order"
    assert order[0].startswith("w_aux") and order[-1].startswith("conv_blocks.0")


class _FakeFlat:
    def __init__(self, names, sd):
        self.names = names
        self.offsets, off = {}, 0
        for n in names:
            self.offsets[n] = off
            off += (sd[n].numel() + 3) // 4 * 4
        self.numel = off
        self.g = torch.zeros(off)


def test_gradient_buckets_partition_the_flat_buffer():
    names, sd = _names(3, 1)
    flat = _FakeFlat(T.backward_order(names, 3, 1), sd)
    sync = T.GradSync(flat, 3, 1, bucket_bytes=64 << 10)
    assert sync.buckets[0][0] == 0 and sync.buckets[-1][1] == flat.numel
    for (s0, e0, st0), (s1, e1, st1) in zip(sync.buckets, sync.buckets[1:]):
        assert e0 == s1 and e0 > s0
        assert sync.stage_order.index(st0) <= sync.stage_order.index(st1)
    assert len(sync.buckets) > 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, sd = _names(2, 1)
    flat = _FakeFlat(T.backward_order(names, 2, 1), sd)
    sync = T.GradSync(flat, 2, 1, bucket_bytes=32 << 10)
    g = torch.Generator().manual_seed(100 + rank)
    flat.g.copy_(torch.randn(flat.numel, generator=g))
    mine = flat.g.clone()
    sync.begin()
    launched = []
    for st in sync.stage_order:                 # Engine.backward announces the stages in this order
        sync.on_stage(st)
        launched.append(sync._next)
    sync.finish()
    others = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(others, mine)
    want = sum(others) / world
    q.put((rank, float((flat.g - want).abs().max()), launched, len(sync.buckets)))
    dist.destroy_process_group()


def test_gradsync_world2_gloo_averages_every_bucket_once():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, launched, nb in res:
        assert err < 1e-6, "rank %d: reduced gradient differs from the mean over ranks" % rank
        assert launched == sorted(launched) and launched[-1] == nb   # buckets fire in order, all of them by the last stage


def _gate_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = [[3, 4, 5, 5, 1, 6, 2, 9], [1, 1, 7, 2, 2, 2, 8, 1]][rank]       # chunks per micro-batch: different on every rank
    gate = T.AccumulationGate(batch_size_grad=10, distributed=True)
    decisions = [gate.add(c) for c in counts]
    q.put((rank, decisions, gate.sum_batch_size))
    dist.destroy_process_group()


def test_accumulation_gate_world2_gloo_decides_identically_on_every_rank():
    """with batch_size_grad > 1 and unequal per-rank chunk counts, a rank that thresholds its LOCAL count
    enters the all-reduce alone.  The gate thresholds the sum over ranks: same decision everywhere, and equal to the
    single-process reference rule (recognition_model.py:81,115-118) applied to the global batch."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_gate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    glob = [a + b for a, b in zip([3, 4, 5, 5, 1, 6, 2, 9], [1, 1, 7, 2, 2, 2, 8, 1])]
    want, acc = [], 0
    for g in glob:
        acc += g
        want.append(acc >= 10)
        if acc >= 10:
            acc = 0
    assert res[0][1] == want and res[1][1] == want
    assert res[0][2] == res[1][2] == acc
    # a local-count rule would have disagreed on this input: rank 0 alone reaches 10 after its third micro-batch
    local0 = [sum([3, 4, 5, 5, 1, 6, 2, 9][:k + 1]) >= 10 for k in range(3)]
    local1 = [sum([1, 1, 7, 2, 2, 2, 8, 1][:k + 1]) >= 10 for k in range(3)]
    assert local0 != local1


def test_accumulation_gate_single_process_matches_the_reference_rule():
    gate = T.AccumulationGate(batch_size_grad=100)
    out = [gate.add(c) for c in [40, 40, 40, 10, 95, 5, 100]]
    assert out == [False, False, True, False, True, False, True]
    assert [T.AccumulationGate(1).add(c) for c in (1, 320)] == [True, True]
    g = T.AccumulationGate(10)
    assert g.add(3, global_chunks=12) is True          # the caller supplied the sum over ranks

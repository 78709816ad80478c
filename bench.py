#!/usr/bin/env python
"""Headline benchmark: train EMG frames/s (forward + hybrid CTC/attention loss + backward + AdamW) of the Silent Speech
Transformer hot path on B200, through libsst.so.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg1]

Workload (BASELINE.json configs[1], "cfg2"): 6 encoder layers (d 768, 8 heads, FFN 3072, rel-pos distance 100) + the
reference's default 6-layer decoder, bf16 compute, 64 utterances x 8000 raw EMG samples per GPU (-> 320 chunks of 1600,
64 x 1000 encoder frames), targets of 80-120 phones, hybrid loss (alpha 0.2, label smoothing 0.1), dropout 0.2, one AdamW
step per batch.  A "step" is one such batch; `value` = encoder frames of all ranks / device time with the batches
resident in HBM; `e2e` = the same step driven from pinned HOST buffers (H2D of the batch and D2H of the losses inside
the timed region).  N > 1 (torchrun): weak scaling, one process per GPU, bucketed NCCL all-reduce overlapped with backward.

`--impl reference` times the reference's CPU implementation of the same step (the oracle port of the PyTorch-CPU
reference: /root/reference does not exist on the GPU box) on the host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train EMG frames/s (fwd+bwd+loss+AdamW)"
UNIT = "frames/s"

WORKLOADS = {
    # name: (n_enc, n_dec, alpha, utterances per GPU, frames per utterance)
    "cfg1": dict(n_enc=6, n_dec=0, alpha=0.2, n_utt=4, frames=200, tgt=(30, 30)),
    "cfg2": dict(n_enc=6, n_dec=6, alpha=0.2, n_utt=64, frames=1000, tgt=(80, 120)),
    "cfg3": dict(n_enc=8, n_dec=4, alpha=0.7, n_utt=64, frames=1000, tgt=(80, 120)),
}


def workload_name(w):
    return ("%s: %d enc + %d dec layers, d768 h8 ffn3072 R100, %d utt x %d samples/GPU, hybrid CTC/label-smoothed CE alpha %.1f, "
            "dropout 0.2, AdamW every step" % (w["name"], w["n_enc"], w["n_dec"], w["n_utt"], w["frames"] * 8, w["alpha"]))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tensor_burst=d["bf16_tflops"],
                    source="measured (MEASURED_PEAKS.json; tensor = sustained cuBLAS bf16)")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's PyTorch-CPU step on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_step_rate(w, steps, warmup, n_utt_sample=2):
    """Times the CPU oracle (oracle/sst_oracle.py, the restated reference step: forward_training + CTC + LabelSmoothingLoss +
    backward + AdamW) on `n_utt_sample` utterances of the workload's length.  Returns (frames/s, ms/step, cores, sample str)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sst_oracle as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg = O.make_cfg(n_enc=w["n_enc"], n_dec=w["n_dec"], rel_dist=100, alpha=w["alpha"], dropout=0.2, dropout_pos=0.2)
    sd = O.synthetic_state_dict(cfg, 0)
    names = O.trainable_names(sd, cfg)
    for n in names:
        sd[n].requires_grad_(True)
    m = {n: torch.zeros_like(sd[n]) for n in names}
    v = {n: torch.zeros_like(sd[n]) for n in names}
    frames = w["frames"]
    tl = (w["tgt"][0] + w["tgt"][1]) // 2
    times = []
    for it in range(warmup + steps):
        batch = O.synthetic_batch(n_utt=n_utt_sample, frames=frames, tgt_len=tl, seed=100 + it)
        t0 = time.perf_counter()
        res = O.train_step_losses(sd, cfg, batch, training=True, shift_r=it % 8)
        res["loss"].backward()
        with torch.no_grad():
            for n in names:
                if sd[n].grad is not None:
                    O.adamw_step(sd[n], sd[n].grad, m[n], v[n], it + 1, 3e-4 * (it + 1) / 1500)
                    sd[n].grad = None
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    fps = n_utt_sample * frames / (ms / 1e3)
    sample = "%d utterances x %d samples (same utterance length and model as the workload), %d warm-up + %d timed steps" % (
        n_utt_sample, frames * 8, warmup, steps)
    return fps, ms, cores, sample


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    fps, ms, cores, sample = cpu_reference_step_rate(w, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(w),
                       "note": "reference CPU path (PyTorch fp32) on the host cores.  /root/reference does not exist on the GPU box, so this is "
                               "the oracle port of it (oracle/sst_oracle.py, pinned to the unmodified reference by tests/golden); each step is a "
                               "2-utterance sample of the workload's 64 (same model, same 8000-sample utterances): the full batch takes ~100 s "
                               "per step on these cores and needs ~60 GB for the (B,H,L,L) tensors the reference materialises"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch.distributed as dist
    import sst_b200  # noqa: F401
    from sst_b200 import lib as L
    from sst_b200 import architecture as A
    from sst_b200.synthetic import make_batch
    from sst_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L.require_device()

    A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=w["n_enc"], num_layers_decoder=w["n_dec"],
                n_heads_encoder=8, n_heads_decoder=8, relative_distance=100, dropout_model=0.2, dropout_pos_emb=0.2,
                sst_dtype=args.dtype)
    torch.manual_seed(0)                                   # identical initial weights on every rank
    model = A.Model(112, 44, 43, dev).to(dev)
    trainer = Trainer(model, alpha_loss=w["alpha"], batch_size_grad=1, seed=rank, distributed=world > 1)

    n_batches = 2
    host = [trainer.prepare(make_batch(w["n_utt"], w["frames"], w["tgt"][0], w["tgt"][1], seed=1234 + rank * 100 + i))
            for i in range(n_batches)]
    resident = [trainer.to_device(h) for h in host]
    pristine = [d["X"].clone() for d in resident]
    frames_per_step = w["n_utt"] * w["frames"]
    h2d = resident[0]["h2d_bytes"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return ms

    chunks_all_ranks = world * resident[0]["X"].shape[0]   # every rank holds an equally sized synthetic batch: no exchange needed

    def resident_step(i):
        d = resident[i % n_batches]
        d["X"].copy_(pristine[i % n_batches])              # the training-time shift mutates x_raw in place (architecture.py:104-108)
        return trainer.step_device(d, global_chunks=chunks_all_ranks)

    def e2e_step(i):
        d = trainer.to_device(host[i % n_batches])         # pinned host -> device, async on the compute stream
        losses = trainer.step_device(d, global_chunks=chunks_all_ranks)
        trainer.fetch_losses(losses)
        return trainer.wait_losses()                       # device -> host read of the step's result

    # ---- device-resident throughput -----------------------------------------------------------------------------------
    for i in range(args.warmup):
        resident_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        losses = resident_step(i)
    e1.record()
    barrier()
    launches = L.launch_count() - n0
    ms_res = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    loss_vals = [float(x) for x in losses.cpu()]

    # ---- end to end from host buffers ---------------------------------------------------------------------------------
    for i in range(2):
        e2e_step(i)
    for _ in trainer.run((host[i % n_batches] for i in range(2)), global_chunks=chunks_all_ranks):      # warm the pipelined entry
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    # Trainer.run: every step uploads its batch from pinned host memory and reads its losses back to the host; the upload of
    # step i+1 and the read-back of step i overlap step i+1's compute
    n_read = sum(1 for _ in trainer.run((host[i % n_batches] for i in range(args.steps)), global_chunks=chunks_all_ranks))
    assert n_read == args.steps
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / args.steps

    # ---- per-kernel-family device times (CUDA events on the launching stream) -----------------------------------------
    with L.Profiler() as prof:
        resident_step(0)
        fam = prof.summary()
        shapes = prof.summary(by_tag=True)
    tot_ms = sum(d["ms"] for d in fam.values())
    pk = peaks()
    kernels = {}
    for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(d["ms"], 3), "share": round(d["ms"] / tot_ms, 4), "launches": d["launches"]}
        if d["flops"] > 0:
            e["tflops"] = round(d["flops"] / d["ms"] / 1e9, 1)
        if d["bytes"] > 0:
            e["gbs"] = round(d["bytes"] / d["ms"] / 1e6, 1)
        kernels[k] = e
    gemm_shapes = [{"shape": k, "ms": round(d["ms"], 3), "n": d["launches"], "tflops": round(d["flops"] / d["ms"] / 1e9, 1)}
                   for k, d in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]) if k.startswith("gemm")][:args.gemm_shapes]
    other_shapes = [{"shape": k, "ms": round(d["ms"], 3), "n": d["launches"], "gbs": round(d["bytes"] / d["ms"] / 1e6, 1)}
                    for k, d in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]) if not k.startswith("gemm")][:args.gemm_shapes]
    dom = max(fam.items(), key=lambda kv: kv[1]["ms"])
    dk, dd = dom
    if dd["flops"] > 0:
        ach = dd["flops"] / dd["ms"] / 1e9
        roof = {"kernel": dk, "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tensor"], "unit": "TFLOP/s",
                "frac": round(ach / pk["tensor"], 4), "traffic": None, "launches_per_step": dd["launches"],
                "avg_launch_ms": round(dd["ms"] / dd["launches"], 4), "peak_source": pk["source"]}
    else:
        ach = dd["bytes"] / dd["ms"] / 1e6
        roof = {"kernel": dk, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s",
                "frac": round(ach / pk["hbm"], 4), "traffic": None, "launches_per_step": dd["launches"],
                "avg_launch_ms": round(dd["ms"] / dd["launches"], 4), "peak_source": pk["source"]}
    # DRAM traffic of the dominant kernel family, per launch: from the committed ncu capture of one step of the same workload
    # (profiles/r02_kernel_traffic.json, tools/traffic_step.py + tools/kernel_traffic.py: dram__bytes_read.sum + dram__bytes_write.sum of
    # every launch, grouped per family); null when the capture does not cover this workload / family
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")))
        ent = tr.get(w["name"], {}).get(dk)
        if ent:
            roof["traffic"] = ent["avg_bytes_per_launch"]
            roof["traffic_note"] = ("ncu dram bytes per launch, family average over the %d launches of one step (this run: %d); algorithmic "
                                    "operand + result bytes per launch here: %d" % (ent["launches_per_step"], dd["launches"],
                                                                                    int(dd["bytes"] / dd["launches"])))
    except (OSError, ValueError, KeyError):
        pass
    # whole-step tensor roofline: algorithmic flops of every GEMM/attention launch / step time
    step_flops = sum(d["flops"] for d in fam.values())

    also = None
    if not args.no_also:
        also = run_also(args, w, world, rank, dev, trainer, model, barrier, max_over_ranks)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fps, ms, cores, sample = cpu_reference_step_rate(w, 2, 1)
            cpu = {"value": round(fps, 1), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": round(ms, 1)}
        line = {
            "metric": METRIC, "value": round(world * frames_per_step / (ms_res / 1e3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(w), "frames_per_step_per_gpu": frames_per_step,
                       "l2": "no flush needed: a step streams >10 GB of activations, far above the 126 MB L2",
                       "parallelism": "dp%d" % world, "loss_last_step": loss_vals},
            "e2e": {"value": round(world * frames_per_step / (ms_e2e / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 12, "ms_per_step": round(ms_e2e, 3)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "step_tensor_roofline": {"algorithmic_tflop_per_step": round(step_flops / 1e12, 3),
                                     "achieved_tflops": round(step_flops / ms_res / 1e9, 1), "peak": pk["tensor"],
                                     "frac": round(step_flops / ms_res / 1e9 / pk["tensor"], 4)},
            "kernels": kernels, "cpu_baseline": cpu, "also": also,
            **({"gemm_shapes": gemm_shapes, "other_shapes": other_shapes} if gemm_shapes else {}),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, timed after the headline (they never enter `value` / `ms_per_step`)
# ---------------------------------------------------------------------------------------------------------------------
def run_also(args, w, world, rank, dev, trainer, model, barrier, max_over_ranks):
    import random
    import sst_b200  # noqa: F401
    from sst_b200 import lib as L
    from sst_b200 import architecture as A
    from sst_b200.synthetic import make_batch, lognormal_lengths
    from sst_b200.train import Trainer, prepare_batch
    from sst_b200.greedy_search import greedy_ids, greedy_ids_cached
    out = {}

    def device_ms(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as ex:                                    # a failing extra must not cost the headline line
            out[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        torch.cuda.synchronize()

    # ---- cfg4 (BASELINE.json configs[3]): 12-step gradient accumulation, batch 200 per GPU -------------------------------------
    def cfg4():
        # reading: batch_size_grad = 200 chunks of 1600 samples PER GPU, reached over 12 micro-batches of 17 chunks (4 utterances x
        # 6800 samples); gradients are summed (recognition_model.py:81,115-118), the all-reduce + AdamW run on the 12th only.
        # No conformer code exists in the reference: the model is the headline workload's.
        n_micro, utt, frames = 12, 4, 850
        hb = [trainer.to_device(trainer.prepare(make_batch(utt, frames, 60, 90, seed=7000 + rank * 100 + i))) for i in range(n_micro)]
        pristine = [d["X"].clone() for d in hb]
        chunks = hb[0]["X"].shape[0]
        trainer.batch_size_grad = 200 * world                     # the gate thresholds the sum over ranks
        trainer.start_epoch()
        steps_before = trainer.flat.step_count

        def cycle(_):
            for i, d in enumerate(hb):
                d["X"].copy_(pristine[i])
                trainer.step_device(d, global_chunks=world * chunks)

        def cycle_graphed(_):
            for i, d in enumerate(hb):
                d["X"].copy_(pristine[i])
                trainer.step_graphed(d, global_chunks=world * chunks)
        cycle(0)
        ms = device_ms(cycle, 2)
        n_opt = trainer.flat.step_count - steps_before
        ms_graphed = None
        if world == 1:                                             # CUDA-graph replay of the micro-steps (single process)
            for _ in range(3):
                cycle_graphed(0)                                   # two eager passes per signature, then the captures
            ms_graphed = device_ms(cycle_graphed, 3)
        trainer.batch_size_grad = 1
        trainer.start_epoch()
        return {"reading": "batch_size_grad = 200 chunks x 1600 samples per GPU reached in 12 micro-batches of %d chunks (%d utt x %d "
                           "samples); summed gradients; all-reduce + AdamW on the 12th micro-step only; model = headline workload "
                           "(no conformer code in the reference)" % (chunks, utt, frames * 8),
                "ms_per_cycle": round(ms, 3), "optimizer_steps": n_opt, "cycles_run": 3,
                "frames_per_s": round(world * n_micro * utt * frames / (ms / 1e3), 1),
                "cuda_graphs": None if ms_graphed is None else {
                    "ms_per_cycle": round(ms_graphed, 3), "frames_per_s": round(n_micro * utt * frames / (ms_graphed / 1e3), 1),
                    "what": "Trainer.step_graphed: each of the 12 micro-batch signatures replayed as one CUDA graph (dropout salt, "
                            "learning rate and Adam bias corrections through device memory).  Small micro-batches are bound by per-kernel "
                            "floors on the device (~11 us per GEMM launch), not by the host's launch rate, so the gain is a few %"}}
    guarded("cfg4_accumulation", cfg4)

    # ---- cfg1 (BASELINE.json configs[0]): the reference's CPU-runnable batch, 4 utterances x 200 frames -- launch-bound on a B200 ---
    def cfg1():
        if world > 1:
            return {"note": "single-process measurement"}
        hb = [trainer.to_device(trainer.prepare(make_batch(4, 200, 28, 32, seed=7600 + i))) for i in range(2)]
        hb = [dict(d) for d in hb]
        res = {}
        for name, fn in (("eager", trainer.step_device), ("cuda_graph", trainer.step_graphed)):
            for i in range(6):
                fn(hb[i % 2], global_chunks=hb[i % 2]["X"].shape[0])
            ms = device_ms(lambda i: fn(hb[i % 2], global_chunks=hb[i % 2]["X"].shape[0]), 20)
            res["ms_per_step_" + name] = round(ms, 3)
        res["frames_per_s_cuda_graph"] = round(800 / (res["ms_per_step_cuda_graph"] / 1e3), 1)
        res["what"] = "headline model, 4 utterances x 1600 samples per step (SURVEY.md 8(d) cfg1 shape), fwd+bwd+loss+AdamW every step"
        return res
    guarded("cfg1_small_batch", cfg1)

    # ---- N1: reference-semantics greedy attention decoding (greedy_search.py:7-53), latency per utterance ------------------------
    def greedy():
        res = {}
        model.eval()
        for B in (1, 16):
            b = make_batch(B, 1000, 10, 20, seed=8000 + rank)
            X = prepare_batch(b)["X"].to(dev)
            max_len = 64
            for name, fn in (("kv_cached", greedy_ids_cached), ("prefix_rerun", greedy_ids)):
                fn(model, b["lengths"], X.clone(), max_len, dev)                 # warm
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ids = fn(model, b["lengths"], X.clone(), max_len, dev)            # returns on the host: includes every sync of the search
                dt = time.perf_counter() - t0
                res["B%d_%s" % (B, name)] = {"ms_per_utterance": round(1e3 * dt / B, 3), "decoder_steps": int(ids.shape[1]) - 1,
                                            "ms_total": round(1e3 * dt, 2)}
        model.train()
        res["note"] = ("encoder forward (1000 frames) + attention-decoder search up to 64 tokens, random-init weights (no </S> before the "
                       "limit), wall clock incl. host synchronisations, rank 0")
        return res
    guarded("greedy_search_latency", greedy)

    # ---- cfg5 (BASELINE.json configs[4]): inference over a 1000-utterance testset_largedev-shaped set ---------------------------
    def cfg5():
        n_utt = 1000
        lens = lognormal_lengths(n_utt, seed=5)
        order = sorted(range(n_utt), key=lambda i: lens[i])
        mine = order[rank::world]                                  # length-sorted round robin: equal work per rank
        batches, cur, tot = [], [], 0
        for i in mine:                                             # <= 64 000 frames per batch (the headline batch size)
            if cur and tot + lens[i] > 64000:
                batches.append(cur); cur, tot = [], 0
            cur.append(i); tot += lens[i]
        if cur:
            batches.append(cur)
        host = []
        for bi, idx in enumerate(batches):
            b = make_batch(lengths=[lens[i] for i in idx], tgt_min=5, tgt_max=10, seed=9000 + bi)
            host.append((prepare_batch(b)["X"].pin_memory(), b["lengths"]))
        model.eval()
        eng = model._packed_engine()
        resident = [(X.to(dev), l) for X, l in host]
        frames_mine = sum(lens[i] for i in mine)

        def run_resident(_):
            for X, l in resident:
                eng.ctc_greedy(X, l)

        def run_e2e(_):
            outs = []
            for X, l in host:
                ids, ln = eng.ctc_greedy(X.to(dev, non_blocking=True), l)
                outs.append((ids.to("cpu", non_blocking=True), ln.to("cpu", non_blocking=True)))
            torch.cuda.synchronize()
            return outs
        run_resident(0)
        ms = device_ms(run_resident, 2)
        eng.cfg["packed"] = False                                  # the reference's pad-to-max layout, for comparison
        run_resident(0)
        ms_padded = device_ms(run_resident, 2)
        eng.cfg["packed"] = True
        run_e2e(0)
        barrier()
        t0 = time.perf_counter()
        run_e2e(0)
        barrier()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
        model.train()
        tot_frames = sum(lens)
        return {"utterances": n_utt, "frames": tot_frames, "batches_per_rank": len(batches), "padded_frames_this_rank":
                sum(len(idx) * max(lens[i] for i in idx) for idx in batches), "frames_this_rank": frames_mine,
                "ms": round(ms, 3), "utterances_per_s": round(n_utt / (ms / 1e3), 1), "frames_per_s": round(tot_frames / (ms / 1e3), 1),
                "ms_padded_layout": round(ms_padded, 3), "layout": "packed (sum(lengths) rows, attention by row offsets; SURVEY.md 8(f) N2)",
                "e2e_ms": round(ms_e2e, 3), "e2e_utterances_per_s": round(n_utt / (ms_e2e / 1e3), 1),
                "what": "eval-mode encoder forward (BN running statistics) + w_aux + CTC best-path decode (sst_ctc_greedy), lengths ~ "
                        "clip(lognormal(450, 0.5), 100, 1500) frames, length-sorted and dealt round-robin to the ranks, <= 64 000 frames "
                        "per batch; e2e = from pinned host buffers incl. the D2H of the decoded ids"}
    guarded("cfg5_inference", cfg5)

    # ---- N2: a ragged TRAINING batch (lognormal lengths, one length bucket would be narrower), packed vs pad-to-max ---------------
    def ragged_training():
        lens = sorted(lognormal_lengths(400, seed=11 + rank))
        lens = lens[100:100 + 110]                                 # a contiguous slice of the length-sorted corpus, ~ one batch
        while sum(lens) > 64000:
            lens.pop()
        b = make_batch(lengths=lens, tgt_min=20, tgt_max=60, seed=9500 + rank)
        d = trainer.to_device(trainer.prepare(b))
        pristine = d["X"].clone()
        chunks = world * d["X"].shape[0]
        res = {"utterances": len(lens), "frames": sum(lens), "padded_frames": len(lens) * max(lens)}

        def step(_):
            d["X"].copy_(pristine)
            trainer.step_device(d, global_chunks=chunks)
        for packed in (True, False):
            trainer.eng.cfg["packed"] = packed
            step(0)
            step(0)
            ms = device_ms(step, 4)
            res["ms_packed" if packed else "ms_padded"] = round(ms, 3)
        trainer.eng.cfg["packed"] = True
        res["frames_per_s_packed"] = round(world * sum(lens) / (res["ms_packed"] / 1e3), 1)
        res["what"] = "one training step (fwd+bwd+loss+AdamW) on a ragged batch per rank; real frames / time"
        return res
    guarded("ragged_training", ragged_training)

    # ---- N4: one beam-search decoder step (BeamSearch.py:111-114), 100 hypotheses of 30 tokens over a 1000-frame memory ----------
    def beam_step():
        from sst_b200.beam_search import BeamDecoder
        model.eval()
        b = make_batch(1, 1000, 10, 20, seed=8100 + rank)
        X = prepare_batch(b)["X"].to(dev)
        n_hyp, t = 100, 30
        hist = torch.randint(0, 40, (n_hyp, t), device=dev, dtype=torch.int64)
        hist[:, 0] = 41
        with torch.no_grad():
            memory, _ = model(b["lengths"], dev, mode='beam_search', part='encoder', x_raw=X)
            dec = BeamDecoder(model, memory)

            def shared(_):
                dec.step_logits(hist)

            def repeat(_):
                model(b["lengths"], dev, mode='beam_search', part='decoder', y=hist, memory=memory.repeat(n_hyp, 1, 1))[:, -1, :-2]
            shared(0); repeat(0)
            ms_s, ms_r = device_ms(shared, 5), device_ms(repeat, 5)
        model.train()
        return {"hypotheses": n_hyp, "prefix_tokens": t, "memory_frames": 1000, "ms_shared_memory": round(ms_s, 3),
                "ms_memory_repeat": round(ms_r, 3), "what": "BeamDecoder (memory projected once, k_off = 0) vs the reference call "
                "pattern memory.repeat(n_hyp, 1, 1) through Model.forward(part='decoder'); identical logits"}
    guarded("beam_decoder_step", beam_step)

    # ---- cfg3 (BASELINE.json configs[2]): 8 encoder + 4 decoder layers, alpha 0.7 -----------------------------------------------
    def cfg3():
        if w["name"] == "cfg3":
            return {"note": "headline workload"}
        w3 = dict(WORKLOADS["cfg3"], name="cfg3")
        A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=w3["n_enc"], num_layers_decoder=w3["n_dec"],
                    n_heads_encoder=8, n_heads_decoder=8, relative_distance=100, dropout_model=0.2, dropout_pos_emb=0.2, sst_dtype=args.dtype)
        torch.manual_seed(0)
        m3 = A.Model(112, 44, 43, dev).to(dev)
        t3 = Trainer(m3, alpha_loss=w3["alpha"], batch_size_grad=1, seed=rank, distributed=world > 1)
        hb = [t3.to_device(t3.prepare(make_batch(w3["n_utt"], w3["frames"], w3["tgt"][0], w3["tgt"][1], seed=4321 + rank * 100 + i)))
              for i in range(2)]
        pristine = [d["X"].clone() for d in hb]
        chunks = world * hb[0]["X"].shape[0]

        def step(i):
            d = hb[i % 2]
            d["X"].copy_(pristine[i % 2])
            t3.step_device(d, global_chunks=chunks)
        for i in range(3):
            step(i)
        ms = device_ms(step, 5)
        return {"workload": workload_name(w3), "ms_per_step": round(ms, 3),
                "frames_per_s": round(world * w3["n_utt"] * w3["frames"] / (ms / 1e3), 1), "steps": 5, "warmup": 3}
    guarded("cfg3_hybrid", cfg3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE.json configurations (cfg3 / cfg4 / cfg5 / greedy latency)")
    ap.add_argument("--gemm-shapes", type=int, default=0, help="also list the N most expensive GEMM shapes of a step")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload], name=args.workload)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, w)
        return
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    args.warmup = max(args.warmup, 3)
    run_ours(args, w)


if __name__ == "__main__":
    main()

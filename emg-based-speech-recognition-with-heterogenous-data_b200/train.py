"""Training step of the reference loop (recognition_model.py:57-64, :76-118, :283-293) on the B200 path.

  * `FlatState`   : every trainable parameter re-homed as a view into ONE fp32 buffer laid out in backward-completion
                    order (heads, decoder N..0, embedding, encoder N..0, w_raw_in, conv 2..0), with matching flat
                    gradient / Adam-moment buffers.  Parameters that never receive a gradient in the reference
                    (rel-pos embeddings, emg_projection: SURVEY.md Q2/Q14) stay outside and are skipped by AdamW exactly
                    as torch.optim.AdamW skips `grad is None`.
  * `GradSync`    : data parallelism, one process per GPU: the flat gradient buffer is cut into ~25 MB buckets that are
                    all-reduced (mean) on a side stream as soon as backward has produced them (NCCL over NVLink; gloo on
                    CPU for the host-logic tests); replaces nn.DataParallel (recognition_model.py:284).
  * `Trainer.step`: schedule_lr -> combine_fixed_length -> H2D -> forward -> CTC + label-smoothed CE -> backward ->
                    summed gradient accumulation until `batch_size_grad` chunks -> fused AdamW.  The three `.item()`
                    syncs of recognition_model.py:108-111 become one async D2H copy of a float[3].
"""
import torch

from . import lib as L
from .data_utils import ChunkStager, combine_fixed_length

PAD = 42


def backward_order(names, n_enc, n_dec):
    """Parameter names sorted by the moment their gradient is complete during Engine.backward."""
    def key(n):
        if n.startswith("w_aux"):
            return (0, 0)
        if n.startswith("w_out"):
            return (1, 0)
        if n.startswith("transformerDecoder.layers."):
            return (2, n_dec - int(n.split(".")[2]))
        if n.startswith("embedding_tgt"):
            return (3, 0)
        if n.startswith("transformerEncoder.layers."):
            return (4, n_enc - int(n.split(".")[2]))
        if n.startswith("w_raw_in"):
            return (5, 0)
        if n.startswith("conv_blocks."):
            return (6, 3 - int(n.split(".")[1]))
        return (7, 0)
    return sorted(names, key=lambda n: (key(n), n))


def stage_of(name, n_enc, n_dec):
    """Backward stage label after which `name`'s gradient is final (matches Engine.backward's on_stage calls)."""
    if name.startswith("w_aux") or name.startswith("w_out"):
        return "heads"
    if name.startswith("transformerDecoder.layers."):
        return "dec%d" % int(name.split(".")[2])
    if name.startswith("embedding_tgt"):
        return "embed"
    if name.startswith("transformerEncoder.layers."):
        return "enc%d" % int(name.split(".")[2])
    if name.startswith("w_raw_in"):
        return "w_raw_in"
    if name.startswith("conv_blocks."):
        return "conv%d" % int(name.split(".")[1])
    return "conv0"


def prepare_batch(example, stager=None):
    """The host half of one training step (recognition_model.py:77, :85-87, :95-97) on a `collate_raw` dict
    (read_emg.py:463-504): X = combine_fixed_length(raw_emg, 1600); decoder input/target = phonemes[:, :-1] / [:, 1:] padded
    with 42; CTC target = phonemes without <S>/</S>.  Returns plain host tensors plus the python-side scalars."""
    pad_seq = torch.nn.utils.rnn.pad_sequence
    X = combine_fixed_length(example['raw_emg'], 200 * 8, stager=stager)
    target = pad_seq(example['phonemes_int'], batch_first=True, padding_value=PAD)
    tgt_in = target[:, :-1].contiguous()
    tgt_out = target[:, 1:].contiguous()
    nmax = target.shape[1]
    ctc_lens = [n - 2 for n in example['phonemes_int_lengths']]
    ctc_tgt = pad_seq([p[1:-1] for p in example['phonemes_int']], batch_first=True, padding_value=PAD).contiguous()
    if ctc_tgt.shape[1] == 0:
        ctc_tgt = torch.full((len(ctc_lens), 1), PAD, dtype=torch.int64)
    host = dict(X=X, tgt_in=tgt_in, tgt_out=tgt_out.view(-1), ctc_tgt=ctc_tgt,
                ctc_lens=torch.tensor(ctc_lens, dtype=torch.int32),
                tgt_lens=torch.tensor([min(n, nmax - 1) for n in example['phonemes_int_lengths']], dtype=torch.int32))
    host['lengths'] = list(example['lengths'])
    host['n_valid'] = int(sum(min(n, nmax) - 1 for n in example['phonemes_int_lengths']))
    return host


class FlatState:
    def __init__(self, model):
        eng = model.engine()
        names = [n for n in model._param_names if model._trainable[n]]
        self.names = backward_order(names, eng.n_enc, eng.n_dec)
        params = dict(model.named_parameters())
        sizes = [(params[n].numel() + 7) // 8 * 8 for n in self.names]          # every view 16-byte aligned, also in the bf16 shadow
        self.offsets, off = {}, 0
        for n, s in zip(self.names, sizes):
            self.offsets[n] = off
            off += s
        self.numel = off
        dev = params[self.names[0]].device
        self.p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.g = torch.zeros(off, dtype=torch.float32, device=dev)
        self.m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.G = {}
        for n in self.names:
            prm = params[n]
            o, k = self.offsets[n], prm.numel()
            view = self.p[o:o + k].view(prm.shape)
            view.copy_(prm.data)
            prm.data = view
            self.G[n] = self.g[o:o + k].view(prm.shape)
            prm.grad = self.G[n]
        model._engine = None                      # parameter storage moved: rebuild the engine's views
        # bf16 shadow of the flat parameters, kept current by the fused AdamW kernel: the GEMM operands of every plain
        # (N_out, K_in) weight are views into it, so the per-step re-packing touches only the per-head / conv permutations
        self.pb = None
        self.shadow = {}
        if model.compute_dtype == torch.bfloat16:
            self.pb = self.p.to(torch.bfloat16)
            for n in self.names:
                o, k = self.offsets[n], params[n].numel()
                self.shadow[n] = self.pb[o:o + k].view(params[n].shape)
        self.step_count = 0
        # gradient sinks for the never-trained parameters (written by nobody, kept so Engine.backward can index them)
        self.G_all = dict(self.G)
        for n in model._param_names:
            if n not in self.G_all:
                self.G_all[n] = None

    def zero_grad(self):
        self.g.zero_()


class AccumulationGate:
    """When to step (recognition_model.py:81,115-118: accumulate micro-batch gradients until the number of 1600-sample chunks
    seen reaches batch_size_grad) -- decided IDENTICALLY on every rank.  Ranks hold different utterances, so their local
    chunk counts differ; a rank deciding from its own count would enter the gradient all-reduce while another keeps
    accumulating (hang, diverged weights).  The count that is thresholded is therefore the sum over ranks -- what
    `len(X)` was for the whole batch under the reference's nn.DataParallel -- exchanged through a host-side (gloo) group so
    that no device synchronisation enters the step.  `global_chunks` short-circuits the exchange when the caller already
    knows every rank's batch (a shared DynamicBatchSampler)."""

    def __init__(self, batch_size_grad, distributed=False):
        self.batch_size_grad = batch_size_grad
        self.sum_batch_size = 0
        self.group = None
        self.dist = None
        if distributed:
            import torch.distributed as dist
            self.dist = dist
            if dist.is_initialized() and dist.get_world_size() > 1:
                self.group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")

    def add(self, local_chunks, global_chunks=None):
        """Account one micro-batch; True when the optimizer must step after it (and the count restarts)."""
        if global_chunks is None:
            if self.group is not None:
                t = torch.tensor([int(local_chunks)], dtype=torch.int64)
                self.dist.all_reduce(t, group=self.group)
                global_chunks = int(t[0])
            else:
                global_chunks = int(local_chunks)
        self.sum_batch_size += global_chunks
        if self.sum_batch_size >= self.batch_size_grad:
            self.sum_batch_size = 0
            return True
        return False


class GradSync:
    def __init__(self, flat, n_enc, n_dec, bucket_bytes=128 << 20, group=None, native_comm=None):
        """native_comm: reduce the buckets through libsst.so's own communicator (sst_comm_*, include/sst.h) instead of
        torch.distributed.all_reduce; default from the environment variable SST_COMM=native.  torch.distributed stays the host
        channel (the unique id travels through it) and the fallback."""
        import os
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.flat = flat
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        if native_comm is None:
            native_comm = os.environ.get("SST_COMM", "") == "native"
        self.comm = None
        if native_comm and self.world > 1 and flat.g.is_cuda:
            def exchange(blob):
                box = [blob]
                dist.broadcast_object_list(box, src=0, group=group)
                return box[0]
            self.comm = L.Comm(dist.get_rank(group), self.world, exchange)
        # buckets = contiguous ranges of the flat gradient buffer, closed at stage boundaries once they reach bucket_bytes
        self.buckets = []            # (start, end, stage after which the bucket is complete)
        start, cur_stage = 0, None
        per = bucket_bytes // 4
        for n in flat.names:
            st = stage_of(n, n_enc, n_dec)
            off = flat.offsets[n]
            if cur_stage is not None and st != cur_stage and off - start >= per:
                self.buckets.append((start, off, cur_stage))
                start = off
            cur_stage = st
        self.buckets.append((start, flat.numel, cur_stage))
        self.stage_order = (["heads"] + ["dec%d" % i for i in reversed(range(n_dec))] + ["embed"] +
                            ["enc%d" % i for i in reversed(range(n_enc))] + ["w_raw_in", "conv2", "conv1", "conv0"])
        self.cuda = flat.g.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self._next = 0
        self._pending = []
        self.trace = False            # record a CUDA-event timeline of the next step (tools/ddp_timeline.py)
        self.timeline = None

    def broadcast_state(self, tensors):
        """Every rank starts from rank 0's parameters / optimizer moments (nn.DataParallel replicated module 0 on every
        forward, recognition_model.py:284); without this ranks agree only if the caller seeded them identically."""
        if self.world == 1:
            return
        for t in tensors:
            if t is not None:
                self.dist.broadcast(t, src=0, group=self.group)

    def begin(self):
        self._next = 0
        self._pending = []
        if self.trace and self.cuda:
            self.timeline = {"t0": self._event(torch.cuda.current_stream()), "stages": [], "buckets": []}
        else:
            self.timeline = None

    @staticmethod
    def _event(stream):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        return ev

    def on_stage(self, stage):
        """Called by Engine.backward when every gradient up to and including `stage` is final."""
        if self.world == 1:
            return
        done = self.stage_order.index(stage)
        if self.timeline is not None:
            self.timeline["stages"].append((stage, self._event(torch.cuda.current_stream())))
        while self._next < len(self.buckets) and self.stage_order.index(self.buckets[self._next][2]) <= done:
            s, e, _ = self.buckets[self._next]
            self._launch(self.flat.g[s:e])
            self._next += 1

    def _launch(self, t):
        dist = self.dist
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=self.timeline is not None)
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                e0 = self._event(self.stream) if self.timeline is not None else None
                if self.comm is not None:
                    self.comm.allreduce(t, average=True)
                elif self.backend == "nccl":
                    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(t, group=self.group)
                    t.div_(self.world)
                if e0 is not None:
                    self.timeline["buckets"].append((t.numel() * 4, ev, e0, self._event(self.stream)))
        else:
            dist.all_reduce(t, group=self.group)
            t.div_(self.world)

    def finish(self):
        """All buckets reduced and visible to the compute stream."""
        if self.world == 1:
            return
        self.on_stage("conv0")
        if self.cuda:
            if self.timeline is not None:
                self.timeline["bwd_end"] = self._event(torch.cuda.current_stream())
            torch.cuda.current_stream().wait_stream(self.stream)
            if self.timeline is not None:
                self.timeline["sync_end"] = self._event(torch.cuda.current_stream())

    def timeline_ms(self):
        """The recorded step as plain numbers (ms since the start of backward): when each stage's gradients were final, when each
        bucket became ready / started / finished its all-reduce, and how long the compute stream then had to wait for the last
        bucket (`exposed_ms`: the only communication time that is not hidden behind backward)."""
        tl = self.timeline
        if tl is None:
            return None
        torch.cuda.synchronize()
        t0 = tl["t0"]
        ms = lambda e: round(t0.elapsed_time(e), 3)      # noqa: E731
        buckets = [dict(mbytes=round(nb / 1e6, 1), ready=ms(r), start=ms(a), end=ms(b)) for nb, r, a, b in tl["buckets"]]
        busy = sum(b["end"] - b["start"] for b in buckets)
        return dict(stages=[(n, ms(e)) for n, e in tl["stages"]], buckets=buckets, backward_ms=ms(tl["bwd_end"]),
                    exposed_ms=round(tl["bwd_end"].elapsed_time(tl["sync_end"]), 3), allreduce_busy_ms=round(busy, 3),
                    mbytes=round(sum(b["mbytes"] for b in buckets), 1))


class Trainer:
    def __init__(self, model, learning_rate=3e-4, learning_rate_warmup=1500, alpha_loss=0.2, batch_size_grad=100,
                 eps_ls=0.1, weight_decay=0.01, seed=0, distributed=False, bucket_bytes=128 << 20):
        self.model = model
        self.lr_target, self.warmup = learning_rate, learning_rate_warmup
        self.lr = learning_rate
        self.alpha, self.eps_ls = alpha_loss, eps_ls
        self.wd = weight_decay
        model.engine()
        self.flat = FlatState(model)
        self.eng = model.engine()
        self.eng.shadow = self.flat.shadow
        self.eng.pack()
        # Model.load_state_dict / weights_changed() write the fp32 masters only: refresh the bf16 shadow the GEMM operands
        # are views of before the next repack
        model._weights_hooks.append(self._refresh_shadow)
        self.sync = GradSync(self.flat, self.eng.n_enc, self.eng.n_dec, bucket_bytes) if distributed else None
        if self.sync is not None:
            self.sync.broadcast_state([self.flat.p, self.flat.m, self.flat.v, self.flat.pb] +
                                      [p.data for n, p in model.named_parameters() if not model._trainable[n]] +
                                      [b for b in model.buffers() if b.is_floating_point()])
            self.eng.pack()
        self.gate = AccumulationGate(batch_size_grad, distributed)
        self.batch_idx = 0
        self.seed = seed
        self.dev = self.flat.p.device
        self._host_loss = torch.zeros(3, dtype=torch.float32).pin_memory() if self.dev.type == "cuda" else torch.zeros(3)
        self._loss_event = None
        self.launch_count = 0

    # ---- checkpoint / resume (the reference saves model.state_dict() only, recognition_model.py:310-312; optimizer, step and
    #      accumulation state are added so that a resumed run continues bit-for-bit) -------------------------------------
    def state_dict(self, dataparallel_prefix=False):
        """{'model': reference-named state_dict (optionally with the 'module.' prefix nn.DataParallel checkpoints carry,
        SURVEY.md section 5), 'optim': flat Adam moments + counters}.  Tensors are copies on the host."""
        pre = "module." if dataparallel_prefix else ""
        model_sd = {pre + k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}
        return {"model": model_sd,
                "optim": {"names": list(self.flat.names), "m": self.flat.m.cpu().clone(), "v": self.flat.v.cpu().clone(),
                          "step_count": self.flat.step_count, "batch_idx": self.batch_idx, "sum_batch_size": self.gate.sum_batch_size,
                          "g": self.flat.g.cpu().clone(), "lr": self.lr}}

    def load_state_dict(self, sd):
        model_sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd["model"].items()}
        self.model.load_state_dict(model_sd)                     # copies into the flat fp32 views (storage is unchanged)
        o = sd.get("optim")
        if o is not None:
            if list(o["names"]) != list(self.flat.names):
                raise L.SstError("optimizer state belongs to a different model configuration")
            self.flat.m.copy_(o["m"]); self.flat.v.copy_(o["v"]); self.flat.g.copy_(o["g"])
            self.flat.step_count, self.batch_idx, self.gate.sum_batch_size = o["step_count"], o["batch_idx"], o["sum_batch_size"]
            self.lr = o["lr"]
        if self.flat.pb is not None:
            self.flat.pb.copy_(self.flat.p)                      # refresh the bf16 shadow of the GEMM operands
        self.eng.pack()
        self.model._weights_version += 1

    def _refresh_shadow(self):
        if self.flat.pb is not None:
            self.flat.pb.copy_(self.flat.p)
        self.eng._packed_version = None

    def start_epoch(self):
        """The reference zeroes the gradients and the accumulated chunk count at the top of every epoch
        (recognition_model.py:67-68), dropping a partial accumulation; call this where the reference enters training_loop()."""
        self.flat.zero_grad()
        self.gate.sum_batch_size = 0

    def schedule_lr(self, iteration):
        """recognition_model.py:57-64."""
        iteration = iteration + 1
        if iteration <= self.warmup:
            self.lr = iteration * self.lr_target / self.warmup

    # ---- host-side batch preparation (recognition_model.py:77,85-87,95-97) ------------------------------------------
    def prepare(self, example, pin=True, reuse=False):
        """collate_raw dict -> host tensors (pinned) ready for an async H2D copy.  The EMG chunks -- all but a few hundred bytes of
        the batch -- are packed directly into page-locked memory (data_utils.ChunkStager).  `reuse`: alternate between two
        long-lived staging buffers instead of allocating one per batch; the returned X is then only valid until the call after
        the next one (what a training loop that prepares batch i+1 while batch i runs needs, Trainer.run)."""
        pinned = pin and self.dev.type == "cuda"
        if reuse:
            if getattr(self, "_stagers", None) is None:
                self._stagers, self._stager_i = [ChunkStager(pin=pinned), ChunkStager(pin=pinned)], 0
            stager = self._stagers[self._stager_i & 1]
            self._stager_i += 1
        else:
            stager = ChunkStager(pin=pinned)
        host = prepare_batch(example, stager=stager)
        if pinned:
            host = {k: (v.pin_memory() if torch.is_tensor(v) and not v.is_pinned() else v) for k, v in host.items()}
        return host

    def to_device(self, host):
        d = {k: (v.to(self.dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
        d['h2d_bytes'] = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))
        return d

    # ---- one micro-batch (recognition_model.py:69-118) -----------------------------------------------------------------
    @property
    def batch_size_grad(self):
        return self.gate.batch_size_grad

    @batch_size_grad.setter
    def batch_size_grad(self, v):
        self.gate.batch_size_grad = v

    @property
    def sum_batch_size(self):
        return self.gate.sum_batch_size

    def step_device(self, dev_batch, shift_r=None, global_chunks=None):
        """forward + loss + backward (+ optimizer when the accumulation threshold is reached) on device-resident inputs.
        Returns the float32[3] device tensor (loss, loss_dec, loss_enc).  `global_chunks`: chunks of this micro-step summed
        over all ranks when the caller knows it (AccumulationGate)."""
        import random
        model = self.model
        model.train()
        self.schedule_lr(self.batch_idx)
        X = dev_batch['X']
        will_step = self.gate.add(X.shape[0], global_chunks)          # the same decision on every rank
        r = random.randrange(8) if shift_r is None else shift_r       # architecture.py:105
        if r > 0:
            L.shift_left(X, X.shape[0], X.shape[1], X.shape[2], r)
        losses = self._micro_step(dev_batch, will_step, self.seed * 1000003 + self.batch_idx, None, None)
        if will_step:
            self.flat.step_count += 1
            model._weights_version += 1
        self.batch_idx += 1
        return losses

    def _micro_step(self, dev_batch, will_step, seed, meta, hyper):
        """The device work of one micro-batch after the input shift (recognition_model.py:90-118).  `hyper`: device float[3]
        (lr, 1 - beta1^t, sqrt(1 - beta2^t)) of the optimizer step instead of by-value scalars -- the form a CUDA graph replays."""
        eng, flat = self.eng, self.flat
        has_dec = eng.n_dec > 0
        _, _, ctx = eng.forward(dev_batch['X'], dev_batch['lengths'], dev_batch['tgt_in'] if has_dec else None,
                                dev_batch['tgt_lens'] if has_dec else None, training=True, seed=seed,
                                ctc=(dev_batch['ctc_tgt'], dev_batch['ctc_lens'], self.alpha if has_dec else 1.0), meta=meta)
        losses = eng.losses(ctx, dev_batch['ctc_tgt'], dev_batch['ctc_lens'], dev_batch['tgt_out'] if has_dec else None,
                            dev_batch['n_valid'], self.alpha, self.eps_ls)
        if self.sync is not None and will_step:
            self.sync.begin()
            eng.backward(ctx, flat.G_all, on_stage=self.sync.on_stage)
            self.sync.finish()
        else:
            eng.backward(ctx, flat.G_all)
        if will_step:                                               # recognition_model.py:115-118
            if hyper is None:
                L.adamw(flat.p, flat.g, flat.m, flat.v, flat.numel, self.lr, 0.9, 0.999, 1e-8, self.wd, flat.step_count + 1, flat.pb)
            else:
                L.adamw_dev(flat.p, flat.g, flat.m, flat.v, flat.numel, hyper, 0.9, 0.999, 1e-8, self.wd, flat.pb)
            flat.zero_grad()
            eng.pack()
        return losses

    # ---- CUDA-graph replay of the micro-step -------------------------------------------------------
    def step_graphed(self, dev_batch, shift_r=None, global_chunks=None, warm=2):
        """step_device with the ~440 launches of a micro-step replayed as ONE CUDA graph: for small batches the step is bound by
        the host's launch rate (16 us per launch through ctypes against 5-10 us of device time), not by the device.  A graph
        freezes shapes and by-value arguments, so
          * one graph per batch SIGNATURE (chunks, utterance lengths, target shapes, number of unpadded targets, step-or-accumulate),
            captured after `warm` ordinary steps of that signature and kept (least recently used of 16 dropped) -- a stream of
            equally shaped batches (fixed-length buckets, the BASELINE.json configs) replays, any other batch simply runs eagerly;
          * what changes from step to step travels through device memory, written in front of the replay (sst_write_scalars): the
            dropout salt every Philox-drawing kernel adds to its seed, the learning rate of the warm-up schedule
            (recognition_model.py:57-64) and Adam's bias corrections (sst_adamw_dev);
          * the inputs are copied into the graph's own buffers; the random input shift (architecture.py:104-108, one
            random.randrange(8) per step, as ever) is applied to that copy by an ordinary launch.
        Single process only: the bucketed all-reduce of GradSync stays on the eager path.  Returns the float32[3] device losses
        (overwritten by the next replay of the same graph)."""
        import random
        if self.sync is not None or self.dev.type != "cuda":
            return self.step_device(dev_batch, shift_r, global_chunks)
        X = dev_batch['X']
        has_dec = self.eng.n_dec > 0
        will_step_peek = self.gate.sum_batch_size + (int(global_chunks) if global_chunks is not None else int(X.shape[0])) >= self.gate.batch_size_grad
        key = (tuple(X.shape), tuple(dev_batch['lengths']), tuple(dev_batch['tgt_in'].shape) if has_dec else None,
               tuple(dev_batch['ctc_tgt'].shape), int(dev_batch['n_valid']), bool(will_step_peek), bool(self.eng.cfg.get("packed", True)))
        st = self._graph_state()
        ent = st["graphs"].get(key)
        if ent is not None and ent["graph"] is not None and ent["plan"] is not getattr(self.eng, "_pack_plan", None):
            ent = st["graphs"][key] = dict(seen=warm, graph=None)      # the packed GEMM operands moved: the recorded pointers are stale
        if ent is None or ent["graph"] is None:
            if ent is None:
                ent = st["graphs"][key] = dict(seen=0, graph=None)
            ent["seen"] += 1
            if ent["seen"] <= warm:
                return self.step_device(dev_batch, shift_r, global_chunks)      # also warms every lazily initialised piece of the path
            self._capture(st, key, ent, dev_batch, will_step_peek)
        ent["tick"] = st["tick"] = st["tick"] + 1
        # ---- replay -------------------------------------------------------------------------------------------------------
        self.model.train()
        self.schedule_lr(self.batch_idx)
        will_step = self.gate.add(X.shape[0], global_chunks)
        assert will_step == will_step_peek
        s = ent["static"]
        for k in ("X", "tgt_in", "tgt_out", "tgt_lens", "ctc_tgt", "ctc_lens"):
            if k in s and torch.is_tensor(s[k]):
                s[k].copy_(dev_batch[k], non_blocking=True)
        r = random.randrange(8) if shift_r is None else shift_r
        if r > 0:
            L.shift_left(s["X"], X.shape[0], X.shape[1], X.shape[2], r)
        t = self.flat.step_count + 1
        L.write_scalars(st["scalars"], "<qfff", (self.seed * 1000003 + self.batch_idx) * 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF,
                        float(self.lr), float(1.0 - 0.9 ** t), float((1.0 - 0.999 ** t) ** 0.5))
        ent["graph"].replay()
        if will_step:
            self.flat.step_count += 1
            self.model._weights_version += 1
            self.eng._packed_version = None
        self.batch_idx += 1
        return ent["losses"]

    def _graph_state(self):
        st = getattr(self, "_graphs", None)
        if st is None:
            scalars = torch.zeros(8, dtype=torch.int32, device=self.dev)          # [salt (8 bytes) | lr, bc1, sqrt bc2 | pad]
            st = self._graphs = dict(graphs={}, tick=0, scalars=scalars, salt=scalars[:2].view(torch.int64),
                                     hyper=scalars[2:5].view(torch.float32), pool=None, stream=torch.cuda.Stream(device=self.dev))
        return st

    def _capture(self, st, key, ent, dev_batch, will_step):
        live = [k for k, e in st["graphs"].items() if e["graph"] is not None]
        if len(live) >= 16:                                                      # drop the least recently replayed graph
            old = min(live, key=lambda k: st["graphs"][k]["tick"])
            del st["graphs"][old]
        static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in dev_batch.items()}
        meta = self.eng.batch_meta(dev_batch['lengths'])
        L.write_scalars(st["scalars"], "<qfff", 0, float(self.lr), 1.0, 1.0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        L.set_dropout_salt(st["salt"])
        # gradients accumulate in flat.g: the capture pass must not be counted -- it is not executed, only recorded
        try:
            with torch.cuda.graph(g, pool=st["pool"], stream=st["stream"]):
                losses = self._micro_step(static, will_step, 0x5353, meta, st["hyper"])
        finally:
            L.set_dropout_salt(None)
        if st["pool"] is None:
            st["pool"] = g.pool()
        # everything the recorded launches point at and the engine may replace later stays referenced by the entry
        ent.update(graph=g, static=static, meta=meta, losses=losses, tick=st["tick"], plan=getattr(self.eng, "_pack_plan", None),
                   ws=L._attn_ws.get(self.dev), pk=dict(self.eng.pk))

    def step(self, example):
        """Public entry: collate_raw dict on the host -> losses; one async D2H of float[3] replaces the three .item() calls."""
        dev_batch = self.to_device(self.prepare(example))
        losses = self.step_device(dev_batch)
        return self.fetch_losses(losses)

    def run(self, host_batches, global_chunks=None):
        """Pipelined public entry for a stream of prepared (pinned) host batches, the way a training loop would drive it:
        the H2D copy of batch i+1 is issued on a copy stream while batch i computes, and the loss of step i is read back
        (async D2H + event) while step i+1 is already enqueued -- every step still uploads its inputs and downloads its
        result, but neither sits on the critical path.  Yields (loss, loss_dec, loss_enc) per step, in order."""
        if self.dev.type != "cuda":
            for h in host_batches:
                yield self.step_host(h)
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._pinned_losses = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream, pinned = self._copy_stream, self._pinned_losses
        main = torch.cuda.current_stream()
        it = iter(host_batches)

        def upload(h):
            # allocated from the copy stream's own pool; record_stream(main) below keeps the blocks alive until the step that
            # consumes them has run, so the upload never has to wait for the compute stream
            with torch.cuda.stream(copy_stream):
                d = self.to_device(h)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return d, ev

        nxt = next(it, None)
        cur = upload(nxt) if nxt is not None else None
        pending = None                              # (slot, event) of the previous step's loss read-back
        k = 0
        while cur is not None:
            nxt = next(it, None)
            d, ev = cur
            main.wait_event(ev)
            for v in d.values():
                if torch.is_tensor(v):
                    v.record_stream(main)
            losses = self.step_device(d, global_chunks=global_chunks)
            cur = upload(nxt) if nxt is not None else None
            slot = k & 1
            pinned[slot].copy_(losses, non_blocking=True)
            lev = torch.cuda.Event()
            lev.record(main)
            if pending is not None:
                yield self._read_losses(*pending)
            pending = (pinned[slot], lev)
            k += 1
        if pending is not None:
            yield self._read_losses(*pending)

    def _read_losses(self, buf, ev):
        ev.synchronize()
        loss_dec, loss_enc = float(buf[1]), float(buf[2])
        if self.eng.n_dec > 0:
            return (1 - self.alpha) * loss_dec + self.alpha * loss_enc, loss_dec, loss_enc
        return loss_enc, 0.0, loss_enc

    def step_host(self, host):
        losses = self.step_device(self.to_device(host))
        self.fetch_losses(losses)
        return self.wait_losses()

    def fetch_losses(self, losses):
        if self.dev.type == "cuda":
            self._host_loss.copy_(losses, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._loss_event = ev
        else:
            self._host_loss.copy_(losses)
        return self._host_loss

    def wait_losses(self):
        if self._loss_event is not None:
            self._loss_event.synchronize()
        a = self._host_loss
        loss_dec, loss_enc = float(a[1]), float(a[2])
        if self.eng.n_dec > 0:
            return (1 - self.alpha) * loss_dec + self.alpha * loss_enc, loss_dec, loss_enc
        return loss_enc, 0.0, loss_enc

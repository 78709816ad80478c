"""Per-rank timeline of one data-parallel training step: when each backward stage finished, when each gradient
bucket became ready / started / finished its NCCL all-reduce, and the communication time that stayed exposed after backward.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/ddp_timeline.py [--bucket-mb 25] [--workload cfg2] [--steps 10]
Prints one JSON document (rank 0): per-rank numbers + the step time with and without the gradient exchange."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import architecture as A
from sst_b200.synthetic import make_batch
from sst_b200.train import Trainer
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--bucket-mb", type=int, default=128)
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
w = bench.WORKLOADS[args.workload]
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=w["n_enc"], num_layers_decoder=w["n_dec"], n_heads_encoder=8,
            n_heads_decoder=8, relative_distance=100, dropout_model=0.2, dropout_pos_emb=0.2, sst_dtype="bf16")
torch.manual_seed(0)
model = A.Model(112, 44, 43, dev).to(dev)
tr = Trainer(model, alpha_loss=w["alpha"], batch_size_grad=1, seed=rank, distributed=True, bucket_bytes=args.bucket_mb << 20)
batches = [tr.to_device(tr.prepare(make_batch(w["n_utt"], w["frames"], w["tgt"][0], w["tgt"][1], seed=1234 + rank * 100 + i))) for i in range(2)]
pristine = [d["X"].clone() for d in batches]
chunks = world * batches[0]["X"].shape[0]


def step(i, exchange=True):
    d = batches[i % 2]
    d["X"].copy_(pristine[i % 2])
    if exchange:
        return tr.step_device(d, global_chunks=chunks)
    sync, tr.sync = tr.sync, None                # same step without the gradient exchange (the ranks then diverge: timing only)
    try:
        return tr.step_device(d, global_chunks=chunks)
    finally:
        tr.sync = sync


def timed(n, exchange):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        step(i, exchange)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


for i in range(3):
    step(i)
ms_with = timed(args.steps, True)
tr.sync.trace = True
step(0)
tl = tr.sync.timeline_ms()
tr.sync.trace = False
ms_without = timed(args.steps, False)
allt = [None] * world
dist.all_gather_object(allt, tl)
if rank == 0:
    doc = {"world": world, "workload": args.workload, "bucket_mb": args.bucket_mb, "ms_per_step_with_exchange": round(ms_with, 3),
           "ms_per_step_without_exchange": round(ms_without, 3), "exchange_cost_ms": round(ms_with - ms_without, 3),
           "n_buckets": len(tl["buckets"]), "gradient_mbytes": tl["mbytes"],
           "per_rank": [{"rank": r, "backward_ms": t["backward_ms"], "exposed_ms": t["exposed_ms"], "allreduce_busy_ms": t["allreduce_busy_ms"],
                         "last_bucket": t["buckets"][-1]} for r, t in enumerate(allt)],
           "rank0_stages": tl["stages"], "rank0_buckets": tl["buckets"]}
    print(json.dumps(doc))
dist.destroy_process_group()

"""Measurement driver (not a pytest file): host-side enqueue time of one cfg2 training step against its device time.
If enqueue >= device time the step is launch-bound and the GPU starves on the small decoder kernels."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import architecture as A
from sst_b200 import lib as L
from sst_b200.synthetic import make_batch
from sst_b200.train import Trainer

dev = torch.device("cuda", 0)
A.configure(model_size=768, feed_forward_layer_size=3072, num_layers_encoder=6, num_layers_decoder=6, n_heads_encoder=8,
            n_heads_decoder=8, relative_distance=100, dropout_model=0.2, dropout_pos_emb=0.2, sst_dtype="bf16")
torch.manual_seed(0)
model = A.Model(112, 44, 43, dev).to(dev)
tr = Trainer(model, alpha_loss=0.2, batch_size_grad=1, seed=0)
d = tr.to_device(tr.prepare(make_batch(64, 1000, 80, 120, seed=1234)))
X0 = d["X"].clone()
for _ in range(3):
    d["X"].copy_(X0)
    tr.step_device(d)
torch.cuda.synchronize()
for rep in range(3):
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    d["X"].copy_(X0)
    tr.step_device(d)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("step %d: host enqueue %.2f ms, device %.2f ms, launches %d" % (rep, (t1 - t0) * 1e3, e0.elapsed_time(e1), L.launch_count() - n0))
# phase split of the enqueue time
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
d["X"].copy_(X0)
tr.step_device(d)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

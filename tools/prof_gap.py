"""Measurement driver (not a pytest file): inter-kernel gap of dependent launches.  50 decoder-sized GEMMs are captured in a
CUDA graph (no CPU launch cost at replay); per-kernel graph time minus the kernel's own duration (ncu: 24.2 us) is what a
kernel boundary costs on the device."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

M, N, K = 7744, 768, 768
x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        L.gemm(x, w, y, M, N, K, K, K, N)
        L.gemm(y, w, x, M, N, K, K, K, N)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
reps = 25
with torch.cuda.graph(g, stream=s):
    for _ in range(reps):
        L.gemm(x, w, y, M, N, K, K, K, N)      # ping-pong: every launch depends on the previous one
        L.gemm(y, w, x, M, N, K, K, K, N)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("graph replay: %.2f us per dependent GEMM launch (7744x768x768)" % (e0.elapsed_time(e1) * 1e3 / (10 * 2 * reps)))
# eager, same chain
e0.record()
for _ in range(10 * reps):
    L.gemm(x, w, y, M, N, K, K, K, N)
    L.gemm(y, w, x, M, N, K, K, K, N)
e1.record()
torch.cuda.synchronize()
print("eager:        %.2f us per launch" % (e0.elapsed_time(e1) * 1e3 / (10 * 2 * reps)))

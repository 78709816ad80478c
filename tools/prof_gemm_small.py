"""Measurement driver (not a pytest file): decoder-sized GEMMs (M = 7744 rows), one launch per (N, K); run under
`ncu --metrics gpu__time_duration.sum` to read pure kernel durations (fixed cost vs K)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

M = 7744
for (N, K) in [(768, 64), (768, 256), (768, 768), (768, 3072), (2304, 768), (3072, 768), (256, 768)]:
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.gemm(x, w, y, M, N, K, K, K, N)
    torch.cuda.synchronize()
print("ok")

"""Profiling driver (not a pytest file): the two FFN GEMMs of the cfg2 encoder layer through the C ABI, 3 times each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

M, D, F = 64000, 768, 3072
x = (torch.randn(M, D, device="cuda") * 0.5).bfloat16()
w1 = (torch.randn(F, D, device="cuda") * 0.05).bfloat16()
w2 = (torch.randn(D, F, device="cuda") * 0.05).bfloat16()
b1 = torch.randn(F, device="cuda")
b2 = torch.randn(D, device="cuda")
h = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
for it in range(3):
    L.gemm(x, w1, h, M, F, D, D, D, F, bias=b1, epilogue=L.EPI_BIAS | L.EPI_RELU | L.EPI_DROPOUT, drop_p=0.2, seed=it)
    L.gemm(h, w2, y, M, D, F, F, F, D, bias=b2, epilogue=L.EPI_BIAS)
torch.cuda.synchronize()
print("ok")

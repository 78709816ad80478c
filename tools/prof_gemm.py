"""Profiling driver (not a pytest file): the cfg2 encoder-layer GEMM shapes through the C ABI.
Kernel order per iteration: FFN1 fwd (bias+relu+dropout), FFN2 fwd (bias), FFN2 dgrad (mulmask), QKV fwd (plain),
out-proj (plain, N=768), FFN1 wgrad (NT_MN split-K)."""
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
wqkv = (torch.randn(3 * D, D, device="cuda") * 0.05).bfloat16()
wo = (torch.randn(D, D, device="cuda") * 0.05).bfloat16()
b1 = torch.randn(F, device="cuda")
b2 = torch.randn(D, device="cuda")
h = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
dh = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
qkv = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
dw1 = torch.zeros(F, D, device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(iters)]
for it in range(iters):
    e = ev[it]
    e[0].record()
    L.gemm(x, w1, h, M, F, D, D, D, F, bias=b1, epilogue=L.EPI_BIAS | L.EPI_RELU | L.EPI_DROPOUT, drop_p=0.2, seed=it)
    e[1].record()
    L.gemm(h, w2, y, M, D, F, F, F, D, bias=b2, epilogue=L.EPI_BIAS)
    e[2].record()
    L.gemm(y, w1, dh, M, F, D, D, D, F, aux=h, ldaux=F, epilogue=L.EPI_MULMASK, mask_scale=1.25)
    e[3].record()
    L.gemm(x, wqkv, qkv, M, 3 * D, D, D, D, 3 * D)
    e[4].record()
    L.gemm(x, wo, y, M, D, D, D, D, D)
    e[5].record()
    L.gemm(dh, x, dw1, F, D, M, F, D, D, layout=L.GEMM_NT_MN)
    e[6].record()
torch.cuda.synchronize()
names = ["ffn1 fwd 64000x3072x768 bias+relu+dropout", "ffn2 fwd 64000x768x3072 bias", "ffn2 dgrad 64000x3072x768 mulmask",
         "qkv 64000x2304x768 plain", "out-proj 64000x768x768 plain", "ffn1 wgrad 3072x768x64000 split-K"]
flops = [2 * M * F * D, 2 * M * F * D, 2 * M * F * D, 2 * M * 3 * D * D, 2 * M * D * D, 2 * M * F * D]
for i, n in enumerate(names):
    ms = min(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(1, iters))
    print("%-45s %8.1f us %7.1f TFLOP/s" % (n, ms * 1e3, flops[i] / ms / 1e9))

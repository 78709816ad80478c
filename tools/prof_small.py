"""Measurement driver (not a pytest file): per-kernel floor of the small-batch step.  Chains of dependent launches are captured in
a CUDA graph (no host launch cost at replay) and timed per launch: GEMMs of decreasing M, LayerNorm, attention forward."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L


def chain(name, fn, reps=40):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %7.2f us per launch" % (name, e0.elapsed_time(e1) * 1e3 / (10 * reps)))


for M in (128, 800, 3400, 7744):
    for N, K in ((768, 768), (3072, 768), (768, 3072)):
        x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        chain("gemm TN %dx%dx%d" % (M, N, K), lambda: L.gemm(x, w, y, M, N, K, K, K, N))
    K = 768
    N = 768
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    dy = (torch.randn(M, N, device="cuda") * 0.5).bfloat16()
    dw = torch.zeros(N, K, device="cuda")
    chain("gemm wgrad %dx%dx%d" % (N, K, M), lambda: L.gemm(dy, x, dw, N, K, M, N, K, K, layout=L.GEMM_NT_MN, epilogue=L.EPI_ACCUM))
    D = 768
    a = torch.randn(M, D, device="cuda").bfloat16()
    r = torch.randn(M, D, device="cuda").bfloat16()
    yy = torch.empty_like(a)
    gam = torch.ones(D, device="cuda"); bet = torch.zeros(D, device="cuda")
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    chain("layernorm_fwd rows %d" % M, lambda: L.layernorm_fwd(L.BF16, M, D, a, r, 0.2, 1, gam, bet, yy, r, mean, rstd))
    dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda")
    ds = torch.empty_like(a); dr = torch.empty_like(a)
    chain("layernorm_bwd rows %d" % M, lambda: L.layernorm_bwd(L.BF16, M, D, a, r, mean, rstd, gam, ds, dr, 0.2, 1, dg, db))
    cs = torch.zeros(D, device="cuda")
    chain("colsum rows %d" % M, lambda: L.colsum_accum(L.BF16, a, M, D, D, cs))

H, dh = 8, 96
for B, Lq in ((4, 200), (4, 850)):
    qkv = (torch.randn(B * Lq, 3 * H * dh, device="cuda") * 0.5).bfloat16()
    E = (torch.randn(H, 199, dh, device="cuda") * 0.1).bfloat16()
    o = torch.empty(B * Lq, H * dh, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(2 * B * H * Lq, device="cuda")
    lens = torch.full((B,), Lq, dtype=torch.int32, device="cuda")
    D = H * dh
    d = L.attn_desc(L.BF16, B, H, Lq, Lq, dh, 3 * D, 3 * D, 3 * D, D, False, True, 100, 1 / math.sqrt(dh), 0.2, 1)
    chain("attn_fwd B%d L%d" % (B, Lq), lambda: L.attn_fwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, lens, o, lse))
    dO = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B * H * Lq, device="cuda")
    chain("attn_bwd B%d L%d (2 kernels)" % (B, Lq), lambda: L.attn_bwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, lens, o, lse, dO,
                                                                      dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta))

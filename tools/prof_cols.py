"""Measurement driver (not a pytest file): the column-wise HBM-bound kernels at the cfg2 shapes, GB/s each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

dev = "cuda"
n, T, C = 320, 800, 768
rows = n * T
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
xa = torch.randn(rows, C, device=dev, generator=g).to(bf)
xb = torch.randn(rows, C, device=dev, generator=g).to(bf)
dout = torch.randn(rows, C, device=dev, generator=g).to(bf)
h = torch.randn(64000, 3072, device=dev, generator=g).to(bf)
stats = torch.empty(2 * C, device=dev, dtype=torch.float64)
mean = torch.zeros(C, device=dev); invstd = torch.ones(C, device=dev); gam = torch.ones(C, device=dev); bet = torch.zeros(C, device=dev)
out = torch.empty(n, T + 2, C, device=dev, dtype=bf)
dxa = torch.empty(n, T + 2, C, device=dev, dtype=bf)
dxb = torch.empty(n, T + 1, C, device=dev, dtype=bf)
dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev); dg2 = torch.zeros(C, device=dev); db2 = torch.zeros(C, device=dev)
red = torch.empty(3 * C, device=dev, dtype=torch.float64)
o1 = torch.zeros(3072, device=dev); o2 = torch.zeros(C, device=dev)
big = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def timeit(name, fn, nbytes, reps=5):
    ts = []
    for _ in range(reps):
        big.zero_()                                           # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("%-34s %8.1f us %7.1f GB/s" % (name, ms * 1e3, nbytes / ms / 1e6))


E = rows * C * 2
timeit("colstats 256000x768", lambda: L.colstats(L.BF16, xa, rows, C, C, stats), E)
timeit("colsum 64000x3072", lambda: L.colsum_accum(L.BF16, h, 64000, 3072, 3072, o1), 64000 * 3072 * 2)
timeit("colsum 64000x768", lambda: L.colsum_accum(L.BF16, xa, 64000, C, C, o2), 64000 * C * 2)
timeit("bn_apply 2br relu", lambda: L.bn_apply(L.BF16, n, T, C, xa, C, (mean, invstd, gam, bet), xb, C, (mean, invstd, gam, bet), True, out, 1, 1), 3 * E)
timeit("bn_apply 1br relu", lambda: L.bn_apply(L.BF16, n, T, C, xa, C, (mean, invstd, gam, bet), None, 0, None, True, out, 1, 1), 2 * E)
timeit("bn_bwd 2br stored mask", lambda: L.bn_bwd(L.BF16, n, T, C, dout, C, out, 1, 1, True, xa, C, mean, invstd, gam, dxa, C, 1, 1, dg, db,
                                                  xb, C, mean, invstd, gam, dxb, C, 0, 1, dg2, db2, red), 10 * E)
timeit("bn_bwd 2br recomputed mask", lambda: L.bn_bwd(L.BF16, n, T, C, dout, C, None, 1, 1, True, xa, C, mean, invstd, gam, dxa, C, 1, 1, dg, db,
                                                      xb, C, mean, invstd, gam, dxb, C, 0, 1, dg2, db2, red, beta_a=bet, beta_b=bet), 8 * E)
timeit("bn_bwd 1br recomputed mask", lambda: L.bn_bwd(L.BF16, n, T, C, dout, C, None, 1, 1, True, xa, C, mean, invstd, gam, dxa, C, 1, 1, dg, db,
                                                      None, 0, None, None, None, None, 0, 0, 0, None, None, red, beta_a=bet), 5 * E)

# LayerNorm(x + dropout(r)) forward / backward at the encoder token count
R = 64000
x = torch.randn(R, C, device=dev, generator=g).to(bf); r = torch.randn(R, C, device=dev, generator=g).to(bf)
yl = torch.empty_like(x); sl = torch.empty_like(x); mu = torch.empty(R, device=dev); rs = torch.empty(R, device=dev)
dsl = torch.empty_like(x); drl = torch.empty_like(x)
EL = R * C * 2
timeit("layernorm_fwd 64000x768 p0.2", lambda: L.layernorm_fwd(L.BF16, R, C, x, r, 0.2, 5, gam, bet, yl, sl, mu, rs), 4 * EL)
timeit("layernorm_bwd 64000x768 p0.2", lambda: L.layernorm_bwd(L.BF16, R, C, x, sl, mu, rs, gam, dsl, drl, 0.2, 5, dg, db), 4 * EL)
timeit("layernorm_bwd 64000x768 p0", lambda: L.layernorm_bwd(L.BF16, R, C, x, sl, mu, rs, gam, dsl, None, 0.0, 5, dg, db), 3 * EL)

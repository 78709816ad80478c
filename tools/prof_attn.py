"""Profiling driver (not a pytest file): cfg2-shaped encoder attention forward + backward through the C ABI, 4 times."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

B, H, Lx, dh, R, p = 64, 8, 1000, 96, 100, 0.2
D = H * dh
g = torch.Generator(device="cuda").manual_seed(3)
qkv = (torch.randn(B * Lx, 3 * D, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
E = (torch.randn(H, 2 * R - 1, dh, device="cuda", generator=g) * dh ** -0.5).to(torch.bfloat16)
dO = torch.randn(B * Lx, D, device="cuda", generator=g).to(torch.bfloat16)
lens = torch.full((B,), Lx, device="cuda", dtype=torch.int32)
o = torch.empty(B * Lx, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(2 * B * H * Lx, device="cuda")
dqkv = torch.empty_like(qkv)
delta = torch.empty(B * H * Lx, device="cuda")
d = L.attn_desc(L.BF16, B, H, Lx, Lx, dh, 3 * D, 3 * D, 3 * D, D, False, True, R, 1 / math.sqrt(dh), p, 1234)
for it in range(4):
    L.attn_fwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, lens, o, lse)
    L.attn_bwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, lens, o, lse, dO, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta)
torch.cuda.synchronize()
print("ok")

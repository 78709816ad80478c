"""Timing driver (not a pytest file): cfg2-shaped encoder attention forward / backward through the C ABI, CUDA events.
   python tools/time_attn.py [p] [iters]      (SST_ATTN_GEN=1: first-generation kernels, SST_ATTN_NSPLIT=2|4)"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, H, Lx, dh, R = 64, 8, 1000, 96, 100
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
fw = run(lambda: L.attn_fwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, lens, o, lse))
bw = run(lambda: L.attn_bwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, lens, o, lse, dO, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta))
f, b = L.attn_work(d)
print("gen %s nsplit %s p %.1f: fwd %.3f ms (%.0f TFLOP/s alg.), bwd %.3f ms (%.0f TFLOP/s alg.)" % (
    os.environ.get("SST_ATTN_GEN", "2"), os.environ.get("SST_ATTN_NSPLIT", "4"), p, fw, f / fw / 1e9, bw, b / bw / 1e9))

"""Timing driver (not a pytest file): cfg2-shaped encoder attention forward / backward through the C ABI, CUDA events.
   python tools/time_attn.py [p] [iters] [Lq] [Lk] [R] [causal]      (SST_ATTN_NSPLIT=2|4; default = the cfg2 encoder shape;
   121 1000 0 0 = decoder cross-attention, 121 121 0 1 = decoder self-attention)"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, H, dh = 64, 8, 96
Lx = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
Lk = int(sys.argv[4]) if len(sys.argv) > 4 else Lx
R = int(sys.argv[5]) if len(sys.argv) > 5 else 100
causal = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
D = H * dh
g = torch.Generator(device="cuda").manual_seed(3)
qkv = (torch.randn(B * max(Lx, Lk), 3 * D, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
E = (torch.randn(H, 2 * max(R, 1) - 1, dh, device="cuda", generator=g) * dh ** -0.5).to(torch.bfloat16) if R > 0 else None
dO = torch.randn(B * Lx, D, device="cuda", generator=g).to(torch.bfloat16)
lens = torch.full((B,), Lx, device="cuda", dtype=torch.int32)
klens = torch.full((B,), Lk, device="cuda", dtype=torch.int32)
o = torch.empty(B * Lx, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(2 * B * H * Lx, device="cuda")
dqkv = torch.empty_like(qkv)
delta = torch.empty(B * H * Lx, device="cuda")
d = L.attn_desc(L.BF16, B, H, Lx, Lk, dh, 3 * D, 3 * D, 3 * D, D, causal, Lx == Lk, R, 1 / math.sqrt(dh), p, 1234)
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
fw = run(lambda: L.attn_fwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, klens, o, lse))
bw = run(lambda: L.attn_bwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E, lens, klens, o, lse, dO, dqkv, dqkv[:, D:], dqkv[:, 2 * D:], delta))
f, b = L.attn_work(d)
print("Lq %d Lk %d R %d causal %d nsplit %s p %.1f: fwd %.3f ms (%.0f TFLOP/s alg.), bwd %.3f ms (%.0f TFLOP/s alg.)" % (
    Lx, Lk, R, causal, os.environ.get("SST_ATTN_NSPLIT", "4"), p, fw, f / fw / 1e9, bw, b / bw / 1e9))

"""Debug (not a pytest file): one tensor-core attention forward against the CUDA-core kernel on identical bf16 inputs."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L
DEV = "cuda"
B, H, Lx, R = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 4, 150, 40))]
p = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
dh = 96; D = H * dh
g = torch.Generator(device=DEV).manual_seed(3)
qkv = (torch.randn(B * Lx, 3 * D, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
E = (torch.randn(H, 2 * max(R, 1) - 1, dh, device=DEV, generator=g) * dh ** -0.5).to(torch.bfloat16)
lens = torch.tensor([Lx] + [max(1, Lx - 49)] * (B - 1), device=DEV, dtype=torch.int32)
outs = []
for simt in (True, False):
    d = L.attn_desc(L.BF16, B, H, Lx, Lx, dh, 3 * D, 3 * D, 3 * D, D, False, True, R, 1 / math.sqrt(dh), p, 1234, force_simt=simt)
    o = torch.zeros(B * Lx, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(2 * B * H * Lx, device=DEV)
    L.attn_fwd(d, qkv, qkv[:, D:], qkv[:, 2 * D:], E if R > 0 else None, lens, lens, o, lse)
    torch.cuda.synchronize()
    outs.append((o.float(), lse[:B * H * Lx] + lse[B * H * Lx:]))
    print("simt" if simt else "tc", "ok", float(o.float().abs().max()))
valid = (torch.arange(Lx, device=DEV)[None, :] < lens[:, None]).reshape(-1)
eo = (outs[0][0] - outs[1][0])[valid].abs().max() / outs[0][0].abs().max()
print("o rel err (valid rows) %.3e" % float(eo))
rows = (outs[0][0] - outs[1][0]).abs().amax(1)
bad = (rows > 2e-2 * outs[0][0].abs().max()).nonzero().flatten()
print("rows off:", bad.numel(), bad[:20].tolist())
if bad.numel():
    import collections
    c = collections.Counter((int(r) // Lx, (int(r) % Lx) // 128) for r in bad.tolist())
    print("(batch, query tile) -> rows off:", sorted(c.items())[:40])
    r0 = int(bad[0]); diff = (outs[0][0][r0] - outs[1][0][r0]).abs().view(H, dh).amax(1)
    print("row", r0, "per-head max err", [round(float(v), 3) for v in diff])

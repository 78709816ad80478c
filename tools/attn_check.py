"""Debug/benchmark script (not a pytest file): tensor-core attention vs the CUDA-core kernel on identical bf16 inputs."""
import math
import sys
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sst_b200  # noqa
from sst_b200 import lib as L

DEV = "cuda"


def run(B, H, Lq, Lk, R, causal, mqr, q_lens, k_lens, p=0.0, bwd=True, timing=False, self_attn=True):
    dh = 96
    D = H * dh
    g = torch.Generator(device=DEV).manual_seed(3)
    qkv = (torch.randn(B * Lq, 3 * D, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
    kv = (torch.randn(B * Lk, 2 * D, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
    E = (torch.randn(H, 2 * max(R, 1) - 1, dh, device=DEV, generator=g) * dh ** -0.5).to(torch.bfloat16)
    if self_attn:
        q_t, k_t, v_t = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        ldq = ldk = ldv = 3 * D
    else:
        q_t, k_t, v_t = qkv[:, :D], kv[:, :D], kv[:, D:]
        ldq, ldk, ldv = 3 * D, 2 * D, 2 * D
    ql = torch.tensor(q_lens, device=DEV, dtype=torch.int32) if q_lens is not None else None
    kl = torch.tensor(k_lens, device=DEV, dtype=torch.int32) if k_lens is not None else None
    scale = 1 / math.sqrt(dh)
    outs = []
    dO = torch.randn(B * Lq, D, device=DEV, generator=g).to(torch.bfloat16)
    for simt in (True, False):
        o = torch.zeros(B * Lq, D, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(2 * B * H * Lq, device=DEV)
        d = L.attn_desc(L.BF16, B, H, Lq, Lk, dh, ldq, ldk, ldv, D, causal, mqr, R, scale, p, 1234, force_simt=simt)
        L.attn_fwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse)
        torch.cuda.synchronize()
        res = [o.float(), lse.clone()]
        if bwd:
            dqkv = torch.zeros_like(qkv); dkv = torch.zeros_like(kv)
            if self_attn:
                dq_t, dk_t, dv_t = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
            else:
                dq_t, dk_t, dv_t = dqkv[:, :D], dkv[:, :D], dkv[:, D:]
            delta = torch.empty(B * H * Lq, device=DEV)
            L.attn_bwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse, dO, dq_t, dk_t, dv_t, delta)
            torch.cuda.synchronize()
            res += [dq_t.float().clone(), dk_t.float().clone(), dv_t.float().clone()]
        if timing and not simt:
            for what in ("fwd", "bwd"):
                if what == "bwd" and not bwd:
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 10
                for it in range(n + 2):
                    if it == 2:
                        e0.record()
                    if what == "fwd":
                        L.attn_fwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse)
                    else:
                        L.attn_bwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse, dO, dq_t, dk_t, dv_t, delta)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                fl = L.attn_work(d)[0 if what == "fwd" else 1]
                print("   %s: %.3f ms  %.1f TFLOP/s algorithmic (band-limited)" % (what, ms, fl / ms / 1e9))
        outs.append(res)
    names = ["o", "lse", "dq", "dk", "dv"]
    ok = True
    for n, a, b_ in zip(names, outs[0], outs[1]):
        if n == "lse":
            nr = B * H * Lq
            # compare m + log l for valid rows
            a = a[:nr] + a[nr:]; b_ = b_[:nr] + b_[nr:]
        if q_lens is not None and n in ("o", "dq") and mqr:
            # padded query rows: garbage-but-finite in both; compare valid rows only
            mask = (torch.arange(Lq, device=DEV)[None, :] < ql[:, None]).reshape(-1)
            if n in ("o", "dq"):
                a = a[mask]; b_ = b_[mask]
        fin = bool(torch.isfinite(b_).all())
        err = float((a - b_).abs().max() / (a.abs().max() + 1e-30))
        flag = "OK" if (err < 2e-2 and fin) else "FAIL"
        ok = ok and flag == "OK"
        print("   %-4s rel err %.3e finite=%s %s" % (n, err, fin, flag))
    return ok


if __name__ == "__main__":
    L.require_device()
    allok = True
    cases = [
        ("enc_band 150 R40", dict(B=2, H=4, Lq=150, Lk=150, R=40, causal=False, mqr=True, q_lens=[150, 101], k_lens=[150, 101])),
        ("enc_short 30 R40", dict(B=2, H=4, Lq=30, Lk=30, R=40, causal=False, mqr=True, q_lens=[30, 17], k_lens=[30, 17])),
        ("enc 200 R100", dict(B=4, H=8, Lq=200, Lk=200, R=100, causal=False, mqr=True, q_lens=[200, 180, 200, 150], k_lens=[200, 180, 200, 150])),
        ("dec_self 21", dict(B=3, H=4, Lq=21, Lk=21, R=0, causal=True, mqr=True, q_lens=[21, 9, 14], k_lens=[21, 9, 14])),
        ("dec_self 131", dict(B=3, H=4, Lq=131, Lk=131, R=0, causal=True, mqr=True, q_lens=[131, 9, 70], k_lens=[131, 9, 70])),
        ("dec_cross 21x77", dict(B=3, H=4, Lq=21, Lk=77, R=0, causal=False, mqr=False, q_lens=None, k_lens=[77, 40, 59], self_attn=False)),
        ("enc 1000 R100 dropout", dict(B=2, H=8, Lq=1000, Lk=1000, R=100, causal=False, mqr=True, q_lens=[1000, 777], k_lens=[1000, 777], p=0.2)),
    ]
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    for name, kw in cases:
        print(name)
        allok = run(**kw) and allok
    print("cfg2 shape timing")
    run(B=64, H=8, Lq=1000, Lk=1000, R=100, causal=False, mqr=True, q_lens=[1000] * 64, k_lens=[1000] * 64, p=0.2, timing=True)
    print("ALL OK" if allok else "SOME FAILED")

"""The tcgen05 attention kernels (attention_tc.cu / attention_tc_bwd.cu) against the CUDA-core kernels (attention_simt.cu) on
IDENTICAL bf16 inputs: forward output, saved softmax statistics, dq / dk / dv -- encoder self-attention with the banded
relative-position bias (several lengths incl. multi-tile L = 1000 with dropout: both paths draw the same Philox mask),
causal decoder self-attention, rectangular cross-attention, ragged lengths, per-position padding masks.
Tolerance 2e-2 of max|ref| (bf16 rounding of P and of the outputs); the fp32 oracle comparison of both paths is in
test_kernels_gpu.py::test_attention_fwd_bwd."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = {
    "enc_band_150_R40": dict(B=2, H=4, Lq=150, Lk=150, R=40, causal=False, mqr=True, q_lens=[150, 101], k_lens=[150, 101]),
    "enc_short_30_R40": dict(B=2, H=4, Lq=30, Lk=30, R=40, causal=False, mqr=True, q_lens=[30, 17], k_lens=[30, 17]),
    "enc_200_R100": dict(B=4, H=8, Lq=200, Lk=200, R=100, causal=False, mqr=True, q_lens=[200, 180, 200, 150], k_lens=[200, 180, 200, 150]),
    "enc_1000_R100_dropout": dict(B=2, H=8, Lq=1000, Lk=1000, R=100, causal=False, mqr=True, q_lens=[1000, 777], k_lens=[1000, 777], p=0.2),
    "enc_1337_R100": dict(B=1, H=8, Lq=1337, Lk=1337, R=100, causal=False, mqr=True, q_lens=[1337], k_lens=[1337]),
    "dec_self_21": dict(B=3, H=4, Lq=21, Lk=21, R=0, causal=True, mqr=True, q_lens=[21, 9, 14], k_lens=[21, 9, 14]),
    "dec_self_131_dropout": dict(B=3, H=4, Lq=131, Lk=131, R=0, causal=True, mqr=True, q_lens=[131, 9, 70], k_lens=[131, 9, 70], p=0.1),
    "dec_cross_21x77": dict(B=3, H=4, Lq=21, Lk=77, R=0, causal=False, mqr=False, q_lens=None, k_lens=[77, 40, 59], self_attn=False),
    "dec_cross_121x1000_dropout": dict(B=2, H=8, Lq=121, Lk=1000, R=0, causal=False, mqr=False, q_lens=None, k_lens=[1000, 640],
                                       self_attn=False, p=0.2),
    "dec_self_pad_positions": dict(B=2, H=4, Lq=40, Lk=40, R=0, causal=True, mqr=True, q_lens=None, k_lens=None, pad_positions=[3, 17, 18]),
}


def _run(L, c, simt):
    B, H, Lq, Lk, R = c["B"], c["H"], c["Lq"], c["Lk"], c["R"]
    dh = 96
    D = H * dh
    p = c.get("p", 0.0)
    self_attn = c.get("self_attn", True)
    g = torch.Generator(device=DEV).manual_seed(3)
    qkv = (torch.randn(B * Lq, 3 * D, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
    kv = (torch.randn(B * Lk, 2 * D, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
    E = (torch.randn(H, 2 * max(R, 1) - 1, dh, device=DEV, generator=g) * dh ** -0.5).to(torch.bfloat16)
    dO = torch.randn(B * Lq, D, device=DEV, generator=g).to(torch.bfloat16)
    if self_attn:
        q_t, k_t, v_t, ldq, ldk, ldv = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], 3 * D, 3 * D, 3 * D
    else:
        q_t, k_t, v_t, ldq, ldk, ldv = qkv[:, :D], kv[:, :D], kv[:, D:], 3 * D, 2 * D, 2 * D
    ql = torch.tensor(c["q_lens"], device=DEV, dtype=torch.int32) if c["q_lens"] is not None else None
    kl = torch.tensor(c["k_lens"], device=DEV, dtype=torch.int32) if c["k_lens"] is not None else None
    pad = None
    if "pad_positions" in c:
        pad = torch.zeros(B, Lq, dtype=torch.uint8, device=DEV)
        pad[:, c["pad_positions"]] = 1
    d = L.attn_desc(L.BF16, B, H, Lq, Lk, dh, ldq, ldk, ldv, D, c["causal"], c["mqr"], R, 1 / math.sqrt(dh), p, 1234,
                    force_simt=simt, q_pad=pad, k_pad=pad)
    o = torch.zeros(B * Lq, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(2 * B * H * Lq, device=DEV)
    L.attn_fwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse)
    dqkv, dkv = torch.zeros_like(qkv), torch.zeros_like(kv)
    if self_attn:
        dq_t, dk_t, dv_t = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
    else:
        dq_t, dk_t, dv_t = dqkv[:, :D], dkv[:, :D], dkv[:, D:]
    delta = torch.empty(B * H * Lq, device=DEV)
    L.attn_bwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse, dO, dq_t, dk_t, dv_t, delta)
    torch.cuda.synchronize()
    nr = B * H * Lq
    return dict(o=o.float(), lse=lse[:nr] + lse[nr:], dq=dq_t.float().clone(), dk=dk_t.float().clone(), dv=dv_t.float().clone())


@pytest.mark.parametrize("name", sorted(CASES))
def test_tensor_core_attention_matches_cuda_core(name):
    import sst_b200  # noqa: F401
    from sst_b200 import lib as L
    L.require_device()
    c = CASES[name]
    ref, got = _run(L, c, True), _run(L, c, False)
    valid_rows = None
    if c["mqr"] and c["q_lens"] is not None:        # padded query rows hold unspecified (finite) values in o / dq
        ql = torch.tensor(c["q_lens"], device=DEV)
        valid_rows = (torch.arange(c["Lq"], device=DEV)[None, :] < ql[:, None]).reshape(-1)
    for key in ("o", "lse", "dq", "dk", "dv"):
        a, b = ref[key], got[key]
        assert bool(torch.isfinite(b).all()), key
        if valid_rows is not None and key in ("o", "dq"):
            a, b = a[valid_rows], b[valid_rows]
        err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        assert err < 2e-2, "%s: %s rel err %.3e" % (name, key, err)

"""Packed (variable-length) attention layouts of include/sst.h (SstAttnDesc.q_off / k_off): entry b owns rows
[off[b], off[b] + len[b]) of the token matrices instead of the padded b*L + t.  Checked against the fp32 torch statement of
transformer.py:177-208 (test_kernels_gpu.ref_attention) on each utterance's own rows, and against the padded call."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


@pytest.fixture(scope="module")
def L():
    import sst_b200  # noqa: F401
    from sst_b200 import lib
    lib.require_device()
    return lib


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def _pack(t_pad, lens, Lmax):
    """(B*Lmax, W) padded token matrix -> packed rows + int64 offsets"""
    rows = [t_pad[b * Lmax:b * Lmax + l] for b, l in enumerate(lens)]
    offs, o = [], 0
    for l in lens:
        offs.append(o)
        o += l
    return torch.cat(rows, 0).contiguous(), torch.tensor(offs, dtype=torch.int64, device=t_pad.device), o


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", ["enc_band_1000", "enc_short", "dec_self", "dec_cross"])
def test_packed_attention_matches_reference_rows(L, dtype, case):
    from test_kernels_gpu import ref_attention
    g = torch.Generator(device=DEV).manual_seed(11)
    H, dh = 4, 96
    D = H * dh
    if case == "enc_band_1000":      # ragged encoder batch at the benchmarked length: whole query tiles of the short entries vanish
        lens_q = [1000, 333, 130, 777]; lens_k = lens_q; R, causal = 100, False
    elif case == "enc_short":
        lens_q = [30, 17, 5]; lens_k = lens_q; R, causal = 40, False
    elif case == "dec_self":
        lens_q = [21, 9, 14]; lens_k = lens_q; R, causal = 0, True
    else:
        lens_q = [21, 9, 14]; lens_k = [700, 64, 333]; R, causal = 0, False
    B, Lq, Lk = len(lens_q), max(lens_q), max(lens_k)
    self_attn = case != "dec_cross"
    scale = 1 / math.sqrt(dh)
    q_pad_t = (torch.randn(B * Lq, 3 * D, device=DEV, generator=g) * 0.7).to(dtype)
    kv_pad_t = (torch.randn(B * Lk, 2 * D, device=DEV, generator=g) * 0.7).to(dtype)
    E = (torch.randn(H, 2 * max(R, 1) - 1, dh, device=DEV, generator=g) * dh ** -0.5).to(dtype)
    dO_pad = torch.randn(B * Lq, D, device=DEV, generator=g).to(dtype)
    qp, q_off, nq = _pack(q_pad_t, lens_q, Lq)
    dOp, _, _ = _pack(dO_pad, lens_q, Lq)
    if self_attn:
        q_t, k_t, v_t = qp[:, :D], qp[:, D:2 * D], qp[:, 2 * D:]
        ldq = ldk = ldv = 3 * D
        k_off, nk = q_off, nq
    else:
        kvp, k_off, nk = _pack(kv_pad_t, lens_k, Lk)
        q_t, k_t, v_t = qp[:, :D], kvp[:, :D], kvp[:, D:]
        ldq, ldk, ldv = 3 * D, 2 * D, 2 * D
    ql = torch.tensor(lens_q, dtype=torch.int32, device=DEV)
    kl = torch.tensor(lens_k, dtype=torch.int32, device=DEV)
    canary = 7.0
    o = torch.full((nq, D), canary, device=DEV, dtype=dtype)
    lse = torch.zeros(2 * B * H * Lq, device=DEV)
    d = L.attn_desc(L.dt(qp), B, H, Lq, Lk, dh, ldq, ldk, ldv, D, causal, True, R, scale, 0.0, 0,
                    q_off=q_off, k_off=k_off, q_rows_total=nq, k_rows_total=nk)
    L.attn_fwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse)
    dq_buf = torch.full_like(qp, canary)
    dkv_buf = torch.full_like(qp if self_attn else kvp, canary)
    if self_attn:
        dq_t, dk_t, dv_t = dq_buf[:, :D], dq_buf[:, D:2 * D], dq_buf[:, 2 * D:]
    else:
        dq_t, dk_t, dv_t = dq_buf[:, :D], dkv_buf[:, :D], dkv_buf[:, D:]
    delta = torch.empty(B * H * Lq, device=DEV)
    L.attn_bwd(d, q_t, k_t, v_t, E if R > 0 else None, ql, kl, o, lse, dOp, dq_t, dk_t, dv_t, delta)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all() and not (o == canary).all(dim=1).any()      # every packed row was written
    # each utterance on its own: the unpadded reference
    for b in range(B):
        lq, lk = lens_q[b], lens_k[b]
        qo, ko = int(q_off[b]), int(k_off[b])

        def heads(t, n):
            return t.float().reshape(1, n, H, dh).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
        qh, kh, vh = heads(q_t[qo:qo + lq], lq), heads(k_t[ko:ko + lk], lk), heads(v_t[ko:ko + lk], lk)
        ref = ref_attention(qh, kh, vh, E.float(), R if lk > R else 0, scale, causal, None, None, False) if not (R > 0 and lk <= R) else \
            ref_attention(qh, kh, vh, E.float(), R, scale, causal, None, None, False)
        ref_tok = ref.permute(0, 2, 1, 3).reshape(lq, D)
        assert rel(o[qo:qo + lq], ref_tok.detach()) < TOL[dtype], (case, b)
        ref_tok.backward(dOp[qo:qo + lq].float())
        tg = TOL[dtype]
        assert rel(dq_t[qo:qo + lq], qh.grad.permute(0, 2, 1, 3).reshape(lq, D)) < tg, (case, b, "dq")
        assert rel(dk_t[ko:ko + lk], kh.grad.permute(0, 2, 1, 3).reshape(lk, D)) < tg, (case, b, "dk")
        assert rel(dv_t[ko:ko + lk], vh.grad.permute(0, 2, 1, 3).reshape(lk, D)) < tg, (case, b, "dv")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_shared_memory_forward(L, dtype):
    """k_off[b] = 0 for every entry: n_hyp decoder queries attend to ONE encoder memory (BeamSearch.py:111 without memory.repeat)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    H, dh, n_hyp, Lm, S = 8, 96, 37, 333, 1
    D = H * dh
    scale = 1 / math.sqrt(dh)
    q = (torch.randn(n_hyp * S, D, device=DEV, generator=g) * 0.7).to(dtype)
    kv = (torch.randn(Lm, 2 * D, device=DEV, generator=g) * 0.7).to(dtype)
    zeros = torch.zeros(n_hyp, dtype=torch.int64, device=DEV)
    kl = torch.full((n_hyp,), Lm, dtype=torch.int32, device=DEV)
    o = torch.empty(n_hyp * S, D, device=DEV, dtype=dtype)
    lse = torch.empty(2 * n_hyp * H * S, device=DEV)
    d = L.attn_desc(L.dt(q), n_hyp, H, S, Lm, dh, D, 2 * D, 2 * D, D, False, False, 0, scale, 0.0, 0, k_off=zeros, k_rows_total=Lm)
    L.attn_fwd(d, q, kv, kv[:, D:], None, None, kl, o, lse)
    # the memory.repeat form through the padded layout
    kv_rep = kv.repeat(n_hyp, 1)
    o2 = torch.empty_like(o)
    d2 = L.attn_desc(L.dt(q), n_hyp, H, S, Lm, dh, D, 2 * D, 2 * D, D, False, False, 0, scale, 0.0, 0)
    L.attn_fwd(d2, q, kv_rep, kv_rep[:, D:], None, None, kl, o2, lse)
    assert torch.equal(o, o2)

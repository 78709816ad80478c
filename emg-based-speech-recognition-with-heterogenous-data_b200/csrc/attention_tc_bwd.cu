// Tensor-core attention backward (bf16 operands, fp32 accumulation), the adjoint of attention_tc.cu.  Logits and
// probabilities are recomputed from q, k, E and the saved (row max, log row sum); nothing of size (L x L) touches HBM.
//
//   kernel 1 (dQ)     one CTA = 128 query rows of one (batch, head), loop over the key tiles (64) inside the band:
//        S = Q K^T, PB = Q E_win^T, dP = dO V^T      -> TMEM
//        per query row (one thread): p = exp(s - max - logsum), dS = p * (keep/(1-pd) * dP - delta)
//        dQ += (scale * dS[unmasked]) K              (A = dS tile from shared memory, B = the K tile read MN-major)
//        dQ += dS_rel E_win                          the bias term: dS scattered to its RELATIVE column j-i (the inverse of
//                                                    the forward skew) in a (128 x 192) shared-memory tile
//      also writes delta_i = dO_i . O_i for kernel 2.  No gradient flows to E.
//   kernel 2 (dK, dV) one CTA = 64 keys of one (batch, head), loop over the query tiles (128) inside the band:
//        same S / PB / dP and per-row math, then, transposed so that M = head dim (padded to 128 TMEM lanes), N = keys:
//        dV^T += dO^T P~        dK^T += Q^T (scale * dS[unmasked])      (A = dO / Q tiles read MN-major, B = P~ / dS tiles)
#include "attention_tc.cuh"
#include <stdlib.h>

namespace sst {

namespace attn_tc {

// S = Q K^T, PB = Q E_win^T, dP = dO V^T for one (query tile, key tile) pair; all operands K-major in shared memory.
template <int DH>
__device__ __forceinline__ void issue_s_pb_dp(uint32_t tmem_s, uint32_t tmem_pb, uint32_t tmem_dp, uint32_t qb, uint32_t dob,
                                              uint32_t kb, uint32_t vb, uint32_t eb, bool has_bias) {
  constexpr int KS = DH / 16;
  constexpr int Q_ATOM = BM * 128, K_ATOM = BN * 128, E_ATOM = PBW * 128;
  const uint32_t id_s = ptx::make_idesc_bf16(BM, BN, 0, 0), id_pb = ptx::make_idesc_bf16(BM, PBW, 0, 0);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
    ptx::umma_bf16(tmem_s, ptx::make_smem_desc_sw128(qb + off * Q_ATOM + in, 0, 1024),
                   ptx::make_smem_desc_sw128(kb + off * K_ATOM + in, 0, 1024), id_s, ks > 0);
  }
  if (has_bias) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
      ptx::umma_bf16(tmem_pb, ptx::make_smem_desc_sw128(qb + off * Q_ATOM + in, 0, 1024),
                     ptx::make_smem_desc_sw128(eb + off * E_ATOM + in, 0, 1024), id_pb, ks > 0);
    }
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
    ptx::umma_bf16(tmem_dp, ptx::make_smem_desc_sw128(dob + off * Q_ATOM + in, 0, 1024),
                   ptx::make_smem_desc_sw128(vb + off * K_ATOM + in, 0, 1024), id_s, ks > 0);
  }
}

// Per-row backward math of the thread's CW keys.  In: U = logits, (Lm, Ll) saved softmax statistics, delta.
// Out: U[x] = dropped-out probability P~ (what multiplies dO in dV), W[x] = dS (gradient w.r.t. the logit).
template <int NSPLIT>
__device__ __forceinline__ void tile_backward_row(const AttnTcParams& p, const RowCtx& rc, uint32_t tDP /* incl. lane base */, int j0,
                                                  int hf, float Lm, float Ll, float delta, float (&U)[Split<NSPLIT>::WIN_LD],
                                                  float (&W)[Split<NSPLIT>::CW]) {
  constexpr int CW = Split<NSPLIT>::CW;
  tmem_load_cols(tDP + (uint32_t)(CW * hf), W);
  if (p.thr) {
    float keep[CW];
    dropout_keep<NSPLIT>(p, rc, j0, hf, keep);
#pragma unroll
    for (int x = 0; x < CW; ++x) {
      const float pr = __expf((U[x] - Lm) - Ll);
      W[x] = pr * (W[x] * keep[x] - delta);
      U[x] = pr * keep[x];
    }
  } else {
#pragma unroll
    for (int x = 0; x < CW; ++x) {
      const float pr = __expf((U[x] - Lm) - Ll);
      W[x] = pr * (W[x] - delta);
      U[x] = pr;
    }
  }
}

}  // namespace attn_tc

// ------------------------------------------------------------------------------------------------------------------
// kernel 1: dQ (and delta)
// ------------------------------------------------------------------------------------------------------------------
template <int DH, int NSPLIT>
__global__ void __launch_bounds__(128 * NSPLIT, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                      const __grid_constant__ CUtensorMap tmDO, const attn_tc::AttnTcParams p) {
  using namespace attn_tc;
  using SP = Split<NSPLIT>;
  constexpr int CW = SP::CW;
  constexpr int NATOM = (DH + 63) / 64;
  constexpr int Q_ATOM = BM * 128, K_ATOM = BN * 128, E_ATOM = PBW * 128;
  constexpr int REL_ATOMS = PBW / 64;
  constexpr int OC = DH / NSPLIT;
  constexpr uint32_t TM_S = 0, TM_PB = 64, TM_DP = 256, TM_DQ = 320;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + NATOM * Q_ATOM;
  uint8_t* sK = sDO + NATOM * Q_ATOM;
  uint8_t* sE = sK + NATOM * K_ATOM;
  uint8_t* sV = sE + NATOM * E_ATOM;              // two buffers; buffer (t & 1) is reused for the dS tile once dP is done
  uint8_t* sRel = sV + 2 * NATOM * K_ATOM;        // (128 x 192) dS in relative coordinates, K-major, 3 swizzle atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRel + REL_ATOMS * Q_ATOM);
  uint64_t* bar_q = bars, *bar_ke = bars + 1, *bar_v = bars + 2 /* [2] */, *bar_s = bars + 4, *bar_dq = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* sdelta = reinterpret_cast<float*>(bars + 8);       // [128]
  static_assert(NATOM * K_ATOM == BM * 128, "the dS tile must fit one V buffer");

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = w & 3, hf = w >> 2;
  const int li = 32 * q + lane;
  const int i0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int i = i0 + li;
  const bool leader = threadIdx.x == 0;
  const bool valid = i < p.Lq;

  if (w == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmDO);
      if (p.R > 0) ptx::prefetch_tmap(&tmE);
      for (int k = 0; k < 6; ++k) ptx::mbar_init(&bars[k], 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;

  int t_lo, t_hi;
  key_tile_range(p, i0, t_lo, t_hi);

  const uint32_t ke_bytes = NATOM * K_ATOM + (p.R > 0 ? NATOM * E_ATOM : 0);
  auto load_ke = [&](int t) {
    ptx::mbar_arrive_expect_tx(bar_ke, ke_bytes);
#pragma unroll
    for (int a = 0; a < NATOM; ++a) ptx::tma_load_2d(sK + a * K_ATOM, &tmK, bar_ke, h * DH + a * 64, b * p.Lk + t * BN);
    if (p.R > 0) {
      const int e0 = (t * BN - i0) - (BM - 1) + (p.R - 1);
#pragma unroll
      for (int a = 0; a < NATOM; ++a) ptx::tma_load_2d(sE + a * E_ATOM, &tmE, bar_ke, a * 64, h * (2 * p.R - 1) + e0);
    }
  };
  auto load_v = [&](int t) {
    const int buf = (t - t_lo) & 1;
    ptx::mbar_arrive_expect_tx(&bar_v[buf], NATOM * K_ATOM);
#pragma unroll
    for (int a = 0; a < NATOM; ++a)
      ptx::tma_load_2d(sV + (buf * NATOM + a) * K_ATOM, &tmV, &bar_v[buf], h * DH + a * 64, b * p.Lk + t * BN);
  };

  if (leader) {
    ptx::mbar_arrive_expect_tx(bar_q, 2 * NATOM * Q_ATOM);
#pragma unroll
    for (int a = 0; a < NATOM; ++a) {
      ptx::tma_load_2d(sQ + a * Q_ATOM, &tmQ, bar_q, h * DH + a * 64, b * p.Lq + i0);
      ptx::tma_load_2d(sDO + a * Q_ATOM, &tmDO, bar_q, h * DH + a * 64, b * p.Lq + i0);
    }
    load_ke(t_lo);
    load_v(t_lo);
  }
  __syncwarp();

  // zero the relative-coordinate dS tile once: a row only ever writes its own 64 columns [127 - li, 191 - li)
  {
    const uint32_t rb = ptx::smem_u32(sRel);
    for (int k = threadIdx.x; k < REL_ATOMS * Q_ATOM / 16; k += SP::THREADS) ptx::st_shared_v4(rb + k * 16, 0u, 0u, 0u, 0u);
  }
  // delta_i = dO_i . O_i (column group 0 computes and shares it), saved softmax statistics
  const RowCtx rc = make_row_ctx(p, b, h, i);
  float Lm = 0.f, Ll = 3.0e38f, delta = 0.f;       // rows that do not exist: p = exp(-inf) = 0
  if (valid) {
    const long nrows = (long)p.B * p.H * p.Lq;
    Lm = p.lse[rc.row_id];
    Ll = p.lse[nrows + rc.row_id];
    if (hf == 0) {
      const __nv_bfloat16* orow = p.o + ((long)b * p.Lq + i) * p.ldo + h * DH;
      const __nv_bfloat16* drow = p.dO + ((long)b * p.Lq + i) * p.ldo + h * DH;
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        const uint4 a4 = *reinterpret_cast<const uint4*>(orow + c * 8), b4 = *reinterpret_cast<const uint4*>(drow + c * 8);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a4);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b4);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float2 fa = __bfloat1622float2(a2[x]), fb = __bfloat1622float2(b2[x]);
          delta = fmaf(fa.x, fb.x, delta);
          delta = fmaf(fa.y, fb.y, delta);
        }
      }
      p.delta[rc.row_id] = delta;
    }
  }
  if (hf == 0) sdelta[li] = delta;
  __syncthreads();
  delta = sdelta[li];

  const uint32_t qb = ptx::smem_u32(sQ), dob = ptx::smem_u32(sDO), kb = ptx::smem_u32(sK), eb = ptx::smem_u32(sE);
  if (leader) {
    ptx::mbar_wait(bar_q, 0);
    ptx::mbar_wait(bar_ke, 0);
    ptx::mbar_wait(&bar_v[0], 0);
    ptx::tc_fence_after();
    issue_s_pb_dp<DH>(tmem + TM_S, tmem + TM_PB, tmem + TM_DP, qb, dob, kb, ptx::smem_u32(sV), eb, p.R > 0);
    ptx::umma_commit(bar_s);
  }
  __syncwarp();

  // shared-memory address of this thread's CW relative columns c = (127 - li + CW*hf) + x in the swizzled dS_rel tile
  uint32_t rel_addr[CW];
  {
    const uint32_t rb = ptx::smem_u32(sRel) + li * 128;
    const int c0 = (BM - 1) - li + CW * hf;
#pragma unroll
    for (int x = 0; x < CW; ++x) {
      const int c = c0 + x;
      rel_addr[x] = rb + (uint32_t)(c >> 6) * Q_ATOM + ((uint32_t)(((c & 63) >> 3) ^ (li & 7)) << 4) + (uint32_t)(c & 7) * 2;
    }
  }
  uint32_t ph_s = 0, ph_ke = 1, ph_dq = 0;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int buf = (t - t_lo) & 1;
    ptx::mbar_wait(bar_s, ph_s);
    ph_s ^= 1u;
    ptx::tc_fence_after();
    if (leader && t < t_hi) load_v(t + 1);
    __syncwarp();

    const bool skip = block_out_of_band<NSPLIT>(p, i0 + 32 * q, t * BN + CW * hf);
    if (!skip) {
      float U[SP::WIN_LD], W[CW];
      uint32_t mbits;
      const bool simple = tile_is_simple(p, rc, t * BN);
      tile_logits<NSPLIT>(p, rc, tmem + TM_S + lane_base, tmem + TM_PB + lane_base, q, hf, lane, t * BN, simple, U, mbits);
      tile_backward_row<NSPLIT>(p, rc, tmem + TM_DP + lane_base, t * BN, hf, Lm, Ll, delta, U, W);
      // bias term: dS at its relative column c = lj - li + 127 (zero outside the band); addresses are tile-independent
      if (p.R > 0) {
        const int d0 = t * BN + CW * hf - i + p.R - 1;
        const uint32_t lim = (uint32_t)(2 * p.R - 1);
#pragma unroll
        for (int x = 0; x < CW; ++x) {
          const float v = ((uint32_t)(d0 + x) < lim) ? W[x] : 0.f;
          const __nv_bfloat16 hv = __float2bfloat16_rn(v);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(rel_addr[x]), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
        }
      }
      // q.k term: scale * dS where the term was not masked
#pragma unroll
      for (int x = 0; x < CW; ++x) W[x] = ((mbits >> x) & 1u) ? 0.f : W[x] * p.scale;
      store_cols_bf16_sw128<CW>(ptx::smem_u32(sV + buf * NATOM * K_ATOM), li, CW * hf, W);
    } else {
      if (p.R > 0) {
#pragma unroll
        for (int x = 0; x < CW; ++x) asm volatile("st.shared.u16 [%0], %1;" ::"r"(rel_addr[x]), "h"((uint16_t)0) : "memory");
      }
      store_zero_cols_sw128<CW>(ptx::smem_u32(sV + buf * NATOM * K_ATOM), li, CW * hf);
    }

    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (leader) {
      ptx::tc_fence_after();
      const uint32_t dsb = ptx::smem_u32(sV + buf * NATOM * K_ATOM), rb = ptx::smem_u32(sRel);
      const uint32_t id_dq = ptx::make_idesc_bf16(BM, DH, 0, 1);
#pragma unroll
      for (int ks = 0; ks < BN / 16; ++ks)
        ptx::umma_bf16(tmem + TM_DQ, ptx::make_smem_desc_sw128(dsb + ks * 32, 0, 1024),
                       ptx::make_smem_desc_sw128(kb + ks * 2048, K_ATOM, 1024), id_dq, (t > t_lo || ks > 0) ? 1u : 0u);
      if (p.R > 0) {
#pragma unroll
        for (int ks = 0; ks < PBW / 16; ++ks)
          ptx::umma_bf16(tmem + TM_DQ, ptx::make_smem_desc_sw128(rb + (ks >> 2) * Q_ATOM + (ks & 3) * 32, 0, 1024),
                         ptx::make_smem_desc_sw128(eb + ks * 2048, E_ATOM, 1024), id_dq, 1u);
      }
      ptx::umma_commit(bar_dq);
      if (t < t_hi) {
        ptx::mbar_wait(bar_dq, ph_dq);             // K, E, dS and dS_rel tiles are free again
        load_ke(t + 1);
        ptx::mbar_wait(bar_ke, ph_ke);
        ph_ke ^= 1u;
        ptx::mbar_wait(&bar_v[buf ^ 1], ((t + 1 - t_lo) >> 1) & 1);
        ptx::tc_fence_after();
        issue_s_pb_dp<DH>(tmem + TM_S, tmem + TM_PB, tmem + TM_DP, qb, dob, kb, ptx::smem_u32(sV + (buf ^ 1) * NATOM * K_ATOM), eb,
                          p.R > 0);
        ptx::umma_commit(bar_s);
      }
    }
    ph_dq ^= 1u;
    __syncwarp();
  }

  ptx::mbar_wait(bar_dq, (uint32_t)((t_hi - t_lo) & 1));
  ptx::tc_fence_after();
  tmem_row_to_global<OC>(tmem + TM_DQ + hf * OC + lane_base, p.dq + ((long)b * p.Lq + i) * p.ldq + h * DH + hf * OC, 1.f, valid);
  ptx::tc_fence_before();
  __syncthreads();
  if (w == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 2: dK, dV
// ------------------------------------------------------------------------------------------------------------------
template <int DH, int NSPLIT>
__global__ void __launch_bounds__(128 * NSPLIT, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                       const __grid_constant__ CUtensorMap tmDO, const attn_tc::AttnTcParams p) {
  using namespace attn_tc;
  using SP = Split<NSPLIT>;
  constexpr int CW = SP::CW;
  constexpr int NATOM = (DH + 63) / 64;
  constexpr int Q_ATOM = BM * 128, K_ATOM = BN * 128, E_ATOM = PBW * 128;
  constexpr uint32_t TM_S = 0, TM_PB = 64, TM_DP = 256, TM_DV = 320, TM_DK = 384;
  static_assert(NATOM == 2, "M = 128 TMEM lanes are fed from two 64-wide head-dim groups");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + NATOM * K_ATOM;
  uint8_t* sQ = sV + NATOM * K_ATOM;
  uint8_t* sDO = sQ + NATOM * Q_ATOM;
  uint8_t* sE = sDO + NATOM * Q_ATOM;
  uint8_t* sP = sE + NATOM * E_ATOM;
  uint8_t* sdS = sP + BM * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + BM * 128);
  uint64_t* bar_kv = bars, *bar_in = bars + 1, *bar_s = bars + 2, *bar_acc = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = w & 3, hf = w >> 2;
  const int li = 32 * q + lane;
  const int j0 = blockIdx.x * BN, h = blockIdx.y, b = blockIdx.z;
  const bool leader = threadIdx.x == 0;

  if (w == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmDO);
      if (p.R > 0) ptx::prefetch_tmap(&tmE);
      for (int k = 0; k < 4; ++k) ptx::mbar_init(&bars[k], 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;

  int q_lo, q_hi;
  query_tile_range(p, j0, q_lo, q_hi);

  const uint32_t in_bytes = 2 * NATOM * Q_ATOM + (p.R > 0 ? NATOM * E_ATOM : 0);
  auto load_in = [&](int u) {
    const int i0 = u * BM;
    ptx::mbar_arrive_expect_tx(bar_in, in_bytes);
#pragma unroll
    for (int a = 0; a < NATOM; ++a) {
      ptx::tma_load_2d(sQ + a * Q_ATOM, &tmQ, bar_in, h * DH + a * 64, b * p.Lq + i0);
      ptx::tma_load_2d(sDO + a * Q_ATOM, &tmDO, bar_in, h * DH + a * 64, b * p.Lq + i0);
    }
    if (p.R > 0) {
      const int e0 = (j0 - i0) - (BM - 1) + (p.R - 1);
#pragma unroll
      for (int a = 0; a < NATOM; ++a) ptx::tma_load_2d(sE + a * E_ATOM, &tmE, bar_in, a * 64, h * (2 * p.R - 1) + e0);
    }
  };

  const uint32_t qb = ptx::smem_u32(sQ), dob = ptx::smem_u32(sDO), kb = ptx::smem_u32(sK), vb = ptx::smem_u32(sV),
                 eb = ptx::smem_u32(sE), pb = ptx::smem_u32(sP), dsb = ptx::smem_u32(sdS);
  if (leader) {
    ptx::mbar_arrive_expect_tx(bar_kv, 2 * NATOM * K_ATOM);
#pragma unroll
    for (int a = 0; a < NATOM; ++a) {
      ptx::tma_load_2d(sK + a * K_ATOM, &tmK, bar_kv, h * DH + a * 64, b * p.Lk + j0);
      ptx::tma_load_2d(sV + a * K_ATOM, &tmV, bar_kv, h * DH + a * 64, b * p.Lk + j0);
    }
    load_in(q_lo);
    ptx::mbar_wait(bar_kv, 0);
    ptx::mbar_wait(bar_in, 0);
    ptx::tc_fence_after();
    issue_s_pb_dp<DH>(tmem + TM_S, tmem + TM_PB, tmem + TM_DP, qb, dob, kb, vb, eb, p.R > 0);
    ptx::umma_commit(bar_s);
  }
  __syncwarp();

  const long nrows = (long)p.B * p.H * p.Lq;
  uint32_t ph_s = 0, ph_in = 1, ph_acc = 0;
  for (int u = q_lo; u <= q_hi; ++u) {
    const int i = u * BM + li;
    const bool valid = i < p.Lq;
    const RowCtx rc = make_row_ctx(p, b, h, i);
    float Lm = 0.f, Ll = 3.0e38f, delta = 0.f;
    if (valid) { Lm = p.lse[rc.row_id]; Ll = p.lse[nrows + rc.row_id]; delta = p.delta[rc.row_id]; }

    ptx::mbar_wait(bar_s, ph_s);
    ph_s ^= 1u;
    ptx::tc_fence_after();

    if (!block_out_of_band<NSPLIT>(p, u * BM + 32 * q, j0 + CW * hf)) {
      float U[SP::WIN_LD], W[CW];
      uint32_t mbits;
      const bool simple = tile_is_simple(p, rc, j0);
      tile_logits<NSPLIT>(p, rc, tmem + TM_S + lane_base, tmem + TM_PB + lane_base, q, hf, lane, j0, simple, U, mbits);
      tile_backward_row<NSPLIT>(p, rc, tmem + TM_DP + lane_base, j0, hf, Lm, Ll, delta, U, W);
#pragma unroll
      for (int x = 0; x < CW; ++x) W[x] = ((mbits >> x) & 1u) ? 0.f : W[x] * p.scale;
      store_cols_bf16_sw128<CW>(pb, li, CW * hf, U);
      store_cols_bf16_sw128<CW>(dsb, li, CW * hf, W);
    } else {
      store_zero_cols_sw128<CW>(pb, li, CW * hf);
      store_zero_cols_sw128<CW>(dsb, li, CW * hf);
    }

    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (leader) {
      ptx::tc_fence_after();
      const uint32_t id_t = ptx::make_idesc_bf16(128, BN, 1, 1);       // M = head dim (two 64-wide groups), N = keys
#pragma unroll
      for (int ks = 0; ks < BM / 16; ++ks)
        ptx::umma_bf16(tmem + TM_DV, ptx::make_smem_desc_sw128(dob + ks * 2048, Q_ATOM, 1024),
                       ptx::make_smem_desc_sw128(pb + ks * 2048, Q_ATOM, 1024), id_t, (u > q_lo || ks > 0) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < BM / 16; ++ks)
        ptx::umma_bf16(tmem + TM_DK, ptx::make_smem_desc_sw128(qb + ks * 2048, Q_ATOM, 1024),
                       ptx::make_smem_desc_sw128(dsb + ks * 2048, Q_ATOM, 1024), id_t, (u > q_lo || ks > 0) ? 1u : 0u);
      ptx::umma_commit(bar_acc);
      if (u < q_hi) {
        ptx::mbar_wait(bar_acc, ph_acc);            // Q, dO, E, P~ and dS tiles are free again
        load_in(u + 1);
        ptx::mbar_wait(bar_in, ph_in);
        ph_in ^= 1u;
        ptx::tc_fence_after();
        issue_s_pb_dp<DH>(tmem + TM_S, tmem + TM_PB, tmem + TM_DP, qb, dob, kb, vb, eb, p.R > 0);
        ptx::umma_commit(bar_s);
      }
    }
    ph_acc ^= 1u;
    __syncwarp();
  }

  ptx::mbar_wait(bar_acc, (uint32_t)((q_hi - q_lo) & 1));
  ptx::tc_fence_after();
  {
    // TMEM lane = head-dim index d, column = key: transposed stores (a warp writes 32 consecutive d of one key row);
    // column group hf stores keys [CW*hf, CW*hf + CW)
    const int dcol = li;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* out = which == 0 ? p.dv : p.dk;
      const long ld = which == 0 ? p.ldv : p.ldk;
      float r[CW];
      tmem_load_cols(tmem + (which == 0 ? TM_DV : TM_DK) + CW * hf + lane_base, r);
      if (dcol < DH) {
#pragma unroll
        for (int x = 0; x < CW; ++x) {
          const int j = j0 + CW * hf + x;
          if (j < p.Lk) out[((long)b * p.Lk + j) * ld + h * DH + dcol] = __float2bfloat16_rn(r[x]);
        }
      }
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (w == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

template <int NSPLIT>
static int attn_bwd_tc_launch_n(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                                const int* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                                float* delta, cudaStream_t st) {
  using namespace attn_tc;
  constexpr int DH = 96;
  constexpr int NATOM = 2;
  AttnTcParams p = make_tc_params(d, q_lens, k_lens);
  p.o = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(o)); p.ldo = d.ldo; p.lse = const_cast<float*>(lse);
  p.dO = reinterpret_cast<const __nv_bfloat16*>(dO); p.delta = delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.ldq = d.ldq; p.ldk = d.ldk; p.ldv = d.ldv;
  CUtensorMap tmQ, tmK, tmV, tmE, tmDO;
  int rc;
  const long HD = (long)d.H * d.dh;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, HD, (long)d.B * d.Lq, d.ldq, 64, BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dO, HD, (long)d.B * d.Lq, d.ldo, 64, BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, HD, (long)d.B * d.Lk, d.ldk, 64, BN))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, HD, (long)d.B * d.Lk, d.ldv, 64, BN))) return rc;
  if (d.rel_dist > 0) {
    if ((rc = make_tmap_bf16_2d(&tmE, E, d.dh, (long)d.H * (2 * d.rel_dist - 1), d.dh, 64, PBW))) return rc;
  } else {
    tmE = tmK;
  }
  constexpr int SMEM_DQ = 2 * NATOM * BM * 128 + NATOM * BN * 128 + NATOM * PBW * 128 + 2 * NATOM * BN * 128 + (PBW / 64) * BM * 128 +
                          1024 + 64 + 128 * 4;
  constexpr int SMEM_DKV = 2 * NATOM * BN * 128 + 2 * NATOM * BM * 128 + NATOM * PBW * 128 + 2 * BM * 128 + 1024 + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<DH, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<DH, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV);
    SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute(attn_bwd_tc): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  attn_bwd_dq_tc_kernel<DH, NSPLIT><<<dim3(cdiv(d.Lq, BM), d.H, d.B), 128 * NSPLIT, SMEM_DQ, st>>>(tmQ, tmK, tmV, tmE, tmDO, p);
  attn_bwd_dkv_tc_kernel<DH, NSPLIT><<<dim3(cdiv(d.Lk, BN), d.H, d.B), 128 * NSPLIT, SMEM_DKV, st>>>(tmQ, tmK, tmV, tmE, tmDO, p);
  return check_launch("attn_bwd_tc", 2);
}

// column groups per tile (threads = 128 * groups): 4 by default, SST_ATTN_NSPLIT=2 selects the 256-thread variant
int attn_tc_nsplit() {
  static int n = 0;
  if (!n) {
    const char* e = getenv("SST_ATTN_NSPLIT");
    n = (e && e[0] == '2') ? 2 : 4;
  }
  return n;
}

int attn_bwd_tc_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                       const int* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                       float* delta, cudaStream_t st) {
  if (attn_tc_nsplit() == 2)
    return attn_bwd_tc_launch_n<2>(d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, st);
  return attn_bwd_tc_launch_n<4>(d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, st);
}

}  // namespace sst

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
//      and hands every (query tile, key tile) pair's P~ and scale*dS[unmasked] to kernel 2 as bf16 (128 x 64) tiles in a
//      workspace (the expensive per-element recompute is NOT repeated).
//   kernel 2 (dK, dV) one CTA = 64 keys of one (batch, head), loop over the query tiles (128) inside the band: a pure
//      TMA -> tcgen05 pipeline (double-buffered), transposed so that M = head dim (padded to 128 TMEM lanes), N = keys:
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
__device__ __forceinline__ void tile_backward_row(const AttnTcParams& p, unsigned long long seed, const RowCtx& rc,
                                                  uint32_t tDP /* incl. lane base */, int j0,
                                                  int hf, float Lm, float Ll, float delta, float (&U)[Split<NSPLIT>::WIN_LD],
                                                  float (&W)[Split<NSPLIT>::CW]) {
  constexpr int CW = Split<NSPLIT>::CW;
  if (tDP != 0xffffffffu) tmem_load_cols(tDP + (uint32_t)(CW * hf), W);      // 0xffffffff: W already holds the dP slice
  // p = exp((s - max) - logsum): two subtractions in fp32 (max can be -1e8, where max + logsum would swallow logsum)
  if (p.thr) {
    float keep[CW];
    dropout_keep<NSPLIT>(p, seed, rc, j0, hf, keep);
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

template <int CW, int N>
__device__ __forceinline__ void store_cols_bf16_global(__nv_bfloat16* dst, const float (&U)[N]) {
#pragma unroll
  for (int c = 0; c < CW / 8; ++c) {
    uint4 o;
    __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int x = 0; x < 4; ++x) o2[x] = __floats2bfloat162_rn(U[c * 8 + 2 * x], U[c * 8 + 2 * x + 1]);
    *reinterpret_cast<uint4*>(dst + c * 8) = o;
  }
}
template <int CW>
__device__ __forceinline__ void store_zero_cols_global(__nv_bfloat16* dst) {
#pragma unroll
  for (int c = 0; c < CW / 8; ++c) *reinterpret_cast<uint4*>(dst + c * 8) = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace attn_tc

// ------------------------------------------------------------------------------------------------------------------
// kernel 1: dQ (and delta)
// ------------------------------------------------------------------------------------------------------------------
// Warp-specialised: 4*NSPLIT compute warps + 1 issuer warp.  Q and dO live in TMEM (A operands read from tensor memory,
// 48 columns each) which leaves shared memory for double-buffered K / E_win / V stages; the compute warps release the
// S / PB / dP accumulators as soon as their slices are in registers, so tile t+1's MMAs and tile t+2's TMA loads run under
// tile t's per-element math.  TMEM: S 0-63 | PB 64-255 | dP 256-319 | dQ 320-415 | Q 416-463 | dO 464-511.
template <int DH, int NSPLIT>
__global__ void __launch_bounds__(128 * NSPLIT + 32, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmE, const attn_tc::AttnTcParams p, const __nv_bfloat16* __restrict__ qg) {
  using namespace attn_tc;
  using SP = Split<NSPLIT>;
  constexpr int CW = SP::CW;
  constexpr int NW = 4 * NSPLIT;
  constexpr int KS = DH / 16;
  constexpr int NATOM = (DH + 63) / 64;
  constexpr int Q_ATOM = BM * 128, K_ATOM = BN * 128, E_ATOM = PBW * 128;
  constexpr int REL_ATOMS = PBW / 64;
  constexpr int OC = DH / NSPLIT;
  constexpr uint32_t TM_S = 0, TM_PB = 64, TM_DP = 256, TM_DQ = 320, TM_Q = 416, TM_DO = TM_Q + DH / 2;
  static_assert(TM_DO + DH / 2 <= 512 && NSPLIT >= 2, "TMEM budget / roles");
  static_assert(NATOM * K_ATOM == BM * 128, "the dS tile must fit one V buffer");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                             // [2]
  uint8_t* sE = sK + 2 * NATOM * K_ATOM;          // [2]
  uint8_t* sV = sE + 2 * NATOM * E_ATOM;          // [2]; buffer (t & 1) is reused for the dS tile once dP(t) is done
  uint8_t* sRel = sV + 2 * NATOM * K_ATOM;        // (128 x 192) dS in relative coordinates, K-major, 3 swizzle atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRel + REL_ATOMS * Q_ATOM);
  uint64_t* bar_ke = bars /* [2] */, *bar_v = bars + 2 /* [2] */, *bar_s = bars + 4, *bar_dq = bars + 5, *bar_sfree = bars + 6,
           *bar_p = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* sdelta = reinterpret_cast<float*>(bars + 10);      // [128]

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  if (!q_tile_exists(p, b, i0)) return;           // packed layout: no such rows (the dK/dV kernel does not visit this tile either)
  const int qb = q_base(p, b), kb = k_base(p, b), n_q = q_rows(p, b);

  // The Q / dO / O rows the compute warps park in TMEM below were written by the forward pass long ago: pull them towards
  // L2 now (no registers involved), so that their latency runs under the barrier / TMEM set-up and the first block barrier.
  if (w < 12 && i0 + 32 * (w & 3) + lane < n_q) {
    const long row = (long)qb + i0 + 32 * (w & 3) + lane;
    const int which = w >> 2;                       // 0: Q, 1: dO, 2: O
    const char* rp = reinterpret_cast<const char*>(which == 0 ? qg + row * p.ldq : (which == 1 ? p.dO : p.o) + row * p.ldo) + h * DH * 2;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 128));      // DH * 2 = 192 bytes: two 128-byte lines cover a row
  }

  if (w == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
      if (p.R > 0) ptx::prefetch_tmap(&tmE);
      for (int k = 0; k < 6; ++k) ptx::mbar_init(&bars[k], 1);
      ptx::mbar_init(bar_sfree, NW);
      ptx::mbar_init(bar_p, NW);
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
  pdl_wait();                                     // dO comes from the kernel right in front of this one (sst_common.cuh)
  pdl_trigger();

  int t_lo, t_hi, t_hi_full;
  key_tile_range(p, i0, t_lo, t_hi_full);
  key_tile_range_valid(p, i0, b, t_lo, t_hi);        // tiles (t_hi, t_hi_full] are pure padding: their hand-off tiles are zeros

  const uint32_t ke_bytes = NATOM * K_ATOM + (p.R > 0 ? NATOM * E_ATOM : 0);
  auto load_ke = [&](int t) {
    const int st = (t - t_lo) & 1;
    ptx::mbar_arrive_expect_tx(&bar_ke[st], ke_bytes);
#pragma unroll
    for (int a = 0; a < NATOM; ++a)
      ptx::tma_load_2d(sK + (st * NATOM + a) * K_ATOM, &tmK, &bar_ke[st], h * DH + a * 64, kb + t * BN);
    if (p.R > 0) {
      const int e0 = (t * BN - i0) - (BM - 1) + (p.R - 1);
#pragma unroll
      for (int a = 0; a < NATOM; ++a)
        ptx::tma_load_2d(sE + (st * NATOM + a) * E_ATOM, &tmE, &bar_ke[st], a * 64, h * (2 * p.R - 1) + e0);
    }
  };
  auto load_v = [&](int t) {
    const int st = (t - t_lo) & 1;
    ptx::mbar_arrive_expect_tx(&bar_v[st], NATOM * K_ATOM);
#pragma unroll
    for (int a = 0; a < NATOM; ++a)
      ptx::tma_load_2d(sV + (st * NATOM + a) * K_ATOM, &tmV, &bar_v[st], h * DH + a * 64, kb + t * BN);
  };
  const bool issuer = (w == NW);
  if (issuer && lane == 0) {        // the first loads do not depend on anything the compute warps set up
    load_ke(t_lo);
    load_v(t_lo);
    if (t_lo < t_hi) { load_ke(t_lo + 1); load_v(t_lo + 1); }
  }

  // ---- one-time setup by the compute warps: Q and dO rows into TMEM, delta, zeroed dS_rel tile ----
  const int q = w & 3, hf = w >> 2;
  const int li = 32 * q + lane;
  const int i = i0 + li;
  const bool valid = !issuer && i < n_q;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const RowCtx rc = make_row_ctx(p, b, h, issuer ? 0 : i);
  float Lm = 0.f, Ll = 3.0e38f, delta = 0.f;       // rows that do not exist: p = exp(-inf) = 0
  if (!issuer) {
    const uint32_t rb = ptx::smem_u32(sRel);
    for (int k = threadIdx.x; k < REL_ATOMS * Q_ATOM / 16; k += SP::THREADS) ptx::st_shared_v4(rb + k * 16, 0u, 0u, 0u, 0u);
    if (valid) {
      const long nrows = (long)p.B * p.H * p.Lq;
      Lm = p.lse[rc.row_id];
      Ll = p.lse[nrows + rc.row_id];
    }
    if (hf < 2) {                                   // column group 0 moves the Q row, group 1 the dO row (and forms delta)
      const __nv_bfloat16* src = (hf == 0 ? qg + ((long)qb + i) * p.ldq : p.dO + ((long)qb + i) * p.ldo) + h * DH;
      const __nv_bfloat16* orow = p.o + ((long)qb + i) * p.ldo + h * DH;
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {           // 32 bf16 = 16 TMEM columns per store
        uint32_t r[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
          if (valid) v4 = *reinterpret_cast<const uint4*>(src + c * 32 + g * 8);
          r[g * 4 + 0] = v4.x; r[g * 4 + 1] = v4.y; r[g * 4 + 2] = v4.z; r[g * 4 + 3] = v4.w;
          if (hf == 1 && valid) {
            const uint4 o4 = *reinterpret_cast<const uint4*>(orow + c * 32 + g * 8);
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&o4);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&v4);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              const float2 fa = __bfloat1622float2(a2[x]), fb = __bfloat1622float2(b2[x]);
              delta = fmaf(fa.x, fb.x, delta);
              delta = fmaf(fa.y, fb.y, delta);
            }
          }
        }
        ptx::tmem_st_32x32b_x16(tmem + (hf == 0 ? TM_Q : TM_DO) + c * 16 + lane_base, r);
      }
      ptx::tmem_st_wait();
      if (hf == 1) {
        sdelta[li] = delta;
        if (valid) p.delta[rc.row_id] = delta;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  if (issuer) {
    // ============================================ issuer (one thread) ============================================
    if (lane == 0) {
      auto issue_s = [&](int st) {   // S = Q K^T, PB = Q E_win^T, dP = dO V^T from stage `st`; A operands from TMEM
        const uint32_t kb = ptx::smem_u32(sK + st * NATOM * K_ATOM), eb = ptx::smem_u32(sE + st * NATOM * E_ATOM),
                       vb = ptx::smem_u32(sV + st * NATOM * K_ATOM);
        const uint32_t id_s = ptx::make_idesc_bf16(BM, BN, 0, 0), id_pb = ptx::make_idesc_bf16(BM, PBW, 0, 0);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
          ptx::umma_bf16_ts(tmem + TM_S, tmem + TM_Q + ks * 8, ptx::make_smem_desc_sw128(kb + off * K_ATOM + in, 0, 1024), id_s, ks > 0);
        }
        if (p.R > 0) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
            ptx::umma_bf16_ts(tmem + TM_PB, tmem + TM_Q + ks * 8, ptx::make_smem_desc_sw128(eb + off * E_ATOM + in, 0, 1024), id_pb, ks > 0);
          }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
          ptx::umma_bf16_ts(tmem + TM_DP, tmem + TM_DO + ks * 8, ptx::make_smem_desc_sw128(vb + off * K_ATOM + in, 0, 1024), id_s, ks > 0);
        }
        ptx::umma_commit(bar_s);
      };
      ptx::mbar_wait(&bar_ke[0], 0);
      ptx::mbar_wait(&bar_v[0], 0);
      ptx::tc_fence_after();
      issue_s(0);
      for (int t = t_lo; t <= t_hi; ++t) {
        const int k = t - t_lo, st = k & 1;
        if (t < t_hi) {
          ptx::mbar_wait(bar_sfree, (uint32_t)(k & 1));     // S / PB / dP slices are in registers
          ptx::mbar_wait(&bar_ke[st ^ 1], (uint32_t)(((k + 1) >> 1) & 1));
          ptx::mbar_wait(&bar_v[st ^ 1], (uint32_t)(((k + 1) >> 1) & 1));
          ptx::tc_fence_after();
          issue_s(st ^ 1);
        }
        ptx::mbar_wait(bar_p, (uint32_t)(k & 1));           // dS tile (in V buffer st) and dS_rel tile written
        ptx::tc_fence_after();
        const uint32_t dsb = ptx::smem_u32(sV + st * NATOM * K_ATOM), rb = ptx::smem_u32(sRel),
                       kb = ptx::smem_u32(sK + st * NATOM * K_ATOM), eb = ptx::smem_u32(sE + st * NATOM * E_ATOM);
        const uint32_t id_dq = ptx::make_idesc_bf16(BM, DH, 0, 1);
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks)
          ptx::umma_bf16(tmem + TM_DQ, ptx::make_smem_desc_sw128(dsb + ks * 32, 0, 1024),
                         ptx::make_smem_desc_sw128(kb + ks * 2048, K_ATOM, 1024), id_dq, (k > 0 || ks > 0) ? 1u : 0u);
        if (p.R > 0) {
#pragma unroll
          for (int ks = 0; ks < PBW / 16; ++ks)
            ptx::umma_bf16(tmem + TM_DQ, ptx::make_smem_desc_sw128(rb + (ks >> 2) * Q_ATOM + (ks & 3) * 32, 0, 1024),
                           ptx::make_smem_desc_sw128(eb + ks * 2048, E_ATOM, 1024), id_dq, 1u);
        }
        ptx::umma_commit(bar_dq);
        if (t + 2 <= t_hi) {
          ptx::mbar_wait(bar_dq, (uint32_t)(k & 1));        // stage st (K, E, dS-in-V) is free again
          load_ke(t + 2);
          load_v(t + 2);
        }
      }
    }
  } else {
    // ============================================ compute warps ==================================================
    delta = sdelta[li];
    const unsigned long long seed = p.thr ? salted(p.seed, p.salt) : 0ull;
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
    // this thread's slice (row li, columns [CW*hf, +CW)) of the query tile's first hand-off tile
    const long tile0 = (((long)b * p.H + h) * p.nQT + blockIdx.x) * p.max_kt;
    __nv_bfloat16* tile_p = p.ws_p + tile0 * BM * BN + li * BN + CW * hf;
    __nv_bfloat16* tile_ds = p.ws_ds + tile0 * BM * BN + li * BN + CW * hf;

    for (int t = t_lo; t <= t_hi; ++t) {
      const int k = t - t_lo, st = k & 1;
      ptx::mbar_wait(bar_s, (uint32_t)(k & 1));
      ptx::tc_fence_after();
      const bool skip = block_out_of_band<NSPLIT>(p, i0 + 32 * q, t * BN + CW * hf);
      float U[SP::WIN_LD], W[CW];
      uint32_t mbits = 0;
      if (!skip) {
        const bool simple = tile_is_simple(p, rc, t * BN);
        tile_logits<NSPLIT>(p, rc, tmem + TM_S + lane_base, tmem + TM_PB + lane_base, q, hf, lane, t * BN, simple, U, mbits);
        tmem_load_cols(tmem + TM_DP + lane_base + (uint32_t)(CW * hf), W);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_sfree);    // the issuer may overwrite S / PB / dP with the next tile

      if (!skip) tile_backward_row<NSPLIT>(p, seed, rc, 0xffffffffu, t * BN, hf, Lm, Ll, delta, U, W);
      if (k > 0) ptx::mbar_wait(bar_dq, (uint32_t)((k - 1) & 1));     // dQ(t-1) has finished reading dS_rel / the dS tile
      if (!skip) {
        // bias term: dS at its relative column c = lj - li + 127 (zero outside the band); addresses are tile-independent
        if (p.R > 0) {
          const uint32_t band = band_bits<CW>(p, rc, t * BN + CW * hf);
#pragma unroll
          for (int x = 0; x < CW; ++x) {
            const float v = ((band >> x) & 1u) ? W[x] : 0.f;
            const __nv_bfloat16 hv = __float2bfloat16_rn(v);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(rel_addr[x]), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
          }
        }
        // q.k term: scale * dS where the term was not masked
#pragma unroll
        for (int x = 0; x < CW; ++x) W[x] = ((mbits >> x) & 1u) ? 0.f : W[x] * p.scale;
        store_cols_bf16_sw128<CW>(ptx::smem_u32(sV + st * NATOM * K_ATOM), li, CW * hf, W);
        store_cols_bf16_global<CW>(tile_p + (long)k * BM * BN, U);
        store_cols_bf16_global<CW>(tile_ds + (long)k * BM * BN, W);
      } else {
        if (p.R > 0) {
#pragma unroll
          for (int x = 0; x < CW; ++x) asm volatile("st.shared.u16 [%0], %1;" ::"r"(rel_addr[x]), "h"((uint16_t)0) : "memory");
        }
        store_zero_cols_sw128<CW>(ptx::smem_u32(sV + st * NATOM * K_ATOM), li, CW * hf);
        store_zero_cols_global<CW>(tile_p + (long)k * BM * BN);
        store_zero_cols_global<CW>(tile_ds + (long)k * BM * BN);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_p);
    }

    for (int t = t_hi + 1; t <= t_hi_full; ++t) {      // the dK/dV kernel walks the full range: padding tiles contribute nothing
      store_zero_cols_global<CW>(tile_p + (long)(t - t_lo) * BM * BN);
      store_zero_cols_global<CW>(tile_ds + (long)(t - t_lo) * BM * BN);
    }
    ptx::mbar_wait(bar_dq, (uint32_t)((t_hi - t_lo) & 1));
    ptx::tc_fence_after();
    tmem_row_to_global<OC>(tmem + TM_DQ + hf * OC + lane_base, p.dq + ((long)qb + i) * p.ldq + h * DH + hf * OC, 1.f, valid);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (w == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 2: dK, dV from the hand-off tiles
// ------------------------------------------------------------------------------------------------------------------
// Persistent: grid = #SMs, every CTA walks the (batch, head, key tile) items blockIdx.x, blockIdx.x + gridDim.x, ...;
// the (item, query tile) pairs form one stream through a two-stage shared-memory ring, the dV^T / dK^T accumulators are
// double-buffered in TMEM so that the epilogue of one item runs under the MMAs of the next.
struct DkvCursor {
  int item, u, q_hi, b, h, kt;
  bool first;     // first query tile of its item
};

template <int DH>
__global__ void __launch_bounds__(160, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmDS,
                       const attn_tc::AttnTcParams p, int n_kt, int n_items) {
  using namespace attn_tc;
  constexpr int NATOM = (DH + 63) / 64;
  constexpr int Q_ATOM = BM * 128;
  constexpr int STAGE = 2 * NATOM * Q_ATOM + 2 * BM * 128;       // Q, dO (two 64-column groups each), P~, dS
  constexpr int OUT_BYTES = BN * DH * 2;                         // one (64 keys x dh) bf16 output tile
  static_assert(NATOM == 2, "M = 128 TMEM lanes are fed from two 64-wide head-dim groups");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOut = smem + 2 * STAGE;                              // dV tile then dK tile, row = key, dh contiguous
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * OUT_BYTES);
  uint64_t* bar_full = bars /* [2] */, *bar_done = bars + 2 /* [2] */, *bar_acc = bars + 4 /* [2] */, *bar_free = bars + 6 /* [2] */;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (w == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmDO); ptx::prefetch_tmap(&tmP); ptx::prefetch_tmap(&tmDS);
      for (int k = 0; k < 6; ++k) ptx::mbar_init(&bars[k], 1);
      ptx::mbar_init(&bar_free[0], 4);
      ptx::mbar_init(&bar_free[1], 4);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                     // the hand-off tiles come from the dQ kernel right in front of this one
  pdl_trigger();

  // Packed layouts: key tiles beyond an entry's keys are not items at all, query tiles beyond its queries are not visited
  // (the dQ kernel wrote no hand-off tiles for them).  The issuer and the epilogue warps walk the same item sequence.
  auto item_exists = [&](int item) {
    if (p.k_off == nullptr) return true;
    const int b = (item / n_kt) / p.H;
    return (item % n_kt) * BN < p.k_lens[b] && (p.q_off == nullptr || p.q_lens[b] > 0);
  };
  auto next_item = [&](int item) {
    item += (int)gridDim.x;
    while (item < n_items && !item_exists(item)) item += (int)gridDim.x;
    return item;
  };
  int first_item = blockIdx.x;
  if (first_item < n_items && !item_exists(first_item)) first_item = next_item(first_item);
  auto open_item = [&](DkvCursor& c) -> bool {        // position the cursor on the first query tile of c.item
    if (c.item >= n_items) return false;
    c.kt = c.item % n_kt;
    const int bh = c.item / n_kt;
    c.h = bh % p.H; c.b = bh / p.H;
    int q_lo;
    query_tile_range(p, c.kt * BN, q_lo, c.q_hi);
    if (p.q_off != nullptr) c.q_hi = min(c.q_hi, (p.q_lens[c.b] - 1) / BM);
    c.u = q_lo; c.first = true;
    return true;
  };
  auto advance = [&](DkvCursor& c) -> bool {
    if (c.u < c.q_hi) { ++c.u; c.first = false; return true; }
    c.item = next_item(c.item);
    return open_item(c);
  };

  if (w == 4) {
    if (lane == 0) {
      auto load = [&](const DkvCursor& c, int g) {
        const int st = g & 1;
        uint8_t* sb = smem + st * STAGE;
        int t_lo, t_hi;
        key_tile_range(p, c.u * BM, t_lo, t_hi);
        const long tile = (((long)c.b * p.H + c.h) * p.nQT + c.u) * p.max_kt + (c.kt - t_lo);
        ptx::mbar_arrive_expect_tx(&bar_full[st], STAGE);
#pragma unroll
        for (int a = 0; a < NATOM; ++a) {
          ptx::tma_load_2d(sb + a * Q_ATOM, &tmQ, &bar_full[st], c.h * DH + a * 64, q_base(p, c.b) + c.u * BM);
          ptx::tma_load_2d(sb + (NATOM + a) * Q_ATOM, &tmDO, &bar_full[st], c.h * DH + a * 64, q_base(p, c.b) + c.u * BM);
        }
        ptx::tma_load_2d(sb + 2 * NATOM * Q_ATOM, &tmP, &bar_full[st], 0, (int)(tile * BM));
        ptx::tma_load_2d(sb + 2 * NATOM * Q_ATOM + BM * 128, &tmDS, &bar_full[st], 0, (int)(tile * BM));
      };
      DkvCursor cur, nxt;
      cur.item = first_item;
      bool have = open_item(cur);
      nxt = cur;
      bool have_next = have && advance(nxt);
      if (have) load(cur, 0);
      if (have_next) load(nxt, 1);
      const uint32_t id_t = ptx::make_idesc_bf16(128, BN, 1, 1);       // M = head dim (two 64-wide groups), N = keys
      int g = 0, n_done_items = 0;
      while (have) {
        const int st = g & 1;
        const int ab = n_done_items & 1;                               // TMEM accumulator buffer of the current item
        if (cur.first && n_done_items >= 2) {                          // the epilogue has drained this buffer (item - 2)
          ptx::mbar_wait(&bar_free[ab], (uint32_t)(((n_done_items >> 1) - 1) & 1));
          ptx::tc_fence_after();
        }
        ptx::mbar_wait(&bar_full[st], (uint32_t)((g >> 1) & 1));
        ptx::tc_fence_after();
        const uint32_t qb = ptx::smem_u32(smem + st * STAGE), dob = qb + NATOM * Q_ATOM, pb = qb + 2 * NATOM * Q_ATOM,
                       dsb = pb + BM * 128;
        const uint32_t t_dv = tmem + ab * 128, t_dk = t_dv + 64;
#pragma unroll
        for (int ks = 0; ks < BM / 16; ++ks)
          ptx::umma_bf16(t_dv, ptx::make_smem_desc_sw128(dob + ks * 2048, Q_ATOM, 1024),
                         ptx::make_smem_desc_sw128(pb + ks * 2048, Q_ATOM, 1024), id_t, (!cur.first || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < BM / 16; ++ks)
          ptx::umma_bf16(t_dk, ptx::make_smem_desc_sw128(qb + ks * 2048, Q_ATOM, 1024),
                         ptx::make_smem_desc_sw128(dsb + ks * 2048, Q_ATOM, 1024), id_t, (!cur.first || ks > 0) ? 1u : 0u);
        ptx::umma_commit(&bar_done[st]);
        const bool last_of_item = cur.u == cur.q_hi;
        if (last_of_item) { ptx::umma_commit(&bar_acc[ab]); ++n_done_items; }
        // move on: `nxt` (already loaded into stage st^1) becomes current; fetch the tile after it into stage st
        cur = nxt; have = have_next;
        if (have) {
          have_next = advance(nxt);
          if (have_next) {
            ptx::mbar_wait(&bar_done[st], (uint32_t)((g >> 1) & 1));    // the MMAs just issued have finished reading stage st
            load(nxt, g + 2);
          }
        }
        ++g;
      }
    }
  } else {
    // epilogue warps: TMEM lane = head-dim index d, column = key.  The (dh x 64) accumulators are transposed through shared
    // memory and leave as 16-byte row pieces (a key row of dV / dK is dh contiguous bf16).
    const uint32_t lane_base = (uint32_t)(w * 32) << 16;
    const int dcol = w * 32 + lane, tid = threadIdx.x;
    int n_items_done = 0;
    for (int item = first_item; item < n_items; item = next_item(item), ++n_items_done) {
      const int kt = item % n_kt, bh = item / n_kt, h = bh % p.H, b = bh / p.H;
      const int j0 = kt * BN, ab = n_items_done & 1;
      const int n_k = k_rows(p, b);
      const long kb = k_base(p, b);
      ptx::mbar_wait(&bar_acc[ab], (uint32_t)((n_items_done >> 1) & 1));
      ptx::tc_fence_after();
      __nv_bfloat16* so = reinterpret_cast<__nv_bfloat16*>(sOut);
#pragma unroll
      for (int which = 0; which < 2; ++which) {
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) {
          float r[16];
          tmem_load_cols(tmem + ab * 128 + which * 64 + c * 16 + lane_base, r);
          if (dcol < DH) {
#pragma unroll
            for (int x = 0; x < 16; ++x) so[(which * BN + c * 16 + x) * DH + dcol] = __float2bfloat16_rn(r[x]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_free[ab]);         // the issuer may reuse this accumulator buffer
      ptx::named_bar_sync(1, 128);
      constexpr int PIECES = DH / 8;                          // 16-byte pieces per key row
      for (int k = tid; k < 2 * BN * PIECES; k += 128) {
        const int which = k / (BN * PIECES), rem = k - which * BN * PIECES, jr = rem / PIECES, pc = rem - jr * PIECES;
        const int j = j0 + jr;
        if (j < n_k) {
          const uint4 v = *reinterpret_cast<const uint4*>(so + (which * BN + jr) * DH + pc * 8);
          __nv_bfloat16* out = which == 0 ? p.dv : p.dk;
          const long ld = which == 0 ? p.ldv : p.ldk;
          *reinterpret_cast<uint4*>(out + (kb + j) * ld + h * DH + pc * 8) = v;
        }
      }
      ptx::named_bar_sync(1, 128);                            // sOut is rewritten by the next item
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (w == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

template <int NSPLIT>
static int attn_bwd_tc_launch_n(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                                const int* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                                float* delta, void* ws, size_t ws_bytes, cudaStream_t st) {
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
  if ((rc = make_tmap_bf16_2d(&tmQ, q, HD, attn_q_rows_total(d), d.ldq, 64, BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dO, HD, attn_q_rows_total(d), d.ldo, 64, BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, HD, attn_k_rows_total(d), d.ldk, 64, BN))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, HD, attn_k_rows_total(d), d.ldv, 64, BN))) return rc;
  if (d.rel_dist > 0) {
    if ((rc = make_tmap_bf16_2d(&tmE, E, d.dh, (long)d.H * (2 * d.rel_dist - 1), d.dh, 64, PBW))) return rc;
  } else {
    tmE = tmK;
  }
  SST_REQUIRE(ws != nullptr && ws_bytes >= attn_bwd_ws_bytes(d), SST_E_ARG, "attn_bwd: workspace of %zu bytes required (got %zu)",
              attn_bwd_ws_bytes(d), ws_bytes);
  p.nQT = cdiv(d.Lq, BM);
  p.max_kt = attn_max_key_tiles(d);
  const long tiles = (long)d.B * d.H * p.nQT * p.max_kt;
  p.ws_p = reinterpret_cast<__nv_bfloat16*>(ws);
  p.ws_ds = p.ws_p + tiles * BM * BN;
  SST_REQUIRE(tiles * BM < (1L << 31), SST_E_ARG, "attn_bwd: hand-off workspace exceeds the TMA row coordinate range");
  CUtensorMap tmP, tmDS;
  if ((rc = make_tmap_bf16_2d(&tmP, p.ws_p, BN, tiles * BM, BN, 64, BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDS, p.ws_ds, BN, tiles * BM, BN, 64, BM))) return rc;
  constexpr int SMEM_DQ = 2 * NATOM * BN * 128 + 2 * NATOM * PBW * 128 + 2 * NATOM * BN * 128 + (PBW / 64) * BM * 128 + 1024 + 128 +
                          128 * 4;
  constexpr int SMEM_DKV = 2 * (2 * NATOM * BM * 128 + 2 * BM * 128) + 2 * BN * DH * 2 + 1024 + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<DH, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV);
    SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute(attn_bwd_tc): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  cudaError_t le = launch_pdl(attn_bwd_dq_tc_kernel<DH, NSPLIT>, dim3(cdiv(d.Lq, BM), d.H, d.B), dim3(128 * NSPLIT + 32), SMEM_DQ, st,
                              tmK, tmV, tmE, p, reinterpret_cast<const __nv_bfloat16*>(q));
  SST_REQUIRE(le == cudaSuccess, SST_E_LAUNCH, "attn_bwd_dq_tc launch: %s", cudaGetErrorString(le));
  {
    const int n_kt = cdiv(d.Lk, BN), n_items = n_kt * d.H * d.B;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    le = launch_pdl(attn_bwd_dkv_tc_kernel<DH>, dim3(grid), dim3(160), SMEM_DKV, st, tmQ, tmDO, tmP, tmDS, p, n_kt, n_items);
    SST_REQUIRE(le == cudaSuccess, SST_E_LAUNCH, "attn_bwd_dkv_tc launch: %s", cudaGetErrorString(le));
  }
  return check_launch("attn_bwd_tc", 2);
}

// column groups per tile (threads = 128 * groups): SST_ATTN_NSPLIT=2|4 forces one; otherwise 4, except for the banded dQ kernel,
// which measures 5 % faster with 2 (cfg2 encoder layer: backward 1.109 -> 1.074 ms; forward 0.420 vs 0.449 ms the other way round)
static int attn_tc_nsplit_env() {
  static int n = -1;
  if (n < 0) {
    const char* e = getenv("SST_ATTN_NSPLIT");
    n = (e && e[0] == '2') ? 2 : (e && e[0] == '4') ? 4 : 0;
  }
  return n;
}
int attn_tc_nsplit() { return attn_tc_nsplit_env() ? attn_tc_nsplit_env() : 4; }

int attn_bwd_tc_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                       const int* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                       float* delta, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int ns = attn_tc_nsplit_env() ? attn_tc_nsplit_env() : (d.rel_dist > 0 && d.Lk > d.rel_dist ? 2 : 4);
  if (ns == 2)
    return attn_bwd_tc_launch_n<2>(d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, ws, ws_bytes, st);
  return attn_bwd_tc_launch_n<4>(d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, ws, ws_bytes, st);
}

size_t attn_bwd_tc_ws_bytes(const SstAttnDesc& d) { return attn_tc::attn_bwd_ws_bytes(d); }

}  // namespace sst

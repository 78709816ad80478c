// Tensor-core attention (bf16 operands, fp32 accumulation): the fused masked / banded relative-position attention of
// include/sst.h on tcgen05, TMA-fed.  Replaces MultiHeadAttention.forward between the projections
// (transformer.py:177-208) and LearnedRelativePositionalEmbedding (transformer.py:260-403).
//
// One CTA = 128 query rows of one (batch, head); key tiles of 64.  4*NSPLIT compute warps + 1 issuer warp (one thread
// drives TMA and tcgen05.mma from mbarriers).  Per key tile the issuer launches:
//     S  (128 x 64)  = Q K^T            tcgen05.mma, accumulator in TMEM columns [0, 64)
//     PB (128 x 192) = Q E_win^T        the relative-position logits against the 191 embedding rows the tile can touch
//                                       (E_win = rows [j0-i0-127+R-1, +192) of E[h]); TMEM columns [64, 256)
// then the threads (one query row x CW keys each, attention_tc.cuh) read S and their PB window from TMEM.  The reference's pad/view "skew"
// (transformer.py:383-395) becomes a per-lane register barrel shift: bias[i][j] = PB[i][(j-j0) - (i-i0) + 127]; the
// warp-uniform part of the shift is folded into the tcgen05.ld column address, the lane part (0..31) is five
// predicated select stages.  Masks are SET to -1e8 and the bias ADDED exactly as the reference does,
// online softmax over key tiles, Philox dropout on the probabilities, P (bf16) goes to 128B-swizzled shared memory and
//     O (128 x dh) += P V               accumulator in TMEM columns [256, 256+dh), rescaled in TMEM when a row max moves.
// Only key tiles that intersect the band |i-j| < R are visited (exact: everything outside has probability 0 in fp32).
// Pipeline: K/E/V tiles are double-buffered in shared memory; the compute warps release the S/PB accumulators as soon
// as they have copied their slices to registers ("s_free"), so the issuer runs tile t+1's S/PB MMAs under tile t's
// exp / dropout / P-store work; P(t) and the O rescale are handed back through "p_ready", PV(t) completion through "pv".
// Persistent: one CTA per SM walks (batch, head, query tile) items.  All barrier phases and shared-memory stages are indexed
// by the CTA's running tile counter, so the issuer starts the NEXT item's Q / K / E / V loads and its first S / PB MMAs while
// the compute warps are still in the current item's last softmax and epilogue -- the ~2 us of TMA + MMA latency at the head
// of every item (a quarter of the time of a 5-6 tile item) is no longer exposed.
// Backward = two kernels of the same shape (dQ per query tile; dK/dV per key tile), see attention_tc_bwd.cu.
#include "attention_tc.cuh"

namespace sst {

template <int DH, int NSPLIT>
__global__ void __launch_bounds__(128 * NSPLIT + 32, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE, const attn_tc::AttnTcParams p) {
  using namespace attn_tc;
  using SP = Split<NSPLIT>;
  constexpr int CW = SP::CW;
  constexpr int NW = 4 * NSPLIT;                 // compute warps
  constexpr int KS = DH / 16;                    // k-steps of the q.k / q.E contractions
  constexpr int NATOM = (DH + 63) / 64;          // 64-column (128-byte) swizzle atoms per row of q / k / E / v
  constexpr int Q_ATOM = BM * 128, K_ATOM = BN * 128, E_ATOM = PBW * 128, V_GRP = BN * 128;
  constexpr int OC = DH / NSPLIT;                // accumulator columns each thread of a row looks after
  constexpr uint32_t TM_S = 0, TM_PB = 64, TM_O = 256;
  static_assert(OC % 8 == 0, "head dim must split into x8 TMEM granules");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NATOM * Q_ATOM;             // [2]
  uint8_t* sE = sK + 2 * NATOM * K_ATOM;         // [2]
  uint8_t* sV = sE + 2 * NATOM * E_ATOM;         // [2]
  uint8_t* sP = sV + 2 * NATOM * V_GRP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + BM * 128);
  uint64_t* bar_q = bars, *bar_ke = bars + 1 /* [2] */, *bar_v = bars + 3 /* [2] */, *bar_s = bars + 5, *bar_sfree = bars + 6,
           *bar_p = bars + 7, *bar_pv = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* sred = reinterpret_cast<float*>(bars + 10);         // [3][NSPLIT][128] row-statistic exchange between column groups
                                                              // (two alternate per tile, the third is the epilogue's)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nQT = (p.Lq + BM - 1) / BM;
  const int n_items = nQT * p.H * p.B;                       // item = (b * H + h) * nQT + query tile
  // items this CTA works on: blockIdx.x, + gridDim.x, ... minus the query tiles a packed layout does not have (both roles walk
  // the same sequence, so the running tile / item counters stay in step)
  auto item_exists = [&](int it) {
    if (p.q_off == nullptr) return true;
    const int bh = it / nQT;
    return q_tile_exists(p, bh / p.H, (it - bh * nQT) * BM);
  };
  auto next_item = [&](int it) {
    it += (int)gridDim.x;
    while (it < n_items && !item_exists(it)) it += (int)gridDim.x;
    return it;
  };
  int first_item = blockIdx.x;
  if (first_item < n_items && !item_exists(first_item)) first_item = next_item(first_item);

  if (w == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
      if (p.R > 0) ptx::prefetch_tmap(&tmE);
      for (int k = 0; k < 6; ++k) ptx::mbar_init(&bars[k], 1);      // q, ke[2], v[2], s: one producer each
      ptx::mbar_init(bar_sfree, NW);
      ptx::mbar_init(bar_p, NW);
      ptx::mbar_init(bar_pv, 1);
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
  pdl_wait();                                     // set-up above overlaps the previous kernel's tail (sst_common.cuh)
  pdl_trigger();

  if (w == NW) {
    // ============================================ issuer (one thread) ============================================
    if (lane == 0) {
      const uint32_t ke_bytes = NATOM * K_ATOM + (p.R > 0 ? NATOM * E_ATOM : 0);
      uint32_t g = 0, n_done = 0;                 // tiles / items this CTA has issued so far: every phase derives from them
      int i0 = 0, h = 0, b = 0, qb = 0, kb = 0;     // qb / kb: first row of the entry in the query- / key-side matrices
      auto load_ke = [&](int t, uint32_t gt) {    // gt = running index of tile t
        const int st = gt & 1;
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
      auto load_v = [&](int t, uint32_t gt) {
        const int st = gt & 1;
        ptx::mbar_arrive_expect_tx(&bar_v[st], NATOM * V_GRP);
#pragma unroll
        for (int a = 0; a < NATOM; ++a)
          ptx::tma_load_2d(sV + (st * NATOM + a) * V_GRP, &tmV, &bar_v[st], h * DH + a * 64, kb + t * BN);
      };
      auto issue_s = [&](int st) {   // S = Q K^T and PB = Q E_win^T from stage `st`
        const uint32_t qb = ptx::smem_u32(sQ), kb = ptx::smem_u32(sK + st * NATOM * K_ATOM), eb = ptx::smem_u32(sE + st * NATOM * E_ATOM);
        const uint32_t id_s = ptx::make_idesc_bf16(BM, BN, 0, 0), id_pb = ptx::make_idesc_bf16(BM, PBW, 0, 0);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
          ptx::umma_bf16(tmem + TM_S, ptx::make_smem_desc_sw128(qb + off * Q_ATOM + in, 0, 1024),
                         ptx::make_smem_desc_sw128(kb + off * K_ATOM + in, 0, 1024), id_s, ks > 0);
        }
        if (p.R > 0) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t off = (ks >> 2), in = (ks & 3) * 32;
            ptx::umma_bf16(tmem + TM_PB, ptx::make_smem_desc_sw128(qb + off * Q_ATOM + in, 0, 1024),
                           ptx::make_smem_desc_sw128(eb + off * E_ATOM + in, 0, 1024), id_pb, ks > 0);
          }
        }
        ptx::umma_commit(bar_s);
      };

      // Head of an item: Q, the first two K / E tiles and the first V tile.  `gh` = running index of the item's first tile.
      // Issued while the PREVIOUS item's last tile is still in its softmax (see below), so the loads and the first S / PB
      // MMAs are off the critical path of every item but a CTA's first.
      auto decode = [&](int item, int& t_lo, int& t_hi) {
        const int bh = item / nQT;
        i0 = (item - bh * nQT) * BM; b = bh / p.H; h = bh - b * p.H;
        qb = q_base(p, b); kb = k_base(p, b);
        key_tile_range_valid(p, i0, b, t_lo, t_hi);
      };
      auto head_loads = [&](int t_lo, int t_hi, uint32_t gh) {
        ptx::mbar_arrive_expect_tx(bar_q, NATOM * Q_ATOM);
#pragma unroll
        for (int a = 0; a < NATOM; ++a) ptx::tma_load_2d(sQ + a * Q_ATOM, &tmQ, bar_q, h * DH + a * 64, qb + i0);
        load_ke(t_lo, gh);
        load_v(t_lo, gh);
        if (t_lo < t_hi) load_ke(t_lo + 1, gh + 1);
      };
      int item = first_item, t_lo = 0, t_hi = -1;
      if (item < n_items) { decode(item, t_lo, t_hi); head_loads(t_lo, t_hi, 0); }
      bool sfree_seen = true;                     // bar_sfree(g - 1) already waited for (nothing to wait for at g = 0)
      while (item < n_items) {
        const int next = next_item(item);
        ptx::mbar_wait(bar_q, n_done & 1u);
        ptx::mbar_wait(&bar_ke[g & 1], (g >> 1) & 1u);
        if (!sfree_seen) ptx::mbar_wait(bar_sfree, (g - 1) & 1u);   // the previous item's last S / PB are in registers
        ptx::tc_fence_after();
        issue_s(g & 1);
        int nt_lo = 0, nt_hi = -1;
        sfree_seen = false;
        for (int t = t_lo; t <= t_hi; ++t) {
          const uint32_t gt = g + (uint32_t)(t - t_lo);
          const int st = gt & 1;
          if (gt > 0) ptx::mbar_wait(bar_pv, (gt - 1) & 1u);  // PV(gt-1) done: its V stage (the one tile gt+1 uses) is free
          if (t < t_hi) {
            load_v(t + 1, gt + 1);
            ptx::mbar_wait(bar_sfree, gt & 1u);               // S / PB copied to registers by every compute warp
            ptx::mbar_wait(&bar_ke[st ^ 1], ((gt + 1) >> 1) & 1u);
            ptx::tc_fence_after();
            issue_s(st ^ 1);
            if (t + 2 <= t_hi) load_ke(t + 2, gt + 2);        // stage `st`: its MMAs finished before bar_s(t) fired
          } else if (next < n_items) {
            // last tile of the item: once its S / PB are in registers, the Q buffer, both K / E stages and the V stage
            // of tile gt + 1 are free -- start the next item's head now, under this tile's softmax
            ptx::mbar_wait(bar_sfree, gt & 1u);
            sfree_seen = true;
            decode(next, nt_lo, nt_hi);                       // (i0, h, b now describe the next item; this one only needs st below)
            head_loads(nt_lo, nt_hi, gt + 1);
          }
          ptx::mbar_wait(bar_p, gt & 1u);                     // P(t) in shared memory, O rescaled
          ptx::mbar_wait(&bar_v[st], (gt >> 1) & 1u);
          ptx::tc_fence_after();
          const uint32_t pb = ptx::smem_u32(sP), vb = ptx::smem_u32(sV + st * NATOM * V_GRP);
          const uint32_t id_o = ptx::make_idesc_bf16(BM, DH, 0, 1);
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks)
            ptx::umma_bf16(tmem + TM_O, ptx::make_smem_desc_sw128(pb + ks * 32, 0, 1024),
                           ptx::make_smem_desc_sw128(vb + ks * 2048, V_GRP, 1024), id_o, (t > t_lo || ks > 0) ? 1u : 0u);
          ptx::umma_commit(bar_pv);
        }
        g += (uint32_t)(t_hi - t_lo + 1);
        ++n_done;
        item = next; t_lo = nt_lo; t_hi = nt_hi;
      }
    }
  } else {
    // ============================================ compute warps ==================================================
    const int q = w & 3, hf = w >> 2;
    const int li = 32 * q + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const unsigned long long seed = p.thr ? salted(p.seed, p.salt) : 0ull;
    uint32_t g = 0;                                 // running tile counter, in step with the issuer's
    // The O rows of a finished item are written out one tile late, from inside the NEXT item's first tile: the wait for the
    // item's last PV (~0.3 us of MMA latency with nothing else to do) then falls behind that tile's logits / exp work.
    bool o_pending = false, o_valid = false;
    float o_scale = 0.f;
    __nv_bfloat16* o_dst = nullptr;
    for (int item = first_item; item < n_items; item = next_item(item)) {
    const int bh = item / nQT;
    const int i0 = (item - bh * nQT) * BM, b = bh / p.H, h = bh - b * p.H;
    int t_lo, t_hi;
    key_tile_range_valid(p, i0, b, t_lo, t_hi);
    const int i = i0 + li;
    const RowCtx rc = make_row_ctx(p, b, h, i);
    float m_run = NEG_BIG, l_run = 0.f;             // l_run: this thread's share (its CW columns) of the row sum

    for (int t = t_lo; t <= t_hi; ++t) {
      const int k = t - t_lo;
      const uint32_t gt = g + (uint32_t)k;
      ptx::mbar_wait(bar_s, gt & 1u);
      ptx::tc_fence_after();

      float U[SP::WIN_LD];
      const bool skip = block_out_of_band<NSPLIT>(p, i0 + 32 * q, t * BN + CW * hf);
      const uint32_t red = ptx::smem_u32(sred) + (uint32_t)((gt & 1) * NSPLIT * 128 + li) * 4;      // this row's slot of group 0
      float mt = NEG_BIG;
      if (!skip) {
        uint32_t mbits;
        const bool simple = tile_is_simple(p, rc, t * BN);
        tile_logits<NSPLIT>(p, rc, tmem + TM_S + lane_base, tmem + TM_PB + lane_base, q, hf, lane, t * BN, simple, U, mbits);
        mt = U[0];
#pragma unroll
        for (int x = 1; x < CW; ++x) mt = fmaxf(mt, U[x]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_sfree);    // the issuer may overwrite S / PB with the next tile

      ptx::st_shared_f32(red + hf * 512, mt);
      ptx::named_bar_sync(1 + q, 32 * NSPLIT);       // the NSPLIT warps that share this lane quarter
#pragma unroll
      for (int g = 0; g < NSPLIT; ++g) mt = fmaxf(mt, ptx::ld_shared_f32(red + g * 512));
      const float m_new = fmaxf(m_run, mt);
      const float alpha = __expf(m_run - m_new);
      float sum = 0.f;
      if (!skip) {
#pragma unroll
        for (int x = 0; x < CW; ++x) { U[x] = __expf(U[x] - m_new); sum += U[x]; }   // exact subtraction first: m can be -1e8
        if (p.thr) {
          float keep[CW];
          dropout_keep<NSPLIT>(p, seed, rc, t * BN, hf, keep);
#pragma unroll
          for (int x = 0; x < CW; ++x) U[x] *= keep[x];
        }
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      if (gt > 0) {                                  // PV(gt-1) done: the P buffer is free and O is complete
        ptx::mbar_wait(bar_pv, (gt - 1) & 1u);
        ptx::tc_fence_after();
        if (o_pending) {                             // k == 0: the previous item's accumulator, before this item's first PV
          tmem_row_to_global<OC>(tmem + TM_O + hf * OC + lane_base, o_dst, o_scale, o_valid);
          o_pending = false;
        }
      }
      if (!skip) store_cols_bf16_sw128<CW>(ptx::smem_u32(sP), li, CW * hf, U);
      else store_zero_cols_sw128<CW>(ptx::smem_u32(sP), li, CW * hf);
      if (k > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int c = 0; c < OC / 8; ++c) {
          uint32_t r[8];
          const uint32_t ta = tmem + TM_O + hf * OC + c * 8 + lane_base;
          ptx::tmem_ld_32x32b_x8(ta, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int x = 0; x < 8; ++x) r[x] = __float_as_uint(__uint_as_float(r[x]) * alpha);
          ptx::tmem_st_32x32b_x8(ta, r);
        }
        ptx::tmem_st_wait();
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_p);
    }

    // row sum = sum of the column groups' shares (third exchange buffer: the next item's first tile reuses the per-tile ones
    // while slower warps of this quarter may still be reading here)
    const int kl = t_hi - t_lo;
    const uint32_t red = ptx::smem_u32(sred) + (uint32_t)(2 * NSPLIT * 128 + li) * 4;
    ptx::st_shared_f32(red + hf * 512, l_run);
    ptx::named_bar_sync(1 + q, 32 * NSPLIT);
    float l_tot = 0.f;
#pragma unroll
    for (int g = 0; g < NSPLIT; ++g) l_tot += ptx::ld_shared_f32(red + g * 512);

    const bool valid = i < q_rows(p, b);
    o_pending = true; o_valid = valid; o_scale = 1.f / l_tot;
    o_dst = p.o + ((long)q_base(p, b) + i) * p.ldo + h * DH + hf * OC;
    if (valid && hf == 0) {
      const long nrows = (long)p.B * p.H * p.Lq;
      p.lse[rc.row_id] = m_run;
      p.lse[nrows + rc.row_id] = __logf(l_tot);
    }
    g += (uint32_t)(kl + 1);
    }  // items
    if (o_pending) {                                // the CTA's last item
      ptx::mbar_wait(bar_pv, (g - 1) & 1u);
      ptx::tc_fence_after();
      tmem_row_to_global<OC>(tmem + TM_O + hf * OC + lane_base, o_dst, o_scale, o_valid);
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
static int attn_fwd_tc_launch_n(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                                const int* k_lens, void* o, float* lse, cudaStream_t st) {
  using namespace attn_tc;
  constexpr int DH = 96;
  AttnTcParams p = make_tc_params(d, q_lens, k_lens);
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = d.ldo; p.lse = lse;
  CUtensorMap tmQ, tmK, tmV, tmE;
  int rc;
  const long HD = (long)d.H * d.dh;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, HD, attn_q_rows_total(d), d.ldq, 64, BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, HD, attn_k_rows_total(d), d.ldk, 64, BN))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, HD, attn_k_rows_total(d), d.ldv, 64, BN))) return rc;
  if (d.rel_dist > 0) {
    if ((rc = make_tmap_bf16_2d(&tmE, E, d.dh, (long)d.H * (2 * d.rel_dist - 1), d.dh, 64, PBW))) return rc;
  } else {
    tmE = tmK;
  }
  constexpr int NATOM = (DH + 63) / 64;
  constexpr int SMEM = NATOM * (BM * 128 + 2 * BN * 128 + 2 * PBW * 128 + 2 * BN * 128) + BM * 128 + 1024 + 128 + 3 * NSPLIT * 128 * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute(attn_fwd_tc): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const long n_items = (long)cdiv(d.Lq, BM) * d.H * d.B;
  const int grid = (int)(n_items < num_sms() ? n_items : num_sms());       // persistent: one CTA per SM
  cudaError_t le = launch_pdl(attn_fwd_tc_kernel<DH, NSPLIT>, dim3(grid), dim3(128 * NSPLIT + 32), SMEM, st, tmQ, tmK, tmV, tmE, p);
  SST_REQUIRE(le == cudaSuccess, SST_E_LAUNCH, "attn_fwd_tc launch: %s", cudaGetErrorString(le));
  return check_launch("attn_fwd_tc");
}

int attn_tc_nsplit();   // attention_tc_bwd.cu: 2 or 4 column groups (SST_ATTN_NSPLIT, default 4)

int attn_fwd_tc_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                       const int* k_lens, void* o, float* lse, cudaStream_t st) {
  if (attn_tc_nsplit() == 2) return attn_fwd_tc_launch_n<2>(d, q, k, v, E, q_lens, k_lens, o, lse, st);
  return attn_fwd_tc_launch_n<4>(d, q, k, v, E, q_lens, k_lens, o, lse, st);
}

}  // namespace sst

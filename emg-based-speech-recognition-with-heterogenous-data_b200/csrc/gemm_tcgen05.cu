// bf16 GEMM on 5th-gen tensor cores: TMA -> 128B-swizzled shared memory -> tcgen05.mma (accumulators in
// TMEM, double buffered) -> tcgen05.ld epilogue.  Persistent, warp-specialised: warp 0 = TMA producer,
// warp 1 = MMA issuer (one thread) + TMEM owner, warps 2-17 = epilogue (TMEM lane quarter = warp & 3, four warps per
// quarter splitting the tile's columns).
//
// CTAS = 2 runs the same roles as a CTA PAIR (cluster of two, tcgen05 cta_group::2): one 256 x BN tile per pair, each CTA
// stages its 128 rows of A and its BN/2 rows of B (a third less shared-memory fill per flop than two independent
// 128 x BN tiles), rank 0 issues the 256-row MMAs for both, the accumulator rows land in each CTA's own TMEM and each
// CTA runs the epilogue of its 128 rows.
//
// One kernel covers every dense contraction of the hot path (include/sst.h, "GEMM family"):
//   TN     : C = A[M,K] * B[N,K]^T, both K-major.  A may be read as up to three row-shifted column windows
//            ("taps"): a k=3 Conv1d over a time-padded channels-last activation is a GEMM whose K blocks come
//            from rows m-1 / m / m+1 (stride 1) or from the (2C)-wide pair-row view (stride 2).
//   NT_MN  : C (+)= A[K,M]^T * B[K,N], both MN-major (weight gradients; K = tokens), split-K with fp32
//            atomics, B optionally row-shifted per N segment (conv taps again).
#include "sst_common.cuh"
#include "sst_ptx.cuh"

namespace sst {

constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_EPI_WARPS = 16;                  // four per TMEM lane quarter, each takes a quarter of the tile's columns
constexpr int G_CGROUPS = G_EPI_WARPS / 4;       // column groups
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;
constexpr int G_A_BYTES = G_BM * G_BK * 2;   // 16 KiB

struct GemmKParams {
  int M, N, K;
  int mode_mn;        // A is MN-major (weight-gradient layout)
  int b_mn;           // B is MN-major (NT_MN and TN_BMN layouts)
  int num_kb;
  int kb_per_seg;     // TN: k-blocks per A segment
  int nseg_cols;      // MN: columns of C per B segment
  int a_row_shift[3], a_col0[3], b_row_shift[3], b_col0[3];
  int splits, kb_per_split;
  int m_blks, n_blks;
  int epilogue;
  float alpha, mask_scale, drop_scale;
  uint32_t drop_thr;
  unsigned long long seed;
  const unsigned long long* salt;
  const float* bias;
  const void* aux;
  long ldaux;
  int aux_f32;
  void* C;
  long ldc;
  int out_f32;
  int atomic_out;
  int remap_P, remap_T, remap_j0;
  int out_seg_cols, out_grp_cols;       // segmented output columns (SstGemmDesc), 0 = plain
  long out_seg_stride, out_grp_off[3];
  void* col_acc; int col_mode, col_grp; // fused column sums (1: float) / BatchNorm statistics (2: double, groups of col_grp columns)
};

// Debug timeline (-DSST_GEMM_TRACE, tools/gemm_trace.py): CTA 0 stamps %globaltimer at its pipeline events.
#ifdef SST_GEMM_TRACE
__device__ unsigned long long g_gemm_trace[64];
__device__ __forceinline__ void trace_stamp(int slot) {
  if (blockIdx.x == 0 && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_gemm_trace[slot] = t;
  }
}
#define SST_TRACE(slot) trace_stamp(slot)
#else
#define SST_TRACE(slot) do { } while (0)
#endif

template <int BN, int CTAS> struct GemmCfg {
  static constexpr int B_ROWS = BN / CTAS;                        // rows of the B tile this CTA stages
  static constexpr int B_BYTES = B_ROWS * G_BK * 2;
  static constexpr int STAGE_BYTES = G_A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = G_EPI_WARPS * 2048;        // per epilogue warp: 32 rows x 64 B of bf16 output
  static constexpr int STAGES_FIT = (227 * 1024 - 1024 - 512 - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;  // 4 (256,1) / 6 (256,2), (128,1) / 8 (128,2)
  static constexpr int TMEM_COLS = 2 * BN;                        // two accumulator buffers
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + STAGING_BYTES;
};

template <int CTAS>
__device__ __forceinline__ void decode_unit(const GemmKParams& p, int u, int rank, int& m0, int& n0, int& kb0, int& kb1) {
  int tile = u / p.splits, split = u - tile * p.splits;
  int mb = tile / p.n_blks, nb = tile - mb * p.n_blks;
  m0 = (mb * CTAS + rank) * G_BM;
  n0 = nb;   // caller multiplies by BN
  kb0 = split * p.kb_per_split;
  kb1 = min(kb0 + p.kb_per_split, p.num_kb);
}

template <int BN, int CTAS>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  using Cfg = GemmCfg<BN, CTAS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) SST_TRACE(0);                                  // kernel entry
  const int rank = CTAS == 2 ? (int)ptx::cluster_ctarank() : 0;        // rank 0 of a pair issues the MMAs
  const int unit0 = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], G_EPI_WARPS * CTAS); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (CTAS == 2) { ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (CTAS == 2) ptx::cluster_sync(); else __syncthreads();       // the peer's barriers exist before anything signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SST_TRACE(1);                                  // barriers, TMEM, cluster sync done
  pdl_wait();                 // everything above ran under the previous kernel's tail; from here on global memory is touched
  pdl_trigger();              // the next kernel's CTAs may take this SM's resources as soon as this CTA leaves

  const int total_units = p.m_blks * p.n_blks * p.splits;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // pair mode: both CTAs' loads are counted on the LEADER's full barrier (it waits for 2 x STAGE_BYTES)
      auto load = [&](void* dst, const CUtensorMap* tm, int st, int x, int y) {
        if (CTAS == 2) ptx::tma_load_2d_pair(dst, tm, ptx::map_to_cta(&full[st], 0), x, y);
        else ptx::tma_load_2d(dst, tm, &full[st], x, y);
      };
      for (int u = unit0; u < total_units; u += unit_stride) {
        int m0, nb, kb0, kb1;
        decode_unit<CTAS>(p, u, rank, m0, nb, kb0, kb1);
        const int n0 = nb * BN;
        const int nB = n0 + rank * Cfg::B_ROWS;                  // this CTA's slice of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u);
          if (u == unit0 && kb == kb0) SST_TRACE(2);                   // first TMA issue
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + G_A_BYTES;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], CTAS * Cfg::STAGE_BYTES);
          if (!p.mode_mn) {
            int seg = kb / p.kb_per_seg;
            int x = p.a_col0[seg] + (kb - seg * p.kb_per_seg) * G_BK;
            load(sA, &tmA, stage, x, m0 + p.a_row_shift[seg]);
          } else {
            const int k0 = kb * G_BK;
#pragma unroll
            for (int i = 0; i < G_BM / 64; ++i) load(sA + i * 8192, &tmA, stage, m0 + i * 64, k0);
          }
          if (!p.b_mn) {
            load(sB, &tmB, stage, kb * G_BK, nB);
          } else {
            const int k0 = kb * G_BK;
            int seg = n0 / p.nseg_cols;
            int nin = nB - seg * p.nseg_cols;
#pragma unroll
            for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
              load(sB + j * 8192, &tmB, stage, p.b_col0[seg] + nin + j * 64, k0 + p.b_row_shift[seg]);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = ptx::make_idesc_bf16(G_BM * CTAS, BN, p.mode_mn, p.b_mn);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = unit0; u < total_units && rank == 0; u += unit_stride) {
      int m0, nb, kb0, kb1;
      decode_unit<CTAS>(p, u, rank, m0, nb, kb0, kb1);
      if (lane == 0) SST_TRACE(40 + 2 * ((u - unit0) / unit_stride));          // issuer arrives at the tile
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1u);
      if (lane == 0) SST_TRACE(41 + 2 * ((u - unit0) / unit_stride));          // accumulator buffer released by the epilogue
      ptx::tc_fence_after();
      if (lane == 0) SST_TRACE(20 + 3 * ((u - unit0) / unit_stride));          // ... and fenced
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full[stage], phase);      // TMA (async proxy) -> MMA (async proxy): the mbarrier is all the ordering
                                                  // needed; a tcgen05 fence here would only stall the issuer
        if (lane == 0 && u == unit0 && kb == kb0) SST_TRACE(3);        // first operands landed
        if (lane == 0 && kb == kb0) SST_TRACE(21 + 3 * ((u - unit0) / unit_stride));       // tile's first operands landed
        if (lane == 0 && kb == kb1 - 1) SST_TRACE(22 + 3 * ((u - unit0) / unit_stride));   // tile's last operands landed
        if (lane == 0) {
          const uint32_t a_base = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + G_A_BYTES;
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k) {
            const uint64_t ad = !p.mode_mn ? ptx::make_smem_desc_sw128(a_base + k * 32, 0, 1024)
                                           : ptx::make_smem_desc_sw128(a_base + k * 2048, 8192, 1024);
            const uint64_t bd = !p.b_mn ? ptx::make_smem_desc_sw128(b_base + k * 32, 0, 1024)
                                        : ptx::make_smem_desc_sw128(b_base + k * 2048, 8192, 1024);
            if (CTAS == 2) ptx::umma_bf16_pair(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else ptx::umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CTAS == 2) {                                       // both CTAs' producers / epilogues are released
            ptx::umma_commit_pair(&empty[stage]);
            if (kb == kb1 - 1) ptx::umma_commit_pair(&tfull[acc]);
          } else {
            ptx::umma_commit(&empty[stage]);
            if (kb == kb1 - 1) ptx::umma_commit(&tfull[acc]);
          }
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int cg = (warp - 2) >> 2;               // column group: chunks [cg * CPW, (cg + 1) * CPW)
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = unit0; u < total_units; u += unit_stride) {
      int m0, nb, kb0, kb1;
      decode_unit<CTAS>(p, u, rank, m0, nb, kb0, kb1);
      const int n0 = nb * BN;
      const int m = m0 + q * 32 + lane;
      // Epilogue inputs that do not depend on the accumulator are produced BEFORE waiting for it, i.e. under the MMA
      // main loop: the dropout keep bits (Philox, eight 16-bit lanes per block) and the ReLU/dropout mask bits of `aux`.
      constexpr int CPW = BN / 32 / G_CGROUPS;      // 32-column chunks per epilogue warp
      const uint32_t stg = ptx::smem_u32(staging) + (uint32_t)(warp - 2) * 2048;
      const unsigned long long seed_eff = (p.epilogue & SST_EPI_DROPOUT) ? salted(p.seed, p.salt) : 0ull;
      Philox4 rnd[4];                               // 16-bit lane e of block i8 decides column i8 * 8 + e of a chunk
      uint32_t mask_bits[CPW];
      auto draw = [&](int lc) {                     // the four Philox blocks of chunk lc (kept in registers until used)
        const int nbase = n0 + (cg * CPW + lc) * 32;
        const unsigned long long e0 = (unsigned long long)m * (unsigned long long)p.N + (unsigned long long)nbase;
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) rnd[i8] = philox4x32(seed_eff, (e0 >> 3) + i8);
      };
      if (p.epilogue & SST_EPI_DROPOUT) draw(0);    // first chunk's bits are produced under the MMA main loop
      if (p.epilogue & SST_EPI_MULMASK) {
        // bits of (aux > 0).  Fast path: the warp reads its 32 x 32 chunk of aux as 8 rows x 64 B per instruction
        // (whole sectors), each lane turns its 8 values into a mask byte, and the bytes are transposed through the
        // warp's staging slice so that lane l ends up with the 32-bit mask of row l.
        const bool aux_vec = !p.aux_f32 && (p.ldaux & 7) == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0;
        const __nv_bfloat16* auxh = reinterpret_cast<const __nv_bfloat16*>(p.aux);
        const int rl0 = lane >> 2, piece = lane & 3;
        uint4 aw[CPW][4];
#pragma unroll
        for (int lc = 0; lc < CPW; ++lc) {                      // every load of the tile is in flight before the first use
          const int nbase = n0 + (cg * CPW + lc) * 32;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int mr = m0 + q * 32 + it * 8 + rl0;
            aw[lc][it] = make_uint4(0u, 0u, 0u, 0u);
            if (aux_vec && nbase + 32 <= p.N && mr < p.M)
              aw[lc][it] = __ldg(reinterpret_cast<const uint4*>(auxh + (long)mr * p.ldaux + nbase + piece * 8));
          }
        }
        {                                                       // next tile's slice of aux: pull it into L2 meanwhile
          const int un = u + unit_stride;
          if (aux_vec && un < total_units) {
            int m0n, nbn, k0n, k1n;
            decode_unit<CTAS>(p, un, rank, m0n, nbn, k0n, k1n);
#pragma unroll
            for (int lc = 0; lc < CPW; ++lc) {
              const int nbase = nbn * BN + (cg * CPW + lc) * 32;
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int mr = m0n + q * 32 + it * 8 + rl0;
                if (nbase + 32 <= p.N && mr < p.M && piece == 0)
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(auxh + (long)mr * p.ldaux + nbase));
              }
            }
          }
        }
#pragma unroll
        for (int lc = 0; lc < CPW; ++lc) {
          const int nbase = n0 + (cg * CPW + lc) * 32;
          uint32_t bits = 0;
          if (aux_vec && nbase + 32 <= p.N) {                   // warp-uniform
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&aw[lc][it]);
              uint32_t byte = 0;
#pragma unroll
              for (int i = 0; i < 8; ++i) byte |= (__bfloat162float(h[i]) > 0.f ? 1u : 0u) << i;
              asm volatile("st.shared.u8 [%0], %1;" ::"r"(stg + lc * 128 + (it * 8 + rl0) * 4 + piece), "r"(byte) : "memory");
            }
            __syncwarp();
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(bits) : "r"(stg + lc * 128 + lane * 4) : "memory");
          } else if (m < p.M && nbase < p.N) {
            const int ncols = min(32, p.N - nbase);
            const long ab = (long)m * p.ldaux + nbase;
            for (int i = 0; i < ncols; ++i)
              bits |= (ld_as_f32(p.aux, ab + i, p.aux_f32 ? SST_F32 : SST_BF16) > 0.f ? 1u : 0u) << i;
          }
          mask_bits[lc] = bits;
        }
        __syncwarp();
      }
      // scale folding: without a bias the mask scale commutes with everything in front of it
      const bool fold_ms = (p.epilogue & SST_EPI_MULMASK) && !(p.epilogue & SST_EPI_BIAS);
      const float alpha_eff = fold_ms ? p.alpha * p.mask_scale : p.alpha;
      const uint32_t thr_hi = p.drop_thr << 16;
      const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0) SST_TRACE(8 + 2 * ((u - unit0) / unit_stride));       // tile's accumulator ready
      bool row_ok = m < p.M;
      int out_row = m;                               // output rows stay far below 2^31 (checked at launch): 32-bit bookkeeping
      if (p.remap_P > 0) {
        int chunk = m / p.remap_P;
        int t = m - chunk * p.remap_P - p.remap_j0;
        row_ok = row_ok && t >= 0 && t < p.remap_T;
        out_row = chunk * p.remap_T + t;
      }
      // Coalesced bf16 stores: the warp's (32 rows x 32 columns) chunk is transposed through shared memory so that one store
      // instruction writes 8 rows x 64 contiguous bytes (full sectors) instead of 32 rows x 16 bytes.  Lane l stores the
      // 16-byte piece (l & 3) of rows it*8 + (l >> 2); the row bookkeeping of those rows comes from their owner lanes.
      const bool staged = !p.atomic_out && !p.out_f32 && !(p.epilogue & SST_EPI_ACCUM) && (p.ldc & 7) == 0 && p.out_seg_cols == 0;
      int st_row[4];                                 // output row, or -1 when the owner lane's row is not stored
#pragma unroll
      for (int it = 0; it < 4; ++it) st_row[it] = __shfl_sync(0xffffffffu, row_ok ? out_row : -1, it * 8 + (lane >> 2));
      // The accumulator chunk of the NEXT column block is requested while this block's staged tile leaves through shared memory
      // (its registers are free by then), so that only the tile's first TMEM load is waited for with nothing else to do.
      uint32_t r[32];
      const uint32_t taddr0 = tmem_base + (uint32_t)(acc * BN + cg * CPW * 32) + ((uint32_t)(q * 32) << 16);
      ptx::tmem_ld_32x32b_x32(taddr0, r);
#pragma unroll
      for (int lc0 = 0; lc0 < CPW; ++lc0) {
        const int c = cg * CPW + lc0;
        ptx::tmem_ld_wait();
        bool next_requested = false;
        const int nbase = n0 + c * 32;
        const bool chunk_staged = staged && nbase + 32 <= p.N;          // warp-uniform
        if ((row_ok || chunk_staged) && nbase < p.N) {
          float v[32];
          const int ncols = min(32, p.N - nbase);
          if (!(p.epilogue & SST_EPI_BIAS)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * alpha_eff;
          } else if (bias_vec && ncols == 32) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + nbase);   // same address in every lane: one broadcast
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), alpha_eff, b.x);
              v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), alpha_eff, b.y);
              v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), alpha_eff, b.z);
              v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), alpha_eff, b.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]), alpha_eff, i < ncols ? __ldg(p.bias + nbase + i) : 0.f);
          }
          if (p.epilogue & SST_EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (p.epilogue & SST_EPI_DROPOUT) {
            // kept iff the element's 16-bit Philox lane >= thr16: the high half is compared in place, the low half after
            // one shift (same decision as philox_keep16)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = philox_keep16_at(rnd[i >> 3], i & 7, thr_hi) ? v[i] * p.drop_scale : 0.f;
            }
            if (lc0 + 1 < CPW) draw(lc0 + 1);
          }
          if (p.epilogue & SST_EPI_MULMASK) {
            const uint32_t mb = mask_bits[lc0];
            if (fold_ms) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = ((mb >> i) & 1u) ? v[i] : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = ((mb >> i) & 1u) ? v[i] * p.mask_scale : 0.f;
            }
          }
          long cb = (long)out_row * p.ldc + nbase;
          if (p.out_seg_cols > 0) {             // the 32-column chunk never straddles a segment (segments are multiples of 32)
            const int grp = nbase / p.out_grp_cols, nin = nbase - grp * p.out_grp_cols;
            const int seg = nin / p.out_seg_cols;
            cb = p.out_grp_off[grp] + (long)seg * p.out_seg_stride + (long)out_row * p.ldc + (nin - seg * p.out_seg_cols);
          }
          if (chunk_staged) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1]);
                w4[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              ptx::st_shared_v4(stg + lane * 64 + ((uint32_t)(j ^ ((lane >> 1) & 3)) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
            __syncwarp();
            if (lc0 + 1 < CPW) {                    // warp-uniform branch (chunk_staged is)
              ptx::tmem_ld_32x32b_x32(taddr0 + (uint32_t)((lc0 + 1) * 32), r);
              next_requested = true;
              __syncwarp();                         // keeps the request in FRONT of the loads / stores below (ptxas sinks it otherwise)
            }
            float cs[8], cq[8];                     // this lane's share of the chunk's column sums (/ sums of squares)
#pragma unroll
            for (int j = 0; j < 8; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rl = it * 8 + (lane >> 2), piece = lane & 3;
              uint4 o;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                           : "r"(stg + rl * 64 + ((uint32_t)(piece ^ ((rl >> 1) & 3)) << 4)));
              if (st_row[it] >= 0) {
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (long)st_row[it] * p.ldc + nbase + piece * 8) = o;
                if (p.col_mode != 0) {              // warp-uniform; the values exactly as stored
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h2[j]);
                    cs[2 * j] += f.x; cs[2 * j + 1] += f.y;
                    if (p.col_mode == 2) { cq[2 * j] = fmaf(f.x, f.x, cq[2 * j]); cq[2 * j + 1] = fmaf(f.y, f.y, cq[2 * j + 1]); }
                  }
                }
              }
            }
            if (p.col_mode != 0) {
              // the 8 lanes that hold the same 8 columns (lane & 3 == piece) add up: afterwards lanes 0..3 own 8 columns each
#pragma unroll
              for (int sh = 4; sh < 32; sh <<= 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], sh);
                  if (p.col_mode == 2) cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], sh);
                }
              }
              if (lane < 4) {
                const int n = nbase + lane * 8;
                if (p.col_mode == 1) {
                  float* dst = reinterpret_cast<float*>(p.col_acc) + n;
                  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    red_add4(dst, cs[0], cs[1], cs[2], cs[3]);
                    red_add4(dst + 4, cs[4], cs[5], cs[6], cs[7]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) atomicAdd(dst + j, cs[j]);
                  }
                } else {
                  const int G = p.col_grp, g = n / G;
                  double* dst = reinterpret_cast<double*>(p.col_acc) + (long)g * 2 * G + (n - g * G);
#pragma unroll
                  for (int j = 0; j < 8; ++j) { atomicAdd(dst + j, (double)cs[j]); atomicAdd(dst + G + j, (double)cq[j]); }
                }
              }
            }
            __syncwarp();
          } else if (!row_ok) {
            // nothing to store for this lane
          } else if (p.atomic_out) {
            float* cp = reinterpret_cast<float*>(p.C) + cb;
            if (ncols == 32 && (p.ldc & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                             "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) atomicAdd(cp + i, v[i]);
            }
          } else if (p.out_f32) {
            float* cp = reinterpret_cast<float*>(p.C) + cb;
            if (ncols == 32 && (p.ldc & 3) == 0) {
              float4* c4 = reinterpret_cast<float4*>(cp);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                if (p.epilogue & SST_EPI_ACCUM) { float4 old = c4[j]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                c4[j] = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncols) cp[i] = v[i] + ((p.epilogue & SST_EPI_ACCUM) ? cp[i] : 0.f);
            }
          } else {
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + cb;
            if (ncols == 32 && (p.ldc & 7) == 0) {
              uint4* c4 = reinterpret_cast<uint4*>(cp);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (p.epilogue & SST_EPI_ACCUM) {
                  uint4 old = c4[j];
                  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&old);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[j * 8 + i] += __bfloat162float(h[i]);
                }
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1]);
                c4[j] = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncols) {
                  float o = v[i] + ((p.epilogue & SST_EPI_ACCUM) ? __bfloat162float(cp[i]) : 0.f);
                  cp[i] = __float2bfloat16_rn(o);
                }
            }
          }
        }
        if (!next_requested && lc0 + 1 < CPW) ptx::tmem_ld_32x32b_x32(taddr0 + (uint32_t)((lc0 + 1) * 32), r);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) SST_TRACE(9 + 2 * ((u - unit0) / unit_stride));       // tile's epilogue (this warp) done
      if (lane == 0) {
        if (CTAS == 2) ptx::mbar_arrive_cluster(ptx::map_to_cta(&tempty[acc], 0));
        else ptx::mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  if (threadIdx.x == 64) SST_TRACE(4);                                 // first epilogue warp at the final barrier
  if (CTAS == 2) ptx::cluster_sync(); else __syncthreads();       // the pair is done with each other's smem / barriers
  if (threadIdx.x == 0) SST_TRACE(5);
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CTAS == 2) ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major matrix (cols contiguous), 128B swizzle, zero fill out of bounds.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long cols, long rows, long ld_elems, int box_cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  SST_REQUIRE(fn != nullptr, SST_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
  SST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, SST_E_ARG, "TMA operand not 16-byte aligned");
  SST_REQUIRE((ld_elems * 2) % 16 == 0, SST_E_ARG, "TMA operand row pitch (%ld elements) not a multiple of 16 bytes", ld_elems);
  SST_REQUIRE(cols > 0 && rows > 0, SST_E_ARG, "empty TMA operand");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SST_REQUIRE(r == CUDA_SUCCESS, SST_E_LAUNCH, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SST_OK;
}

template <int BN, int CTAS>
static int launch_bn(const SstGemmDesc& d, const void* A, const void* B, GemmKParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CTAS>;
  p.m_blks = cdiv(d.M, G_BM * CTAS);
  p.n_blks = cdiv(d.N, BN);
  const int tiles = p.m_blks * p.n_blks;
  const int sms = num_sms();
  int splits = 1;
  if (p.mode_mn) {
    // split K (= tokens) until the machine is full, but keep >= 16 k-blocks per unit: every unit pays a pipeline fill and
    // a (BM x BN) fp32 atomic epilogue
    splits = (2 * sms / CTAS) / tiles;
    const int max_by_k = p.num_kb / 16;
    if (splits > max_by_k) splits = max_by_k;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = cdiv(p.num_kb, splits);
  p.splits = cdiv(p.num_kb, p.kb_per_split);
  p.atomic_out = p.splits > 1;
  if (p.atomic_out) {
    SST_REQUIRE(d.out_dtype == SST_F32 && (d.epilogue & ~SST_EPI_ACCUM) == 0, SST_E_ARG,
                "split-K GEMM needs fp32 output and no epilogue besides ACCUM");
    if (!(d.epilogue & SST_EPI_ACCUM)) {
      cudaError_t e = cudaMemset2DAsync(p.C, (size_t)d.ldc * 4, 0, (size_t)d.N * 4, (size_t)d.M, st);
      SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
    }
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!p.mode_mn) {
    if ((rc = make_tmap_bf16_2d(&tmA, A, d.a_cols, d.a_rows, d.lda, G_BK, G_BM))) return rc;
  } else {
    if ((rc = make_tmap_bf16_2d(&tmA, A, d.a_cols, d.a_rows, d.lda, 64, G_BK))) return rc;
  }
  if (!p.b_mn) {
    if ((rc = make_tmap_bf16_2d(&tmB, B, d.K, d.N, d.ldb, G_BK, Cfg::B_ROWS))) return rc;
  } else {
    if ((rc = make_tmap_bf16_2d(&tmB, B, d.b_cols, d.b_rows, d.ldb, 64, G_BK))) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int units = tiles * p.splits;
  const int slots = sms / CTAS;                                  // CTAs (or CTA pairs = TPCs) that can be resident
  const int grid = (units < slots ? units : slots) * CTAS;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, CTAS>, tmA, tmB, p);
  SST_REQUIRE(le == cudaSuccess, SST_E_LAUNCH, "gemm_tcgen05 launch: %s", cudaGetErrorString(le));
  return check_launch("gemm_tcgen05");
}

#ifdef SST_GEMM_TRACE
extern "C" int sst_debug_gemm_trace(unsigned long long* out_host) {
  return cudaMemcpyFromSymbol(out_host, g_gemm_trace, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif

int launch_gemm_tcgen05(const SstGemmDesc& d, const void* A, const void* B, void* C, const void* bias, const void* aux,
                        cudaStream_t st) {
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)d.M; p.N = (int)d.N; p.K = (int)d.K;
  p.mode_mn = d.layout == SST_GEMM_NT_MN;
  p.b_mn = d.layout != SST_GEMM_TN;
  p.num_kb = cdiv(d.K, G_BK);
  const int nseg = d.n_seg > 0 ? d.n_seg : 1;
  SST_REQUIRE(nseg <= 3, SST_E_ARG, "n_seg must be <= 3");
  for (int s = 0; s < 3; ++s) {
    p.a_row_shift[s] = d.a_row_shift[s]; p.a_col0[s] = d.a_col0[s];
    p.b_row_shift[s] = d.b_row_shift[s]; p.b_col0[s] = d.b_col0[s];
  }
  if (!p.mode_mn) {
    SST_REQUIRE(nseg == 1 || (d.K % nseg == 0 && (d.K / nseg) % G_BK == 0), SST_E_ARG,
                "segmented A needs K/n_seg to be a multiple of %d", G_BK);
    p.kb_per_seg = nseg == 1 ? p.num_kb : (int)(d.K / nseg / G_BK);
    p.nseg_cols = (int)d.N;
    if (p.b_mn) { p.b_row_shift[0] = 0; p.b_col0[0] = 0; }      // TN_BMN: B is a plain (K, N) row-major matrix
  } else {
    p.kb_per_seg = p.num_kb;
    p.nseg_cols = (int)(d.N / nseg);
    SST_REQUIRE(nseg == 1 || (d.N % nseg == 0 && p.nseg_cols % 256 == 0), SST_E_ARG,
                "segmented B needs N/n_seg to be a multiple of 256");
  }
  p.epilogue = d.epilogue;
  p.alpha = d.alpha; p.mask_scale = d.mask_scale;
  p.drop_thr = drop_threshold16(d.drop_p);
  p.drop_scale = d.drop_p < 1.f ? 1.f / (1.f - d.drop_p) : 0.f;
  p.seed = d.seed;
  p.salt = dropout_salt();
  p.bias = reinterpret_cast<const float*>(bias);
  p.aux = aux; p.ldaux = d.ldaux; p.aux_f32 = d.aux_dtype == SST_F32;
  p.C = C; p.ldc = d.ldc; p.out_f32 = d.out_dtype == SST_F32;
  p.remap_P = d.remap_P; p.remap_T = d.remap_T; p.remap_j0 = d.remap_j0;
  p.col_acc = d.col_acc; p.col_mode = d.col_acc != nullptr ? d.col_acc_mode : 0;
  p.col_grp = d.col_acc_grp > 0 ? d.col_acc_grp : (int)d.N;
  if (p.col_mode != 0) {
    SST_REQUIRE((p.col_mode == 1 || p.col_mode == 2) && !p.mode_mn && d.out_dtype == SST_BF16 && !(d.epilogue & SST_EPI_ACCUM) &&
                d.N % 32 == 0 && d.ldc % 8 == 0 && d.out_seg_cols == 0 && p.col_grp % 8 == 0 && d.N % p.col_grp == 0, SST_E_ARG,
                "col_acc needs a bf16 TN result without ACCUM / segments, N %% 32 == 0, ldc %% 8 == 0 and groups of whole 8-column pieces");
    if (p.col_mode == 2) {
      cudaError_t e = cudaMemsetAsync(d.col_acc, 0, sizeof(double) * 2 * (size_t)d.N, st);
      SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
    }
  }
  p.out_seg_cols = (int)d.out_seg_cols; p.out_grp_cols = (int)d.out_grp_cols; p.out_seg_stride = d.out_seg_stride;
  for (int s = 0; s < 3; ++s) p.out_grp_off[s] = d.out_grp_off[s];
  if (d.out_seg_cols > 0)
    SST_REQUIRE(d.out_dtype == SST_F32 && d.out_seg_cols % 32 == 0 && d.out_grp_cols % d.out_seg_cols == 0 && d.N % 32 == 0 &&
                cdiv(d.N, d.out_grp_cols) <= 3 && d.ldc % 4 == 0 && d.out_seg_stride % 4 == 0, SST_E_ARG,
                "segmented output needs fp32 C, segments / groups of whole 32-column chunks and at most 3 groups");
  if (d.epilogue & SST_EPI_DROPOUT) SST_REQUIRE(d.N % 8 == 0, SST_E_ARG, "dropout epilogue needs N %% 8 == 0");
  // CTA pairs (256-row tiles) whenever there are at least two row blocks; SST_GEMM_CTAS=1 forces single-CTA tiles
  static const int ctas_env = [] { const char* e = getenv("SST_GEMM_CTAS"); return e ? atoi(e) : 2; }();
  const int ctas = (ctas_env == 2 && d.M > G_BM) ? 2 : 1;
  // tile width: 256 columns whenever N has more than 128.  Measured with SST_GEMM_TRACE: a 64-wide k-block of a pair tile
  // takes ~0.45 us whether the tile is 128 or 256 columns wide (MMA rate paces the wide tile, operand latency the narrow
  // one), so a 128-wide tile costs almost as much as a 256-wide one and an extra wave of them never pays.
  const bool wide = p.mode_mn ? (d.N % 256 == 0) : (d.N > 128);
  if (ctas == 2) return wide ? launch_bn<256, 2>(d, A, B, p, st) : launch_bn<128, 2>(d, A, B, p, st);
  return wide ? launch_bn<256, 1>(d, A, B, p, st) : launch_bn<128, 1>(d, A, B, p, st);
}

}  // namespace sst

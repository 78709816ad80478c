// Shared pieces of the tensor-core attention kernels (attention_tc.cu forward, attention_tc_bwd.cu backward):
// tile geometry, the per-row logit evaluation (mask SET -1e8, relative-position bias via the register barrel shift),
// dropout and the swizzled shared-memory stores that turn register rows into tcgen05 operands.
#pragma once
#include "sst_common.cuh"
#include "sst_ptx.cuh"
#include <utility>

namespace sst {

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long cols, long rows, long ld_elems, int box_cols, int box_rows);

namespace attn_tc {

constexpr int BM = 128;            // query rows per tile  (TMEM lanes, one thread each)
constexpr int BN = 64;             // keys per tile
constexpr int PBW = 192;           // relative offsets a (BM x BN) tile can touch: BM + BN - 1 = 191, padded to the MMA N step
constexpr float NEG_MASK = -1e8f;  // the reference's masked_fill / out-of-range value (transformer.py:181-196, :354-357)
constexpr float NEG_BIG = -3.0e38f;  // keys that do not exist (j >= Lk): excluded from the softmax altogether

struct AttnTcParams {
  int B, H, Lq, Lk, Lkp;           // Lkp = Lk rounded up to 4: pitch of the dropout counter space
  int causal, mask_q_rows, R, band;
  float scale;
  uint32_t thr; float dscale; unsigned long long seed;
  const int* q_lens; const int* k_lens;
  __nv_bfloat16* o; long ldo;
  float* lse;
  // backward only
  const __nv_bfloat16* dO;                 // pitch ldo
  float* delta;                            // [B*H*Lq]  dO_i . O_i   (written by the dQ kernel, read by the dK/dV kernel)
  __nv_bfloat16 *dq, *dk, *dv; long ldq, ldk, ldv;
};

inline AttnTcParams make_tc_params(const SstAttnDesc& d, const int* q_lens, const int* k_lens) {
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.H = d.H; p.Lq = d.Lq; p.Lk = d.Lk; p.Lkp = (d.Lk + 3) & ~3;
  p.causal = d.causal; p.mask_q_rows = d.mask_q_rows; p.R = d.rel_dist;
  p.band = d.rel_dist > 0 && d.Lk > d.rel_dist;
  p.scale = d.scale;
  p.thr = d.drop_p > 0.f ? drop_threshold(d.drop_p) : 0u;
  p.dscale = d.drop_p < 1.f ? 1.f / (1.f - d.drop_p) : 0.f;
  p.seed = d.seed;
  p.q_lens = q_lens; p.k_lens = k_lens;
  return p;
}

// key tiles [t_lo, t_hi] a query tile starting at i0 has to visit
__device__ __forceinline__ void key_tile_range(const AttnTcParams& p, int i0, int& t_lo, int& t_hi) {
  int j_lo = 0, j_hi = p.Lk - 1;
  if (p.band) {
    const int i_last = min(i0 + BM - 1, p.Lq - 1);
    j_lo = max(0, i0 - (p.R - 1));
    j_hi = min(p.Lk - 1, i_last + p.R - 1);
  }
  t_lo = j_lo / BN;
  t_hi = j_hi / BN;
}
// query tiles [q_lo, q_hi] (of BM rows) a key tile starting at j0 receives contributions from
__device__ __forceinline__ void query_tile_range(const AttnTcParams& p, int j0, int& q_lo, int& q_hi) {
  int i_lo = 0, i_hi = p.Lq - 1;
  if (p.band) {
    const int j_last = min(j0 + BN - 1, p.Lk - 1);
    i_lo = max(0, j0 - (p.R - 1));
    i_hi = min(p.Lq - 1, j_last + p.R - 1);
  }
  q_lo = i_lo / BM;
  q_hi = i_hi / BM;
}

struct RowCtx {
  long row_id;      // (b*H + h)*Lq + i
  int i;            // query index inside the utterance
  int klen;         // keys j >= klen are padding (masked, -1e8)
  bool rowmask;     // whole query row is padding (masked, -1e8)
};

__device__ __forceinline__ RowCtx make_row_ctx(const AttnTcParams& p, int b, int h, int i) {
  RowCtx rc;
  rc.row_id = ((long)b * p.H + h) * p.Lq + i;
  rc.i = i;
  rc.klen = p.k_lens ? p.k_lens[b] : p.Lk;
  rc.rowmask = p.mask_q_rows && p.q_lens && i >= p.q_lens[b];
  return rc;
}

// compile-time loop: indices stay constants, so the row arrays live in registers whatever the unroller's thresholds say
template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// U[x] <- U[x + K] for the lanes whose shift amount has bit K set
template <int K>
__device__ __forceinline__ void shift_stage(float (&U)[96], int s) {
  const bool sh = (s & K) != 0;
  static_for<BN + K - 1>([&](auto x) { U[x] = sh ? U[x + K] : U[x]; });
}

// Relative-position logits of one query row against the 64 keys of the tile: U[lj] = PB[li][lj - li + 127], lj in [0, 64).
// The thread's TMEM lane is row li = 32*w + lane.  Columns [96 - 32*w, 96 - 32*w + 96) of PB are loaded (warp-uniform
// address), leaving a lane-dependent left shift by s = 31 - lane, done as five select stages (16, 8, 4, 2, 1).
__device__ __forceinline__ void load_bias_window(uint32_t tPB /* incl. lane base */, int w, int lane, float (&U)[96]) {
  const uint32_t cb = 96 - 32 * w;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(tPB + cb + c * 32, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int x = 0; x < 32; ++x) U[c * 32 + x] = __uint_as_float(r[x]);
  }
  const int s = 31 - lane;
  shift_stage<16>(U, s); shift_stage<8>(U, s); shift_stage<4>(U, s); shift_stage<2>(U, s); shift_stage<1>(U, s);
}

// logits of one query row against the tile's keys, exactly as MultiHeadAttention.forward builds them:
//   s = masked ? -1e8 : q.k * scale;   s += |j - i| < R ? q.E[j-i+R-1] : -1e8   (transformer.py:177-204, Q3/Q9)
// `masked_out[lj>>5]` bit (lj&31) reports that the q.k term was overwritten (its gradient is zero).
template <bool WANT_MASK>
__device__ __forceinline__ void tile_logits(const AttnTcParams& p, const RowCtx& rc, uint32_t tS, uint32_t tPB, int w, int lane,
                                            int j0, float (&U)[96], uint32_t* masked_out = nullptr) {
  if (p.R > 0) load_bias_window(tPB, w, lane, U);
#pragma unroll
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(tS + c * 32, r);
    ptx::tmem_ld_wait();
    uint32_t mbits = 0;
#pragma unroll
    for (int x = 0; x < 32; ++x) {
      const int lj = c * 32 + x;
      const int j = j0 + lj;
      const bool masked = rc.rowmask || j >= rc.klen || (p.causal && j > rc.i);
      float s = masked ? NEG_MASK : __uint_as_float(r[x]) * p.scale;
      if (p.R > 0) {
        const int rel = j - rc.i;
        s += (rel > -p.R && rel < p.R) ? U[lj] : NEG_MASK;
      }
      U[lj] = j < p.Lk ? s : NEG_BIG;
      if (WANT_MASK) mbits |= (masked ? 1u : 0u) << x;
    }
    if (WANT_MASK) masked_out[c] = mbits;
  }
}

// dropout on the probabilities of one row (counter = row_id * Lkp + j, four keys per Philox block; the same stream as
// the CUDA-core kernels of attention_simt.cu)
__device__ __forceinline__ void apply_dropout(const AttnTcParams& p, const RowCtx& rc, int j0, float (&U)[96]) {
  const unsigned long long base = ((unsigned long long)rc.row_id * p.Lkp + j0) >> 2;
#pragma unroll
  for (int g = 0; g < BN / 4; ++g) {
    const Philox4 r = philox4x32_10(p.seed, base + g);
    U[4 * g + 0] = r.x >= p.thr ? U[4 * g + 0] * p.dscale : 0.f;
    U[4 * g + 1] = r.y >= p.thr ? U[4 * g + 1] * p.dscale : 0.f;
    U[4 * g + 2] = r.z >= p.thr ? U[4 * g + 2] * p.dscale : 0.f;
    U[4 * g + 3] = r.w >= p.thr ? U[4 * g + 3] * p.dscale : 0.f;
  }
}

// One 64-element row (128 bytes of bf16) into a K-major SWIZZLE_128B operand tile: 16-byte chunk c of row r lives at
// chunk c ^ (r & 7).  `sbase` must be 1024-byte aligned.
template <int N>
__device__ __forceinline__ void store_row_bf16_sw128(uint32_t sbase, int row, const float (&U)[N]) {
  static_assert(N >= BN, "row array too short");
  const uint32_t rbase = sbase + row * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t w4[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(U[c * 8 + 2 * x], U[c * 8 + 2 * x + 1]);
      w4[x] = *reinterpret_cast<uint32_t*>(&h2);
    }
    ptx::st_shared_v4(rbase + ((uint32_t)(c ^ (row & 7)) << 4), w4[0], w4[1], w4[2], w4[3]);
  }
}

}  // namespace attn_tc
}  // namespace sst

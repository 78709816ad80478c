// Shared pieces of the tensor-core attention kernels (attention_tc.cu forward, attention_tc_bwd.cu backward):
// tile geometry, the per-row logit evaluation (mask SET -1e8, relative-position bias via the register barrel shift),
// dropout and the swizzled shared-memory stores that turn register rows into tcgen05 operands.
//
// Thread mapping (all three kernels): a (128 query x 64 key) tile lives in TMEM with the query row on the lane axis.
// Warp w reads lane quarter q = w & 3 (the only one tcgen05.ld lets it touch) and column group hf = w >> 2: with
// NSPLIT column groups a thread owns CW = 64 / NSPLIT consecutive keys of ONE query row, so 4 * NSPLIT warps share a
// tile (more warps per scheduler hide the TMEM-load / MUFU / Philox latencies of the per-element math).
#pragma once
#include "sst_common.cuh"
#include "sst_ptx.cuh"
#include <utility>

namespace sst {

int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long cols, long rows, long ld_elems, int box_cols, int box_rows);

namespace attn_tc {

constexpr int BM = 128;            // query rows per tile  (TMEM lanes)
constexpr int BN = 64;             // keys per tile
constexpr int PBW = 192;           // relative offsets a (BM x BN) tile can touch: BM + BN - 1 = 191, padded to the MMA N step
constexpr float NEG_MASK = -1e8f;  // the reference's masked_fill / out-of-range value (transformer.py:181-196, :354-357)
constexpr float NEG_BIG = -3.0e38f;  // keys that do not exist (j >= Lk): excluded from the softmax altogether

struct AttnTcParams;

struct AttnTcParams {
  int B, H, Lq, Lk, Lkp;           // Lkp = Lk rounded up to 8: pitch of the dropout counter space
  int causal, mask_q_rows, R, band;
  float scale;
  uint32_t thr; float dscale; unsigned long long seed;   // thr: 16-bit keep threshold (0 = no dropout)
  const unsigned long long* salt;                         // added to the seed when registered (sst_common.cuh)
  const int* q_lens; const int* k_lens;
  const uint8_t* q_pad; const uint8_t* k_pad;   // optional per-position padding masks (OR-ed with the length masks)
  const long long* q_off; const long long* k_off;   // packed layouts (include/sst.h): first row of entry b, or null = b*Lq / b*Lk
  __nv_bfloat16* o; long ldo;
  float* lse;
  // backward only
  const __nv_bfloat16* dO;                 // pitch ldo
  float* delta;                            // [B*H*Lq]  dO_i . O_i   (written by the dQ kernel, read by the dK/dV kernel)
  __nv_bfloat16 *dq, *dk, *dv; long ldq, ldk, ldv;
  // (query tile, key tile) hand-off between the two backward kernels: bf16 (128 x 64) tiles of P~ and of scale*dS, row-major,
  // tile index = ((b*H + h)*nQT + u)*max_kt + (kt - first key tile of query tile u)
  __nv_bfloat16 *ws_p, *ws_ds;
  int nQT, max_kt;
};

// key tiles a query tile can touch at most (size of one query tile's slot row in the hand-off workspace)
inline int attn_max_key_tiles(const SstAttnDesc& d) {
  const int nkt = (d.Lk + BN - 1) / BN;
  if (!(d.rel_dist > 0 && d.Lk > d.rel_dist)) return nkt;
  const int m = (BM + 2 * d.rel_dist - 3) / BN + 2;
  return m < nkt ? m : nkt;
}
inline size_t attn_bwd_ws_bytes(const SstAttnDesc& d) {
  const size_t tiles = (size_t)d.B * d.H * ((d.Lq + BM - 1) / BM) * attn_max_key_tiles(d);
  return 2 * tiles * BM * BN * sizeof(__nv_bfloat16);
}

inline AttnTcParams make_tc_params(const SstAttnDesc& d, const int* q_lens, const int* k_lens) {
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  p.B = d.B; p.H = d.H; p.Lq = d.Lq; p.Lk = d.Lk; p.Lkp = (d.Lk + 7) & ~7;
  p.causal = d.causal; p.mask_q_rows = d.mask_q_rows; p.R = d.rel_dist;
  p.band = d.rel_dist > 0 && d.Lk > d.rel_dist;
  p.scale = d.scale;
  p.thr = d.drop_p > 0.f ? drop_threshold16(d.drop_p) : 0u;
  p.dscale = d.drop_p < 1.f ? 1.f / (1.f - d.drop_p) : 0.f;
  p.seed = d.seed;
  p.salt = dropout_salt();
  p.q_lens = q_lens; p.k_lens = k_lens;
  p.q_pad = d.q_pad; p.k_pad = d.k_pad;
  p.q_off = reinterpret_cast<const long long*>(d.q_off); p.k_off = reinterpret_cast<const long long*>(d.k_off);
  return p;
}

// rows of the token matrices the tensor maps may touch: the padded B*L, or what a packed layout says it holds
inline long attn_q_rows_total(const SstAttnDesc& d) { return d.q_off ? d.q_rows_total : (long)d.B * d.Lq; }
inline long attn_k_rows_total(const SstAttnDesc& d) { return d.k_off ? d.k_rows_total : (long)d.B * d.Lk; }

// first row of batch entry b in the query-side (q, o, dO, dq) / key-side (k, v, dk, dv) matrices and the rows that exist for it
__device__ __forceinline__ int q_base(const AttnTcParams& p, int b) { return p.q_off ? (int)p.q_off[b] : b * p.Lq; }
__device__ __forceinline__ int k_base(const AttnTcParams& p, int b) { return p.k_off ? (int)p.k_off[b] : b * p.Lk; }
__device__ __forceinline__ int q_rows(const AttnTcParams& p, int b) { return p.q_off ? min(p.q_lens[b], p.Lq) : p.Lq; }
__device__ __forceinline__ int k_rows(const AttnTcParams& p, int b) { return p.k_off ? min(p.k_lens[b], p.Lk) : p.Lk; }
// packed query side: a (batch, query tile) pair whose first row lies beyond the entry's length has no rows at all
__device__ __forceinline__ bool q_tile_exists(const AttnTcParams& p, int b, int i0) { return p.q_off == nullptr || i0 < p.q_lens[b]; }

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
// The same range without the key tiles that lie entirely in utterance b's padding (j >= k_lens[b]): every key there is
// masked, so its probability is exactly 0 in fp32 for any query row that has an unmasked key -- skipping those tiles changes
// nothing.  Only done for query tiles without a masked row: a fully masked row spreads uniformly over ALL keys in the
// reference (transformer.py:185-187), padding included, and stays bit-compatible with that here.
__device__ __forceinline__ void key_tile_range_valid(const AttnTcParams& p, int i0, int b, int& t_lo, int& t_hi) {
  key_tile_range(p, i0, t_lo, t_hi);
  // (packed query side: rows beyond the length do not exist, so there is no masked row to keep compatible)
  const bool rows_unmasked = p.q_off != nullptr || !p.mask_q_rows ||
                             (p.q_pad == nullptr && (p.q_lens == nullptr || i0 + BM <= p.q_lens[b]));
  if (p.k_lens != nullptr && rows_unmasked) {
    const int klen = p.k_lens[b];
    // klen == 0 (an empty memory): EVERY key is masked and the reference's softmax is uniform over all Lk of them -- no
    // tile may be skipped then
    if (klen > 0) t_hi = max(t_lo, min(t_hi, (klen - 1) / BN));      // (klen - 1) / BN: last tile that holds a real key
  }
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
  const uint8_t* kpad;   // per-key padding flags of this utterance (nullable)
};

__device__ __forceinline__ RowCtx make_row_ctx(const AttnTcParams& p, int b, int h, int i) {
  RowCtx rc;
  rc.row_id = ((long)b * p.H + h) * p.Lq + i;
  rc.i = i;
  rc.klen = p.k_lens ? p.k_lens[b] : p.Lk;
  rc.rowmask = p.mask_q_rows && ((p.q_lens && i >= p.q_lens[b]) || (p.q_pad && i < p.Lq && p.q_pad[(long)b * p.Lq + i]));
  rc.kpad = p.k_pad ? p.k_pad + (long)b * p.Lk : nullptr;
  return rc;
}

// compile-time loop: indices stay constants, so the row arrays live in registers whatever the unroller's thresholds say
template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

template <int NSPLIT> struct Split {
  static constexpr int CW = BN / NSPLIT;                  // keys per thread
  static constexpr int WIN = CW + 31;                     // bias window before the lane shift
  static constexpr int WIN_LD = (WIN + 15) / 16 * 16;     // columns actually loaded (x16 granules)
  static constexpr int THREADS = 128 * NSPLIT;
  static_assert(CW == 16 || CW == 32, "one or two 16-column TMEM loads per thread");
};

// `n16` x16 loads issued back to back, one wait
template <int N>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, float (&out)[N]) {
  static_assert(N % 16 == 0, "x16 granules");
  uint32_t r[N / 16][16];
#pragma unroll
  for (int c = 0; c < N / 16; ++c) ptx::tmem_ld_32x32b_x16(taddr + c * 16, r[c]);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < N / 16; ++c)
#pragma unroll
    for (int x = 0; x < 16; ++x) out[c * 16 + x] = __uint_as_float(r[c][x]);
}

// U[x] <- U[x + K] for the lanes whose shift amount has bit K set
template <int K, int CW, int N>
__device__ __forceinline__ void shift_stage(float (&U)[N], int s) {
  const bool sh = (s & K) != 0;
  static_for<CW + K - 1>([&](auto x) { U[x] = sh ? U[x + K] : U[x]; });
}

// Relative-position logits of one query row against the thread's CW keys: U[x] = PB[li][lj - li + 127], lj = CW*hf + x.
// Row li = 32*q + lane.  Columns [96 - 32*q + CW*hf, + WIN) of PB are loaded (warp-uniform address), leaving a
// lane-dependent left shift by s = 31 - lane, done as five select stages (16, 8, 4, 2, 1).
template <int NSPLIT>
__device__ __forceinline__ void load_bias_window(uint32_t tPB /* incl. lane base */, int q, int hf, int lane,
                                                 float (&U)[Split<NSPLIT>::WIN_LD]) {
  constexpr int CW = Split<NSPLIT>::CW;
  tmem_load_cols(tPB + (uint32_t)(96 - 32 * q + CW * hf), U);
  const int s = 31 - lane;
  shift_stage<16, CW>(U, s); shift_stage<8, CW>(U, s); shift_stage<4, CW>(U, s); shift_stage<2, CW>(U, s); shift_stage<1, CW>(U, s);
}

// bit x set  <=>  key jb + x lies inside the band of query rc.i: -R < (jb + x) - i < R  (one interval of x, built once per tile)
template <int CW>
__device__ __forceinline__ uint32_t band_bits(const AttnTcParams& p, const RowCtx& rc, int jb) {
  const int d = jb - rc.i;                               // rel = d + x
  const int lo = max(0, -p.R + 1 - d), hi = min(CW, p.R - d);       // x in [lo, hi)
  if (hi <= lo) return 0u;
  const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return upto_hi & ~((1u << lo) - 1u);
}

// logits of one query row against the thread's keys, exactly as MultiHeadAttention.forward builds them:
//   s = masked ? -1e8 : q.k * scale;   s += |j - i| < R ? q.E[j-i+R-1] : -1e8   (transformer.py:177-204, Q3/Q9)
// Bit x of `mbits` reports that the q.k term of key x was overwritten (its gradient is zero).
// `simple` (warp-uniform): no causal mask, every key of the tile exists and is unpadded, no padded row in the warp.
template <int NSPLIT>
__device__ __forceinline__ void tile_logits(const AttnTcParams& p, const RowCtx& rc, uint32_t tS, uint32_t tPB, int q, int hf, int lane,
                                            int j0, bool simple, float (&U)[Split<NSPLIT>::WIN_LD], uint32_t& mbits) {
  constexpr int CW = Split<NSPLIT>::CW;
  if (p.R > 0) load_bias_window<NSPLIT>(tPB, q, hf, lane, U);
  float sv[CW];
  tmem_load_cols(tS + (uint32_t)(CW * hf), sv);
  const int jb = j0 + CW * hf;                          // first key of this thread
  const uint32_t band = band_bits<CW>(p, rc, jb);       // bit x: key jb + x is inside the band |i - j| < R
  mbits = 0;
  if (simple) {
#pragma unroll
    for (int x = 0; x < CW; ++x) {
      float s = sv[x] * p.scale;
      if (p.R > 0) s += ((band >> x) & 1u) ? U[x] : NEG_MASK;
      U[x] = s;
    }
  } else {
#pragma unroll
    for (int x = 0; x < CW; ++x) {
      const int j = jb + x;
      const bool masked = rc.rowmask || j >= rc.klen || (p.causal && j > rc.i) || (rc.kpad && j < p.Lk && rc.kpad[j]);
      float s = masked ? NEG_MASK : sv[x] * p.scale;
      if (p.R > 0) s += ((band >> x) & 1u) ? U[x] : NEG_MASK;
      U[x] = j < p.Lk ? s : NEG_BIG;
      mbits |= (masked ? 1u : 0u) << x;
    }
  }
}

// The (32 query rows of this warp) x (CW keys of this thread's column group) block lies entirely outside the band
// |i - j| < R: every logit there is <= -1e8 below an in-band one, its probability is exactly 0 in fp32 and the block
// needs no TMEM load, shift, exp or Philox at all.  Warp-uniform.
template <int NSPLIT>
__device__ __forceinline__ bool block_out_of_band(const AttnTcParams& p, int iw0 /* first row of the warp */, int jb /* first key */) {
  constexpr int CW = Split<NSPLIT>::CW;
  return p.R > 0 && (jb - (iw0 + 31) >= p.R || iw0 - (jb + CW - 1) >= p.R);
}

template <int CW>
__device__ __forceinline__ void store_zero_cols_sw128(uint32_t sbase, int row, int col0) {
  const uint32_t rbase = sbase + row * 128;
#pragma unroll
  for (int c = 0; c < CW / 8; ++c) ptx::st_shared_v4(rbase + ((uint32_t)(((col0 >> 3) + c) ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
}

__device__ __forceinline__ bool tile_is_simple(const AttnTcParams& p, const RowCtx& rc, int j0) {
  const bool mine = !p.causal && !rc.rowmask && !rc.kpad && (j0 + BN <= min(rc.klen, p.Lk));
  return __all_sync(0xffffffffu, mine);
}

// keep factors (0 or 1/(1-pd)) of the thread's CW keys: counter = row_id * Lkp + j, eight keys per Philox block, the same
// stream as the CUDA-core kernels of attention_simt.cu
template <int NSPLIT>
__device__ __forceinline__ void dropout_keep(const AttnTcParams& p, unsigned long long seed, const RowCtx& rc, int j0, int hf,
                                             float (&keep)[Split<NSPLIT>::CW]) {
  constexpr int CW = Split<NSPLIT>::CW;
  const unsigned long long base = ((unsigned long long)rc.row_id * p.Lkp + j0 + CW * hf) >> 3;
#pragma unroll
  for (int g = 0; g < CW / 8; ++g) {
    const Philox4 r = philox4x32(seed, base + g);
    const uint32_t thr_hi = p.thr << 16;
#pragma unroll
    for (int e = 0; e < 8; ++e) keep[g * 8 + e] = philox_keep16_at(r, e, thr_hi) ? p.dscale : 0.f;
  }
}

// CW consecutive elements of one row (bf16) into a K-major SWIZZLE_128B operand tile whose rows are 64 elements
// (128 bytes): 16-byte chunk c of row r lives at chunk c ^ (r & 7).  `sbase` must be 1024-byte aligned.
template <int CW, int N>
__device__ __forceinline__ void store_cols_bf16_sw128(uint32_t sbase, int row, int col0, const float (&U)[N]) {
  static_assert(N >= CW, "row array too short");
  const uint32_t rbase = sbase + row * 128;
#pragma unroll
  for (int c = 0; c < CW / 8; ++c) {
    uint32_t w4[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(U[c * 8 + 2 * x], U[c * 8 + 2 * x + 1]);
      w4[x] = *reinterpret_cast<uint32_t*>(&h2);
    }
    ptx::st_shared_v4(rbase + ((uint32_t)(((col0 >> 3) + c) ^ (row & 7)) << 4), w4[0], w4[1], w4[2], w4[3]);
  }
}

// TMEM accumulator row slice [c0, c0 + NC) -> bf16 global row (scaled); every lane loads (warp-collective), `valid` stores
template <int NC>
__device__ __forceinline__ void tmem_row_to_global(uint32_t taddr, __nv_bfloat16* dst, float mul, bool valid) {
  static_assert(NC % 8 == 0, "x8 granules");
#pragma unroll
  for (int c = 0; c < NC / 8; ++c) {
    uint32_t r[8];
    ptx::tmem_ld_32x32b_x8(taddr + c * 8, r);
    ptx::tmem_ld_wait();
    if (valid) {
      uint4 o4;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o4);
#pragma unroll
      for (int x = 0; x < 4; ++x) o2[x] = __floats2bfloat162_rn(__uint_as_float(r[2 * x]) * mul, __uint_as_float(r[2 * x + 1]) * mul);
      *reinterpret_cast<uint4*>(dst + c * 8) = o4;
    }
    __syncwarp();
  }
}

}  // namespace attn_tc
}  // namespace sst

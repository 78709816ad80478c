// CUDA-core attention (fp32 math; float or bf16 storage): the fp32 parity-mode implementation of the fused
// masked / relative-position attention of include/sst.h, and the device-side cross-check for the tensor-core
// kernel.  One warp per query row (forward, dq) or per key row (dk, dv); logits are recomputed in backward from
// the saved log-sum-exp, nothing of size (L x L) is ever written to memory.
#include "vec.cuh"

namespace sst {

struct AttnP {
  int B, H, Lq, Lk, Lkp, dh;      // Lkp = Lk rounded up to 8: pitch of the dropout counter space (shared with attention_tc)
  long ldq, ldk, ldv, ldo;
  int causal, mask_q_rows, R;
  float scale;
  uint32_t thr; float dscale; unsigned long long seed;
  const unsigned long long* salt;
  const int* q_lens; const int* k_lens;
  const uint8_t* q_pad; const uint8_t* k_pad;
  const long long* q_off; const long long* k_off;     // packed layouts (include/sst.h): first row of entry b, or null
};

// first row of batch entry b in the query-side / key-side matrices, and the rows that exist for it
__device__ __forceinline__ long q_base(const AttnP& p, int b) { return p.q_off ? (long)p.q_off[b] : (long)b * p.Lq; }
__device__ __forceinline__ long k_base(const AttnP& p, int b) { return p.k_off ? (long)p.k_off[b] : (long)b * p.Lk; }
__device__ __forceinline__ int q_rows(const AttnP& p, int b) { return p.q_off ? min(p.q_lens[b], p.Lq) : p.Lq; }
__device__ __forceinline__ int k_rows(const AttnP& p, int b) { return p.k_off ? min(p.k_lens[b], p.Lk) : p.Lk; }

template <typename T>
__device__ __forceinline__ float dot_row(const float* __restrict__ a_smem, const T* __restrict__ row, int dh) {
  float acc = 0.f;
  for (int a = 0; a < dh; a += 8) {
    float v[8];
    Vec8<T>::load(row + a, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(a_smem[a + j], v[j], acc);
  }
  return acc;
}

// logit of (i, j) given the already computed q.k and q.E dots; `masked` reports whether the q.k term was overwritten.
__device__ __forceinline__ float make_logit(const AttnP& p, int b, int i, int j, float qk, float qe, bool& masked) {
  masked = (p.causal && j > i) || (p.k_lens && j >= p.k_lens[b]) || (p.mask_q_rows && p.q_lens && i >= p.q_lens[b]) ||
           (p.k_pad && p.k_pad[(long)b * p.Lk + j]) || (p.mask_q_rows && p.q_pad && p.q_pad[(long)b * p.Lq + i]);
  float s = masked ? -1e8f : qk * p.scale;
  if (p.R > 0) {
    int rel = j - i;
    s += (rel > -p.R && rel < p.R) ? qe : -1e8f;
  }
  return s;
}

// keys a query row visits.  Packed layouts: keys j >= k_rows(b) do not exist in memory (their probability is exactly 0 in the
// padded layout as well, where they are masked with -1e8 next to at least one real key)
__device__ __forceinline__ void key_range(const AttnP& p, int b, int i, int& lo, int& hi) {
  lo = 0; hi = p.Lk - 1;
  if (p.R > 0 && p.Lk > p.R) { lo = max(0, i - p.R + 1); hi = min(p.Lk - 1, i + p.R - 1); }
  hi = min(hi, k_rows(p, b) - 1);
}

template <typename T>
__global__ void __launch_bounds__(128)
attn_fwd_simt(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ E,
              T* __restrict__ o, float* __restrict__ lse, const AttnP p, int maxk) {
  extern __shared__ float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = sm + w * (p.dh + maxk);
  float* sc = qs + p.dh;
  const long row_id = (long)blockIdx.x * 4 + w;
  const long nrows = (long)p.B * p.H * p.Lq;
  if (row_id >= nrows) return;
  const int i = (int)(row_id % p.Lq);
  const int h = (int)((row_id / p.Lq) % p.H);
  const int b = (int)(row_id / ((long)p.Lq * p.H));
  if (i >= q_rows(p, b)) return;                 // packed layout: the row does not exist
  const long qb = q_base(p, b), kb = k_base(p, b);
  const T* qrow = q + (qb + i) * p.ldq + h * p.dh;
  for (int a = lane; a < p.dh; a += 32) qs[a] = to_f32(qrow[a]);
  __syncwarp();
  int lo, hi;
  key_range(p, b, i, lo, hi);
  const int nk = hi - lo + 1;
  float mx = -INFINITY;
  for (int jj = lane; jj < nk; jj += 32) {
    const int j = lo + jj;
    const float qk = dot_row(qs, k + (kb + j) * p.ldk + h * p.dh, p.dh);
    float qe = 0.f;
    const int rel = j - i;
    if (p.R > 0 && rel > -p.R && rel < p.R) qe = dot_row(qs, E + ((long)h * (2 * p.R - 1) + rel + p.R - 1) * p.dh, p.dh);
    bool masked;
    const float s = make_logit(p, b, i, j, qk, qe, masked);
    sc[jj] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int jj = lane; jj < nk; jj += 32) { float e = __expf(sc[jj] - mx); sc[jj] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  // saved as (max, log-sum) pairs: a fully masked row has max = -1e8, where max + log(sum) would round the sum away
  if (lane == 0) { lse[row_id] = mx; lse[nrows + row_id] = __logf(sum); }
  for (int jj = lane; jj < nk; jj += 32) {
    float pr = sc[jj] * inv;
    if (p.thr) pr = philox_keep16(salted(p.seed, p.salt), (unsigned long long)row_id * p.Lkp + (lo + jj), p.thr) ? pr * p.dscale : 0.f;
    sc[jj] = pr;
  }
  __syncwarp();
  for (int a = lane; a < p.dh; a += 32) {
    float acc = 0.f;
    const T* vp = v + (kb + lo) * p.ldv + h * p.dh + a;
    for (int jj = 0; jj < nk; ++jj) acc = fmaf(sc[jj], to_f32(vp[(long)jj * p.ldv]), acc);
    o[(qb + i) * p.ldo + h * p.dh + a] = from_f32<T>(acc);
  }
}

// dq (one warp per query row); also writes delta[row] = dO_i . O_i
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dq_simt(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ E,
                 const T* __restrict__ o, const T* __restrict__ dO, const float* __restrict__ lse, float* __restrict__ delta,
                 T* __restrict__ dq, const AttnP p, int maxk) {
  extern __shared__ float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = sm + w * (2 * p.dh + 2 * maxk);
  float* dos = qs + p.dh;
  float* dsq = dos + p.dh;       // ds * scale for unmasked q.k terms
  float* dsb = dsq + maxk;       // ds for the bias term
  const long row_id = (long)blockIdx.x * 4 + w;
  if (row_id >= (long)p.B * p.H * p.Lq) return;
  const int i = (int)(row_id % p.Lq);
  const int h = (int)((row_id / p.Lq) % p.H);
  const int b = (int)(row_id / ((long)p.Lq * p.H));
  if (i >= q_rows(p, b)) return;
  const long tok = q_base(p, b) + i, kb = k_base(p, b);
  float dl = 0.f;
  for (int a = lane; a < p.dh; a += 32) {
    qs[a] = to_f32(q[tok * p.ldq + h * p.dh + a]);
    float d = to_f32(dO[tok * p.ldo + h * p.dh + a]);
    dos[a] = d;
    dl = fmaf(d, to_f32(o[tok * p.ldo + h * p.dh + a]), dl);
  }
  dl = warp_sum(dl);
  if (lane == 0) delta[row_id] = dl;
  __syncwarp();
  const float Lm = lse[row_id], Ll = lse[(long)p.B * p.H * p.Lq + row_id];
  int lo, hi;
  key_range(p, b, i, lo, hi);
  const int nk = hi - lo + 1;
  for (int jj = lane; jj < nk; jj += 32) {
    const int j = lo + jj;
    const float qk = dot_row(qs, k + (kb + j) * p.ldk + h * p.dh, p.dh);
    float qe = 0.f;
    const int rel = j - i;
    const bool inband = p.R > 0 && rel > -p.R && rel < p.R;
    if (inband) qe = dot_row(qs, E + ((long)h * (2 * p.R - 1) + rel + p.R - 1) * p.dh, p.dh);
    bool masked;
    const float s = make_logit(p, b, i, j, qk, qe, masked);
    const float pr = __expf((s - Lm) - Ll);
    float dp = dot_row(dos, v + (kb + j) * p.ldv + h * p.dh, p.dh);
    if (p.thr) dp = philox_keep16(salted(p.seed, p.salt), (unsigned long long)row_id * p.Lkp + j, p.thr) ? dp * p.dscale : 0.f;
    const float ds = pr * (dp - dl);
    dsq[jj] = masked ? 0.f : ds * p.scale;
    dsb[jj] = inband ? ds : 0.f;
  }
  __syncwarp();
  for (int a = lane; a < p.dh; a += 32) {
    float acc = 0.f;
    const T* kp = k + (kb + lo) * p.ldk + h * p.dh + a;
    for (int jj = 0; jj < nk; ++jj) acc = fmaf(dsq[jj], to_f32(kp[(long)jj * p.ldk]), acc);
    if (p.R > 0) {
      for (int jj = 0; jj < nk; ++jj) {
        const int rel = lo + jj - i;
        if (rel > -p.R && rel < p.R) acc = fmaf(dsb[jj], to_f32(E[((long)h * (2 * p.R - 1) + rel + p.R - 1) * p.dh + a]), acc);
      }
    }
    dq[tok * p.ldq + h * p.dh + a] = from_f32<T>(acc);
  }
}

// dk, dv (one warp per key row)
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_simt(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ E,
                  const T* __restrict__ dO, const float* __restrict__ lse, const float* __restrict__ delta,
                  T* __restrict__ dk, T* __restrict__ dv, const AttnP p, int maxq) {
  extern __shared__ float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ks = sm + w * (2 * p.dh + 2 * maxq);
  float* vs = ks + p.dh;
  float* pw = vs + p.dh;         // dropped-out probabilities
  float* dsw = pw + maxq;        // ds * scale (0 where the q.k term was masked)
  const long row_id = (long)blockIdx.x * 4 + w;
  if (row_id >= (long)p.B * p.H * p.Lk) return;
  const int j = (int)(row_id % p.Lk);
  const int h = (int)((row_id / p.Lk) % p.H);
  const int b = (int)(row_id / ((long)p.Lk * p.H));
  if (j >= k_rows(p, b)) return;                 // packed layout: the key does not exist
  const long tokk = k_base(p, b) + j, qb = q_base(p, b);
  for (int a = lane; a < p.dh; a += 32) {
    ks[a] = to_f32(k[tokk * p.ldk + h * p.dh + a]);
    vs[a] = to_f32(v[tokk * p.ldv + h * p.dh + a]);
  }
  __syncwarp();
  int lo = 0, hi = p.Lq - 1;
  if (p.R > 0 && p.Lk > p.R) { lo = max(0, j - p.R + 1); hi = min(p.Lq - 1, j + p.R - 1); }
  hi = min(hi, q_rows(p, b) - 1);
  const int nq = max(hi - lo + 1, 0);
  for (int ii = lane; ii < nq; ii += 32) {
    const int i = lo + ii;
    const long tokq = qb + i;
    const long rid = ((long)b * p.H + h) * p.Lq + i;
    const T* qrow = q + tokq * p.ldq + h * p.dh;
    const float qk = dot_row(ks, qrow, p.dh);
    float qe = 0.f;
    const int rel = j - i;
    if (p.R > 0 && rel > -p.R && rel < p.R) {
      const T* er = E + ((long)h * (2 * p.R - 1) + rel + p.R - 1) * p.dh;
      for (int a = 0; a < p.dh; ++a) qe = fmaf(to_f32(qrow[a]), to_f32(er[a]), qe);
    }
    bool masked;
    const float s = make_logit(p, b, i, j, qk, qe, masked);
    const float pr = __expf((s - lse[rid]) - lse[(long)p.B * p.H * p.Lq + rid]);
    float dp = dot_row(vs, dO + tokq * p.ldo + h * p.dh, p.dh);
    float keep = 1.f;
    if (p.thr) keep = philox_keep16(salted(p.seed, p.salt), (unsigned long long)rid * p.Lkp + j, p.thr) ? p.dscale : 0.f;
    dp *= keep;
    const float ds = pr * (dp - delta[rid]);
    pw[ii] = pr * keep;
    dsw[ii] = masked ? 0.f : ds * p.scale;
  }
  __syncwarp();
  for (int a = lane; a < p.dh; a += 32) {
    float accv = 0.f, acck = 0.f;
    const T* dop = dO + (qb + lo) * p.ldo + h * p.dh + a;
    const T* qp = q + (qb + lo) * p.ldq + h * p.dh + a;
    for (int ii = 0; ii < nq; ++ii) {
      accv = fmaf(pw[ii], to_f32(dop[(long)ii * p.ldo]), accv);
      acck = fmaf(dsw[ii], to_f32(qp[(long)ii * p.ldq]), acck);
    }
    dv[tokk * p.ldv + h * p.dh + a] = from_f32<T>(accv);
    dk[tokk * p.ldk + h * p.dh + a] = from_f32<T>(acck);
  }
}

static AttnP make_params(const SstAttnDesc& d, const int* q_lens, const int* k_lens) {
  AttnP p;
  p.B = d.B; p.H = d.H; p.Lq = d.Lq; p.Lk = d.Lk; p.Lkp = (d.Lk + 7) & ~7; p.dh = d.dh;
  p.ldq = d.ldq; p.ldk = d.ldk; p.ldv = d.ldv; p.ldo = d.ldo;
  p.causal = d.causal; p.mask_q_rows = d.mask_q_rows; p.R = d.rel_dist;
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

int attn_fwd_simt_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                         const int* k_lens, void* o, float* lse, cudaStream_t st) {
  AttnP p = make_params(d, q_lens, k_lens);
  const int maxk = (d.rel_dist > 0 && d.Lk > d.rel_dist) ? 2 * d.rel_dist - 1 : d.Lk;
  const size_t smem = (size_t)4 * (d.dh + maxk) * sizeof(float);
  SST_REQUIRE(smem <= 200 * 1024, SST_E_ARG, "attention: Lk=%d too long for the CUDA-core kernel", d.Lk);
  const long rows = (long)d.B * d.H * d.Lq;
  const int grid = (int)((rows + 3) / 4);
  if (d.dtype == SST_F32) {
    cudaFuncSetAttribute(attn_fwd_simt<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_fwd_simt<float><<<grid, 128, smem, st>>>((const float*)q, (const float*)k, (const float*)v, (const float*)E, (float*)o, lse, p, maxk);
  } else {
    cudaFuncSetAttribute(attn_fwd_simt<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_fwd_simt<__nv_bfloat16><<<grid, 128, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                          (const __nv_bfloat16*)E, (__nv_bfloat16*)o, lse, p, maxk);
  }
  return check_launch("attn_fwd_simt");
}

int attn_bwd_simt_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                         const int* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                         float* delta, cudaStream_t st) {
  AttnP p = make_params(d, q_lens, k_lens);
  const bool band = d.rel_dist > 0 && d.Lk > d.rel_dist;
  const int maxk = band ? 2 * d.rel_dist - 1 : d.Lk;
  const int maxq = band ? 2 * d.rel_dist - 1 : d.Lq;
  const size_t smem1 = (size_t)4 * (2 * d.dh + 2 * maxk) * sizeof(float);
  const size_t smem2 = (size_t)4 * (2 * d.dh + 2 * maxq) * sizeof(float);
  SST_REQUIRE(smem1 <= 200 * 1024 && smem2 <= 200 * 1024, SST_E_ARG, "attention: sequence too long for the CUDA-core kernel");
  const long rows_q = (long)d.B * d.H * d.Lq, rows_k = (long)d.B * d.H * d.Lk;
  if (d.dtype == SST_F32) {
    cudaFuncSetAttribute(attn_bwd_dq_simt<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(attn_bwd_dkv_simt<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_bwd_dq_simt<float><<<(int)((rows_q + 3) / 4), 128, smem1, st>>>((const float*)q, (const float*)k, (const float*)v, (const float*)E,
        (const float*)o, (const float*)dO, lse, delta, (float*)dq, p, maxk);
    attn_bwd_dkv_simt<float><<<(int)((rows_k + 3) / 4), 128, smem2, st>>>((const float*)q, (const float*)k, (const float*)v, (const float*)E,
        (const float*)dO, lse, delta, (float*)dk, (float*)dv, p, maxq);
  } else {
    typedef __nv_bfloat16 bf;
    cudaFuncSetAttribute(attn_bwd_dq_simt<bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(attn_bwd_dkv_simt<bf>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attn_bwd_dq_simt<bf><<<(int)((rows_q + 3) / 4), 128, smem1, st>>>((const bf*)q, (const bf*)k, (const bf*)v, (const bf*)E,
        (const bf*)o, (const bf*)dO, lse, delta, (bf*)dq, p, maxk);
    attn_bwd_dkv_simt<bf><<<(int)((rows_k + 3) / 4), 128, smem2, st>>>((const bf*)q, (const bf*)k, (const bf*)v, (const bf*)E,
        (const bf*)dO, lse, delta, (bf*)dk, (bf*)dv, p, maxq);
  }
  return check_launch("attn_bwd_simt", 2);
}

int attn_fwd_tc_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                       const int* k_lens, void* o, float* lse, cudaStream_t st);

int attn_bwd_tc_launch(const SstAttnDesc& d, const void* q, const void* k, const void* v, const void* E, const int* q_lens,
                       const int* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                       float* delta, void* ws, size_t ws_bytes, cudaStream_t st);
size_t attn_bwd_tc_ws_bytes(const SstAttnDesc& d);

static bool use_tc(const SstAttnDesc& d) { return d.dtype == SST_BF16 && !d.force_simt && d.dh == 96; }

}  // namespace sst

using namespace sst;

extern "C" {

static int attn_check(const SstAttnDesc* d, const void* q, const void* k, const void* v, const void* E) {
  SST_REQUIRE(d && q && k && v, SST_E_ARG, "attention: null argument");
  SST_REQUIRE(d->dh % 8 == 0 && d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0, SST_E_ARG,
              "attention: dh and pitches must be multiples of 8");
  SST_REQUIRE(d->rel_dist == 0 || (E != nullptr && d->Lq == d->Lk && !d->causal), SST_E_ARG,
              "attention: relative-position bias needs E and unmasked self-attention (Lq == Lk)");
  return SST_OK;
}
// packed layouts need the lengths that say which rows exist
static int attn_check_packed(const SstAttnDesc* d, const int32_t* q_lens, const int32_t* k_lens) {
  SST_REQUIRE((d->q_off == nullptr || q_lens != nullptr) && (d->k_off == nullptr || k_lens != nullptr), SST_E_ARG,
              "attention: q_off / k_off need q_lens / k_lens");
  SST_REQUIRE((d->q_off == nullptr || d->q_pad == nullptr) && (d->k_off == nullptr || d->k_pad == nullptr), SST_E_ARG,
              "attention: per-position padding masks are indexed b*L + t and do not combine with a packed layout");
  return SST_OK;
}

int sst_attn_fwd(const SstAttnDesc* d, const void* q, const void* k, const void* v, const void* E, const int32_t* q_lens,
                 const int32_t* k_lens, void* o, float* lse, void* stream) {
  int rc = attn_check(d, q, k, v, E);
  if (rc) return rc;
  if ((rc = attn_check_packed(d, q_lens, k_lens))) return rc;
  if (d->B * d->H * d->Lq == 0) return SST_OK;
  if (use_tc(*d)) return attn_fwd_tc_launch(*d, q, k, v, E, q_lens, k_lens, o, lse, reinterpret_cast<cudaStream_t>(stream));
  return attn_fwd_simt_launch(*d, q, k, v, E, q_lens, k_lens, o, lse, reinterpret_cast<cudaStream_t>(stream));
}

size_t sst_attn_bwd_workspace_bytes(const SstAttnDesc* d) {
  if (!d || !use_tc(*d) || d->B * d->H * d->Lq == 0) return 0;
  return attn_bwd_tc_ws_bytes(*d);
}

int sst_attn_bwd(const SstAttnDesc* d, const void* q, const void* k, const void* v, const void* E, const int32_t* q_lens,
                 const int32_t* k_lens, const void* o, const float* lse, const void* dO, void* dq, void* dk, void* dv,
                 float* delta, void* ws, size_t ws_bytes, void* stream) {
  int rc = attn_check(d, q, k, v, E);
  if (rc) return rc;
  if ((rc = attn_check_packed(d, q_lens, k_lens))) return rc;
  if (d->B * d->H * d->Lq == 0) return SST_OK;
  if (use_tc(*d))
    return attn_bwd_tc_launch(*d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, ws, ws_bytes,
                              reinterpret_cast<cudaStream_t>(stream));
  return attn_bwd_simt_launch(*d, q, k, v, E, q_lens, k_lens, o, lse, dO, dq, dk, dv, delta, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"

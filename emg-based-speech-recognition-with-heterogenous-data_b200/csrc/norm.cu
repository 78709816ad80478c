// Bandwidth-bound normalisation kernels (128-bit vectorised HBM access, fp32 statistics):
//   LayerNorm(x + dropout(r)) forward / backward      -- transformer.py:59-63, :124-133 (post-LN, eps 1e-5)
// (the per-channel BatchNorm / column-sum kernels live in colnorm.cu)
#include "colstream.cuh"

namespace sst {

constexpr int LN_MAXV = 8;   // up to 8 vectors of 8 per lane -> D <= 2048
// cp.async ring (colstream.cuh).  D = 768 (EXACT): 4 stages = three rows per warp in flight, two 256-thread blocks per SM;
// the generic variant (up to 8 vectors per lane) runs 2 stages in 128-thread blocks so that its ring still fits.
template <bool EXACT> struct LnCfg { static constexpr int STAGES = EXACT ? 4 : 2; static constexpr int THREADS = EXACT ? 256 : 128; };

// this lane's slot of (stage 0, tensor 0, vector 0); slot (t, i) of a stage sits (t * NV + i) planes further
template <typename T>
__device__ __forceinline__ uint32_t ln_ring(void* smem, uint32_t& plane) {
  plane = blockDim.x * ColSlot<T>::BYTES;
  return (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * ColSlot<T>::BYTES;
}

// 8 dropout keep decisions for elements idx .. idx+7 (idx % 8 == 0): ONE Philox block, eight 16-bit lanes
// (sst_common.cuh philox_keep16 -- the same stream the host mirror in tests/helpers.py draws)
__device__ __forceinline__ void keep8_16(unsigned long long seed, unsigned long long idx, uint32_t thr16, bool (&k)[8]) {
  const Philox4 r = philox4x32(seed, idx >> 3);
  const uint32_t thr_hi = thr16 << 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) k[e] = philox_keep16_at(r, e, thr_hi);
}

// One warp per row, NV vectors of 8 columns per lane (column c = (i*32 + lane)*8, the same columns for every row the
// warp visits, so gamma / beta and the dgamma / dbeta partials live in registers).  EXACT: D == NV*256 (no predicates).
template <typename T, int NV, bool EXACT>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ r, T* __restrict__ y, T* __restrict__ s_out,
              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long rows, int D, float eps, uint32_t thr, float dscale, unsigned long long seed,
              const unsigned long long* __restrict__ salt) {
  extern __shared__ __align__(16) uint8_t ln_smem[];
  pdl_wait();
  pdl_trigger();
  if (thr != 0) seed = salted(seed, salt);
  const int lane = threadIdx.x & 31;
  const int warp = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  const int nwarps = (int)(gridDim.x * (blockDim.x >> 5));
  const int NT = r != nullptr ? 2 : 1;
  uint32_t plane;
  const uint32_t ring = ln_ring<T>(ln_smem, plane);
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (EXACT || c < D) { load8_f32(gamma + c, g[i]); load8_f32(beta + c, b[i]); }
  }
  const float invD = 1.f / D;
  stream_rows<1, LnCfg<EXACT>::STAGES>(warp, (int)rows, nwarps, ring, (uint32_t)(NT * NV) * plane,
      [&](uint32_t a0, int, int row, bool valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 8;
          if (EXACT || c < D) {
            ColSlot<T>::issue(a0 + i * plane, x + (long)row * D + c, valid);
            if (r != nullptr) ColSlot<T>::issue(a0 + (NV + i) * plane, r + (long)row * D + c, valid);
          }
        }
      },
      [&](uint32_t a0, int, int row) {
    float v[NV][8], rv[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
        ColSlot<T>::read(a0 + i * plane, v[i]);
        if (r != nullptr) ColSlot<T>::read(a0 + (NV + i) * plane, rv[i]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
        if (r != nullptr) {
          if (thr != 0) {
            bool k[8];
            keep8_16(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[i][j] = k[j] ? rv[i][j] * dscale : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += rv[i][j];
        }
        if (s_out != nullptr) {
          // the saved pre-norm sum is what backward normalises: round it to T first so both see one value
          Vec8<T>::store(s_out + (long)row * D + c, v[i]);
          if (sizeof(T) == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = __bfloat162float(__float2bfloat16_rn(v[i][j]));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
    }
    const float mean = warp_sum(sum) * invD;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float d = v[i][j] - mean; sq += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * invD + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[i][j] + b[i][j];
        Vec8<T>::store(y + (long)row * D + c, o);
      }
    }
      });
}

// ds = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat));  dr = ds * keep/(1-p);
// dgamma += sum_rows dy*xhat, dbeta += sum_rows dy: per-lane register partials; at the end every warp parks its 2*D partials in the
// (by then idle) ring, the block adds them up and issues one 16-byte reduction per four columns (no shared-memory atomics: eight
// warps adding into the same 1536 shared words was most of the kernel's ~9 us floor on decoder-sized inputs)
template <typename T, int NV, bool EXACT>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ s, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ gamma, T* __restrict__ ds, T* __restrict__ dr,
              float* __restrict__ dgamma, float* __restrict__ dbeta, long rows, int D, uint32_t thr, float dscale,
              unsigned long long seed, const unsigned long long* __restrict__ salt) {
  extern __shared__ __align__(16) uint8_t ln_smem[];   // [2][D] floats of column partials, then the ring
  pdl_wait();
  pdl_trigger();
  if (dr != nullptr) seed = salted(seed, salt);
  const int lane = threadIdx.x & 31;
  const int warp = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  const int nwarps = (int)(gridDim.x * (blockDim.x >> 5));
  uint32_t plane;
  const uint32_t ring = ln_ring<T>(ln_smem, plane);
  float ag[NV][8], ab[NV][8], gm[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (EXACT || c < D) load8_f32(gamma + c, gm[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[i][j] = 0.f; ab[i][j] = 0.f; }
  }
  const float invD = 1.f / D;
  stream_rows<1, LnCfg<EXACT>::STAGES>(warp, (int)rows, nwarps, ring, (uint32_t)(2 * NV) * plane,
      [&](uint32_t a0, int, int row, bool valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = (i * 32 + lane) * 8;
          if (EXACT || c < D) {
            ColSlot<T>::issue(a0 + i * plane, dy + (long)row * D + c, valid);
            ColSlot<T>::issue(a0 + (NV + i) * plane, s + (long)row * D + c, valid);
          }
        }
      },
      [&](uint32_t a0, int, int row) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[NV][8], g[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) { ColSlot<T>::read(a0 + i * plane, g[i]); ColSlot<T>::read(a0 + (NV + i) * plane, xh[i]); }
    }
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mean) * rstd;
          ag[i][j] = fmaf(g[i][j], xh[i][j], ag[i][j]);
          ab[i][j] += g[i][j];
          g[i][j] *= gm[i][j];
          c1 += g[i][j];
          c2 = fmaf(g[i][j], xh[i][j], c2);
        }
      }
    }
    c1 = warp_sum(c1) * invD;
    c2 = warp_sum(c2) * invD;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - c1 - xh[i][j] * c2);
        Vec8<T>::store(ds + (long)row * D + c, o);
        if (dr != nullptr) {
          bool k[8];
          keep8_16(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = k[j] ? o[j] * dscale : 0.f;
          Vec8<T>::store(dr + (long)row * D + c, o);
        }
      }
    }
      });
  cp_async_wait<0>();
  __syncthreads();                                   // every warp has left the ring: it now holds [warp][dgamma D | dbeta D] partials
  float* part = reinterpret_cast<float*>(ln_smem);
  const int nw = (int)(blockDim.x >> 5);
  float* mine = part + (size_t)(threadIdx.x >> 5) * 2 * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (EXACT || c < D) {
      *reinterpret_cast<float4*>(mine + c) = make_float4(ag[i][0], ag[i][1], ag[i][2], ag[i][3]);
      *reinterpret_cast<float4*>(mine + c + 4) = make_float4(ag[i][4], ag[i][5], ag[i][6], ag[i][7]);
      *reinterpret_cast<float4*>(mine + D + c) = make_float4(ab[i][0], ab[i][1], ab[i][2], ab[i][3]);
      *reinterpret_cast<float4*>(mine + D + c + 4) = make_float4(ab[i][4], ab[i][5], ab[i][6], ab[i][7]);
    }
  }
  __syncthreads();
  const bool vec = ((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0;
  for (int i = threadIdx.x * 4; i < 2 * D; i += blockDim.x * 4) {      // D % 8 == 0: a group of four never straddles dgamma | dbeta
    float4 s = *reinterpret_cast<const float4*>(part + i);
    for (int w = 1; w < nw; ++w) {
      const float4 t = *reinterpret_cast<const float4*>(part + (size_t)w * 2 * D + i);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    float* dst = i < D ? dgamma + i : dbeta + (i - D);
    if (vec) red_add4(dst, s.x, s.y, s.z, s.w);                         // four columns per reduction (sst_common.cuh)
    else { atomicAdd(dst, s.x); atomicAdd(dst + 1, s.y); atomicAdd(dst + 2, s.z); atomicAdd(dst + 3, s.w); }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Column-wise kernels over channels-last matrices.  Common shape: thread (tx, ty) owns the 8-column vector c = 8*tx and
// walks rows r0 + ty, r0 + ty + TY, ... of its block's row range, four rows (= four independent 128-bit loads per
// operand) per iteration; partial sums are combined across ty in shared memory so that a block issues ONE atomic per
// column.  No per-element integer division: the only division is row -> (chunk, t), once per row, in 32 bits.
// ---------------------------------------------------------------------------------------------------------
static int ew_grid(long total, int threads) {
  long blocks = (total + threads - 1) / threads;
  long cap = (long)num_sms() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace sst

using namespace sst;

extern "C" {

int sst_layernorm_fwd(int dtype, int64_t rows, int D, const void* x, const void* r, float drop_p, uint64_t seed,
                      const float* gamma, const float* beta, void* y, void* s_out, float* mean, float* rstd, float eps,
                      void* stream) {
  SST_REQUIRE(D % 8 == 0 && D <= LN_MAXV * 256, SST_E_ARG, "layernorm: D=%d must be a multiple of 8 and <= %d", D, LN_MAXV * 256);
  if (rows <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thr = (r != nullptr && drop_p > 0.f) ? drop_threshold16(drop_p) : 0u;
  const float dscale = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  SST_REQUIRE(rows < (1L << 31), SST_E_ARG, "layernorm: too many rows");
  long blocks = (rows + 7) / 8;
  const long cap = (long)num_sms() * 2;
  const int grid = (int)(blocks < cap ? blocks : cap);
  // ring: LN_STAGES x (x [+ r]) x NV vectors x 256 threads
#define SST_LN_FWD(T_, NV_, EX_)                                                                                             \
  do {                                                                                                                       \
    constexpr int TH = LnCfg<EX_>::THREADS;                                                                                  \
    const size_t smem = (size_t)LnCfg<EX_>::STAGES * (r != nullptr ? 2 : 1) * NV_ * TH * ColSlot<T_>::BYTES;                   \
    static bool attr = false;                                                                                                \
    if (!attr) {                                                                                                             \
      cudaError_t e = cudaFuncSetAttribute(ln_fwd_kernel<T_, NV_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));                          \
      attr = true;                                                                                                           \
    }                                                                                                                        \
    SST_REQUIRE(smem <= 227 * 1024, SST_E_ARG, "layernorm: D=%d needs %zu bytes of shared memory", D, smem);                  \
    launch_pdl(ln_fwd_kernel<T_, NV_, EX_>, dim3(grid), dim3(TH), smem, st, (const T_*)x, (const T_*)r, (T_*)y, (T_*)s_out, gamma, beta, \
               mean, rstd, rows, D, eps, thr, dscale, (unsigned long long)seed, dropout_salt());                               \
  } while (0)
  if (dtype == SST_F32) { if (D == 768) SST_LN_FWD(float, 3, true); else SST_LN_FWD(float, LN_MAXV, false); }
  else { if (D == 768) SST_LN_FWD(__nv_bfloat16, 3, true); else SST_LN_FWD(__nv_bfloat16, LN_MAXV, false); }
#undef SST_LN_FWD
  return check_launch("layernorm_fwd");
}

int sst_layernorm_bwd(int dtype, int64_t rows, int D, const void* dy, const void* s, const float* mean, const float* rstd,
                      const float* gamma, void* ds, void* dr, float drop_p, uint64_t seed, float* dgamma, float* dbeta,
                      void* stream) {
  SST_REQUIRE(D % 8 == 0 && D <= LN_MAXV * 256, SST_E_ARG, "layernorm: D=%d must be a multiple of 8 and <= %d", D, LN_MAXV * 256);
  if (rows <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thr = (dr != nullptr && drop_p > 0.f) ? drop_threshold16(drop_p) : 0u;
  if (thr == 0) dr = nullptr;
  const float dscale = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  // two resident blocks per SM and no more blocks than that: every block ends with 2*D global atomics (dgamma, dbeta)
  long blocks = (rows + 7) / 8;                  // one row per warp at least: decoder-sized inputs spread over every SM
  long cap = (long)num_sms() * 2;
  const int grid = (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  SST_REQUIRE(rows < (1L << 31), SST_E_ARG, "layernorm: too many rows");
#define SST_LN_BWD(T_, NV_, EX_)                                                                                             \
  do {                                                                                                                       \
    constexpr int TH = LnCfg<EX_>::THREADS;                                                                                  \
    size_t smem = (size_t)LnCfg<EX_>::STAGES * 2 * NV_ * TH * ColSlot<T_>::BYTES;                                           \
    if (smem < (size_t)(TH / 32) * 2 * D * sizeof(float)) smem = (size_t)(TH / 32) * 2 * D * sizeof(float);   /* tail partials */ \
    static bool attr = false;                                                                                                \
    if (!attr) {                                                                                                             \
      cudaError_t e = cudaFuncSetAttribute(ln_bwd_kernel<T_, NV_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));                          \
      attr = true;                                                                                                           \
    }                                                                                                                        \
    SST_REQUIRE(smem <= 227 * 1024, SST_E_ARG, "layernorm: D=%d needs %zu bytes of shared memory", D, smem);                  \
    launch_pdl(ln_bwd_kernel<T_, NV_, EX_>, dim3(grid), dim3(TH), smem, st, (const T_*)dy, (const T_*)s, mean, rstd, gamma, (T_*)ds,  \
               (T_*)dr, dgamma, dbeta, rows, D, thr, dscale, (unsigned long long)seed, dropout_salt());                        \
  } while (0)
  if (dtype == SST_F32) { if (D == 768) SST_LN_BWD(float, 3, true); else SST_LN_BWD(float, LN_MAXV, false); }
  else { if (D == 768) SST_LN_BWD(__nv_bfloat16, 3, true); else SST_LN_BWD(__nv_bfloat16, LN_MAXV, false); }
#undef SST_LN_BWD
  return check_launch("layernorm_bwd");
}

}  // extern "C"

// Bandwidth-bound normalisation kernels (128-bit vectorised HBM access, fp32 statistics):
//   LayerNorm(x + dropout(r)) forward / backward      -- transformer.py:59-63, :124-133 (post-LN, eps 1e-5)
//   per-channel batch statistics, BatchNorm apply (+ReLU, + second normalised branch) and backward
//                                                      -- architecture.py:27,29,33,40-48 (training-mode BN)
#include "vec.cuh"

namespace sst {

constexpr int LN_MAXV = 8;   // up to 8 vectors of 8 per lane -> D <= 2048

template <typename T>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ r, T* __restrict__ y, T* __restrict__ s_out,
              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long rows, int D, float eps, uint32_t thr, float dscale, unsigned long long seed) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (long row = warp; row < rows; row += nwarps) {
    float v[LN_MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < D) {
        Vec8<T>::load(x + row * D + c, v[i]);
        if (r != nullptr) {
          float rv[8];
          Vec8<T>::load(r + row * D + c, rv);
          if (thr != 0) {
            bool k[8];
            keep8(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = k[j] ? rv[j] * dscale : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += rv[j];
        }
        if (s_out != nullptr) {
          // the saved pre-norm sum is what backward normalises: round it to T first so both see one value
          Vec8<T>::store(s_out + row * D + c, v[i]);
          if (sizeof(T) == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = __bfloat162float(__float2bfloat16_rn(v[i][j]));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[i][j];
      }
    }
    const float mean = warp_sum(sum) / D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < D) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float d = v[i][j] - mean; sq += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / D + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < D) {
        float g[8], b[8], o[8];
        load8_f32(gamma + c, g);
        load8_f32(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        Vec8<T>::store(y + row * D + c, o);
      }
    }
  }
}

// ds = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat));  dr = ds * keep/(1-p);
// dgamma += sum_rows dy*xhat, dbeta += sum_rows dy  (per-warp register partials -> smem -> one atomic per block/column)
template <typename T>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ s, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ gamma, T* __restrict__ ds, T* __restrict__ dr,
              float* __restrict__ dgamma, float* __restrict__ dbeta, long rows, int D, uint32_t thr, float dscale,
              unsigned long long seed) {
  extern __shared__ float red[];   // [2][D]
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float ag[LN_MAXV][8], ab[LN_MAXV][8];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[i][j] = 0.f; ab[i][j] = 0.f; }

  for (long row = warp; row < rows; row += nwarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[LN_MAXV][8], g[LN_MAXV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < D) {
        float dv[8], sv[8], gm[8];
        Vec8<T>::load(dy + row * D + c, dv);
        Vec8<T>::load(s + row * D + c, sv);
        load8_f32(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (sv[j] - mean) * rstd;
          ag[i][j] += dv[j] * xh[i][j];
          ab[i][j] += dv[j];
          g[i][j] = dv[j] * gm[j];
          c1 += g[i][j];
          c2 += g[i][j] * xh[i][j];
        }
      }
    }
    c1 = warp_sum(c1) / D;
    c2 = warp_sum(c2) / D;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < D) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - c1 - xh[i][j] * c2);
        Vec8<T>::store(ds + row * D + c, o);
        if (dr != nullptr) {
          bool k[8];
          keep8(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = k[j] ? o[j] * dscale : 0.f;
          Vec8<T>::store(dr + row * D + c, o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < D) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&red[c + j], ag[i][j]); atomicAdd(&red[D + c + j], ab[i][j]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// column statistics over a (rows, C) matrix with row pitch ld: stats[0][c] += sum x, stats[1][c] += sum x^2 (double)
// thread = one 8-column vector, blockDim.y rows in flight; fp32 partials over <= 64 rows, double across.
template <typename T>
__global__ void __launch_bounds__(512)
colstats_kernel(const T* __restrict__ x, long rows, int C, long ld, double* __restrict__ stats, long rows_per_block) {
  const int c = threadIdx.x * 8;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  double s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.0; q[j] = 0.0; }
  for (long rb = r0 + threadIdx.y; rb < r1; rb += (long)blockDim.y * 64) {
    float fs[8], fq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { fs[j] = 0.f; fq[j] = 0.f; }
    for (int it = 0; it < 64; ++it) {
      long r = rb + (long)it * blockDim.y;
      if (r >= r1) break;
      float v[8];
      Vec8<T>::load(x + r * ld + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { fs[j] += v[j]; fq[j] = fmaf(v[j], v[j], fq[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += fs[j]; q[j] += fq[j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { atomicAdd(stats + c + j, s[j]); atomicAdd(stats + C + c + j, q[j]); }
}

// out[c] += sum_rows x[r][c]  (fp32 partials per thread, one fp32 atomic per thread/column)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long rows, int C, long ld, float* __restrict__ out, long rows_per_block) {
  const int c = threadIdx.x * 8;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  double s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.0;
  if (c < C) {
    for (long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      float v[8];
      Vec8<T>::load(x + r * ld + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += (double)v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c + j < C) atomicAdd(out + c + j, (float)s[j]);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, long count, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int training) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (training) {
    double m = stats[c] / (double)count;
    double var = stats[C + c] / (double)count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
      double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean[c] = running_mean[c];
    invstd[c] = 1.f / sqrtf(running_var[c] + eps);
  }
}

struct BnBranch {
  const void* x;          // conv output, row pitch ld
  long ld;
  const float* mean; const float* invstd; const float* gamma; const float* beta;
};

// out[(chunk, lead + t), :] = act( bnA(xa) [+ bnB(xb)] ), halo rows written as zero.  out rows pitch C.
template <typename T>
__global__ void bn_apply_kernel(BnBranch a, BnBranch b, int has_b, int relu, T* __restrict__ out, long n_chunks, int Tlen,
                                int C, int lead, int trail) {
  const int P = Tlen + lead + trail;
  const int cv = C / 8;
  const long total = n_chunks * P * (long)cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    const long prow = i / cv;
    const long chunk = prow / P;
    const int t = (int)(prow - chunk * P) - lead;
    float o[8];
    if (t < 0 || t >= Tlen) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    } else {
      const long row = chunk * Tlen + t;
      float xv[8], m[8], is[8], g[8], be[8];
      Vec8<T>::load(reinterpret_cast<const T*>(a.x) + row * a.ld + c, xv);
      load8_f32(a.mean + c, m); load8_f32(a.invstd + c, is); load8_f32(a.gamma + c, g); load8_f32(a.beta + c, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (xv[j] - m[j]) * is[j] * g[j] + be[j];
      if (has_b) {
        Vec8<T>::load(reinterpret_cast<const T*>(b.x) + row * b.ld + c, xv);
        load8_f32(b.mean + c, m); load8_f32(b.invstd + c, is); load8_f32(b.gamma + c, g); load8_f32(b.beta + c, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += (xv[j] - m[j]) * is[j] * g[j] + be[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
      }
    }
    Vec8<T>::store(out + prow * C + c, o);
  }
}

// backward pass 1: g = dout * (y > 0);  red[0][c] += g, red[1][c] += g*xhat_a, red[2][c] += g*xhat_b   (double)
template <typename T>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(const T* __restrict__ dout, long ld_dout, const T* __restrict__ y, int y_lead, int y_trail,
                                     int relu, BnBranch a, BnBranch b, int has_b, long n_chunks, int Tlen, int C,
                                     double* __restrict__ red, long rows_per_block) {
  const int c = threadIdx.x * 8;
  const long rows = n_chunks * Tlen;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  const int Py = Tlen + y_lead + y_trail;
  float ma[8], ia[8], mb[8], ib[8];
  load8_f32(a.mean + c, ma); load8_f32(a.invstd + c, ia);
  if (has_b) { load8_f32(b.mean + c, mb); load8_f32(b.invstd + c, ib); }
  double s0[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = s1[j] = s2[j] = 0.0; }
  for (long rb = r0 + threadIdx.y; rb < r1; rb += (long)blockDim.y * 32) {
    float f0[8], f1[8], f2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { f0[j] = f1[j] = f2[j] = 0.f; }
    for (int it = 0; it < 32; ++it) {
      long r = rb + (long)it * blockDim.y;
      if (r >= r1) break;
      float g[8], xv[8];
      Vec8<T>::load(dout + r * ld_dout + c, g);
      if (relu) {
        const long chunk = r / Tlen;
        const long prow = chunk * Py + (r - chunk * Tlen) + y_lead;
        float yv[8];
        Vec8<T>::load(y + prow * C + c, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
      }
      Vec8<T>::load(reinterpret_cast<const T*>(a.x) + r * a.ld + c, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) { f0[j] += g[j]; f1[j] = fmaf(g[j], (xv[j] - ma[j]) * ia[j], f1[j]); }
      if (has_b) {
        Vec8<T>::load(reinterpret_cast<const T*>(b.x) + r * b.ld + c, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) f2[j] = fmaf(g[j], (xv[j] - mb[j]) * ib[j], f2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] += f0[j]; s1[j] += f1[j]; s2[j] += f2[j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(red + c + j, s0[j]);
    atomicAdd(red + C + c + j, s1[j]);
    if (has_b) atomicAdd(red + 2 * C + c + j, s2[j]);
  }
}

struct BnGradOut {
  void* dx;     // gradient w.r.t. the conv output, written in a time-padded layout (zero halos)
  long ld;      // row pitch of dx
  int lead, trail;
  float* dgamma; float* dbeta;   // accumulated (+=)
};

// backward pass 2: dx = gamma*invstd*(g - sum_g/N - xhat*sum_gxhat/N) for each branch, padded layouts; block 0 adds dgamma/dbeta.
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dout, long ld_dout, const T* __restrict__ y, int y_lead, int y_trail,
                                    int relu, BnBranch a, BnBranch b, int has_b, BnGradOut ga, BnGradOut gb, long n_chunks,
                                    int Tlen, int C, const double* __restrict__ red) {
  const int cv = C / 8;
  const long rows = n_chunks * Tlen;
  const double invN = 1.0 / (double)rows;
  const int Py = Tlen + y_lead + y_trail;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      ga.dgamma[c] += (float)red[C + c];
      ga.dbeta[c] += (float)red[c];
      if (has_b) { gb.dgamma[c] += (float)red[2 * C + c]; gb.dbeta[c] += (float)red[c]; }
    }
  }
  // pass over the larger of the two padded extents; each branch guards its own range
  const int Pa = Tlen + ga.lead + ga.trail;
  const int Pb = has_b ? Tlen + gb.lead + gb.trail : 0;
  const int Pm = Pa > Pb ? Pa : Pb;
  const long total = n_chunks * Pm * (long)cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    const long prow = i / cv;
    const long chunk = prow / Pm;
    const int pp = (int)(prow - chunk * Pm);
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      if (br == 1 && !has_b) break;
      const BnBranch& bx = br == 0 ? a : b;
      const BnGradOut& go = br == 0 ? ga : gb;
      const int Pbr = br == 0 ? Pa : Pb;
      if (pp >= Pbr) continue;
      const int t = pp - go.lead;
      float o[8];
      if (t < 0 || t >= Tlen) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      } else {
        const long r = chunk * Tlen + t;
        float g[8], xv[8], m[8], is[8], gm[8];
        Vec8<T>::load(dout + r * ld_dout + c, g);
        if (relu) {
          float yv[8];
          Vec8<T>::load(y + (chunk * Py + t + y_lead) * C + c, yv);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
        }
        Vec8<T>::load(reinterpret_cast<const T*>(bx.x) + r * bx.ld + c, xv);
        load8_f32(bx.mean + c, m); load8_f32(bx.invstd + c, is); load8_f32(bx.gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - m[j]) * is[j];
          const float sg = (float)(red[c + j] * invN);
          const float sgx = (float)(red[(br + 1) * C + c + j] * invN);
          o[j] = gm[j] * is[j] * (g[j] - sg - xh * sgx);
        }
      }
      Vec8<T>::store(reinterpret_cast<T*>(go.dx) + (chunk * Pbr + pp) * go.ld + c, o);
    }
  }
}

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
  const uint32_t thr = (r != nullptr && drop_p > 0.f) ? drop_threshold(drop_p) : 0u;
  const float dscale = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  const int grid = ew_grid(rows * 32, 256);
  if (dtype == SST_F32)
    ln_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)r, (float*)y, (float*)s_out, gamma, beta, mean, rstd,
                                               rows, D, eps, thr, dscale, seed);
  else
    ln_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)r, (__nv_bfloat16*)y,
                                                       (__nv_bfloat16*)s_out, gamma, beta, mean, rstd, rows, D, eps, thr, dscale, seed);
  return check_launch("layernorm_fwd");
}

int sst_layernorm_bwd(int dtype, int64_t rows, int D, const void* dy, const void* s, const float* mean, const float* rstd,
                      const float* gamma, void* ds, void* dr, float drop_p, uint64_t seed, float* dgamma, float* dbeta,
                      void* stream) {
  SST_REQUIRE(D % 8 == 0 && D <= LN_MAXV * 256, SST_E_ARG, "layernorm: D=%d must be a multiple of 8 and <= %d", D, LN_MAXV * 256);
  if (rows <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thr = (dr != nullptr && drop_p > 0.f) ? drop_threshold(drop_p) : 0u;
  if (thr == 0) dr = nullptr;
  const float dscale = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  long blocks = (rows + 7) / 8;
  long cap = (long)num_sms() * 4;
  const int grid = (int)(blocks < cap ? blocks : cap);
  const size_t smem = (size_t)2 * D * sizeof(float);
  if (dtype == SST_F32)
    ln_bwd_kernel<float><<<grid, 256, smem, st>>>((const float*)dy, (const float*)s, mean, rstd, gamma, (float*)ds, (float*)dr, dgamma,
                                                  dbeta, rows, D, thr, dscale, seed);
  else
    ln_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)s, mean, rstd, gamma,
                                                          (__nv_bfloat16*)ds, (__nv_bfloat16*)dr, dgamma, dbeta, rows, D, thr,
                                                          dscale, seed);
  return check_launch("layernorm_bwd");
}

int sst_colstats(int dtype, const void* x, int64_t rows, int C, int64_t ld, double* stats, void* stream) {
  SST_REQUIRE(C % 8 == 0 && C / 8 <= 512 && ld % 8 == 0, SST_E_ARG, "colstats: C=%d, ld=%ld must be multiples of 8", C, (long)ld);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st);
  SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
  if (rows <= 0) return SST_OK;
  const int tx = C / 8;
  int ty = 256 / tx; if (ty < 1) ty = 1; if (ty > 8) ty = 8;
  long nblk = (long)num_sms() * 4;
  long rpb = (rows + nblk - 1) / nblk;
  if (rpb < ty) rpb = ty;
  nblk = (rows + rpb - 1) / rpb;
  dim3 block(tx, ty);
  if (dtype == SST_F32) colstats_kernel<float><<<(int)nblk, block, 0, st>>>((const float*)x, rows, C, ld, stats, rpb);
  else colstats_kernel<__nv_bfloat16><<<(int)nblk, block, 0, st>>>((const __nv_bfloat16*)x, rows, C, ld, stats, rpb);
  return check_launch("colstats");
}

int sst_colsum_accum(int dtype, const void* x, int64_t rows, int C, int64_t ld, float* out, void* stream) {
  SST_REQUIRE(ld % 8 == 0 && C <= ld && (C + 7) / 8 <= 512, SST_E_ARG, "colsum: pitch %ld must be a multiple of 8 and >= C=%d", (long)ld, C);
  if (rows <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tx = (C + 7) / 8;
  int ty = 512 / tx; if (ty < 1) ty = 1; if (ty > 16) ty = 16;
  long nblk = (long)num_sms() * 2;
  long rpb = (rows + nblk - 1) / nblk;
  if (rpb < ty) rpb = ty;
  nblk = (rows + rpb - 1) / rpb;
  dim3 block(tx, ty);
  if (dtype == SST_F32) colsum_kernel<float><<<(int)nblk, block, 0, st>>>((const float*)x, rows, C, ld, out, rpb);
  else colsum_kernel<__nv_bfloat16><<<(int)nblk, block, 0, st>>>((const __nv_bfloat16*)x, rows, C, ld, out, rpb);
  return check_launch("colsum_accum");
}

int sst_bn_finalize(const double* stats, int64_t count, int C, float eps, float momentum, float* mean, float* invstd,
                    float* running_mean, float* running_var, int training, void* stream) {
  SST_REQUIRE(training || (running_mean && running_var), SST_E_ARG, "bn_finalize: eval mode needs running buffers");
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats, count, C, eps, momentum, mean, invstd,
                                                                                  running_mean, running_var, training);
  return check_launch("bn_finalize");
}

int sst_bn_apply(int dtype, int64_t n_chunks, int T, int C, const void* xa, int64_t lda, const float* mean_a, const float* invstd_a,
                 const float* gamma_a, const float* beta_a, const void* xb, int64_t ldb, const float* mean_b, const float* invstd_b,
                 const float* gamma_b, const float* beta_b, int relu, void* out, int lead, int trail, void* stream) {
  SST_REQUIRE(C % 8 == 0 && lda % 8 == 0 && (xb == nullptr || ldb % 8 == 0), SST_E_ARG, "bn_apply: C and pitches must be multiples of 8");
  BnBranch a{xa, lda, mean_a, invstd_a, gamma_a, beta_a};
  BnBranch b{xb, ldb, mean_b, invstd_b, gamma_b, beta_b};
  const long total = n_chunks * (long)(T + lead + trail) * (C / 8);
  if (total <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid(total, 256);
  if (dtype == SST_F32) bn_apply_kernel<float><<<grid, 256, 0, st>>>(a, b, xb != nullptr, relu, (float*)out, n_chunks, T, C, lead, trail);
  else bn_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(a, b, xb != nullptr, relu, (__nv_bfloat16*)out, n_chunks, T, C, lead, trail);
  return check_launch("bn_apply");
}

/* Backward of out = act(bnA(xa) [+ bnB(xb)]).  `red` is a caller-provided double[3*C] scratch. */
int sst_bn_bwd(int dtype, int64_t n_chunks, int T, int C, const void* dout, int64_t ld_dout, const void* y, int y_lead, int y_trail,
               int relu, const void* xa, int64_t lda, const float* mean_a, const float* invstd_a, const float* gamma_a,
               void* dxa, int64_t ld_dxa, int lead_a, int trail_a, float* dgamma_a, float* dbeta_a,
               const void* xb, int64_t ldb, const float* mean_b, const float* invstd_b, const float* gamma_b,
               void* dxb, int64_t ld_dxb, int lead_b, int trail_b, float* dgamma_b, float* dbeta_b, double* red, void* stream) {
  SST_REQUIRE(C % 8 == 0 && C / 8 <= 512, SST_E_ARG, "bn_bwd: C=%d must be a multiple of 8", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int has_b = xb != nullptr;
  BnBranch a{xa, lda, mean_a, invstd_a, gamma_a, nullptr};
  BnBranch b{xb, ldb, mean_b, invstd_b, gamma_b, nullptr};
  BnGradOut ga{dxa, ld_dxa, lead_a, trail_a, dgamma_a, dbeta_a};
  BnGradOut gb{dxb, ld_dxb, lead_b, trail_b, dgamma_b, dbeta_b};
  cudaError_t e = cudaMemsetAsync(red, 0, sizeof(double) * 3 * C, st);
  SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
  const long rows = n_chunks * T;
  if (rows <= 0) return SST_OK;
  const int tx = C / 8;
  int ty = 256 / tx; if (ty < 1) ty = 1; if (ty > 8) ty = 8;
  long nblk = (long)num_sms() * 4;
  long rpb = (rows + nblk - 1) / nblk;
  if (rpb < ty) rpb = ty;
  nblk = (rows + rpb - 1) / rpb;
  dim3 block(tx, ty);
  const int Pa = T + lead_a + trail_a, Pb = has_b ? T + lead_b + trail_b : 0;
  const long total = n_chunks * (long)(Pa > Pb ? Pa : Pb) * (C / 8);
  const int grid2 = ew_grid(total, 256);
  if (dtype == SST_F32) {
    bn_bwd_reduce_kernel<float><<<(int)nblk, block, 0, st>>>((const float*)dout, ld_dout, (const float*)y, y_lead, y_trail, relu, a, b,
                                                             has_b, n_chunks, T, C, red, rpb);
    bn_bwd_apply_kernel<float><<<grid2, 256, 0, st>>>((const float*)dout, ld_dout, (const float*)y, y_lead, y_trail, relu, a, b, has_b,
                                                      ga, gb, n_chunks, T, C, red);
  } else {
    bn_bwd_reduce_kernel<__nv_bfloat16><<<(int)nblk, block, 0, st>>>((const __nv_bfloat16*)dout, ld_dout, (const __nv_bfloat16*)y, y_lead,
                                                                     y_trail, relu, a, b, has_b, n_chunks, T, C, red, rpb);
    bn_bwd_apply_kernel<__nv_bfloat16><<<grid2, 256, 0, st>>>((const __nv_bfloat16*)dout, ld_dout, (const __nv_bfloat16*)y, y_lead,
                                                              y_trail, relu, a, b, has_b, ga, gb, n_chunks, T, C, red);
  }
  return check_launch("bn_bwd", 2);
}

}  // extern "C"

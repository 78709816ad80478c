// Bandwidth-bound normalisation kernels (128-bit vectorised HBM access, fp32 statistics):
//   LayerNorm(x + dropout(r)) forward / backward      -- transformer.py:59-63, :124-133 (post-LN, eps 1e-5)
//   per-channel batch statistics, BatchNorm apply (+ReLU, + second normalised branch) and backward
//                                                      -- architecture.py:27,29,33,40-48 (training-mode BN)
#include "vec.cuh"

namespace sst {

constexpr int LN_MAXV = 8;   // up to 8 vectors of 8 per lane -> D <= 2048

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
              float* __restrict__ rstd_out, long rows, int D, float eps, uint32_t thr, float dscale, unsigned long long seed) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (EXACT || c < D) { load8_f32(gamma + c, g[i]); load8_f32(beta + c, b[i]); }
  }
  const float invD = 1.f / D;
  for (long row = warp; row < rows; row += nwarps) {
    float v[NV][8], rv[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) {
        Vec8<T>::load(x + row * D + c, v[i]);
        if (r != nullptr) Vec8<T>::load(r + row * D + c, rv[i]);
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
        Vec8<T>::store(y + row * D + c, o);
      }
    }
  }
}

// ds = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat));  dr = ds * keep/(1-p);
// dgamma += sum_rows dy*xhat, dbeta += sum_rows dy  (per-lane register partials -> smem -> one atomic per block/column)
template <typename T, int NV, bool EXACT>
__global__ void __launch_bounds__(256, 2)
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
  float ag[NV][8], ab[NV][8], gm[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (EXACT || c < D) load8_f32(gamma + c, gm[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[i][j] = 0.f; ab[i][j] = 0.f; }
  }
  const float invD = 1.f / D;
  for (long row = warp; row < rows; row += nwarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[NV][8], g[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (EXACT || c < D) { Vec8<T>::load(dy + row * D + c, g[i]); Vec8<T>::load(s + row * D + c, xh[i]); }
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
        Vec8<T>::store(ds + row * D + c, o);
        if (dr != nullptr) {
          bool k[8];
          keep8_16(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = k[j] ? o[j] * dscale : 0.f;
          Vec8<T>::store(dr + row * D + c, o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (EXACT || c < D) {
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
// Column-wise kernels over channels-last matrices.  Common shape: thread (tx, ty) owns the 8-column vector c = 8*tx and
// walks rows r0 + ty, r0 + ty + TY, ... of its block's row range, four rows (= four independent 128-bit loads per
// operand) per iteration; partial sums are combined across ty in shared memory so that a block issues ONE atomic per
// column.  No per-element integer division: the only division is row -> (chunk, t), once per row, in 32 bits.
// ---------------------------------------------------------------------------------------------------------
constexpr int COL_UNR = 4;

// v[j] (ty == 0) += sum over ty > 0 of v[j];  buf: (blockDim.y - 1) * 8 * blockDim.x elements
template <typename A>
__device__ __forceinline__ void reduce_over_ty(A (&v)[8], A* buf) {
  const int W = blockDim.x * 8, c = threadIdx.x * 8;
  if (threadIdx.y > 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) buf[(threadIdx.y - 1) * W + c + j] = v[j];
  }
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int y = 0; y < (int)blockDim.y - 1; ++y)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += buf[y * W + c + j];
  }
  __syncthreads();
}

// stats[0][c] += sum x, stats[1][c] += sum x^2 (double; fp32 partials over <= 64 rows)
template <typename T>
__global__ void __launch_bounds__(512)
colstats_kernel(const T* __restrict__ x, long rows, int C, long ld, double* __restrict__ stats, long rows_per_block) {
  extern __shared__ double dbuf[];
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  double s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.0; q[j] = 0.0; }
  for (long rb = r0 + threadIdx.y; rb < r1; rb += (long)TY * COL_UNR * 16) {
    float fs[8], fq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { fs[j] = 0.f; fq[j] = 0.f; }
    for (int it = 0; it < 16; ++it) {
      const long ra = rb + (long)it * TY * COL_UNR;
      if (ra >= r1) break;
      float v[COL_UNR][8];
#pragma unroll
      for (int u = 0; u < COL_UNR; ++u) {
        const long r = ra + (long)u * TY;
        if (r < r1) Vec8<T>::load(x + r * ld + c, v[u]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < COL_UNR; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) { fs[j] += v[u][j]; fq[j] = fmaf(v[u][j], v[u][j], fq[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += fs[j]; q[j] += fq[j]; }
  }
  reduce_over_ty(s, dbuf);
  reduce_over_ty(q, dbuf);
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(stats + c + j, s[j]); atomicAdd(stats + C + c + j, q[j]); }
  }
}

// out[c] += sum_rows x[r][c]  (bias gradients)
template <typename T>
__global__ void __launch_bounds__(512)
colsum_kernel(const T* __restrict__ x, long rows, int C, long ld, float* __restrict__ out, long rows_per_block) {
  extern __shared__ double dbuf[];
  float* fbuf = reinterpret_cast<float*>(dbuf);
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (c < C) {
    for (long ra = r0 + threadIdx.y; ra < r1; ra += (long)TY * COL_UNR) {
      float v[COL_UNR][8];
#pragma unroll
      for (int u = 0; u < COL_UNR; ++u) {
        const long r = ra + (long)u * TY;
        if (r < r1) Vec8<T>::load(x + r * ld + c, v[u]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < COL_UNR; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[u][j];
    }
  }
  reduce_over_ty(s, fbuf);
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c + j < C) atomicAdd(out + c + j, s[j]);
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
// One row of 8-column vectors per (tx, ty); per-column affine constants (scale, shift) live in registers.
template <typename T>
__global__ void __launch_bounds__(512)
bn_apply_kernel(BnBranch a, BnBranch b, int has_b, int relu, T* __restrict__ out, long n_chunks, int Tlen, int C, int lead,
                int trail, long prows_per_block) {
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const unsigned P = Tlen + lead + trail;
  const long prows = n_chunks * (long)P;
  const long p0 = (long)blockIdx.x * prows_per_block;
  const long p1 = min(prows, p0 + prows_per_block);
  float ma[8], sa[8], ha[8], mb[8], sb[8], hb[8];        // out = (x - mean) * (invstd * gamma) + beta, per branch
  {
    float is[8], g[8];
    load8_f32(a.mean + c, ma); load8_f32(a.invstd + c, is); load8_f32(a.gamma + c, g); load8_f32(a.beta + c, ha);
#pragma unroll
    for (int j = 0; j < 8; ++j) { sa[j] = is[j] * g[j]; mb[j] = 0.f; sb[j] = 0.f; hb[j] = 0.f; }
    if (has_b) {
      load8_f32(b.mean + c, mb); load8_f32(b.invstd + c, is); load8_f32(b.gamma + c, g); load8_f32(b.beta + c, hb);
#pragma unroll
      for (int j = 0; j < 8; ++j) sb[j] = is[j] * g[j];
    }
  }
  for (long pa = p0 + threadIdx.y; pa < p1; pa += (long)TY * COL_UNR) {
    float xa[COL_UNR][8], xb[COL_UNR][8];
    int tt[COL_UNR];
#pragma unroll
    for (int u = 0; u < COL_UNR; ++u) {
      const long prow = pa + (long)u * TY;
      tt[u] = -1;
      if (prow < p1) {
        const unsigned chunk = (unsigned)prow / P;
        const int t = (int)((unsigned)prow - chunk * P) - lead;
        if (t >= 0 && t < Tlen) {
          tt[u] = t;
          const long row = (long)chunk * Tlen + t;
          Vec8<T>::load(reinterpret_cast<const T*>(a.x) + row * a.ld + c, xa[u]);
          if (has_b) Vec8<T>::load(reinterpret_cast<const T*>(b.x) + row * b.ld + c, xb[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < COL_UNR; ++u) {
      const long prow = pa + (long)u * TY;
      if (prow >= p1) continue;
      float o[8];
      if (tt[u] < 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(xa[u][j] - ma[j], sa[j], ha[j]);
        if (has_b) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += fmaf(xb[u][j] - mb[j], sb[j], hb[j]);
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
      }
      Vec8<T>::store(out + prow * C + c, o);
    }
  }
}

// backward pass 1: g = dout * (y > 0);  red[0][c] += g, red[1][c] += g*xhat_a, red[2][c] += g*xhat_b   (double)
template <typename T>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(const T* __restrict__ dout, long ld_dout, const T* __restrict__ y, int y_lead, int y_trail,
                     int relu, BnBranch a, BnBranch b, int has_b, long n_chunks, int Tlen, int C,
                     double* __restrict__ red, long rows_per_block) {
  extern __shared__ double dbuf[];
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const long rows = n_chunks * Tlen;
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  const int Py = Tlen + y_lead + y_trail;
  float ma[8], ia[8], mb[8], ib[8];
  load8_f32(a.mean + c, ma); load8_f32(a.invstd + c, ia);
  if (has_b) { load8_f32(b.mean + c, mb); load8_f32(b.invstd + c, ib); }
  // fp32 partials per thread (the launch keeps a thread's share below ~512 rows), double from the block reduction on
  float f0[8], f1[8], f2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f0[j] = f1[j] = f2[j] = 0.f; }
  for (long ra = r0 + threadIdx.y; ra < r1; ra += (long)TY * 2) {
    float g[2][8], yv[2][8], xa[2][8], xb[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long r = ra + (long)u * TY;
      if (r < r1) {
        Vec8<T>::load(dout + r * ld_dout + c, g[u]);
        if (relu) {
          const unsigned chunk = (unsigned)r / (unsigned)Tlen;
          const long prow = (long)chunk * Py + (r - (long)chunk * Tlen) + y_lead;
          Vec8<T>::load(y + prow * C + c, yv[u]);
        }
        Vec8<T>::load(reinterpret_cast<const T*>(a.x) + r * a.ld + c, xa[u]);
        if (has_b) Vec8<T>::load(reinterpret_cast<const T*>(b.x) + r * b.ld + c, xb[u]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { g[u][j] = 0.f; yv[u][j] = 1.f; xa[u][j] = 0.f; xb[u][j] = 0.f; }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = (relu && !(yv[u][j] > 0.f)) ? 0.f : g[u][j];
        f0[j] += gg;
        f1[j] = fmaf(gg, (xa[u][j] - ma[j]) * ia[j], f1[j]);
        if (has_b) f2[j] = fmaf(gg, (xb[u][j] - mb[j]) * ib[j], f2[j]);
      }
    }
  }
  double s0[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = f0[j]; s1[j] = f1[j]; s2[j] = f2[j]; }
  reduce_over_ty(s0, dbuf);
  reduce_over_ty(s1, dbuf);
  if (has_b) reduce_over_ty(s2, dbuf);
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(red + c + j, s0[j]);
      atomicAdd(red + C + c + j, s1[j]);
      if (has_b) atomicAdd(red + 2 * C + c + j, s2[j]);
    }
  }
}

struct BnGradOut {
  void* dx;     // gradient w.r.t. the conv output, written in a time-padded layout (zero halos)
  long ld;      // row pitch of dx
  int lead, trail;
  float* dgamma; float* dbeta;   // accumulated (+=)
};

// backward pass 2: dx = gamma*invstd*(g - sum_g/N - xhat*sum_gxhat/N) for each branch, padded layouts; block 0 adds
// dgamma/dbeta.  Per column the expression is affine in (g, x - mean): constants in registers.
template <typename T>
__global__ void __launch_bounds__(512)
bn_bwd_apply_kernel(const T* __restrict__ dout, long ld_dout, const T* __restrict__ y, int y_lead, int y_trail,
                    int relu, BnBranch a, BnBranch b, int has_b, BnGradOut ga, BnGradOut gb, long n_chunks,
                    int Tlen, int C, const double* __restrict__ red, long prows_per_block) {
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const long rows = n_chunks * Tlen;
  const double invN = 1.0 / (double)rows;
  const int Py = Tlen + y_lead + y_trail;
  if (blockIdx.x == 0) {
    for (int cc = threadIdx.y * blockDim.x + threadIdx.x; cc < C; cc += blockDim.x * blockDim.y) {
      ga.dgamma[cc] += (float)red[C + cc];
      ga.dbeta[cc] += (float)red[cc];
      if (has_b) { gb.dgamma[cc] += (float)red[2 * C + cc]; gb.dbeta[cc] += (float)red[cc]; }
    }
  }
  float A[2][8], Bx[2][8], K[2][8], Mn[2][8];           // dx = A*g + K - (x - mean)*Bx
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    if (br == 1 && !has_b) break;
    const BnBranch& bx = br == 0 ? a : b;
    float m[8], is[8], gm[8];
    load8_f32(bx.mean + c, m); load8_f32(bx.invstd + c, is); load8_f32(bx.gamma + c, gm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = (float)(red[c + j] * invN);
      const float sgx = (float)(red[(br + 1) * C + c + j] * invN);
      const float k1 = gm[j] * is[j];
      A[br][j] = k1;
      Bx[br][j] = k1 * sgx * is[j];
      K[br][j] = -k1 * sg;
      Mn[br][j] = m[j];
    }
  }
  // pass over the larger of the two padded extents; each branch guards its own range
  const int Pa = Tlen + ga.lead + ga.trail;
  const int Pb = has_b ? Tlen + gb.lead + gb.trail : 0;
  const unsigned Pm = Pa > Pb ? Pa : Pb;
  const long prows = n_chunks * (long)Pm;
  const long p0 = (long)blockIdx.x * prows_per_block;
  const long p1 = min(prows, p0 + prows_per_block);
  for (long pa = p0 + threadIdx.y; pa < p1; pa += (long)TY * 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long prow = pa + (long)u * TY;
      if (prow >= p1) continue;
      const unsigned chunk = (unsigned)prow / Pm;
      const int pp = (int)((unsigned)prow - chunk * Pm);
      // the gradient row (shared by both branches) -- loaded once
      float g[8];
      bool have_g = false;
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
          const long r = (long)chunk * Tlen + t;
          if (!have_g || ga.lead != gb.lead) {
            Vec8<T>::load(dout + r * ld_dout + c, g);
            if (relu) {
              float yv[8];
              Vec8<T>::load(y + ((long)chunk * Py + t + y_lead) * C + c, yv);
#pragma unroll
              for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
            }
            have_g = true;
          }
          float xv[8];
          Vec8<T>::load(reinterpret_cast<const T*>(bx.x) + r * bx.ld + c, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(-(xv[j] - Mn[br][j]), Bx[br][j], fmaf(A[br][j], g[j], K[br][j]));
        }
        Vec8<T>::store(reinterpret_cast<T*>(go.dx) + ((long)chunk * Pbr + pp) * go.ld + c, o);
      }
    }
  }
}

// launch shape of the column-wise kernels: block (tx, ty) with ty as large as 512 threads / 48 KB of reduction scratch allow,
// `per_sm` blocks per SM, contiguous row ranges of `rpb` rows per block
struct ColLaunch { dim3 block; int grid; long rpb; };
static ColLaunch col_launch(long rows, int tx, int per_sm, long max_rpb = 1L << 40) {
  int ty = 512 / tx; if (ty < 1) ty = 1; if (ty > 8) ty = 8;
  while (ty > 1 && (size_t)(ty - 1) * tx * 8 * sizeof(double) > 48 * 1024) --ty;
  long nblk = (long)num_sms() * per_sm;
  long rpb = (rows + nblk - 1) / nblk;
  if (rpb > max_rpb) rpb = max_rpb;
  const long quantum = (long)ty * COL_UNR;
  rpb = (rpb + quantum - 1) / quantum * quantum;
  nblk = (rows + rpb - 1) / rpb;
  ColLaunch cl;
  cl.block = dim3(tx, ty);
  cl.grid = (int)(nblk > 0 ? nblk : 1);
  cl.rpb = rpb;
  return cl;
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
  const uint32_t thr = (r != nullptr && drop_p > 0.f) ? drop_threshold16(drop_p) : 0u;
  const float dscale = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  long blocks = (rows + 7) / 8;
  const long cap = (long)num_sms() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
#define SST_LN_FWD(T_, NV_, EX_)                                                                                             \
  ln_fwd_kernel<T_, NV_, EX_><<<grid, 256, 0, st>>>((const T_*)x, (const T_*)r, (T_*)y, (T_*)s_out, gamma, beta, mean, rstd, rows, D, \
                                                    eps, thr, dscale, seed)
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
  long blocks = (rows + 63) / 64;
  long cap = (long)num_sms() * 2;
  const int grid = (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
  const size_t smem = (size_t)2 * D * sizeof(float);
#define SST_LN_BWD(T_, NV_, EX_)                                                                                             \
  ln_bwd_kernel<T_, NV_, EX_><<<grid, 256, smem, st>>>((const T_*)dy, (const T_*)s, mean, rstd, gamma, (T_*)ds, (T_*)dr, dgamma, dbeta, \
                                                       rows, D, thr, dscale, seed)
  if (dtype == SST_F32) { if (D == 768) SST_LN_BWD(float, 3, true); else SST_LN_BWD(float, LN_MAXV, false); }
  else { if (D == 768) SST_LN_BWD(__nv_bfloat16, 3, true); else SST_LN_BWD(__nv_bfloat16, LN_MAXV, false); }
#undef SST_LN_BWD
  return check_launch("layernorm_bwd");
}

int sst_colstats(int dtype, const void* x, int64_t rows, int C, int64_t ld, double* stats, void* stream) {
  SST_REQUIRE(C % 8 == 0 && C / 8 <= 512 && ld % 8 == 0, SST_E_ARG, "colstats: C=%d, ld=%ld must be multiples of 8", C, (long)ld);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st);
  SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
  if (rows <= 0) return SST_OK;
  const ColLaunch cl = col_launch(rows, C / 8, 2);
  const size_t sm = (size_t)(cl.block.y - 1) * C * sizeof(double);
  if (dtype == SST_F32) colstats_kernel<float><<<cl.grid, cl.block, sm, st>>>((const float*)x, rows, C, ld, stats, cl.rpb);
  else colstats_kernel<__nv_bfloat16><<<cl.grid, cl.block, sm, st>>>((const __nv_bfloat16*)x, rows, C, ld, stats, cl.rpb);
  return check_launch("colstats");
}

int sst_colsum_accum(int dtype, const void* x, int64_t rows, int C, int64_t ld, float* out, void* stream) {
  SST_REQUIRE(ld % 8 == 0 && C <= ld && (C + 7) / 8 <= 512, SST_E_ARG, "colsum: pitch %ld must be a multiple of 8 and >= C=%d", (long)ld, C);
  if (rows <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tx = (C + 7) / 8;
  const ColLaunch cl = col_launch(rows, tx, 2);
  const size_t sm = (size_t)(cl.block.y - 1) * tx * 8 * sizeof(float);
  if (dtype == SST_F32) colsum_kernel<float><<<cl.grid, cl.block, sm, st>>>((const float*)x, rows, C, ld, out, cl.rpb);
  else colsum_kernel<__nv_bfloat16><<<cl.grid, cl.block, sm, st>>>((const __nv_bfloat16*)x, rows, C, ld, out, cl.rpb);
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
  SST_REQUIRE(C / 8 <= 512, SST_E_ARG, "bn_apply: C=%d too wide", C);
  const ColLaunch cl = col_launch(n_chunks * (long)(T + lead + trail), C / 8, 4);
  if (dtype == SST_F32)
    bn_apply_kernel<float><<<cl.grid, cl.block, 0, st>>>(a, b, xb != nullptr, relu, (float*)out, n_chunks, T, C, lead, trail, cl.rpb);
  else
    bn_apply_kernel<__nv_bfloat16><<<cl.grid, cl.block, 0, st>>>(a, b, xb != nullptr, relu, (__nv_bfloat16*)out, n_chunks, T, C, lead,
                                                                 trail, cl.rpb);
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
  const ColLaunch c1 = col_launch(rows, C / 8, 3, 2048);
  const size_t sm = (size_t)(c1.block.y - 1) * C * sizeof(double);
  const int Pa = T + lead_a + trail_a, Pb = has_b ? T + lead_b + trail_b : 0;
  const ColLaunch c2 = col_launch(n_chunks * (long)(Pa > Pb ? Pa : Pb), C / 8, 4);
  if (dtype == SST_F32) {
    bn_bwd_reduce_kernel<float><<<c1.grid, c1.block, sm, st>>>((const float*)dout, ld_dout, (const float*)y, y_lead, y_trail, relu, a, b,
                                                               has_b, n_chunks, T, C, red, c1.rpb);
    bn_bwd_apply_kernel<float><<<c2.grid, c2.block, 0, st>>>((const float*)dout, ld_dout, (const float*)y, y_lead, y_trail, relu, a, b,
                                                             has_b, ga, gb, n_chunks, T, C, red, c2.rpb);
  } else {
    bn_bwd_reduce_kernel<__nv_bfloat16><<<c1.grid, c1.block, sm, st>>>((const __nv_bfloat16*)dout, ld_dout, (const __nv_bfloat16*)y,
                                                                       y_lead, y_trail, relu, a, b, has_b, n_chunks, T, C, red, c1.rpb);
    bn_bwd_apply_kernel<__nv_bfloat16><<<c2.grid, c2.block, 0, st>>>((const __nv_bfloat16*)dout, ld_dout, (const __nv_bfloat16*)y, y_lead,
                                                                     y_trail, relu, a, b, has_b, ga, gb, n_chunks, T, C, red, c2.rpb);
  }
  return check_launch("bn_bwd", 2);
}

}  // extern "C"

// CUDA-core GEMM with fp32 accumulation: the fp32 "parity mode" engine (1e-4 relative against the reference
// rules out bf16/tf32 tensor-core products) and the on-device cross-check of the
// tcgen05 kernel.  Implements exactly the SstGemmDesc semantics of include/sst.h for float or bf16 operands.
#include "sst_common.cuh"

namespace sst {

struct SimtParams {
  int M, N, K;
  int mode_mn;
  int b_plain_mn;     // TN_BMN: B is a plain (K, N) row-major matrix
  long lda, ldb, ldc, ldaux;
  long a_rows, b_rows;
  int n_seg, kseg, nsegc;
  int a_row_shift[3], a_col0[3], b_row_shift[3], b_col0[3];
  int epilogue;
  float alpha, mask_scale, drop_scale;
  uint32_t drop_thr;
  unsigned long long seed;
  const unsigned long long* salt;
  const float* bias;
  const void* aux;
  int aux_dtype;
  void* C;
  int out_dtype;
  int remap_P, remap_T, remap_j0;
  int out_seg_cols, out_grp_cols;
  long out_seg_stride, out_grp_off[3];
};

template <typename T>
__device__ __forceinline__ float load_a(const T* A, const SimtParams& p, int m, int k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (!p.mode_mn) {
    int s = k / p.kseg;
    long row = (long)m + p.a_row_shift[s];
    if (row < 0 || row >= p.a_rows) return 0.f;
    return to_f32(A[row * p.lda + p.a_col0[s] + (k - s * p.kseg)]);
  }
  return to_f32(A[(long)k * p.lda + m]);
}
template <typename T>
__device__ __forceinline__ float load_b(const T* B, const SimtParams& p, int n, int k) {
  if (n >= p.N || k >= p.K) return 0.f;
  if (p.b_plain_mn) return to_f32(B[(long)k * p.ldb + n]);
  if (!p.mode_mn) return to_f32(B[(long)n * p.ldb + k]);
  int s = n / p.nsegc;
  long row = (long)k + p.b_row_shift[s];
  if (row < 0 || row >= p.b_rows) return 0.f;
  return to_f32(B[row * p.ldb + p.b_col0[s] + (n - s * p.nsegc)]);
}

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, const SimtParams p) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  // Two-level accumulation: each 16-deep K tile is summed into `part`, which is then folded into `acc` with Kahan
  // compensation.  Weight gradients contract over thousands of tokens with heavy sign cancellation; a plain running
  // fp32 sum there loses ~sqrt(K) more digits than the blocked sums of the reference's CPU BLAS.
  float acc[4][4], comp[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; comp[i][j] = 0.f; }

  for (int k0 = 0; k0 < p.K; k0 += TK) {
    // 64x16 elements per operand, 256 threads -> 4 each.  In TN mode k is the contiguous index, in MN mode m/n is.
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * 256;
      int kk, mm;
      if (!p.mode_mn) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      sA[kk][mm] = load_a(A, p, m0 + mm, k0 + kk);
      sB[kk][mm] = load_b(B, p, n0 + mm, k0 + kk);
    }
    __syncthreads();
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = __fsub_rn(part[i][j], comp[i][j]);
        const float t = __fadd_rn(acc[i][j], y);
        comp[i][j] = __fsub_rn(__fsub_rn(t, acc[i][j]), y);
        acc[i][j] = t;
      }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    long out_row = m;
    if (p.remap_P > 0) {
      int chunk = m / p.remap_P;
      int t = m - chunk * p.remap_P - p.remap_j0;
      if (t < 0 || t >= p.remap_T) continue;
      out_row = (long)chunk * p.remap_T + t;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.epilogue & SST_EPI_BIAS) v += p.bias[n];
      if (p.epilogue & SST_EPI_RELU) v = fmaxf(v, 0.f);
      if (p.epilogue & SST_EPI_DROPOUT)
        v = philox_keep16(salted(p.seed, p.salt), (unsigned long long)m * (unsigned long long)p.N + n, p.drop_thr) ? v * p.drop_scale : 0.f;
      if (p.epilogue & SST_EPI_MULMASK)
        v *= (ld_as_f32(p.aux, (long)m * p.ldaux + n, p.aux_dtype) > 0.f) ? p.mask_scale : 0.f;
      long ci = out_row * p.ldc + n;
      if (p.out_seg_cols > 0) {
        const int grp = n / p.out_grp_cols, nin = n - grp * p.out_grp_cols;
        const int seg = nin / p.out_seg_cols;
        ci = p.out_grp_off[grp] + (long)seg * p.out_seg_stride + out_row * p.ldc + (nin - seg * p.out_seg_cols);
      }
      if (p.epilogue & SST_EPI_ACCUM) v += ld_as_f32(p.C, ci, p.out_dtype);
      st_from_f32(p.C, ci, p.out_dtype, v);
    }
  }
}

int launch_gemm_simt(const SstGemmDesc& d, const void* A, const void* B, void* C, const void* bias, const void* aux,
                     cudaStream_t st) {
  SimtParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)d.M; p.N = (int)d.N; p.K = (int)d.K;
  p.mode_mn = d.layout == SST_GEMM_NT_MN;
  p.b_plain_mn = d.layout == SST_GEMM_TN_BMN;
  p.lda = d.lda; p.ldb = d.ldb; p.ldc = d.ldc; p.ldaux = d.ldaux;
  p.a_rows = d.a_rows; p.b_rows = d.b_rows;
  p.n_seg = d.n_seg > 0 ? d.n_seg : 1;
  SST_REQUIRE(p.n_seg <= 3, SST_E_ARG, "n_seg must be <= 3");
  p.kseg = p.mode_mn ? (int)d.K : (int)(d.K / p.n_seg);
  p.nsegc = p.mode_mn ? (int)(d.N / p.n_seg) : (int)d.N;
  if (p.kseg < 1) p.kseg = 1;
  if (p.nsegc < 1) p.nsegc = 1;
  for (int s = 0; s < 3; ++s) {
    p.a_row_shift[s] = d.a_row_shift[s]; p.a_col0[s] = d.a_col0[s];
    p.b_row_shift[s] = d.b_row_shift[s]; p.b_col0[s] = d.b_col0[s];
  }
  p.epilogue = d.epilogue;
  p.alpha = d.alpha; p.mask_scale = d.mask_scale;
  p.drop_thr = drop_threshold16(d.drop_p);
  p.drop_scale = d.drop_p < 1.f ? 1.f / (1.f - d.drop_p) : 0.f;
  SST_REQUIRE(d.col_acc == nullptr || d.col_acc_mode == 0, SST_E_UNSUPPORTED, "col_acc is a tensor-core epilogue feature: use sst_colsum_accum / sst_colstats");
  p.seed = d.seed;
  p.salt = dropout_salt();
  p.bias = reinterpret_cast<const float*>(bias);
  p.aux = aux; p.aux_dtype = d.aux_dtype;
  p.C = C; p.out_dtype = d.out_dtype;
  p.remap_P = d.remap_P; p.remap_T = d.remap_T; p.remap_j0 = d.remap_j0;
  p.out_seg_cols = (int)d.out_seg_cols; p.out_grp_cols = (int)d.out_grp_cols; p.out_seg_stride = d.out_seg_stride;
  for (int s = 0; s < 3; ++s) p.out_grp_off[s] = d.out_grp_off[s];
  if (d.out_seg_cols > 0)
    SST_REQUIRE(d.out_grp_cols % d.out_seg_cols == 0 && cdiv(d.N, d.out_grp_cols) <= 3, SST_E_ARG, "segmented output: bad group / segment sizes");
  dim3 grid(cdiv(d.N, 64), cdiv(d.M, 64));
  if (d.dtype == SST_F32)
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(A), reinterpret_cast<const float*>(B), p);
  else
    gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                                         reinterpret_cast<const __nv_bfloat16*>(B), p);
  return check_launch("gemm_simt");
}

}  // namespace sst

// Loss kernels of the hybrid CTC / attention objective (recognition_model.py:93-107):
//   log-softmax rows                              (F.log_softmax, recognition_model.py:93)
//   CTC alpha/beta recursion, one warp per utterance, log space, gradient written w.r.t. the raw logits
//                                                 (F.ctc_loss blank=43, reduction 'mean', recognition_model.py:98)
//   label-smoothed cross entropy with the sum-exp regulariser, forward + gradient in one pass
//                                                 (LabelSmoothingLoss.py:13-15)
#include "vec.cuh"

namespace sst {

__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + __logf(__expf(a - m) + __expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(fmaxf(a, b), c);
  if (m == -INFINITY) return -INFINITY;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// one warp per row; out fp32 with pitch C
template <typename T>
__global__ void log_softmax_kernel(const T* __restrict__ x, long ld, float* __restrict__ out, long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, to_f32(x[row * ld + c]));
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(to_f32(x[row * ld + c]) - mx);
  s = warp_sum(s);
  const float l = mx + logf(s);
  for (int c = lane; c < C; c += 32) out[row * C + c] = to_f32(x[row * ld + c]) - l;
}

constexpr int CTC_MAXSPL = 16;   // states per lane -> 2S+1 <= 512

// lp: (B, L, C) fp32 log-probs.  alpha_ws: (B, L, NS) fp32.  grad: (B*L, ldg) in T.  One warp per utterance.
template <typename T>
__global__ void __launch_bounds__(32)
ctc_kernel(const float* __restrict__ lp, const long* __restrict__ targets, int tgt_pitch, const int* __restrict__ in_lens,
           const int* __restrict__ tgt_lens, int L, int C, int blank, int NS, float* __restrict__ alpha_ws,
           float* __restrict__ nll, T* __restrict__ grad, long ldg, float gcoef, int B) {
  extern __shared__ float occ[];   // [C]
  const int b = blockIdx.x, lane = threadIdx.x;
  const int Tn = min(in_lens[b], L);
  const int S = tgt_lens[b];
  const int n = 2 * S + 1;
  int spl = (NS + 31) / 32;
  if (spl < 2) spl = 2;
  const float* lpb = lp + (long)b * L * C;
  float* aw = alpha_ws + (long)b * L * NS;
  int lab[CTC_MAXSPL];
  bool skip_a[CTC_MAXSPL], skip_b[CTC_MAXSPL];
#pragma unroll
  for (int u = 0; u < CTC_MAXSPL; ++u) {
    const int s = lane * spl + u;
    lab[u] = blank; skip_a[u] = false; skip_b[u] = false;
    if (u < spl && s < n && (s & 1)) {
      const int l = (int)targets[(long)b * tgt_pitch + (s >> 1)];
      lab[u] = l;
      skip_a[u] = s >= 2 && l != (int)targets[(long)b * tgt_pitch + (s >> 1) - 1];
      skip_b[u] = s + 2 < n && l != (int)targets[(long)b * tgt_pitch + (s >> 1) + 1];
    }
  }
  float a[CTC_MAXSPL];
  float ll = -INFINITY;
  if (Tn > 0) {
    // ---- alpha ----
#pragma unroll
    for (int u = 0; u < CTC_MAXSPL; ++u) {
      const int s = lane * spl + u;
      a[u] = -INFINITY;
      if (u < spl && s < n && s <= 1) a[u] = lpb[lab[u]];
      if (u < spl && s < n) aw[s] = a[u];
    }
    for (int t = 1; t < Tn; ++t) {
      float emit[CTC_MAXSPL];
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) emit[u] = (u < spl) ? __ldg(lpb + (long)t * C + lab[u]) : 0.f;
      // neighbours from the previous lane: its last and second-to-last state
      float last = a[0], last2 = a[0];
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) { if (u == spl - 1) last = a[u]; if (u == spl - 2) last2 = a[u]; }
      float p1 = __shfl_up_sync(0xffffffffu, last, 1), p2 = __shfl_up_sync(0xffffffffu, last2, 1);
      if (lane == 0) { p1 = -INFINITY; p2 = -INFINITY; }
      float na[CTC_MAXSPL];
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) {
        if (u < spl) {
          const float m1 = u >= 1 ? a[u >= 1 ? u - 1 : 0] : p1;
          const float m2 = u >= 2 ? a[u >= 2 ? u - 2 : 0] : (u == 1 ? p1 : p2);
          na[u] = (skip_a[u] ? lse3(a[u], m1, m2) : lse2(a[u], m1)) + emit[u];
        }
      }
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) {
        const int s = lane * spl + u;
        if (u < spl) {
          a[u] = s < n ? na[u] : -INFINITY;
          if (s < n) aw[(long)t * NS + s] = a[u];
        }
      }
    }
    // log-likelihood = lse(alpha[T-1][n-1], alpha[T-1][n-2])
    float mine = -INFINITY;
#pragma unroll
    for (int u = 0; u < CTC_MAXSPL; ++u) {
      const int s = lane * spl + u;
      if (u < spl && (s == n - 1 || s == n - 2)) mine = lse2(mine, a[u]);
    }
    // combine across lanes in log space
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = lse2(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    ll = mine;
  }
  if (lane == 0) nll[b] = -ll;
  const float gscale = gcoef / (float)(max(S, 1)) / (float)B;

  // ---- beta + gradient ----
  float be[CTC_MAXSPL];
#pragma unroll
  for (int u = 0; u < CTC_MAXSPL; ++u) {
    const int s = lane * spl + u;
    be[u] = -INFINITY;
    if (Tn > 0 && u < spl && s < n && s >= n - 2) be[u] = lpb[(long)(Tn - 1) * C + lab[u]];
  }
  for (int t = L - 1; t >= 0; --t) {
    T* grow = grad + ((long)b * L + t) * ldg;
    if (t >= Tn) {
      for (int c = lane; c < ldg; c += 32) grow[c] = from_f32<T>(0.f);
      continue;
    }
    for (int c = lane; c < C; c += 32) occ[c] = 0.f;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < CTC_MAXSPL; ++u) {
      const int s = lane * spl + u;
      if (u < spl && s < n) {
        const float v = aw[(long)t * NS + s] + be[u];
        if (v > -INFINITY) atomicAdd(&occ[lab[u]], __expf(v - __ldg(lpb + (long)t * C + lab[u]) - ll));
      }
    }
    __syncwarp();
    for (int c = lane; c < ldg; c += 32) {
      float g = 0.f;
      if (c < C) g = (__expf(lpb[(long)t * C + c]) - occ[c]) * gscale;
      grow[c] = from_f32<T>(g);
    }
    __syncwarp();
    if (t > 0) {
      float emit[CTC_MAXSPL];
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) emit[u] = (u < spl) ? __ldg(lpb + (long)(t - 1) * C + lab[u]) : 0.f;
      float first = be[0], second = be[1];
      float n1 = __shfl_down_sync(0xffffffffu, first, 1), n2 = __shfl_down_sync(0xffffffffu, second, 1);
      if (lane == 31) { n1 = -INFINITY; n2 = -INFINITY; }
      float nb[CTC_MAXSPL];
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) {
        if (u < spl) {
          const float m1 = (u + 1 < spl) ? be[u + 1 < CTC_MAXSPL ? u + 1 : 0] : n1;
          const float m2 = (u + 2 < spl) ? be[u + 2 < CTC_MAXSPL ? u + 2 : 0] : (u + 1 < spl ? n1 : n2);
          nb[u] = (skip_b[u] ? lse3(be[u], m1, m2) : lse2(be[u], m1)) + emit[u];
        }
      }
#pragma unroll
      for (int u = 0; u < CTC_MAXSPL; ++u) {
        const int s = lane * spl + u;
        if (u < spl) be[u] = s < n ? nb[u] : -INFINITY;
      }
    }
  }
}

// loss_out[slot] = coef * mean_b( nll_b / max(tgt_len_b, 1) )   (deterministic: one warp, fixed order)
__global__ void ctc_finalize_kernel(const float* __restrict__ nll, const int* __restrict__ tgt_lens, int B, float* loss_out) {
  double s = 0.0;
  for (int b = threadIdx.x; b < B; b += 32) s += (double)nll[b] / (double)max(tgt_lens[b], 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *loss_out = (float)(s / (double)B);
}

// one warp per row: ce_row (0 when target == ignore), sumexp_row, and the gradient
template <typename T, typename TG>
__global__ void ce_sumexp_kernel(const T* __restrict__ x, long ld, const long* __restrict__ target, long rows, int C, int ignore,
                                 float eps, float inv_nvalid, float inv_S, float gcoef, float* __restrict__ row_ce,
                                 float* __restrict__ row_se, TG* __restrict__ grad, long ldg) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long tg = target[row];
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, to_f32(x[row * ld + c]));
  mx = warp_max(mx);
  float s = 0.f, se = 0.f;
  for (int c = lane; c < C; c += 32) { float v = to_f32(x[row * ld + c]); s += expf(v - mx); se += expf(v); }
  s = warp_sum(s);
  se = warp_sum(se);
  const float l = mx + logf(s);
  const bool valid = tg != ignore;
  if (lane == 0) {
    row_ce[row] = valid ? l - to_f32(x[row * ld + tg]) : 0.f;
    row_se[row] = se;
  }
  if (grad != nullptr) {
    for (int c = lane; c < ldg; c += 32) {
      float g = 0.f;
      if (c < C) {
        const float v = to_f32(x[row * ld + c]);
        if (valid) g = (1.f - eps) * inv_nvalid * (expf(v - l) - (c == tg ? 1.f : 0.f));
        g += eps * inv_S * expf(v);
        g *= gcoef;
      }
      grad[row * ldg + c] = from_f32<TG>(g);
    }
  }
}

__global__ void ce_finalize_kernel(const float* __restrict__ row_ce, const float* __restrict__ row_se, long rows, float eps,
                                   float inv_nvalid, float inv_S, float* loss_out) {
  double a = 0.0, b = 0.0;
  for (long r = threadIdx.x; r < rows; r += 32) { a += row_ce[r]; b += row_se[r]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (threadIdx.x == 0) *loss_out = (float)((1.0 - eps) * a * inv_nvalid + eps * inv_S * b);
}

}  // namespace sst

using namespace sst;

extern "C" {

/* logits: (B*L, ld) in dtype; lp_ws: float[B*L*C]; alpha_ws: float[B*L*(2*Smax+1)]; nll: float[B];
 * grad: (B*L, ldg) in dtype (columns >= C zeroed) = gcoef * d(mean_b nll_b/len_b)/d logits;  loss_out: float[1]. */
int sst_ctc_loss(int logits_dtype, int grad_dtype, int B, int L, int C, int blank, const void* logits, int64_t ld,
                 const int64_t* targets, int Smax, const int32_t* in_lens, const int32_t* tgt_lens, float gcoef, float* lp_ws,
                 float* alpha_ws, float* nll, void* grad, int64_t ldg, float* loss_out, void* stream) {
  SST_REQUIRE(2 * Smax + 1 <= 32 * CTC_MAXSPL, SST_E_ARG, "ctc: target length %d too long (max %d)", Smax, (32 * CTC_MAXSPL - 1) / 2);
  SST_REQUIRE(C <= 1024 && blank < C && ldg >= C, SST_E_ARG, "ctc: bad class count / gradient pitch");
  if (B <= 0 || L <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long rows = (long)B * L;
  const int NS = 2 * Smax + 1;
  if (logits_dtype == SST_F32) log_softmax_kernel<float><<<(int)((rows + 7) / 8), 256, 0, st>>>((const float*)logits, ld, lp_ws, rows, C);
  else log_softmax_kernel<__nv_bfloat16><<<(int)((rows + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)logits, ld, lp_ws, rows, C);
  if (grad_dtype == SST_F32)
    ctc_kernel<float><<<B, 32, C * sizeof(float), st>>>(lp_ws, (const long*)targets, Smax, in_lens, tgt_lens, L, C, blank, NS, alpha_ws,
                                                        nll, (float*)grad, ldg, gcoef, B);
  else
    ctc_kernel<__nv_bfloat16><<<B, 32, C * sizeof(float), st>>>(lp_ws, (const long*)targets, Smax, in_lens, tgt_lens, L, C, blank, NS,
                                                                alpha_ws, nll, (__nv_bfloat16*)grad, ldg, gcoef, B);
  ctc_finalize_kernel<<<1, 32, 0, st>>>(nll, tgt_lens, B, loss_out);
  return check_launch("ctc_loss", 3);
}

/* logits (rows, ld) with rows = B*S; loss = (1-eps)*CE(ignore_index, mean over non-ignored) + eps/S * sum(exp(logits)).
 * row_ws: float[2*rows].  grad (rows, ldg) = gcoef * dloss/dlogits (nullable). */
int sst_ce_sumexp_loss(int logits_dtype, int grad_dtype, int64_t rows, int S, int C, const void* logits, int64_t ld,
                       const int64_t* target, int ignore, float eps, int64_t n_valid, float gcoef, float* row_ws, void* grad,
                       int64_t ldg, float* loss_out, void* stream) {
  if (rows <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float inv_nv = n_valid > 0 ? 1.f / (float)n_valid : 0.f;
  const float inv_S = 1.f / (float)S;
  float* rce = row_ws; float* rse = row_ws + rows;
  const int grid = (int)((rows + 7) / 8);
  typedef __nv_bfloat16 bf;
#define SST_CE_LAUNCH(TL, TGR) ce_sumexp_kernel<TL, TGR><<<grid, 256, 0, st>>>((const TL*)logits, ld, (const long*)target, rows, C, \
      ignore, eps, inv_nv, inv_S, gcoef, rce, rse, (TGR*)grad, ldg)
  if (logits_dtype == SST_F32 && grad_dtype == SST_F32) SST_CE_LAUNCH(float, float);
  else if (logits_dtype == SST_F32) SST_CE_LAUNCH(float, bf);
  else if (grad_dtype == SST_F32) SST_CE_LAUNCH(bf, float);
  else SST_CE_LAUNCH(bf, bf);
#undef SST_CE_LAUNCH
  ce_finalize_kernel<<<1, 32, 0, st>>>(rce, rse, rows, eps, inv_nv, inv_S, loss_out);
  return check_launch("ce_sumexp_loss", 2);
}

}  // extern "C"

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

// CTC forward-backward lattice of one utterance per CTA, log space (Graves 2006 eq. 6-11): warp 0 runs the alpha recursion
// over t = 0 .. T-1, warp 1 the beta recursion over t = T-1 .. 0, concurrently; the 2S+1 lattice states are spread over
// the 32 lanes (SPL consecutive states per lane, neighbours by warp shuffle).  The utterance's log-probabilities are
// staged once in shared memory (all threads) so that the serial loops never wait on HBM.  alpha and beta (each already
// containing the emission at t) go to the workspace; ctc_grad_kernel turns them into the gradient in parallel over (b, t).
template <int SPL, bool SMEM_LP>
__global__ void __launch_bounds__(128)
ctc_lattice_kernel(const float* __restrict__ lp, const long* __restrict__ targets, int tgt_pitch, const int* __restrict__ in_lens,
                   const int* __restrict__ tgt_lens, int L, int C, int blank, int NS, float* __restrict__ alpha_ws,
                   float* __restrict__ beta_ws, float* __restrict__ nll) {
  extern __shared__ float slp[];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tn = min(in_lens[b], L);
  const int S = tgt_lens[b];
  const int n = 2 * S + 1;
  const float* lpb = lp + (long)b * L * C;
  if (SMEM_LP) {
    const long tot = (long)Tn * C;
    if ((((uintptr_t)lpb) & 15) == 0) {
      const float4* src = reinterpret_cast<const float4*>(lpb);
      float4* dst = reinterpret_cast<float4*>(slp);
      for (long k = threadIdx.x; k < tot / 4; k += blockDim.x) dst[k] = __ldg(src + k);
      for (long k = (tot & ~3L) + threadIdx.x; k < tot; k += blockDim.x) slp[k] = __ldg(lpb + k);
    } else {
      for (long k = threadIdx.x; k < tot; k += blockDim.x) slp[k] = __ldg(lpb + k);
    }
    __syncthreads();
  }
  if (warp > 1) return;
  if (Tn <= 0) { if (warp == 0 && lane == 0) nll[b] = INFINITY; return; }
  const float* LP = SMEM_LP ? slp : lpb;
  int lab[SPL];
  bool skip[SPL];      // alpha warp: transition from s-2 allowed; beta warp: transition to s+2 allowed
#pragma unroll
  for (int u = 0; u < SPL; ++u) {
    const int s = lane * SPL + u;
    lab[u] = blank; skip[u] = false;
    if (s < n && (s & 1)) {
      const int l = (int)targets[(long)b * tgt_pitch + (s >> 1)];
      lab[u] = l;
      if (warp == 0) skip[u] = s >= 2 && l != (int)targets[(long)b * tgt_pitch + (s >> 1) - 1];
      else skip[u] = s + 2 < n && l != (int)targets[(long)b * tgt_pitch + (s >> 1) + 1];
    }
  }
  float a[SPL];
  if (warp == 0) {
    float* aw = alpha_ws + (long)b * L * NS;
#pragma unroll
    for (int u = 0; u < SPL; ++u) {
      const int s = lane * SPL + u;
      a[u] = (s < n && s <= 1) ? LP[lab[u]] : -INFINITY;
      if (s < n) aw[s] = a[u];
    }
    for (int t = 1; t < Tn; ++t) {
      float emit[SPL];
#pragma unroll
      for (int u = 0; u < SPL; ++u) emit[u] = LP[(long)t * C + lab[u]];
      float p1 = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1), p2 = __shfl_up_sync(0xffffffffu, a[SPL - 2], 1);
      if (lane == 0) { p1 = -INFINITY; p2 = -INFINITY; }
      float na[SPL];
#pragma unroll
      for (int u = 0; u < SPL; ++u) {
        const float m1 = u >= 1 ? a[u >= 1 ? u - 1 : 0] : p1;
        const float m2 = u >= 2 ? a[u >= 2 ? u - 2 : 0] : (u == 1 ? p1 : p2);
        na[u] = (skip[u] ? lse3(a[u], m1, m2) : lse2(a[u], m1)) + emit[u];
      }
#pragma unroll
      for (int u = 0; u < SPL; ++u) {
        const int s = lane * SPL + u;
        a[u] = s < n ? na[u] : -INFINITY;
        if (s < n) aw[(long)t * NS + s] = a[u];
      }
    }
    // log-likelihood = lse(alpha[T-1][n-1], alpha[T-1][n-2])
    float mine = -INFINITY;
#pragma unroll
    for (int u = 0; u < SPL; ++u) {
      const int s = lane * SPL + u;
      if (s == n - 1 || s == n - 2) mine = lse2(mine, a[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = lse2(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    if (lane == 0) nll[b] = -mine;
  } else {
    float* bw = beta_ws + (long)b * L * NS;
#pragma unroll
    for (int u = 0; u < SPL; ++u) {
      const int s = lane * SPL + u;
      a[u] = (s < n && s >= n - 2) ? LP[(long)(Tn - 1) * C + lab[u]] : -INFINITY;
      if (s < n) bw[(long)(Tn - 1) * NS + s] = a[u];
    }
    for (int t = Tn - 2; t >= 0; --t) {
      float emit[SPL];
#pragma unroll
      for (int u = 0; u < SPL; ++u) emit[u] = LP[(long)t * C + lab[u]];
      float n1 = __shfl_down_sync(0xffffffffu, a[0], 1), n2 = __shfl_down_sync(0xffffffffu, a[1], 1);
      if (lane == 31) { n1 = -INFINITY; n2 = -INFINITY; }
      float nb[SPL];
#pragma unroll
      for (int u = 0; u < SPL; ++u) {
        const float m1 = (u + 1 < SPL) ? a[u + 1 < SPL ? u + 1 : 0] : n1;
        const float m2 = (u + 2 < SPL) ? a[u + 2 < SPL ? u + 2 : 0] : (u + 1 < SPL ? n1 : n2);
        nb[u] = (skip[u] ? lse3(a[u], m1, m2) : lse2(a[u], m1)) + emit[u];
      }
#pragma unroll
      for (int u = 0; u < SPL; ++u) {
        const int s = lane * SPL + u;
        a[u] = s < n ? nb[u] : -INFINITY;
        if (s < n) bw[(long)t * NS + s] = a[u];
      }
    }
  }
}

// gradient w.r.t. the raw logits, one warp per (utterance, frame):
//   g[c] = gscale_b * ( softmax[c] - sum_{s: label(s) = c} exp(alpha[t][s] + beta[t][s] - lp[t][c] - ll) ),  0 for t >= T_b
template <typename T>
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ lp, const long* __restrict__ targets, int tgt_pitch, const int* __restrict__ in_lens,
                const int* __restrict__ tgt_lens, int L, int C, int blank, int NS, const float* __restrict__ alpha_ws,
                const float* __restrict__ beta_ws, const float* __restrict__ nll, T* __restrict__ grad, long ldg, float gcoef, int B) {
  extern __shared__ float occ_all[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  float* occ = occ_all + wi * C;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + wi;
  if (row >= (long)B * L) return;
  const int b = (int)(row / L), t = (int)(row - (long)b * L);
  T* grow = grad + row * ldg;
  const int Tn = min(in_lens[b], L);
  if (t >= Tn) {
    for (int c = lane; c < ldg; c += 32) grow[c] = from_f32<T>(0.f);
    return;
  }
  const int S = tgt_lens[b];
  const int n = 2 * S + 1;
  const float ll = -nll[b];
  const float gscale = gcoef / (float)(max(S, 1)) / (float)B;
  const float* lpr = lp + row * C;
  for (int c = lane; c < C; c += 32) occ[c] = 0.f;
  __syncwarp();
  const float* ar = alpha_ws + row * NS;
  const float* br = beta_ws + row * NS;
  for (int s = lane; s < n; s += 32) {
    const int l = (s & 1) ? (int)targets[(long)b * tgt_pitch + (s >> 1)] : blank;
    const float v = ar[s] + br[s];
    if (v > -INFINITY) atomicAdd(&occ[l], __expf(v - lpr[l] - ll));
  }
  __syncwarp();
  for (int c = lane; c < ldg; c += 32) {
    float g = 0.f;
    if (c < C) g = (__expf(lpr[c]) - occ[c]) * gscale;
    grow[c] = from_f32<T>(g);
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

// CTC best-path decode of one utterance per block: arg-max per frame (lowest index wins ties, as torch.argmax), merge
// repeats, drop blanks; order-preserving compaction with a block-wide exclusive scan per 256-frame chunk.
template <typename T>
__global__ void __launch_bounds__(256)
ctc_greedy_kernel(const T* __restrict__ x, long ld, const int* __restrict__ in_lens, int L, int C, int blank,
                  int* __restrict__ out_ids, int* __restrict__ out_lens) {
  __shared__ int s_am[257];        // arg-max of the chunk's frames, [0] = last frame of the previous chunk
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int Tn = in_lens ? min(in_lens[b], L) : L;
  if (tid == 0) { s_am[0] = -1; s_base = 0; }
  for (int k = tid; k < L; k += 256) out_ids[(long)b * L + k] = -1;
  __syncthreads();
  for (int f0 = 0; f0 < Tn; f0 += 256) {
    const int f = f0 + tid;
    int am = -1;
    if (f < Tn) {
      const T* row = x + ((long)b * L + f) * ld;
      float best = to_f32(row[0]);
      am = 0;
      for (int c = 1; c < C; ++c) {
        const float v = to_f32(row[c]);
        if (v > best) { best = v; am = c; }
      }
    }
    s_am[tid + 1] = am;
    __syncthreads();
    const int keep = (f < Tn && am != blank && am != s_am[tid]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < w; ++k) off += s_warp[k];
    if (keep) out_ids[(long)b * L + off + within] = am;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int k = 0; k < 8; ++k) tot += s_warp[k];
      s_base += tot;
      s_am[0] = s_am[min(256, Tn - f0)];
    }
    __syncthreads();
  }
  if (tid == 0) out_lens[b] = s_base;
}

// One step of the greedy attention-decoder search (greedy_search.py:22-37) for every sample: arg-max of the last position's
// logits (lowest index wins ties), append to the prefix, latch "has produced </S>", count the finished samples.
// One warp per sample, one block (searches run on a handful of utterances).
template <typename T>
__global__ void __launch_bounds__(1024)
greedy_pick_kernel(const T* __restrict__ logits, long row_stride, int B, int C, long long* __restrict__ tokens, long tok_sb, long tok_sp,
                   int pos, int eos, unsigned char* __restrict__ done, int* __restrict__ n_done) {
  __shared__ int s_cnt;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int b = w; b < B; b += nw) {
    const T* row = logits + (long)b * row_stride;
    float best = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = to_f32(row[c]);
      if (v > best || (v == best && c < am)) { best = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, am, o);
      if (ov > best || (ov == best && oi < am)) { best = ov; am = oi; }
    }
    if (lane == 0) {
      if (am == 0x7fffffff) am = 0;                   // every logit NaN: torch.argmax also answers with an index, not a fault
      tokens[(long)b * tok_sb + (long)pos * tok_sp] = am;
      const unsigned char d = done[b] | (am == eos ? 1 : 0);
      done[b] = d;
      if (d) atomicAdd(&s_cnt, 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_done = s_cnt;
}

}  // namespace sst

using namespace sst;

extern "C" {

/* logits: (B*L, ld) in dtype; lp_ws: float[B*L*C]; alpha_ws: float[2*B*L*(2*Smax+1)] (alpha, then beta); nll: float[B];
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
  float* beta_ws = alpha_ws + rows * NS;
  {
    const size_t lp_bytes = (size_t)L * C * sizeof(float);
    const bool in_smem = lp_bytes <= 200 * 1024;
    const size_t sm = in_smem ? lp_bytes : 0;
    const int spl = NS <= 64 ? 2 : NS <= 128 ? 4 : NS <= 256 ? 8 : 16;
#define SST_CTC_LAUNCH(SPL_, SM_)                                                                                          \
    do {                                                                                                                   \
      if (SM_) cudaFuncSetAttribute(ctc_lattice_kernel<SPL_, SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      ctc_lattice_kernel<SPL_, SM_><<<B, 128, sm, st>>>(lp_ws, (const long*)targets, Smax, in_lens, tgt_lens, L, C, blank, NS, \
                                                        alpha_ws, beta_ws, nll);                                           \
    } while (0)
    if (in_smem) {
      if (spl == 2) SST_CTC_LAUNCH(2, true); else if (spl == 4) SST_CTC_LAUNCH(4, true);
      else if (spl == 8) SST_CTC_LAUNCH(8, true); else SST_CTC_LAUNCH(16, true);
    } else {
      if (spl == 2) SST_CTC_LAUNCH(2, false); else if (spl == 4) SST_CTC_LAUNCH(4, false);
      else if (spl == 8) SST_CTC_LAUNCH(8, false); else SST_CTC_LAUNCH(16, false);
    }
#undef SST_CTC_LAUNCH
  }
  {
    const int wpb = 8;
    const int grid = (int)((rows + wpb - 1) / wpb);
    const size_t sm = (size_t)wpb * C * sizeof(float);
    if (grad_dtype == SST_F32)
      ctc_grad_kernel<float><<<grid, wpb * 32, sm, st>>>(lp_ws, (const long*)targets, Smax, in_lens, tgt_lens, L, C, blank, NS, alpha_ws,
                                                         beta_ws, nll, (float*)grad, ldg, gcoef, B);
    else
      ctc_grad_kernel<__nv_bfloat16><<<grid, wpb * 32, sm, st>>>(lp_ws, (const long*)targets, Smax, in_lens, tgt_lens, L, C, blank, NS,
                                                                 alpha_ws, beta_ws, nll, (__nv_bfloat16*)grad, ldg, gcoef, B);
  }
  ctc_finalize_kernel<<<1, 32, 0, st>>>(nll, tgt_lens, B, loss_out);
  return check_launch("ctc_loss", 4);
}

/* CTC best-path decode (BASELINE.json config 5; the reference has no CTC decode of its own):
 * out_ids int32 (B, L) = collapsed label sequence padded with -1, out_lens int32[B]. */
int sst_ctc_greedy(int logits_dtype, int B, int L, int C, int blank, const void* logits, int64_t ld, const int32_t* in_lens,
                   int32_t* out_ids, int32_t* out_lens, void* stream) {
  SST_REQUIRE(C >= 1 && blank < C && out_ids && out_lens, SST_E_ARG, "ctc_greedy: bad arguments");
  if (B <= 0 || L <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (logits_dtype == SST_F32) ctc_greedy_kernel<float><<<B, 256, 0, st>>>((const float*)logits, ld, in_lens, L, C, blank, out_ids, out_lens);
  else ctc_greedy_kernel<__nv_bfloat16><<<B, 256, 0, st>>>((const __nv_bfloat16*)logits, ld, in_lens, L, C, blank, out_ids, out_lens);
  return check_launch("ctc_greedy");
}

/* tokens int64: tokens[b*tok_stride_b + pos*tok_stride_pos] = arg-max_c logits[b * row_stride + c]; done[b] |= (that == eos);
 * n_done[0] = #done. */
int sst_greedy_pick(int logits_dtype, int B, int C, const void* logits, int64_t row_stride, int64_t* tokens, int64_t tok_stride_b,
                    int64_t tok_stride_pos, int pos, int eos, uint8_t* done, int32_t* n_done, void* stream) {
  SST_REQUIRE(logits && tokens && done && n_done && C >= 1 && pos >= 0, SST_E_ARG, "greedy_pick: bad arguments");
  if (B <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int threads = B >= 32 ? 1024 : 32 * B;
  if (logits_dtype == SST_F32)
    greedy_pick_kernel<float><<<1, threads, 0, st>>>((const float*)logits, row_stride, B, C, reinterpret_cast<long long*>(tokens), tok_stride_b, tok_stride_pos, pos, eos, done, n_done);
  else
    greedy_pick_kernel<__nv_bfloat16><<<1, threads, 0, st>>>((const __nv_bfloat16*)logits, row_stride, B, C, reinterpret_cast<long long*>(tokens),
                                                             tok_stride_b, tok_stride_pos, pos, eos, done, n_done);
  return check_launch("greedy_pick");
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

// Column-wise (per-channel) HBM-bound kernels of the convolutional front end and the bias gradients:
//   per-channel batch statistics, BatchNorm apply (+ReLU, + second normalised branch) and backward
//                                                      -- architecture.py:27,29,33,40-48 (training-mode BN)
//   column sums (bias gradients of the Linear layers)  -- transformer.py / architecture.py nn.Linear biases
//
// All of them walk a (rows, C) matrix with one thread per 8 consecutive channels, so the per-channel constants and partial
// sums live in registers.  That register footprint is what used to starve them: with loads landing in registers a thread
// could only keep 2-4 rows in flight and the kernels sat at ~45 % of HBM bandwidth, latency-bound.  Here every thread
// streams its own 16-byte pieces through shared memory with cp.async (LDGSTS): a ring of S stages of U rows per tensor, so
// (S-1)*U rows per tensor are in flight per thread whatever the register pressure -- ~100 KB per SM, enough to cover
// HBM latency at full bandwidth.  A thread only ever reads back what it copied itself, so the ring needs no block barrier.
#include "colstream.cuh"

namespace sst {

constexpr int COL_STAGES = 8;                          // ring depth: 7 steps in flight per thread

struct ColGeom {
  uint32_t ring, plane, stage_bytes; int nt;
  __device__ __forceinline__ uint32_t slot(uint32_t slot0, int u, int t) const { return slot0 + (uint32_t)(u * nt + t) * plane; }
};
template <typename T>
__device__ __forceinline__ ColGeom col_geom(void* ring_smem, int U, int nt) {
  ColGeom g;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
  g.plane = (uint32_t)nth * ColSlot<T>::BYTES;
  g.ring = (uint32_t)__cvta_generic_to_shared(ring_smem) + (uint32_t)tid * ColSlot<T>::BYTES;
  g.stage_bytes = (uint32_t)(U * nt) * g.plane;
  g.nt = nt;
  return g;
}

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

// Dynamic shared memory of these kernels: [reduction scratch | sign constants | ring]; sizes are passed by the launcher.
struct ColSmem { int red_bytes, sgn_bytes, nt; };

// ---- stats[0][c] += sum x, stats[1][c] += sum x^2 (double; fp32 partials over <= 64 rows) --------------------------
template <typename T>
__global__ void __launch_bounds__(512)
colstats_kernel(const T* __restrict__ x, long rows, int C, long ld, double* __restrict__ stats, long rows_per_block, ColSmem sm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_wait();                 // launched with launch_pdl (sst_common.cuh): nothing of global memory is touched before this
  pdl_trigger();
  double* dbuf = reinterpret_cast<double*>(smem_raw);
  constexpr int U = 2;
  const ColGeom cg = col_geom<T>(smem_raw + sm.red_bytes, U, sm.nt);
  const int c = threadIdx.x * 8, TY = blockDim.y;
  double s[8], q[8];
  float fs[8], fq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.0; q[j] = 0.0; fs[j] = 0.f; fq[j] = 0.f; }
  int pending = 0;
  stream_rows<U, COL_STAGES>((int)blockIdx.x * TY + (int)threadIdx.y, (int)rows, (int)gridDim.x * TY, cg.ring, cg.stage_bytes,
      [&](uint32_t a0, int u, int r, bool valid) { ColSlot<T>::issue(cg.slot(a0, u, 0), valid ? x + (long)r * ld + c : x, valid); },
      [&](uint32_t a0, int u, int) {
        float v[8];
        ColSlot<T>::read(cg.slot(a0, u, 0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { fs[j] += v[j]; fq[j] = fmaf(v[j], v[j], fq[j]); }
        if (++pending == 64) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s[j] += fs[j]; q[j] += fq[j]; fs[j] = 0.f; fq[j] = 0.f; }
          pending = 0;
        }
      });
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] += fs[j]; q[j] += fq[j]; }
  reduce_over_ty(s, dbuf);
  reduce_over_ty(q, dbuf);
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(stats + c + j, s[j]); atomicAdd(stats + C + c + j, q[j]); }
  }
}

// ---- out[c] += sum_rows x[r][c]  (bias gradients) --------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(512)
colsum_kernel(const T* __restrict__ x, long rows, int C, long ld, float* __restrict__ out, long rows_per_block, ColSmem sm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_wait();                 // launched with launch_pdl (sst_common.cuh): nothing of global memory is touched before this
  pdl_trigger();
  float* fbuf = reinterpret_cast<float*>(smem_raw);
  constexpr int U = 2;
  const ColGeom cg = col_geom<T>(smem_raw + sm.red_bytes, U, sm.nt);
  const int c = threadIdx.x * 8, TY = blockDim.y;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  stream_rows<U, COL_STAGES>((int)blockIdx.x * TY + (int)threadIdx.y, (int)rows, (int)gridDim.x * TY, cg.ring, cg.stage_bytes,
      [&](uint32_t a0, int u, int r, bool valid) { ColSlot<T>::issue(cg.slot(a0, u, 0), valid ? x + (long)r * ld + c : x, valid); },
      [&](uint32_t a0, int u, int) {
        float v[8];
        ColSlot<T>::read(cg.slot(a0, u, 0), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
      });
  reduce_over_ty(s, fbuf);
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c + j >= C) s[j] = 0.f;
    if (c + 8 <= C && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      red_add4(out + c, s[0], s[1], s[2], s[3]);
      red_add4(out + c + 4, s[4], s[5], s[6], s[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c + j < C) atomicAdd(out + c + j, s[j]);
    }
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
// Per-column affine constants (scale, shift) live in registers; the padded output rows are the streamed index.
template <typename T>
__global__ void __launch_bounds__(512)
bn_apply_kernel(BnBranch a, BnBranch b, int has_b, int relu, T* __restrict__ out, long n_chunks, int Tlen, int C, int lead,
                int trail, long prows_per_block, ColSmem sm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_wait();                 // launched with launch_pdl (sst_common.cuh): nothing of global memory is touched before this
  pdl_trigger();
  constexpr int U = 1;
  const ColGeom cg = col_geom<T>(smem_raw, U, sm.nt);
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const unsigned P = Tlen + lead + trail;
  const long prows = n_chunks * (long)P;
  const T* xa = reinterpret_cast<const T*>(a.x);
  const T* xb = reinterpret_cast<const T*>(b.x);
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
  auto row_of = [&](int prow, int& row) {                // padded row -> data row; false on a halo row
    const unsigned chunk = (unsigned)prow / P;
    const int t = (int)((unsigned)prow - chunk * P) - lead;
    row = (int)chunk * Tlen + t;
    return t >= 0 && t < Tlen;
  };
  stream_rows<U, COL_STAGES>((int)blockIdx.x * TY + (int)threadIdx.y, (int)prows, (int)gridDim.x * TY, cg.ring, cg.stage_bytes,
      [&](uint32_t a0, int u, int prow, bool valid) {
        int row = 0;
        const bool real = valid && row_of(prow, row);
        ColSlot<T>::issue(cg.slot(a0, u, 0), real ? xa + (long)row * a.ld + c : xa, real);
        if (has_b) ColSlot<T>::issue(cg.slot(a0, u, 1), real ? xb + (long)row * b.ld + c : xb, real);
      },
      [&](uint32_t a0, int u, int prow) {
        int row;
        float o[8];
        if (!row_of(prow, row)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = 0.f;
        } else {
          float va[8];
          ColSlot<T>::read(cg.slot(a0, u, 0), va);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(va[j] - ma[j], sa[j], ha[j]);
          if (has_b) {
            float vb[8];
            ColSlot<T>::read(cg.slot(a0, u, 1), vb);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += fmaf(vb[j] - mb[j], sb[j], hb[j]);
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
          }
        }
        Vec8<T>::store(out + (long)prow * C + c, o);
      });
}

// The ReLU mask of out = relu(bnA(xa) [+ bnB(xb)]) without reading `out` back: bn_apply_kernel's expression, evaluated
// again on the same inputs with the same operations, has the same sign bit for bit (and a positive fp32 value never rounds
// to a bf16 zero), so each backward pass reads one tensor less.  The four per-column constants (scale_a, shift_a, scale_b,
// shift_b) are kept in shared memory as sgn[4][C] -- in registers they would not fit beside the two branches' statistics.
struct BnSign {
  const float* sgn; int C;
  __device__ __forceinline__ void fill(float* dst, int C_, const BnBranch& a, const BnBranch& b, int has_b, int c) {   // + __syncthreads
    sgn = dst; C = C_;
    if (threadIdx.y != 0) return;
    float is[8], g[8], h[8];
    load8_f32(a.invstd + c, is); load8_f32(a.gamma + c, g); load8_f32(a.beta + c, h);
#pragma unroll
    for (int j = 0; j < 8; ++j) { dst[c + j] = is[j] * g[j]; dst[C + c + j] = h[j]; dst[2 * C + c + j] = 0.f; dst[3 * C + c + j] = 0.f; }
    if (has_b) {
      load8_f32(b.invstd + c, is); load8_f32(b.gamma + c, g); load8_f32(b.beta + c, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) { dst[2 * C + c + j] = is[j] * g[j]; dst[3 * C + c + j] = h[j]; }
    }
  }
  // bit j: out[c + j] > 0 given da = xa - mean_a, db = xb - mean_b
  __device__ __forceinline__ uint32_t positive8(int c, const float (&da)[8], const float (&db)[8], int has_b) const {
    float k[8], h[8], o[8];
    load8_f32(sgn + c, k); load8_f32(sgn + C + c, h);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(da[j], k[j], h[j]);
    if (has_b) {
      load8_f32(sgn + 2 * C + c, k); load8_f32(sgn + 3 * C + c, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += fmaf(db[j], k[j], h[j]);
    }
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) bits |= (o[j] > 0.f ? 1u : 0u) << j;
    return bits;
  }
};

// backward pass 1: g = dout * (out > 0);  red[0][c] += g, red[1][c] += g*xhat_a, red[2][c] += g*xhat_b   (double)
// y == nullptr with relu set: the mask is recomputed from xa / xb (BnSign).  Ring tensors: 0 dout, 1 xa, 2 xb, 3 y.
template <typename T>
__global__ void __launch_bounds__(512)
bn_bwd_reduce_kernel(const T* __restrict__ dout, long ld_dout, const T* __restrict__ y, int y_lead, int y_trail,
                     int relu, BnBranch a, BnBranch b, int has_b, long n_chunks, int Tlen, int C,
                     double* __restrict__ red, long rows_per_block, ColSmem sm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_wait();                 // launched with launch_pdl (sst_common.cuh): nothing of global memory is touched before this
  pdl_trigger();
  double* dbuf = reinterpret_cast<double*>(smem_raw);
  constexpr int U = 1;
  const ColGeom cg = col_geom<T>(smem_raw + sm.red_bytes + sm.sgn_bytes, U, sm.nt);
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const long rows = n_chunks * Tlen;
  const int Py = Tlen + y_lead + y_trail;
  const T* xa = reinterpret_cast<const T*>(a.x);
  const T* xb = reinterpret_cast<const T*>(b.x);
  const bool resign = relu && y == nullptr, use_y = relu && y != nullptr;
  BnSign sg;
  if (resign) {
    sg.fill(reinterpret_cast<float*>(smem_raw + sm.red_bytes), C, a, b, has_b, c);
    __syncthreads();
  }
  float ma[8], ia[8], mb[8], ib[8];
  load8_f32(a.mean + c, ma); load8_f32(a.invstd + c, ia);
  if (has_b) { load8_f32(b.mean + c, mb); load8_f32(b.invstd + c, ib); }
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { mb[j] = 0.f; ib[j] = 0.f; }
  }
  // fp32 partials per thread (the launch keeps a thread's share to a few hundred rows), double from the block reduction on
  float f0[8], f1[8], f2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f0[j] = f1[j] = f2[j] = 0.f; }
  stream_rows<U, COL_STAGES>((int)blockIdx.x * TY + (int)threadIdx.y, (int)rows, (int)gridDim.x * TY, cg.ring, cg.stage_bytes,
      [&](uint32_t a0, int u, int r, bool valid) {
        ColSlot<T>::issue(cg.slot(a0, u, 0), valid ? dout + (long)r * ld_dout + c : dout, valid);
        ColSlot<T>::issue(cg.slot(a0, u, 1), valid ? xa + (long)r * a.ld + c : xa, valid);
        if (has_b) ColSlot<T>::issue(cg.slot(a0, u, 2), valid ? xb + (long)r * b.ld + c : xb, valid);
        if (use_y) {
          const int chunk = r / Tlen;
          const long prow = (long)chunk * Py + (r - chunk * Tlen) + y_lead;
          ColSlot<T>::issue(cg.slot(a0, u, has_b ? 3 : 2), valid ? y + prow * C + c : y, valid);
        }
      },
      [&](uint32_t a0, int u, int) {
        float g[8], da[8], db[8];
        ColSlot<T>::read(cg.slot(a0, u, 0), g);
        ColSlot<T>::read(cg.slot(a0, u, 1), da);
#pragma unroll
        for (int j = 0; j < 8; ++j) { da[j] -= ma[j]; db[j] = 0.f; }
        if (has_b) {
          ColSlot<T>::read(cg.slot(a0, u, 2), db);
#pragma unroll
          for (int j = 0; j < 8; ++j) db[j] -= mb[j];
        }
        uint32_t pos = 0xffu;
        if (resign) pos = sg.positive8(c, da, db, has_b);
        else if (use_y) {
          float yv[8];
          ColSlot<T>::read(cg.slot(a0, u, has_b ? 3 : 2), yv);
          pos = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) pos |= (yv[j] > 0.f ? 1u : 0u) << j;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = ((pos >> j) & 1u) ? g[j] : 0.f;
          f0[j] += gg;
          f1[j] = fmaf(gg, da[j] * ia[j], f1[j]);
          if (has_b) f2[j] = fmaf(gg, db[j] * ib[j], f2[j]);
        }
      });
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

// backward pass 2: dx = gamma*invstd*(g - sum_g/N - xhat*sum_gxhat/N) for each branch; block 0 adds dgamma/dbeta.
// Per column the expression is affine in (g, x - mean): constants in registers.  The streamed index is the DATA row; each
// branch writes it at its own padded position and the thread that owns a chunk's first row also zeroes that chunk's halo
// rows (lead + trail of them per branch).
template <typename T>
__global__ void __launch_bounds__(512)
bn_bwd_apply_kernel(const T* __restrict__ dout, long ld_dout, const T* __restrict__ y, int y_lead, int y_trail,
                    int relu, BnBranch a, BnBranch b, int has_b, BnGradOut ga, BnGradOut gb, long n_chunks,
                    int Tlen, int C, const double* __restrict__ red, long rows_per_block, ColSmem sm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_wait();                 // launched with launch_pdl (sst_common.cuh): nothing of global memory is touched before this
  pdl_trigger();
  constexpr int U = 1;
  const ColGeom cg = col_geom<T>(smem_raw + sm.sgn_bytes, U, sm.nt);
  const int c = threadIdx.x * 8, TY = blockDim.y;
  const long rows = n_chunks * Tlen;
  const double invN = 1.0 / (double)rows;
  const int Py = Tlen + y_lead + y_trail;
  const T* xa = reinterpret_cast<const T*>(a.x);
  const T* xb = reinterpret_cast<const T*>(b.x);
  const bool resign = relu && y == nullptr, use_y = relu && y != nullptr;
  if (blockIdx.x == 0) {
    for (int cc = threadIdx.y * blockDim.x + threadIdx.x; cc < C; cc += blockDim.x * blockDim.y) {
      ga.dgamma[cc] += (float)red[C + cc];
      ga.dbeta[cc] += (float)red[cc];
      if (has_b) { gb.dgamma[cc] += (float)red[2 * C + cc]; gb.dbeta[cc] += (float)red[cc]; }
    }
  }
  BnSign sg;
  if (resign) {
    sg.fill(reinterpret_cast<float*>(smem_raw), C, a, b, has_b, c);
    __syncthreads();
  }
  float A[2][8], Bx[2][8], K[2][8], Mn[2][8];           // dx = A*g + K - (x - mean)*Bx
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    if (br == 1 && !has_b) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { A[1][j] = Bx[1][j] = K[1][j] = Mn[1][j] = 0.f; }
      break;
    }
    const BnBranch& bx = br == 0 ? a : b;
    float m[8], is[8], gm[8];
    load8_f32(bx.mean + c, m); load8_f32(bx.invstd + c, is); load8_f32(bx.gamma + c, gm);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mg = (float)(red[c + j] * invN);
      const float mgx = (float)(red[(br + 1) * C + c + j] * invN);
      const float k1 = gm[j] * is[j];
      A[br][j] = k1;
      Bx[br][j] = k1 * mgx * is[j];
      K[br][j] = -k1 * mg;
      Mn[br][j] = m[j];
    }
  }
  const int Pa = Tlen + ga.lead + ga.trail, Pb = Tlen + gb.lead + gb.trail;
  T* dxa = reinterpret_cast<T*>(ga.dx);
  T* dxb = reinterpret_cast<T*>(gb.dx);
  stream_rows<U, COL_STAGES>((int)blockIdx.x * TY + (int)threadIdx.y, (int)rows, (int)gridDim.x * TY, cg.ring, cg.stage_bytes,
      [&](uint32_t a0, int u, int r, bool valid) {
        ColSlot<T>::issue(cg.slot(a0, u, 0), valid ? dout + (long)r * ld_dout + c : dout, valid);
        ColSlot<T>::issue(cg.slot(a0, u, 1), valid ? xa + (long)r * a.ld + c : xa, valid);
        if (has_b) ColSlot<T>::issue(cg.slot(a0, u, 2), valid ? xb + (long)r * b.ld + c : xb, valid);
        if (use_y) {
          const int chunk = r / Tlen;
          ColSlot<T>::issue(cg.slot(a0, u, has_b ? 3 : 2), valid ? y + ((long)chunk * Py + (r - chunk * Tlen) + y_lead) * C + c : y, valid);
        }
      },
      [&](uint32_t a0, int u, int r) {
        const long chunk = r / Tlen;
        const int t = r - (int)chunk * Tlen;
        float g[8], da[8], db[8];
        ColSlot<T>::read(cg.slot(a0, u, 0), g);
        ColSlot<T>::read(cg.slot(a0, u, 1), da);
#pragma unroll
        for (int j = 0; j < 8; ++j) { da[j] -= Mn[0][j]; db[j] = 0.f; }
        if (has_b) {
          ColSlot<T>::read(cg.slot(a0, u, 2), db);
#pragma unroll
          for (int j = 0; j < 8; ++j) db[j] -= Mn[1][j];
        }
        if (relu) {
          uint32_t pos;
          if (resign) pos = sg.positive8(c, da, db, has_b);
          else {
            float yv[8];
            ColSlot<T>::read(cg.slot(a0, u, has_b ? 3 : 2), yv);
            pos = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) pos |= (yv[j] > 0.f ? 1u : 0u) << j;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = ((pos >> j) & 1u) ? g[j] : 0.f;
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(-da[j], Bx[0][j], fmaf(A[0][j], g[j], K[0][j]));
        Vec8<T>::store(dxa + (chunk * Pa + t + ga.lead) * ga.ld + c, o);
        if (has_b) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(-db[j], Bx[1][j], fmaf(A[1][j], g[j], K[1][j]));
          Vec8<T>::store(dxb + (chunk * Pb + t + gb.lead) * gb.ld + c, o);
        }
        if (t == 0) {                                    // this thread zeroes the chunk's halo rows (its 8 columns)
          float z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = 0.f;
          for (int h = 0; h < ga.lead; ++h) Vec8<T>::store(dxa + (chunk * Pa + h) * ga.ld + c, z);
          for (int h = 0; h < ga.trail; ++h) Vec8<T>::store(dxa + (chunk * Pa + ga.lead + Tlen + h) * ga.ld + c, z);
          if (has_b) {
            for (int h = 0; h < gb.lead; ++h) Vec8<T>::store(dxb + (chunk * Pb + h) * gb.ld + c, z);
            for (int h = 0; h < gb.trail; ++h) Vec8<T>::store(dxb + (chunk * Pb + gb.lead + Tlen + h) * gb.ld + c, z);
          }
        }
      });
}

// ---- launch shapes ---------------------------------------------------------------------------------------------------
// block (tx, ty): tx = C/8 column threads, ty rows as large as `max_threads` allows; ONE block per SM (the ring is what
// hides latency, not occupancy), contiguous row ranges of `rpb` rows per block; as many ring stages (2..4) as fit.
struct ColLaunch { dim3 block; int grid; long rpb; ColSmem sm; size_t smem; };
static ColLaunch col_launch(long rows, int tx, size_t red_elem_bytes, size_t sgn_bytes, int U, int nt, size_t elem_bytes,
                            long max_rows_per_thread = 1L << 40) {
  const size_t per_thread = (size_t)COL_STAGES * U * nt * 8 * elem_bytes;        // ring bytes per thread
  int ty = 512 / tx; if (ty < 1) ty = 1; if (ty > 16) ty = 16;
  int blocks_per_sm = 512 / (tx * ty); if (blocks_per_sm < 1) blocks_per_sm = 1; if (blocks_per_sm > 4) blocks_per_sm = 4;
  const size_t budget = (size_t)224 * 1024 / blocks_per_sm;
  auto total = [&](int ty_) { return (size_t)(ty_ - 1) * tx * 8 * red_elem_bytes + sgn_bytes + per_thread * tx * ty_; };
  while (ty > 1 && total(ty) > budget) --ty;
  long nblk = (long)num_sms() * blocks_per_sm;
  const long min_blk = (rows + max_rows_per_thread * ty - 1) / (max_rows_per_thread * ty);   // fp32 partials stay short
  if (nblk < min_blk) nblk = min_blk;
  const long max_blk = (rows + ty - 1) / ty;
  if (nblk > max_blk) nblk = max_blk;
  ColLaunch cl;
  cl.block = dim3(tx, ty);
  cl.grid = (int)(nblk > 0 ? nblk : 1);
  cl.rpb = 0;
  cl.sm.red_bytes = (int)((size_t)(ty - 1) * tx * 8 * red_elem_bytes);
  cl.sm.sgn_bytes = (int)sgn_bytes;
  cl.sm.nt = nt;
  cl.smem = total(ty);
  return cl;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  SST_REQUIRE(bytes <= 227 * 1024, SST_E_ARG, "column kernel needs %zu bytes of shared memory", bytes);
  return SST_OK;
}

}  // namespace sst

using namespace sst;

// opt in to the large dynamic shared memory once per instantiation, then launch
#define SST_COL_LAUNCH(KERNEL, T_, ...)                                                    \
  do {                                                                                     \
    int rc_ = opt_in_smem(KERNEL<T_>, cl.smem);                                            \
    if (rc_) return rc_;                                                                   \
    cudaError_t le_ = launch_pdl(KERNEL<T_>, dim3(cl.grid), dim3(cl.block), cl.smem, st, __VA_ARGS__);   \
    SST_REQUIRE(le_ == cudaSuccess, SST_E_LAUNCH, #KERNEL " launch: %s", cudaGetErrorString(le_)); \
  } while (0)

extern "C" {

int sst_colstats(int dtype, const void* x, int64_t rows, int C, int64_t ld, double* stats, void* stream) {
  SST_REQUIRE(C % 8 == 0 && C / 8 <= 512 && ld % 8 == 0, SST_E_ARG, "colstats: C=%d, ld=%ld must be multiples of 8", C, (long)ld);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st);
  SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
  if (rows <= 0) return SST_OK;
  SST_REQUIRE(rows < (1L << 31), SST_E_ARG, "colstats: too many rows");
  const size_t esz = dtype == SST_F32 ? 4 : 2;
  const ColLaunch cl = col_launch(rows, C / 8, sizeof(double), 0, 2, 1, esz);
  if (dtype == SST_F32) SST_COL_LAUNCH(colstats_kernel, float, (const float*)x, rows, C, ld, stats, cl.rpb, cl.sm);
  else SST_COL_LAUNCH(colstats_kernel, __nv_bfloat16, (const __nv_bfloat16*)x, rows, C, ld, stats, cl.rpb, cl.sm);
  return check_launch("colstats");
}

int sst_colsum_accum(int dtype, const void* x, int64_t rows, int C, int64_t ld, float* out, void* stream) {
  SST_REQUIRE(ld % 8 == 0 && C <= ld && (C + 7) / 8 <= 512, SST_E_ARG, "colsum: pitch %ld must be a multiple of 8 and >= C=%d", (long)ld, C);
  if (rows <= 0) return SST_OK;
  SST_REQUIRE(rows < (1L << 31), SST_E_ARG, "colsum: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int tx = (C + 7) / 8;
  const size_t esz = dtype == SST_F32 ? 4 : 2;
  const ColLaunch cl = col_launch(rows, tx, sizeof(float), 0, 2, 1, esz);
  if (dtype == SST_F32) SST_COL_LAUNCH(colsum_kernel, float, (const float*)x, rows, C, ld, out, cl.rpb, cl.sm);
  else SST_COL_LAUNCH(colsum_kernel, __nv_bfloat16, (const __nv_bfloat16*)x, rows, C, ld, out, cl.rpb, cl.sm);
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
  const long prows = n_chunks * (long)(T + lead + trail);
  if (prows <= 0) return SST_OK;
  SST_REQUIRE(prows < (1L << 31), SST_E_ARG, "bn_apply: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SST_REQUIRE(C / 8 <= 512, SST_E_ARG, "bn_apply: C=%d too wide", C);
  const size_t esz = dtype == SST_F32 ? 4 : 2;
  const int has_b = xb != nullptr;
  const ColLaunch cl = col_launch(prows, C / 8, 0, 0, 1, 1 + has_b, esz);
  if (dtype == SST_F32) SST_COL_LAUNCH(bn_apply_kernel, float, a, b, has_b, relu, (float*)out, n_chunks, T, C, lead, trail, cl.rpb, cl.sm);
  else SST_COL_LAUNCH(bn_apply_kernel, __nv_bfloat16, a, b, has_b, relu, (__nv_bfloat16*)out, n_chunks, T, C, lead, trail, cl.rpb, cl.sm);
  return check_launch("bn_apply");
}

/* Backward of out = act(bnA(xa) [+ bnB(xb)]).  `red` is a caller-provided double[3*C] scratch. */
int sst_bn_bwd(int dtype, int64_t n_chunks, int T, int C, const void* dout, int64_t ld_dout, const void* y, int y_lead, int y_trail,
               int relu, const void* xa, int64_t lda, const float* mean_a, const float* invstd_a, const float* gamma_a,
               const float* beta_a, void* dxa, int64_t ld_dxa, int lead_a, int trail_a, float* dgamma_a, float* dbeta_a,
               const void* xb, int64_t ldb, const float* mean_b, const float* invstd_b, const float* gamma_b,
               const float* beta_b, void* dxb, int64_t ld_dxb, int lead_b, int trail_b, float* dgamma_b, float* dbeta_b,
               double* red, void* stream) {
  SST_REQUIRE(C % 8 == 0 && C / 8 <= 512, SST_E_ARG, "bn_bwd: C=%d must be a multiple of 8", C);
  SST_REQUIRE(!relu || y != nullptr || (beta_a != nullptr && (xb == nullptr || beta_b != nullptr)), SST_E_ARG,
              "bn_bwd: the ReLU mask needs either the forward output y or the BN shifts (beta) to recompute its sign");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int has_b = xb != nullptr;
  BnBranch a{xa, lda, mean_a, invstd_a, gamma_a, beta_a};
  BnBranch b{xb, ldb, mean_b, invstd_b, gamma_b, beta_b};
  BnGradOut ga{dxa, ld_dxa, lead_a, trail_a, dgamma_a, dbeta_a};
  BnGradOut gb{dxb, ld_dxb, lead_b, trail_b, dgamma_b, dbeta_b};
  cudaError_t e = cudaMemsetAsync(red, 0, sizeof(double) * 3 * C, st);
  SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
  const long rows = n_chunks * T;
  if (rows <= 0) return SST_OK;
  SST_REQUIRE(rows < (1L << 31) && n_chunks * (long)(T + 4) < (1L << 31), SST_E_ARG, "bn_bwd: too many rows");
  const size_t esz = dtype == SST_F32 ? 4 : 2;
  const size_t sgn_bytes = (relu && y == nullptr) ? (size_t)4 * C * sizeof(float) : 0;
  const int nt = 2 + has_b + ((relu && y != nullptr) ? 1 : 0);      // dout, xa, [xb], [y]
  {
    const ColLaunch cl = col_launch(rows, C / 8, sizeof(double), sgn_bytes, 1, nt, esz, 512);    // <= 512 rows per fp32 partial
    if (dtype == SST_F32)
      SST_COL_LAUNCH(bn_bwd_reduce_kernel, float, (const float*)dout, ld_dout, (const float*)y, y_lead, y_trail, relu, a, b, has_b,
                     n_chunks, T, C, red, cl.rpb, cl.sm);
    else
      SST_COL_LAUNCH(bn_bwd_reduce_kernel, __nv_bfloat16, (const __nv_bfloat16*)dout, ld_dout, (const __nv_bfloat16*)y, y_lead, y_trail,
                     relu, a, b, has_b, n_chunks, T, C, red, cl.rpb, cl.sm);
  }
  {
    const ColLaunch cl = col_launch(rows, C / 8, 0, sgn_bytes, 1, nt, esz);
    if (dtype == SST_F32)
      SST_COL_LAUNCH(bn_bwd_apply_kernel, float, (const float*)dout, ld_dout, (const float*)y, y_lead, y_trail, relu, a, b, has_b, ga, gb,
                     n_chunks, T, C, red, cl.rpb, cl.sm);
    else
      SST_COL_LAUNCH(bn_bwd_apply_kernel, __nv_bfloat16, (const __nv_bfloat16*)dout, ld_dout, (const __nv_bfloat16*)y, y_lead, y_trail,
                     relu, a, b, has_b, ga, gb, n_chunks, T, C, red, cl.rpb, cl.sm);
  }
  return check_launch("bn_bwd", 2);
}

}  // extern "C"

// Small data-movement / optimiser kernels around the hot path:
//   shift_left        in-place time shift augmentation               (architecture.py:104-108)
//   im2col_first      raw 8-channel EMG -> (rows, 32) patch matrix so the first ResBlock's k=3 s=2 conv and its 1x1
//                     s=2 residual conv run as ONE tensor-core GEMM   (architecture.py:26,32,55)
//   gather_rows_pad / scatter_rows   decollate_tensor + pad_sequence(42.0) and its adjoint
//                                                                    (data_utils.py:176-185, architecture.py:116-117)
//   embed_posenc fwd/bwd   nn.Embedding(padding_idx) + batch-indexed sinusoid/768 + dropout
//                                                                    (architecture.py:126-127, transformer.py:431-435, Q10)
//   permute3_cast     weight packing / gradient unpacking between the reference's parameter layouts and GEMM operands
//   adamw             fused AdamW over a flat fp32 parameter buffer  (recognition_model.py:293; torch defaults)
#include "vec.cuh"

namespace sst {

// one block per chunk, in place: every thread first reads its (shifted) elements, the block synchronises, then writes
constexpr int SHIFT_PER_THREAD = 64;
__global__ void __launch_bounds__(256)
shift_left_kernel(float* __restrict__ x, long n_chunks, int Tlen, int Cc, int r) {
  float* p = x + (long)blockIdx.x * Tlen * Cc;
  const int n = Tlen * Cc, lim = (Tlen - r) * Cc, off = r * Cc;
  for (int base = 0; base < n; base += 256 * SHIFT_PER_THREAD) {     // chunks longer than 16 K elements: sequential passes
    float v[SHIFT_PER_THREAD];                                       // pass k only reads elements >= its own range: safe
#pragma unroll
    for (int k = 0; k < SHIFT_PER_THREAD; ++k) {
      const int e = base + k * 256 + threadIdx.x;
      v[k] = (e < lim) ? p[e + off] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SHIFT_PER_THREAD; ++k) {
      const int e = base + k * 256 + threadIdx.x;
      if (e < n) p[e] = v[k];
    }
    __syncthreads();
  }
}

// x: (n, Tin, 8) fp32 -> col: (n*Tin/2, 32):  [k*8 + c] = x[2t+k-1][c] (zero outside), [24 + c] = x[2t][c]
template <typename T>
__global__ void im2col_first_kernel(const float* __restrict__ x, T* __restrict__ col, long n_chunks, int Tin) {
  const int To = Tin / 2;
  const long total = n_chunks * To * 4;      // one thread per (row, 8-wide group)
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int grp = (int)(i & 3);
    const long row = i >> 2;
    const long chunk = row / To;
    const int t = (int)(row - chunk * To);
    const int tin = grp < 3 ? 2 * t + grp - 1 : 2 * t;
    float v[8];
    if (tin < 0 || tin >= Tin) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    } else {
      load8_f32(x + (chunk * Tin + tin) * 8, v);
    }
    Vec8<T>::store(col + row * 32 + grp * 8, v);
  }
}

// out[(b, t), :] = t < lens[b] ? in[(offs[b] + t), :] : fill
template <typename T>
__global__ void gather_rows_pad_kernel(const T* __restrict__ in, T* __restrict__ out, const long* __restrict__ offs,
                                       const int* __restrict__ lens, int B, int Lmax, int D, float fill) {
  const int dv = D / 8;
  const long total = (long)B * Lmax * dv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv) * 8;
    const long row = i / dv;
    const int b = (int)(row / Lmax), t = (int)(row % Lmax);
    float v[8];
    if (t < lens[b]) Vec8<T>::load(in + (offs[b] + t) * D + c, v);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fill;
    }
    Vec8<T>::store(out + row * D + c, v);
  }
}

// adjoint: din[(offs[b] + t), :] = dout[(b, t), :] for t < lens[b]   (din pre-zeroed by the caller)
template <typename T>
__global__ void scatter_rows_kernel(const T* __restrict__ dout, T* __restrict__ din, const long* __restrict__ offs,
                                    const int* __restrict__ lens, int B, int Lmax, int D) {
  const int dv = D / 8;
  const long total = (long)B * Lmax * dv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv) * 8;
    const long row = i / dv;
    const int b = (int)(row / Lmax), t = (int)(row % Lmax);
    if (t < lens[b]) {
      float v[8];
      Vec8<T>::load(dout + row * D + c, v);
      Vec8<T>::store(din + (offs[b] + t) * D + c, v);
    }
  }
}

// out[(b,s), :] = dropout( W[y[b,s], :] + pe[b, :] / D )
template <typename T>
__global__ void embed_posenc_kernel(const long* __restrict__ y, const float* __restrict__ W, const float* __restrict__ pe,
                                    T* __restrict__ out, int B, int S, int D, uint32_t thr, float dscale, unsigned long long seed,
                                    const unsigned long long* __restrict__ salt) {
  if (thr) seed = salted(seed, salt);
  const int dv = D / 8;
  const long total = (long)B * S * dv;
  const float invD = 1.f / (float)D;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv) * 8;
    const long row = i / dv;
    const int b = (int)(row / S);
    float w[8], p[8], o[8];
    load8_f32(W + y[row] * D + c, w);
    load8_f32(pe + (long)b * D + c, p);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = w[j] + invD * p[j];
    if (thr) {
      bool k[8];
      keep8(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = k[j] ? o[j] * dscale : 0.f;
    }
    Vec8<T>::store(out + row * D + c, o);
  }
}

// dW[y[row], :] += keep * dout[row, :]   (rows with y == pad_idx skipped: nn.Embedding padding_idx)
template <typename T>
__global__ void embed_bwd_kernel(const long* __restrict__ y, const T* __restrict__ dout, float* __restrict__ dW, int B, int S, int D,
                                 int pad_idx, uint32_t thr, float dscale, unsigned long long seed,
                                 const unsigned long long* __restrict__ salt) {
  if (thr) seed = salted(seed, salt);
  const int dv = D / 8;
  const long total = (long)B * S * dv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % dv) * 8;
    const long row = i / dv;
    const long tok = y[row];
    if (tok == pad_idx) continue;
    float g[8];
    Vec8<T>::load(dout + row * D + c, g);
    if (thr) {
      bool k[8];
      keep8(seed, (unsigned long long)row * D + c, thr, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = k[j] ? g[j] * dscale : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dW + tok * D + c + j, g[j]);
  }
}

// out[i*o0 + j*o1 + k*o2] (+)= in[i*s0 + j*s1 + k*s2]   for (i,j,k) in d0 x d1 x d2
template <typename TI, typename TO>
__global__ void permute3_cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long d0, long d1, long d2, long s0, long s1,
                                     long s2, long o0, long o1, long o2, int accumulate) {
  const long total = d0 * d1 * d2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long k = i % d2, j = (i / d2) % d1, a = i / (d2 * d1);
    const long oi = a * o0 + j * o1 + k * o2;
    float v = to_f32(in[a * s0 + j * s1 + k * s2]);
    if (accumulate) v += to_f32(out[oi]);
    out[oi] = from_f32<TO>(v);
  }
}

// Same mapping for the transposing cases (unit input stride along one of (j, k), unit output stride along the other):
// a 32 x 32 tile of (j, k) goes through shared memory so that both the global read and the global write are coalesced.
// IN_J: the input is contiguous along j (s1 == 1) and the output along k (o2 == 1); otherwise the reverse.
template <typename TI, typename TO, bool IN_J>
__global__ void __launch_bounds__(256)
permute3_tiled_kernel(const TI* __restrict__ in, TO* __restrict__ out, long d1, long d2, long s0, long s1, long s2, long o0, long o1,
                      long o2, int accumulate) {
  __shared__ float tile[32][33];
  const long a = blockIdx.z, j0 = (long)blockIdx.y * 32, k0 = (long)blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    // read: tx runs along the input's contiguous dim
    const long j = IN_J ? j0 + tx : j0 + r, k = IN_J ? k0 + r : k0 + tx;
    if (j < d1 && k < d2) tile[IN_J ? r : tx][IN_J ? tx : r] = to_f32(in[a * s0 + j * s1 + k * s2]);   // tile[k - k0][j - j0] when IN_J
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    // write: tx runs along the output's contiguous dim
    const long j = IN_J ? j0 + r : j0 + tx, k = IN_J ? k0 + tx : k0 + r;
    if (j < d1 && k < d2) {
      const long oi = a * o0 + j * o1 + k * o2;
      float v = tile[IN_J ? tx : r][IN_J ? r : tx];
      if (accumulate) v += to_f32(out[oi]);
      out[oi] = from_f32<TO>(v);
    }
  }
}

// ---- batched form: one launch walks a table of permutes (the whole weight repack of an optimizer step) -------------
// Block b belongs to the item whose [first_block, first_block + nblocks) contains it (binary search over <= a few hundred
// items).  mode 0: flat grid-stride over the item's elements with its own blocks; 1 / 2: the 32 x 32 tiled transposes
// (input contiguous along j / along k), block -> (a, j tile, k tile).
__device__ __forceinline__ float ld_any(const void* p, long i, int dtype) { return ld_as_f32(p, i, dtype); }
__device__ __forceinline__ void st_any(void* p, long i, int dtype, float v) { st_from_f32(p, i, dtype, v); }

// modes 3 / 4: 64 x 64 tiles, 16-byte global accesses on both sides, bf16 output (the big weight transposes of the repack: the
// 32 x 32 scalar tiles above move 2 bytes per thread and instruction and pay the item lookup per 1024 elements).
// c = the input's contiguous dimension (j in mode 3, k in mode 4), r = the other one, which is the output's contiguous dimension.
// Shared tile: row r, 16-byte chunk (c / 8) XOR (r / 8 % 8) -- both the vector stores of the read phase and the 2-byte column
// gathers of the write phase are bank-conflict free.
constexpr int PT_TILES_PER_BLOCK = 4;
template <typename TI>
__device__ __forceinline__ void permute_tile64(const SstPermuteItem& it, bool in_j, long a, long j0, long k0, uint16_t (*tile)[64]) {
  const TI* in = reinterpret_cast<const TI*>(it.in);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(it.out);
  const long dC = in_j ? it.d1 : it.d2, dR = in_j ? it.d2 : it.d1;
  const long c0 = in_j ? j0 : k0, r0 = in_j ? k0 : j0;
  const long s_r = in_j ? it.s2 : it.s1, o_c = in_j ? it.o1 : it.o2;
  const int t = threadIdx.x;
  {
    const int cv = t & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rl = (t >> 3) + 32 * h;
      const long r = r0 + rl, c = c0 + 8 * cv;
      uint4 w = make_uint4(0u, 0u, 0u, 0u);
      if (r < dR && c < dC) {
        float v[8];
        Vec8<TI>::load(in + a * it.s0 + r * s_r + c, v);
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      }
      *reinterpret_cast<uint4*>(&tile[rl][((cv ^ (rl >> 3)) & 7) * 8]) = w;
    }
  }
  __syncthreads();
  {
    const int rv = t & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cl = (t >> 3) + 32 * h;
      const long c = c0 + cl, r = r0 + 8 * rv;
      if (c < dC && r < dR) {
        uint4 w;
        uint16_t* hw = reinterpret_cast<uint16_t*>(&w);
#pragma unroll
        for (int e = 0; e < 8; ++e) hw[e] = tile[8 * rv + e][(((cl >> 3) ^ rv) & 7) * 8 + (cl & 7)];
        *reinterpret_cast<uint4*>(out + a * it.o0 + c * o_c + r) = w;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
permute3_batch_kernel(const SstPermuteItem* __restrict__ items, int n_items) {
  __shared__ __align__(16) float tile[32][33];
  __shared__ __align__(16) uint16_t tile64[64][64];
  __shared__ int s_item;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;
    const int b = (int)blockIdx.x;
    while (lo < hi) {                                       // last item with first_block <= b
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].first_block <= b) lo = mid; else hi = mid - 1;
    }
    s_item = lo;
  }
  __syncthreads();
  const SstPermuteItem it = items[s_item];
  const int lb = (int)blockIdx.x - it.first_block;
  if (it.mode == 0) {
    const long total = it.d0 * it.d1 * it.d2;
    for (long i = (long)lb * 256 + threadIdx.x; i < total; i += (long)it.nblocks * 256) {
      const long k = i % it.d2, j = (i / it.d2) % it.d1, a = i / (it.d2 * it.d1);
      const long oi = a * it.o0 + j * it.o1 + k * it.o2;
      float v = ld_any(it.in, a * it.s0 + j * it.s1 + k * it.s2, it.in_dtype);
      if (it.accumulate) v += ld_any(it.out, oi, it.out_dtype);
      st_any(it.out, oi, it.out_dtype, v);
    }
    return;
  }
  if (it.mode >= 3) {
    const bool inj = it.mode == 3;
    const long tk = (it.d2 + 63) / 64, tj = (it.d1 + 63) / 64, tiles = tk * tj * it.d0;
    for (long tl = (long)lb * PT_TILES_PER_BLOCK; tl < tiles && tl < (long)(lb + 1) * PT_TILES_PER_BLOCK; ++tl) {
      const long a = tl / (tk * tj), rem = tl - a * tk * tj;
      const long j0 = (rem / tk) * 64, k0 = (rem % tk) * 64;
      if (it.in_dtype == SST_F32) permute_tile64<float>(it, inj, a, j0, k0, tile64);
      else permute_tile64<__nv_bfloat16>(it, inj, a, j0, k0, tile64);
    }
    return;
  }
  const bool in_j = it.mode == 1;
  const int bx = (int)((it.d2 + 31) / 32), by = (int)((it.d1 + 31) / 32);
  const long a = lb / (bx * by);
  const int rem = lb - (int)a * bx * by;
  const long j0 = (long)(rem / bx) * 32, k0 = (long)(rem % bx) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const long j = in_j ? j0 + tx : j0 + r, k = in_j ? k0 + r : k0 + tx;
    if (j < it.d1 && k < it.d2) tile[in_j ? r : tx][in_j ? tx : r] = ld_any(it.in, a * it.s0 + j * it.s1 + k * it.s2, it.in_dtype);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const long j = in_j ? j0 + r : j0 + tx, k = in_j ? k0 + tx : k0 + r;
    if (j < it.d1 && k < it.d2) {
      const long oi = a * it.o0 + j * it.o1 + k * it.o2;
      float v = tile[in_j ? tx : r][in_j ? r : tx];
      if (it.accumulate) v += ld_any(it.out, oi, it.out_dtype);
      st_any(it.out, oi, it.out_dtype, v);
    }
  }
}

// hyper != nullptr: (lr, bias correction 1, sqrt of bias correction 2) come from device memory -- the values a captured graph
// replays must not be baked into it (sst_adamw_dev)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                             float lr, float beta1, float beta2, float eps, float wd, float bc1, float sqrt_bc2,
                             __nv_bfloat16* __restrict__ p_bf16, const float* __restrict__ hyper) {
  if (hyper != nullptr) { lr = hyper[0]; bc1 = hyper[1]; sqrt_bc2 = hyper[2]; }
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long)gridDim.x * blockDim.x * 4) {
    if (i + 4 <= n) {
      float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        P[j] *= (1.f - lr * wd);
        M[j] = beta1 * M[j] + (1.f - beta1) * G[j];
        V[j] = beta2 * V[j] + (1.f - beta2) * G[j] * G[j];
        const float denom = sqrtf(V[j]) / sqrt_bc2 + eps;
        P[j] -= (lr / bc1) * (M[j] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
      if (p_bf16 != nullptr) {        // bf16 shadow of the updated parameters: the tensor-core GEMM operands, no separate cast pass
        uint2 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
        o2[0] = __floats2bfloat162_rn(pp.x, pp.y);
        o2[1] = __floats2bfloat162_rn(pp.z, pp.w);
        *reinterpret_cast<uint2*>(p_bf16 + i) = o;
      }
    } else {
      for (long e = i; e < n; ++e) {
        float P = p[e] * (1.f - lr * wd);
        float M = beta1 * m[e] + (1.f - beta1) * g[e];
        float V = beta2 * v[e] + (1.f - beta2) * g[e] * g[e];
        P -= (lr / bc1) * (M / (sqrtf(V) / sqrt_bc2 + eps));
        p[e] = P; m[e] = M; v[e] = V;
        if (p_bf16 != nullptr) p_bf16[e] = __float2bfloat16_rn(P);
      }
    }
  }
}

// ---- exact (erf) GELU + dropout, forward and backward: HBM-bound, one 16-byte vector of 8 elements per thread and trip ------------
// y = keep ? gelu(x) / (1 - p) : 0 with gelu(x) = x/2 (1 + erf(x / sqrt 2)) (torch's default F.gelu); keep decision of element
// (row, col) = 16-bit Philox lane of index row*cols + col (philox_keep16: one Philox block per 8-element vector).
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * __expf(-0.5f * x * x);
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
gelu_dropout_kernel(const T* __restrict__ x, long ldx, const T* __restrict__ dy, long lddy, T* __restrict__ out, long ldo, long rows, int cols,
                    uint32_t thr16, float dscale, unsigned long long seed, const unsigned long long* __restrict__ salt) {
  if (thr16) seed = salted(seed, salt);
  const int vpr = cols >> 3;                                          // vectors per row
  const long total = rows * vpr;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long)gridDim.x * blockDim.x) {
    const long r = v / vpr;
    const int c = (int)(v - r * vpr) << 3;
    float a[8], g[8];
    Vec8<T>::load(x + r * ldx + c, a);
    if (BWD) Vec8<T>::load(dy + r * lddy + c, g);
    Philox4 rn;
    if (thr16) rn = philox4x32(seed, (unsigned long long)(r * cols + c) >> 3);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float k = (thr16 == 0 || philox_lane16(rn, e) >= thr16) ? dscale : 0.f;
      a[e] = BWD ? g[e] * gelu_erf_grad(a[e]) * k : gelu_erf(a[e]) * k;
    }
    Vec8<T>::store(out + r * ldo + c, a);
  }
}

static int ew_grid2(long total, int threads) {
  long blocks = (total + threads - 1) / threads;
  long cap = (long)num_sms() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace sst

using namespace sst;

extern "C" {

int sst_shift_left(float* x, int64_t n_chunks, int T, int Cc, int r, void* stream) {
  SST_REQUIRE(r >= 0 && r < T, SST_E_ARG, "shift_left: bad shift %d", r);
  if (r == 0 || n_chunks <= 0) return SST_OK;
  SST_REQUIRE((long)r * Cc <= 256L * SHIFT_PER_THREAD, SST_E_ARG, "shift_left: shift %d too large for the in-place passes", r);
  shift_left_kernel<<<(int)n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n_chunks, T, Cc, r);
  return check_launch("shift_left");
}

int sst_im2col_first(int out_dtype, const float* x, void* col, int64_t n_chunks, int Tin, void* stream) {
  SST_REQUIRE(Tin % 2 == 0, SST_E_ARG, "im2col_first: Tin must be even");
  if (n_chunks <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid2(n_chunks * (Tin / 2) * 4, 256);
  if (out_dtype == SST_F32) im2col_first_kernel<float><<<grid, 256, 0, st>>>(x, (float*)col, n_chunks, Tin);
  else im2col_first_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)col, n_chunks, Tin);
  return check_launch("im2col_first");
}

int sst_gather_rows_pad(int dtype, const void* in, void* out, const int64_t* offs, const int32_t* lens, int B, int Lmax, int D,
                        float fill, void* stream) {
  SST_REQUIRE(D % 8 == 0, SST_E_ARG, "gather_rows_pad: D must be a multiple of 8");
  if (B <= 0 || Lmax <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid2((long)B * Lmax * (D / 8), 256);
  if (dtype == SST_F32) gather_rows_pad_kernel<float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, (const long*)offs, lens, B, Lmax, D, fill);
  else gather_rows_pad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, (const long*)offs, lens, B, Lmax, D, fill);
  return check_launch("gather_rows_pad");
}

int sst_scatter_rows(int dtype, const void* dout, void* din, const int64_t* offs, const int32_t* lens, int B, int Lmax, int D,
                     void* stream) {
  SST_REQUIRE(D % 8 == 0, SST_E_ARG, "scatter_rows: D must be a multiple of 8");
  if (B <= 0 || Lmax <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ew_grid2((long)B * Lmax * (D / 8), 256);
  if (dtype == SST_F32) scatter_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)dout, (float*)din, (const long*)offs, lens, B, Lmax, D);
  else scatter_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dout, (__nv_bfloat16*)din, (const long*)offs, lens, B, Lmax, D);
  return check_launch("scatter_rows");
}

int sst_embed_posenc_fwd(int out_dtype, const int64_t* y, const float* W, const float* pe, void* out, int B, int S, int D,
                         float drop_p, uint64_t seed, void* stream) {
  SST_REQUIRE(D % 8 == 0, SST_E_ARG, "embed_posenc: D must be a multiple of 8");
  if (B <= 0 || S <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thr = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float ds = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  const int grid = ew_grid2((long)B * S * (D / 8), 256);
  if (out_dtype == SST_F32) embed_posenc_kernel<float><<<grid, 256, 0, st>>>((const long*)y, W, pe, (float*)out, B, S, D, thr, ds, seed, dropout_salt());
  else embed_posenc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const long*)y, W, pe, (__nv_bfloat16*)out, B, S, D, thr, ds, seed, dropout_salt());
  return check_launch("embed_posenc_fwd");
}

int sst_embed_bwd(int dtype, const int64_t* y, const void* dout, float* dW, int B, int S, int D, int pad_idx, float drop_p,
                  uint64_t seed, void* stream) {
  SST_REQUIRE(D % 8 == 0, SST_E_ARG, "embed_bwd: D must be a multiple of 8");
  if (B <= 0 || S <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thr = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float ds = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  const int grid = ew_grid2((long)B * S * (D / 8), 256);
  if (dtype == SST_F32) embed_bwd_kernel<float><<<grid, 256, 0, st>>>((const long*)y, (const float*)dout, dW, B, S, D, pad_idx, thr, ds, seed, dropout_salt());
  else embed_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const long*)y, (const __nv_bfloat16*)dout, dW, B, S, D, pad_idx, thr, ds, seed, dropout_salt());
  return check_launch("embed_bwd");
}

int sst_permute3_cast(int in_dtype, int out_dtype, const void* in, void* out, int64_t d0, int64_t d1, int64_t d2, int64_t s0,
                      int64_t s1, int64_t s2, int64_t o0, int64_t o1, int64_t o2, int accumulate, void* stream) {
  const long total = d0 * d1 * d2;
  if (total <= 0) return SST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  typedef __nv_bfloat16 bf;
  const bool in_j = (s1 == 1 && o2 == 1 && s2 != 1), in_k = (s2 == 1 && o1 == 1 && o2 != 1);
  if ((in_j || in_k) && d1 >= 32 && d2 >= 32 && d0 <= 65535 && (d1 + 31) / 32 <= 65535) {
    const dim3 grid((unsigned)((d2 + 31) / 32), (unsigned)((d1 + 31) / 32), (unsigned)d0), block(32, 8);
#define SST_P3T(TI_, TO_)                                                                                                    \
    do {                                                                                                                     \
      if (in_j) permute3_tiled_kernel<TI_, TO_, true><<<grid, block, 0, st>>>((const TI_*)in, (TO_*)out, d1, d2, s0, s1, s2, o0, o1, o2, accumulate); \
      else permute3_tiled_kernel<TI_, TO_, false><<<grid, block, 0, st>>>((const TI_*)in, (TO_*)out, d1, d2, s0, s1, s2, o0, o1, o2, accumulate);    \
    } while (0)
    if (in_dtype == SST_F32 && out_dtype == SST_F32) SST_P3T(float, float);
    else if (in_dtype == SST_F32) SST_P3T(float, bf);
    else if (out_dtype == SST_F32) SST_P3T(bf, float);
    else SST_P3T(bf, bf);
#undef SST_P3T
    return check_launch("permute3_cast");
  }
  const int grid = ew_grid2(total, 256);
  if (in_dtype == SST_F32 && out_dtype == SST_F32) permute3_cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, d0, d1, d2, s0, s1, s2, o0, o1, o2, accumulate);
  else if (in_dtype == SST_F32) permute3_cast_kernel<float, bf><<<grid, 256, 0, st>>>((const float*)in, (bf*)out, d0, d1, d2, s0, s1, s2, o0, o1, o2, accumulate);
  else if (out_dtype == SST_F32) permute3_cast_kernel<bf, float><<<grid, 256, 0, st>>>((const bf*)in, (float*)out, d0, d1, d2, s0, s1, s2, o0, o1, o2, accumulate);
  else permute3_cast_kernel<bf, bf><<<grid, 256, 0, st>>>((const bf*)in, (bf*)out, d0, d1, d2, s0, s1, s2, o0, o1, o2, accumulate);
  return check_launch("permute3_cast");
}

int sst_permute3_plan(SstPermuteItem* items, int n_items, int* total_blocks) {
  SST_REQUIRE(items != nullptr && total_blocks != nullptr && n_items > 0, SST_E_ARG, "permute3_plan: empty table");
  long nb = 0;
  for (int i = 0; i < n_items; ++i) {
    SstPermuteItem& it = items[i];
    const long total = it.d0 * it.d1 * it.d2;
    SST_REQUIRE(total > 0, SST_E_ARG, "permute3_plan: item %d is empty", i);
    const bool in_j = (it.s1 == 1 && it.o2 == 1 && it.s2 != 1), in_k = (it.s2 == 1 && it.o1 == 1 && it.o2 != 1);
    it.first_block = (int)nb;
    const long s_r = in_j ? it.s2 : it.s1, o_c = in_j ? it.o1 : it.o2;
    const bool vec_ok = (in_j || in_k) && it.out_dtype == SST_BF16 && !it.accumulate && it.d1 % 8 == 0 && it.d2 % 8 == 0 && it.d1 >= 16 &&
                        it.d2 >= 16 && s_r % 8 == 0 && o_c % 8 == 0 && (it.d0 == 1 || (it.s0 % 8 == 0 && it.o0 % 8 == 0)) &&
                        ((uintptr_t)it.in & 31) == 0 && ((uintptr_t)it.out & 15) == 0;
    if (vec_ok) {
      it.mode = in_j ? 3 : 4;
      const long tiles = ((it.d2 + 63) / 64) * ((it.d1 + 63) / 64) * it.d0;
      it.nblocks = (int)((tiles + PT_TILES_PER_BLOCK - 1) / PT_TILES_PER_BLOCK);
    } else if ((in_j || in_k) && it.d1 >= 32 && it.d2 >= 32) {
      it.mode = in_j ? 1 : 2;
      it.nblocks = (int)(((it.d2 + 31) / 32) * ((it.d1 + 31) / 32) * it.d0);
    } else {
      it.mode = 0;
      long b = (total + 2047) / 2048;                        // eight elements per thread
      it.nblocks = (int)(b < 1 ? 1 : (b > 4096 ? 4096 : b));
    }
    nb += it.nblocks;
    SST_REQUIRE(nb < (1L << 31), SST_E_ARG, "permute3_plan: too many blocks");
  }
  *total_blocks = (int)nb;
  return SST_OK;
}

int sst_permute3_cast_batch(const SstPermuteItem* items_dev, int n_items, int total_blocks, void* stream) {
  if (n_items <= 0 || total_blocks <= 0) return SST_OK;
  permute3_batch_kernel<<<total_blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(items_dev, n_items);
  return check_launch("permute3_cast_batch");
}

static int gelu_launch(bool bwd, int dtype, int64_t rows, int cols, const void* x, int64_t ldx, const void* dy, int64_t lddy, void* out,
                       int64_t ldo, float drop_p, uint64_t seed, void* stream) {
  SST_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && (!bwd || lddy % 8 == 0), SST_E_ARG, "gelu: columns and pitches must be multiples of 8");
  SST_REQUIRE(drop_p >= 0.f && drop_p < 1.f, SST_E_ARG, "gelu: dropout probability out of range");
  if (rows <= 0 || cols <= 0) return SST_OK;
  const uint32_t thr = drop_p > 0.f ? drop_threshold16(drop_p) : 0u;
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int grid = ew_grid2(rows * (cols / 8), 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define SST_GELU(T, B) gelu_dropout_kernel<T, B><<<grid, 256, 0, st>>>((const T*)x, ldx, (const T*)dy, lddy, (T*)out, ldo, rows, cols, thr, dscale, seed, dropout_salt())
  if (dtype == SST_F32) { if (bwd) SST_GELU(float, true); else SST_GELU(float, false); }
  else { if (bwd) SST_GELU(__nv_bfloat16, true); else SST_GELU(__nv_bfloat16, false); }
#undef SST_GELU
  return check_launch(bwd ? "gelu_dropout_bwd" : "gelu_dropout_fwd");
}

int sst_gelu_dropout_fwd(int dtype, int64_t rows, int cols, const void* x, int64_t ldx, float drop_p, uint64_t seed, void* y, int64_t ldy,
                         void* stream) {
  return gelu_launch(false, dtype, rows, cols, x, ldx, nullptr, 0, y, ldy, drop_p, seed, stream);
}

int sst_gelu_dropout_bwd(int dtype, int64_t rows, int cols, const void* dy, int64_t lddy, const void* x, int64_t ldx, float drop_p,
                         uint64_t seed, void* dx, int64_t lddx, void* stream) {
  return gelu_launch(true, dtype, rows, cols, x, ldx, dy, lddy, dx, lddx, drop_p, seed, stream);
}

int sst_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
              int64_t step, void* p_bf16, void* stream) {
  SST_REQUIRE(step >= 1, SST_E_ARG, "adamw: step is 1-based");
  SST_REQUIRE(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)v & 15) == 0, SST_E_ARG,
              "adamw: buffers must be 16-byte aligned");
  if (n <= 0) return SST_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int grid = ew_grid2((n + 3) / 4, 256);
  SST_REQUIRE(p_bf16 == nullptr || ((uintptr_t)p_bf16 & 7) == 0, SST_E_ARG, "adamw: bf16 shadow must be 8-byte aligned");
  adamw_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, (float)bc1,
                                                                         (float)sqrt(bc2), reinterpret_cast<__nv_bfloat16*>(p_bf16), nullptr);
  return check_launch("adamw");
}

int sst_adamw_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1, float beta2, float eps,
                  float wd, void* p_bf16, void* stream) {
  SST_REQUIRE(hyper != nullptr, SST_E_ARG, "adamw_dev: hyper (device float[3]: lr, 1 - beta1^t, sqrt(1 - beta2^t)) required");
  if (n <= 0) return SST_OK;
  const int grid = ew_grid2((n + 3) / 4, 256);
  SST_REQUIRE(p_bf16 == nullptr || ((uintptr_t)p_bf16 & 7) == 0, SST_E_ARG, "adamw: bf16 shadow must be 8-byte aligned");
  adamw_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, 0.f, beta1, beta2, eps, wd, 1.f, 1.f,
                                                                         reinterpret_cast<__nv_bfloat16*>(p_bf16), hyper);
  return check_launch("adamw_dev");
}

// ---- by-value scalars into device memory: the per-step values (dropout salt, learning rate, bias corrections) a replayed CUDA
//      graph reads.  The bytes travel as a kernel PARAMETER, i.e. they are captured when this call returns: no pinned staging buffer
//      whose lifetime the caller would have to manage.
struct ScalarBlob { unsigned int w[16]; };
__global__ void write_scalars_kernel(unsigned int* __restrict__ dst, ScalarBlob b, int nwords) {
  if ((int)threadIdx.x < nwords) dst[threadIdx.x] = b.w[threadIdx.x];
}

int sst_write_scalars(void* dst, const void* src_host, int nbytes, void* stream) {
  SST_REQUIRE(dst != nullptr && src_host != nullptr && nbytes > 0 && nbytes <= 64 && nbytes % 4 == 0 && ((uintptr_t)dst & 3) == 0,
              SST_E_ARG, "write_scalars: 4..64 bytes, a multiple of 4, 4-byte aligned destination");
  ScalarBlob b;
  memset(&b, 0, sizeof(b));
  memcpy(&b, src_host, (size_t)nbytes);
  write_scalars_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned int*>(dst), b, nbytes / 4);
  return check_launch("write_scalars");
}

}  // extern "C"

// bf16 GEMM on 5th-gen tensor cores: TMA -> 128B-swizzled shared memory -> tcgen05.mma (accumulators in
// TMEM, double buffered) -> tcgen05.ld epilogue.  Persistent, warp-specialised: warp 0 = TMA producer,
// warp 1 = MMA issuer (one thread) + TMEM owner, warps 2-17 = epilogue (TMEM lane quarter = warp & 3, four warps per
// quarter splitting the tile's columns).
//
// One kernel covers every dense contraction of the hot path (include/sst.h, "GEMM family"):
//   TN     : C = A[M,K] * B[N,K]^T, both K-major.  A may be read as up to three row-shifted column windows
//            ("taps"): a k=3 Conv1d over a time-padded channels-last activation is a GEMM whose K blocks come
//            from rows m-1 / m / m+1 (stride 1) or from the (2C)-wide pair-row view (stride 2).
//   NT_MN  : C (+)= A[K,M]^T * B[K,N], both MN-major (weight gradients; K = tokens), split-K with fp32
//            atomics, B optionally row-shifted per N segment (conv taps again).
#include "sst_common.cuh"
#include "sst_ptx.cuh"

namespace sst {

constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_EPI_WARPS = 16;                  // four per TMEM lane quarter, each takes a quarter of the tile's columns
constexpr int G_CGROUPS = G_EPI_WARPS / 4;       // column groups
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;
constexpr int G_A_BYTES = G_BM * G_BK * 2;   // 16 KiB

struct GemmKParams {
  int M, N, K;
  int mode_mn;        // A is MN-major (weight-gradient layout)
  int b_mn;           // B is MN-major (NT_MN and TN_BMN layouts)
  int num_kb;
  int kb_per_seg;     // TN: k-blocks per A segment
  int nseg_cols;      // MN: columns of C per B segment
  int a_row_shift[3], a_col0[3], b_row_shift[3], b_col0[3];
  int splits, kb_per_split;
  int m_blks, n_blks;
  int epilogue;
  float alpha, mask_scale, drop_scale;
  uint32_t drop_thr;
  unsigned long long seed;
  const float* bias;
  const void* aux;
  long ldaux;
  int aux_f32;
  void* C;
  long ldc;
  int out_f32;
  int atomic_out;
  int remap_P, remap_T, remap_j0;
};

template <int BN> struct GemmCfg {
  static constexpr int B_BYTES = BN * G_BK * 2;
  static constexpr int STAGE_BYTES = G_A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;                       // two accumulator buffers
  static constexpr int STAGING_BYTES = G_EPI_WARPS * 2048;        // per epilogue warp: 32 rows x 64 B of bf16 output
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STAGING_BYTES;
};

__device__ __forceinline__ void decode_unit(const GemmKParams& p, int u, int& m0, int& n0, int& kb0, int& kb1) {
  int tile = u / p.splits, split = u - tile * p.splits;
  int mb = tile / p.n_blks, nb = tile - mb * p.n_blks;
  m0 = mb * G_BM;
  n0 = nb;   // caller multiplies by BN
  kb0 = split * p.kb_per_split;
  kb1 = min(kb0 + p.kb_per_split, p.num_kb);
}

template <int BN>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull[a], 1); ptx::mbar_init(&tempty[a], G_EPI_WARPS); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_units = p.m_blks * p.n_blks * p.splits;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        int m0, nb, kb0, kb1;
        decode_unit(p, u, m0, nb, kb0, kb1);
        const int n0 = nb * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + G_A_BYTES;
          ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          if (!p.mode_mn) {
            int seg = kb / p.kb_per_seg;
            int x = p.a_col0[seg] + (kb - seg * p.kb_per_seg) * G_BK;
            ptx::tma_load_2d(sA, &tmA, &full[stage], x, m0 + p.a_row_shift[seg]);
          } else {
            const int k0 = kb * G_BK;
#pragma unroll
            for (int i = 0; i < G_BM / 64; ++i) ptx::tma_load_2d(sA + i * 8192, &tmA, &full[stage], m0 + i * 64, k0);
          }
          if (!p.b_mn) {
            ptx::tma_load_2d(sB, &tmB, &full[stage], kb * G_BK, n0);
          } else {
            const int k0 = kb * G_BK;
            int seg = n0 / p.nseg_cols;
            int nin = n0 - seg * p.nseg_cols;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sB + j * 8192, &tmB, &full[stage], p.b_col0[seg] + nin + j * 64, k0 + p.b_row_shift[seg]);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t idesc = ptx::make_idesc_bf16(G_BM, BN, p.mode_mn, p.b_mn);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      int m0, nb, kb0, kb1;
      decode_unit(p, u, m0, nb, kb0, kb1);
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + G_A_BYTES;
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k) {
            const uint64_t ad = !p.mode_mn ? ptx::make_smem_desc_sw128(a_base + k * 32, 0, 1024)
                                           : ptx::make_smem_desc_sw128(a_base + k * 2048, 8192, 1024);
            const uint64_t bd = !p.b_mn ? ptx::make_smem_desc_sw128(b_base + k * 32, 0, 1024)
                                        : ptx::make_smem_desc_sw128(b_base + k * 2048, 8192, 1024);
            ptx::umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty[stage]);
          if (kb == kb1 - 1) ptx::umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int cg = (warp - 2) >> 2;               // column group: chunks [cg * CPW, (cg + 1) * CPW)
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      int m0, nb, kb0, kb1;
      decode_unit(p, u, m0, nb, kb0, kb1);
      const int n0 = nb * BN;
      const int m = m0 + q * 32 + lane;
      // Epilogue inputs that do not depend on the accumulator are produced BEFORE waiting for it, i.e. under the MMA
      // main loop: the dropout keep bits (Philox, eight 16-bit lanes per block) and the ReLU/dropout mask bits of `aux`.
      constexpr int CPW = BN / 32 / G_CGROUPS;      // 32-column chunks per epilogue warp
      uint32_t keep_bits[CPW], mask_bits[CPW];
      if (p.epilogue & SST_EPI_DROPOUT) {
#pragma unroll
        for (int lc = 0; lc < CPW; ++lc) {
          const int nbase = n0 + (cg * CPW + lc) * 32;
          const unsigned long long e0 = (unsigned long long)m * (unsigned long long)p.N + (unsigned long long)nbase;
          uint32_t bits = 0;
#pragma unroll
          for (int i8 = 0; i8 < 4; ++i8) {
            const Philox4 rr = philox4x32_10(p.seed, (e0 >> 3) + i8);
#pragma unroll
            for (int e = 0; e < 8; ++e) bits |= (philox_lane16(rr, e) >= p.drop_thr ? 1u : 0u) << (i8 * 8 + e);
          }
          keep_bits[lc] = bits;
        }
      }
      if (p.epilogue & SST_EPI_MULMASK) {
#pragma unroll
        for (int lc = 0; lc < CPW; ++lc) {
          const int nbase = n0 + (cg * CPW + lc) * 32;
          uint32_t bits = 0;
          if (m < p.M && nbase < p.N) {
            const int ncols = min(32, p.N - nbase);
            const long ab = (long)m * p.ldaux + nbase;
            if (!p.aux_f32 && ncols == 32 && (p.ldaux & 7) == 0) {
              const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + ab);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 w = __ldg(ap + j);
                const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&w);
#pragma unroll
                for (int i = 0; i < 8; ++i) bits |= (__bfloat162float(h[i]) > 0.f ? 1u : 0u) << (j * 8 + i);
              }
            } else {
              for (int i = 0; i < ncols; ++i)
                bits |= (ld_as_f32(p.aux, ab + i, p.aux_f32 ? SST_F32 : SST_BF16) > 0.f ? 1u : 0u) << i;
            }
          }
          mask_bits[lc] = bits;
        }
      }
      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();
      bool row_ok = m < p.M;
      long out_row = m;
      if (p.remap_P > 0) {
        int chunk = m / p.remap_P;
        int t = m - chunk * p.remap_P - p.remap_j0;
        row_ok = row_ok && t >= 0 && t < p.remap_T;
        out_row = (long)chunk * p.remap_T + t;
      }
      // Coalesced bf16 stores: the warp's (32 rows x 32 columns) chunk is transposed through shared memory so that one store
      // instruction writes 8 rows x 64 contiguous bytes (full sectors) instead of 32 rows x 16 bytes.  Lane l stores the
      // 16-byte piece (l & 3) of rows it*8 + (l >> 2); the row bookkeeping of those rows comes from their owner lanes.
      const bool staged = !p.atomic_out && !p.out_f32 && !(p.epilogue & SST_EPI_ACCUM) && (p.ldc & 7) == 0;
      long st_row[4];
      bool st_ok[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int src = it * 8 + (lane >> 2);
        st_row[it] = __shfl_sync(0xffffffffu, out_row, src);
        st_ok[it] = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, src) != 0;
      }
      const uint32_t stg = ptx::smem_u32(staging) + (uint32_t)(warp - 2) * 2048;
#pragma unroll
      for (int lc0 = 0; lc0 < CPW; ++lc0) {
        const int c = cg * CPW + lc0;
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + c * 32) + ((uint32_t)(q * 32) << 16);
        ptx::tmem_ld_32x32b_x32(taddr, r);
        ptx::tmem_ld_wait();
        const int nbase = n0 + c * 32;
        const bool chunk_staged = staged && nbase + 32 <= p.N;          // warp-uniform
        if ((row_ok || chunk_staged) && nbase < p.N) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
          const int ncols = min(32, p.N - nbase);
          if (p.epilogue & SST_EPI_BIAS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __ldg(p.bias + nbase + i);
          }
          if (p.epilogue & SST_EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (p.epilogue & SST_EPI_DROPOUT) {
            const uint32_t kb = keep_bits[lc0];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = ((kb >> i) & 1u) ? v[i] * p.drop_scale : 0.f;
          }
          if (p.epilogue & SST_EPI_MULMASK) {
            const uint32_t mb = mask_bits[lc0];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= ((mb >> i) & 1u) ? p.mask_scale : 0.f;
          }
          const long cb = out_row * p.ldc + nbase;
          if (chunk_staged) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1]);
                w4[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              ptx::st_shared_v4(stg + lane * 64 + ((uint32_t)(j ^ ((lane >> 1) & 3)) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rl = it * 8 + (lane >> 2), piece = lane & 3;
              uint4 o;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                           : "r"(stg + rl * 64 + ((uint32_t)(piece ^ ((rl >> 1) & 3)) << 4)));
              if (st_ok[it])
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + st_row[it] * p.ldc + nbase + piece * 8) = o;
            }
            __syncwarp();
          } else if (!row_ok) {
            // nothing to store for this lane
          } else if (p.atomic_out) {
            float* cp = reinterpret_cast<float*>(p.C) + cb;
            if (ncols == 32 && (p.ldc & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                             "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) atomicAdd(cp + i, v[i]);
            }
          } else if (p.out_f32) {
            float* cp = reinterpret_cast<float*>(p.C) + cb;
            if (ncols == 32 && (p.ldc & 3) == 0) {
              float4* c4 = reinterpret_cast<float4*>(cp);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                if (p.epilogue & SST_EPI_ACCUM) { float4 old = c4[j]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                c4[j] = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncols) cp[i] = v[i] + ((p.epilogue & SST_EPI_ACCUM) ? cp[i] : 0.f);
            }
          } else {
            __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + cb;
            if (ncols == 32 && (p.ldc & 7) == 0) {
              uint4* c4 = reinterpret_cast<uint4*>(cp);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (p.epilogue & SST_EPI_ACCUM) {
                  uint4 old = c4[j];
                  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&old);
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[j * 8 + i] += __bfloat162float(h[i]);
                }
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1]);
                c4[j] = o;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncols) {
                  float o = v[i] + ((p.epilogue & SST_EPI_ACCUM) ? __bfloat162float(cp[i]) : 0.f);
                  cp[i] = __float2bfloat16_rn(o);
                }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major matrix (cols contiguous), 128B swizzle, zero fill out of bounds.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, long cols, long rows, long ld_elems, int box_cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  SST_REQUIRE(fn != nullptr, SST_E_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
  SST_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, SST_E_ARG, "TMA operand not 16-byte aligned");
  SST_REQUIRE((ld_elems * 2) % 16 == 0, SST_E_ARG, "TMA operand row pitch (%ld elements) not a multiple of 16 bytes", ld_elems);
  SST_REQUIRE(cols > 0 && rows > 0, SST_E_ARG, "empty TMA operand");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SST_REQUIRE(r == CUDA_SUCCESS, SST_E_LAUNCH, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SST_OK;
}

template <int BN>
static int launch_bn(const SstGemmDesc& d, const void* A, const void* B, GemmKParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  p.n_blks = cdiv(d.N, BN);
  const int tiles = p.m_blks * p.n_blks;
  const int sms = num_sms();
  int splits = 1;
  if (p.mode_mn) {
    // split K (= tokens) until the machine is full, but keep >= 16 k-blocks per unit: every unit pays a pipeline fill and
    // a (BM x BN) fp32 atomic epilogue
    splits = (2 * sms) / tiles;
    const int max_by_k = p.num_kb / 16;
    if (splits > max_by_k) splits = max_by_k;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = cdiv(p.num_kb, splits);
  p.splits = cdiv(p.num_kb, p.kb_per_split);
  p.atomic_out = p.splits > 1;
  if (p.atomic_out) {
    SST_REQUIRE(d.out_dtype == SST_F32 && (d.epilogue & ~SST_EPI_ACCUM) == 0, SST_E_ARG,
                "split-K GEMM needs fp32 output and no epilogue besides ACCUM");
    if (!(d.epilogue & SST_EPI_ACCUM)) {
      cudaError_t e = cudaMemset2DAsync(p.C, (size_t)d.ldc * 4, 0, (size_t)d.N * 4, (size_t)d.M, st);
      SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "memset: %s", cudaGetErrorString(e));
    }
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!p.mode_mn) {
    if ((rc = make_tmap_bf16_2d(&tmA, A, d.a_cols, d.a_rows, d.lda, G_BK, G_BM))) return rc;
  } else {
    if ((rc = make_tmap_bf16_2d(&tmA, A, d.a_cols, d.a_rows, d.lda, 64, G_BK))) return rc;
  }
  if (!p.b_mn) {
    if ((rc = make_tmap_bf16_2d(&tmB, B, d.K, d.N, d.ldb, G_BK, BN))) return rc;
  } else {
    if ((rc = make_tmap_bf16_2d(&tmB, B, d.b_cols, d.b_rows, d.ldb, 64, G_BK))) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    SST_REQUIRE(e == cudaSuccess, SST_E_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const int units = tiles * p.splits;
  const int grid = units < sms ? units : sms;
  gemm_tcgen05_kernel<BN><<<grid, G_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, p);
  return check_launch("gemm_tcgen05");
}

int launch_gemm_tcgen05(const SstGemmDesc& d, const void* A, const void* B, void* C, const void* bias, const void* aux,
                        cudaStream_t st) {
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)d.M; p.N = (int)d.N; p.K = (int)d.K;
  p.mode_mn = d.layout == SST_GEMM_NT_MN;
  p.b_mn = d.layout != SST_GEMM_TN;
  p.num_kb = cdiv(d.K, G_BK);
  const int nseg = d.n_seg > 0 ? d.n_seg : 1;
  SST_REQUIRE(nseg <= 3, SST_E_ARG, "n_seg must be <= 3");
  for (int s = 0; s < 3; ++s) {
    p.a_row_shift[s] = d.a_row_shift[s]; p.a_col0[s] = d.a_col0[s];
    p.b_row_shift[s] = d.b_row_shift[s]; p.b_col0[s] = d.b_col0[s];
  }
  if (!p.mode_mn) {
    SST_REQUIRE(nseg == 1 || (d.K % nseg == 0 && (d.K / nseg) % G_BK == 0), SST_E_ARG,
                "segmented A needs K/n_seg to be a multiple of %d", G_BK);
    p.kb_per_seg = nseg == 1 ? p.num_kb : (int)(d.K / nseg / G_BK);
    p.nseg_cols = (int)d.N;
    if (p.b_mn) { p.b_row_shift[0] = 0; p.b_col0[0] = 0; }      // TN_BMN: B is a plain (K, N) row-major matrix
  } else {
    p.kb_per_seg = p.num_kb;
    p.nseg_cols = (int)(d.N / nseg);
    SST_REQUIRE(nseg == 1 || (d.N % nseg == 0 && p.nseg_cols % 256 == 0), SST_E_ARG,
                "segmented B needs N/n_seg to be a multiple of 256");
  }
  p.m_blks = cdiv(d.M, G_BM);
  p.epilogue = d.epilogue;
  p.alpha = d.alpha; p.mask_scale = d.mask_scale;
  p.drop_thr = drop_threshold16(d.drop_p);
  p.drop_scale = d.drop_p < 1.f ? 1.f / (1.f - d.drop_p) : 0.f;
  p.seed = d.seed;
  p.bias = reinterpret_cast<const float*>(bias);
  p.aux = aux; p.ldaux = d.ldaux; p.aux_f32 = d.aux_dtype == SST_F32;
  p.C = C; p.ldc = d.ldc; p.out_f32 = d.out_dtype == SST_F32;
  p.remap_P = d.remap_P; p.remap_T = d.remap_T; p.remap_j0 = d.remap_j0;
  if (d.epilogue & SST_EPI_DROPOUT) SST_REQUIRE(d.N % 8 == 0, SST_E_ARG, "dropout epilogue needs N %% 8 == 0");
  // tile width by wave quantisation: time ~ waves * BN; 256-wide tiles reuse the A tile twice as long, so 128 must win by 10 %
  if (d.N <= 128) return launch_bn<128>(d, A, B, p, st);
  if (p.mode_mn) return (d.N % 256 == 0) ? launch_bn<256>(d, A, B, p, st) : launch_bn<128>(d, A, B, p, st);
  const long sms = num_sms();
  const long t256 = (long)p.m_blks * cdiv(d.N, 256), t128 = (long)p.m_blks * cdiv(d.N, 128);
  const long c256 = ((t256 + sms - 1) / sms) * 256, c128 = ((t128 + sms - 1) / sms) * 128;
  if (c128 * 10 < c256 * 9) return launch_bn<128>(d, A, B, p, st);
  return launch_bn<256>(d, A, B, p, st);
}

}  // namespace sst

// Standalone check (not part of libsst.so) of the three hardware assumptions behind attention_tc2.cu:
//   (1) fence.proxy.async.global and add.rn.f32.f16 (FHADD .H0/.H1) execute on sm_100a;
//   (2) a 3-D tensor map whose row pitch is (NE - 8) elements delivers the skewed bias tile;
//   (3) SWIZZLE_128B is a function of the ABSOLUTE shared-memory address (class blocks 17 rows apart start at phase rho).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_skew_test tma_skew_test.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../emg-based-speech-recognition-with-heterogenous-data_b200/csrc/sst_ptx.cuh"
using namespace sst;

__global__ void k_fence(float* out) {
  out[threadIdx.x] = 1.f;
  ptx::fence_proxy_async_global();
  out[threadIdx.x + 32] = 2.f;
}
__global__ void k_fhadd(const uint32_t* w, float* out) {
  out[threadIdx.x] = ptx::add_f16lo(w[threadIdx.x], 1.f) + 100.f * ptx::add_f16hi(w[threadIdx.x], 2.f);
}

__global__ void __launch_bounds__(128, 1)
k_skew(const __grid_constant__ CUtensorMap tmB, int c0, int blk, int cls_rows, int abs_swz, float* out /*[128][64]*/) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * 17 * 128);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, 8 * 16 * 128);
    for (int rho = 0; rho < 8; ++rho) ptx::tma_load_3d(smem + rho * cls_rows * 128, &tmB, bar, c0 - rho, 0, blk + rho);
  }
  {
    uint32_t spins = 0;
    bool ok = true;
    while (!ptx::mbar_try_wait(bar, 0)) { if (++spins > 200000u) { ok = false; break; } }
    if (!ok) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
  }
  const int li = threadIdx.x;
  const uint32_t brow = cls_rows * (li & 7) + (li >> 3);
  const uint32_t phase = abs_swz ? (brow & 7u) : ((li >> 3) & 7u);
  for (int c = 0; c < 8; ++c) {
    const uint4 v = ptx::ld_shared_v4(ptx::smem_u32(smem) + brow * 128 + (((uint32_t)c ^ phase) << 4));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    for (int e = 0; e < 4; ++e) {
      out[li * 64 + c * 8 + 2 * e] = __low2float(h[e]);
      out[li * 64 + c * 8 + 2 * e + 1] = __high2float(h[e]);
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  float* out; cudaMalloc(&out, 128 * 64 * 4);
  k_fence<<<1, 32>>>(out);
  printf("fence.proxy.async.global: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  uint32_t hw[32]; for (int i = 0; i < 32; ++i) { __half2 h = __floats2half2_rn(0.5f * i, 3.f); hw[i] = *reinterpret_cast<uint32_t*>(&h); }
  uint32_t* dw; cudaMalloc(&dw, 128); cudaMemcpy(dw, hw, 128, cudaMemcpyHostToDevice);
  k_fhadd<<<1, 32>>>(dw, out);
  float ho[32]; cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(ho, out, 128, cudaMemcpyDeviceToHost);
  printf("add.rn.f32.f16: %s, out[5] = %g (expect %g)\n", cudaGetErrorString(e), ho[5], 0.5f * 5 + 1.f + 100.f * 5.f);

  const int NE = 208, R = 100, pad = 5, RP = R - 1 + pad, n_blocks = 16;
  std::vector<__half> h((size_t)n_blocks * 16 * NE + 256);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr);
  PFN_encodeTiled fn = (PFN_encodeTiled)fnp;
  __half* ds; cudaMalloc(&ds, h.size() * 2);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)NE + 120, 16, (cuuint64_t)n_blocks};
  cuuint64_t strides[2] = {(cuuint64_t)(NE - 8) * 2, (cuuint64_t)16 * NE * 2};
  cuuint32_t box[3] = {64, 16, 1}, estr[3] = {1, 1, 1};
  const int variant = getenv("VARIANT") ? atoi(getenv("VARIANT")) : 3;
  if (variant == 1) { dims[0] = NE; strides[0] = NE * 2; }
  if (variant == 2) { dims[0] = NE - 8; }
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ds, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d encode: %d\n", variant, (int)r);
  cudaFuncSetAttribute(k_skew, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const int blk = 8;     // second slot
  for (int mode = 0; mode < 2; ++mode) {              // 0: value = column, 1: value = row id (rho*16 + k)
    for (size_t b = 0; b < (size_t)n_blocks; ++b)
      for (int k = 0; k < 16; ++k)
        for (int c = 0; c < NE; ++c) h[(b * 16 + k) * NE + c] = __float2half(mode == 0 ? (float)c : (float)((b % 8) * 16 + k));
    cudaMemcpy(ds, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    for (int cls = 16; cls <= 17; ++cls)
      for (int abs_swz = 0; abs_swz < 2; ++abs_swz)
        for (int tcase = 0; tcase < 3; ++tcase) {
          const int i0 = 128, j0 = tcase == 0 ? 128 : tcase == 1 ? 64 : 256;     // diagonal tile, tile to the left, tile to the right
          const int c0 = j0 - i0 + RP;
          cudaMemset(out, 0, 128 * 64 * 4);
          k_skew<<<1, 128, 32768>>>(tm, c0, blk, cls, abs_swz, out);
          cudaError_t ee = cudaDeviceSynchronize();
          std::vector<float> o(128 * 64);
          cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0, checked = 0;
          for (int li = 0; li < 128; ++li)
            for (int x = 0; x < 64; ++x) {
              const int c = (j0 + x) - (i0 + li) + RP;          // slab column the reference formula wants
              if (c < 0 || c >= NE) continue;                   // out of band: unspecified
              const float want = mode == 0 ? (float)c : (float)((li & 7) * 16 + (li >> 3));
              ++checked;
              if (o[li * 64 + x] != want) ++bad;
            }
          printf("mode %d cls %d abs_swz %d tile %d: %s%s, %d / %d wrong\n", mode, cls, abs_swz, tcase, cudaGetErrorString(ee),
                 o[0] == -12345.f ? " (TMA never completed)" : "", bad, checked);
        }
  }
  return 0;
}

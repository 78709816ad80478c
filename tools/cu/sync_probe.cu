// Micro-benchmark (not part of libsst.so): cycles per synchronisation primitive the attention kernels execute once per tile and warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_probe sync_probe.cu && ./sync_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ float red[512];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");   // phase 0 of bar[0] done
  __syncthreads();
  const int N = 2000;
  long long t0, t1;
  uint32_t ok = 0, acc = 0;
  // 1. try_wait (suspend hint 20000 ns) on a phase that completed long ago
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"(0u), "r"(20000u) : "memory");
    acc += ok;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / N;
  // 2. test_wait on the same
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
    acc += ok;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = (t1 - t0) / N;
  // 3. arrive (lane 0 of every warp, as the kernels do) on a barrier that never completes
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = (t1 - t0) / N;
  // 4. named barrier over the 4 warps that share a lane quarter (ids 1..4, 128 threads), with the shared-memory exchange around it
  t0 = clock64();
  float m = (float)threadIdx.x;
  for (int i = 0; i < N; ++i) {
    red[(w >> 2) * 128 + (w & 3) * 32 + lane] = m;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (w & 3)), "r"(blockDim.x >= 512 ? 128 : 32) : "memory");
    float v = 0.f;
    for (int g = 0; g < 4; ++g) v = fmaxf(v, red[g * 128 + (w & 3) * 32 + lane]);
    m = v + 1.f;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = (t1 - t0) / N;
  // 5. fence.proxy.async + tcgen05 fences as executed before the P-ready arrive
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = (t1 - t0) / N;
  if (acc == 0xdeadbeef || m < 0.f) out[7] = acc;
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  for (int threads : {32, 512}) {      // 32: one warp alone (its named barrier spans just itself); 512: 16 warps
    probe<<<1, threads>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%3d threads (%s): try_wait(done) %lld cyc | test_wait(done) %lld | syncwarp+arrive %lld | st+bar.sync(128)+4 ld %lld | proxy fence + tcgen05 fence + syncwarp %lld   [%s]\n",
           threads, threads == 32 ? "one warp alone" : "16 warps, as a CTA of the forward kernel", h[0], h[1], h[2], h[3], h[4], cudaGetErrorString(e));
  }
  return 0;
}

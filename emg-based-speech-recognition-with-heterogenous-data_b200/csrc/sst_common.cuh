// Shared device/host helpers for libsst.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sst.h"

namespace sst {

// ---- error plumbing (thread-local message, negative return codes; include/sst.h) ----------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what, int n_kernels = 1);   // counts launches; cudaPeekAtLastError -> SST_E_LAUNCH

#define SST_REQUIRE(cond, code, ...)                         \
  do { if (!(cond)) { ::sst::set_error(__VA_ARGS__); return (code); } } while (0)

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
int num_sms();

// ---- programmatic dependent launch ---------------------------------------------------------------
// A step is ~440 back-to-back launches on one stream.  Kernels launched through launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining (its CTAs have all passed pdl_trigger() or exited): block scheduling, barrier / TMEM set-up
// and tensor-map prefetch then run under the predecessor's tail instead of after it.  EVERY such kernel calls pdl_wait() before it
// touches global memory (loads, stores, TMA): that returns once the predecessor has completed and its writes are visible -- by
// induction everything earlier in the stream has, too.  Both instructions are no-ops in a kernel launched the ordinary way.
// SST_PDL=0 turns the launch attribute off (bench.py measures both).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- dropout salt -------------------------------------------------------------------------------------
// Dropout seeds cross the C ABI by value, which a captured CUDA graph would replay unchanged.  Every kernel that draws from the
// Philox stream therefore adds *salt (a device uint64 the caller registered with sst_set_dropout_salt, or nothing when there is
// none) to its seed: a replayed graph sees a fresh value the caller wrote in front of the replay (sst_write_scalars).
const unsigned long long* dropout_salt();       // host: the registered device pointer (may be null)
__device__ __forceinline__ unsigned long long salted(unsigned long long seed, const unsigned long long* salt) {
  return salt != nullptr ? seed + *salt : seed;
}

// ---- column partials into a global accumulator ------------------------------------------------------------
// Every block of a column-reduction kernel ends by adding its partial sums to the same C addresses, and same-address atomics are
// served one after the other (~80 ns each: 148 blocks cost ~10 us whatever the input size).  One 16-byte vector reduction
// carries four columns per operation.
__device__ __forceinline__ void red_add4(float* dst, float a, float b, float c, float d) {       // dst 16-byte aligned
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- dtype helpers --------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_as_f32(const void* p, long i, int dtype) {
  return dtype == SST_F32 ? ((const float*)p)[i] : __bfloat162float(((const __nv_bfloat16*)p)[i]);
}
__device__ __forceinline__ void st_from_f32(void* p, long i, int dtype, float v) {
  if (dtype == SST_F32) ((float*)p)[i] = v; else ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}

// ---- warp / block reductions ----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- Philox4x32 counter RNG: the one dropout stream shared by every kernel -------------------------
// keep(seed, idx) is a pure function of (seed, element index) so forward and backward kernels (and the
// SIMT and tensor-core variants of one op) regenerate identical masks without storing them.
// Seven rounds: the smallest Philox4x32 variant that passes BigCrush (Salmon et al., SC'11, "Philox4x32-7");
// the RNG is the largest per-element cost of the fused softmax and of the FFN dropout epilogue, and the three
// extra rounds of the -10 default buy nothing a dropout mask can use.  tests/helpers.py mirrors it on the host.
constexpr int PHILOX_ROUNDS = 7;
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ Philox4 philox4x32(uint64_t seed, uint64_t ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5353542du, c3 = 0x62323030u;
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    // one 32x32 -> 64 product per multiplier (IMAD.WIDE.U32): half the multiplier instructions of mulhi + mullo
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
// keep-threshold: element kept iff rnd >= thr, thr = p * 2^32
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
__host__ __device__ __forceinline__ bool philox_keep(uint64_t seed, uint64_t idx, uint32_t thr) {
  Philox4 r = philox4x32(seed, idx >> 2);
  uint32_t v = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return v >= thr;
}

// 16-bit variant used by the attention kernels (eight keep decisions per Philox block instead of four: the RNG is the
// largest per-element cost of the fused softmax): element kept iff its 16-bit lane >= thr16, thr16 = round(p * 2^16).
__host__ __device__ __forceinline__ uint32_t drop_threshold16(float p) {
  double t = (double)p * 65536.0 + 0.5;
  return t >= 65535.0 ? 0xffffu : (uint32_t)t;
}
__host__ __device__ __forceinline__ uint32_t philox_lane16(const Philox4& r, int sub) {
  const uint32_t w = (sub >> 1) == 0 ? r.x : (sub >> 1) == 1 ? r.y : (sub >> 1) == 2 ? r.z : r.w;
  return (sub & 1) ? (w >> 16) : (w & 0xffffu);
}
__host__ __device__ __forceinline__ bool philox_keep16(uint64_t seed, uint64_t idx, uint32_t thr16) {
  return philox_lane16(philox4x32(seed, idx >> 3), (int)(idx & 7)) >= thr16;
}
// The same decision for a compile-time lane `sub` without extracting it: the high half of a word is compared in place
// against thr16 << 16, the low half after one shift.
__host__ __device__ __forceinline__ bool philox_keep16_at(const Philox4& r, int sub, uint32_t thr16_hi) {
  const uint32_t w = (sub >> 1) == 0 ? r.x : (sub >> 1) == 1 ? r.y : (sub >> 1) == 2 ? r.z : r.w;
  return ((sub & 1) ? w : (w << 16)) >= thr16_hi;
}

}  // namespace sst

// 128-bit vectorised global access: 8 consecutive elements of float or bf16 <-> float[8].
#pragma once
#include "sst_common.cuh"

namespace sst {

template <typename T> struct Vec8;

template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 w = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 w;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = w;
  }
};

__device__ __forceinline__ void load8_f32(const float* p, float (&v)[8]) { Vec8<float>::load(p, v); }

// 8 dropout keep bits for elements idx .. idx+7 (idx % 4 == 0): two Philox blocks.
__device__ __forceinline__ void keep8(unsigned long long seed, unsigned long long idx, uint32_t thr, bool (&k)[8]) {
  Philox4 a = philox4x32(seed, idx >> 2), b = philox4x32(seed, (idx >> 2) + 1);
  k[0] = a.x >= thr; k[1] = a.y >= thr; k[2] = a.z >= thr; k[3] = a.w >= thr;
  k[4] = b.x >= thr; k[5] = b.y >= thr; k[6] = b.z >= thr; k[7] = b.w >= thr;
}

}  // namespace sst

// Per-thread cp.async (LDGSTS) ring: the streaming skeleton of the HBM-bound row kernels (colnorm.cu, norm.cu).
// Every thread copies its own 16-byte pieces of the rows it will process into its own shared-memory slots, S - 1 steps
// ahead; it only ever reads back what it copied itself, so the ring needs no block barrier, and the bytes in flight per SM
// (~100-190 KB) no longer depend on how many registers the kernel's constants and partial sums take.
#pragma once
#include "vec.cuh"

namespace sst {

// ---- per-thread cp.async ring ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
  if (valid) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");   // invalid rows are never read back
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// One 8-element slot: 16 B (bf16) or 32 B (fp32).  Slot (stage, u, t) of a thread lives at
//   ring + stage * stage_bytes + (u * nt + t) * plane + tid * SLOT,     plane = nthreads * SLOT, stage_bytes = U * nt * plane.
template <typename T> struct ColSlot {
  static constexpr int BYTES = (int)sizeof(T) * 8;
  static __device__ __forceinline__ void issue(uint32_t a, const T* g, bool valid) {
    cp_async16(a, g, valid);
    if (BYTES == 32) cp_async16(a + 16, reinterpret_cast<const char*>(g) + 16, valid);
  }
  static __device__ __forceinline__ void read(uint32_t a, float (&v)[8]) {
    if (BYTES == 16) {
      uint4 w;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(a));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    } else {
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a + 16));
    }
  }
};

// The loop every streaming kernel runs: this thread owns rows first + k*stride (k = 0, 1, ...) below `end`, U of them per
// step.  `issue(slot0, u, row, valid)` queues the copies of one row into the step's slots (slot0 = the thread's slot of
// (stage, u = 0, t = 0)), `use(slot0, u, row)` consumes a landed, valid row.  All state is a handful of running
// counters: the loop body is a few instructions beside the kernel's own work.
template <int U, int S, typename Issue, typename Use>
__device__ __forceinline__ void stream_rows(int first, int end, int stride, uint32_t ring, uint32_t stage_bytes, Issue&& issue, Use&& use) {
  const int mine = first < end ? (end - first + stride - 1) / stride : 0;
  const int nsteps = (mine + U - 1) / U;
  const int step_rows = U * stride;
  int rq = first, sq = 0, left = nsteps;
  uint32_t aq = ring;
  auto queue = [&]() {
    if (left > 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = rq + u * stride;
        issue(aq, u, r, r < end);
      }
      --left;
    }
    cp_async_commit();                                  // one group per step, empty ones included: the wait counts groups
    rq += step_rows;
    aq += stage_bytes;
    if (++sq == S) { sq = 0; aq = ring; }
  };
#pragma unroll
  for (int s = 0; s < S - 1; ++s) queue();
  int ru = first, su = 0;
  uint32_t au = ring;
  for (int step = 0; step < nsteps; ++step) {
    queue();
    cp_async_wait<S - 1>();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = ru + u * stride;
      if (r < end) use(au, u, r);
    }
    ru += step_rows;
    au += stage_bytes;
    if (++su == S) { su = 0; au = ring; }
  }
}

}  // namespace sst

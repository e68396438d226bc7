// Warp programming model used by the fused step kernel.
//
// One environment is simulated by one warp.  Code is written as
//   * uniform code     - executed identically by all 32 lanes (scalars in registers),
//   * WARP_FOR(i, n)   - a parallel loop: lane l handles i = l, l+32, ...   (work items must be independent),
//   * warp_sum/max/or  - reductions of a per-lane partial accumulated inside a WARP_FOR,
//   * WARP_SYNC()      - makes shared-memory writes of a phase visible to the next phase.
// The same source also compiles as plain C++ (tests/hostcheck): WARP_FOR becomes a sequential loop and
// the reductions are identities, which executes exactly the same arithmetic per work item.  That build
// exists only so the kernel's mathematics can be debugged without a GPU; it is not part of the package.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define UR3E_HD __host__ __device__ __forceinline__
#define UR3E_D __device__ __forceinline__
// big phase functions: one out-of-line copy each, so the step's code stays resident in the SM's instruction cache
#define UR3E_PHASE __host__ __device__ __noinline__
#else
#define UR3E_HD inline
#define UR3E_D inline
#define UR3E_PHASE inline
#endif

#if defined(__CUDA_ARCH__)
#define UR3E_LANE ((int)(threadIdx.x & 31))
#define WARP_FOR(i, n) for (int i = UR3E_LANE; i < (n); i += 32)
#define WARP_SYNC() __syncwarp()
#define BLOCK_SYNC() __syncthreads()
#define IF_LANE0 if (UR3E_LANE == 0)
#define UR3E_LDG(x) __ldg(&(x))
// per-lane private array (registers); the host build keeps one copy per emulated lane
#define LANE_ARRAY(T, name, N) T name[N]
#define LA(name, lane) name
namespace ur3e {
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
__device__ __forceinline__ int warp_or(int v) { return (int)__reduce_or_sync(0xffffffffu, (unsigned)v); }
__device__ __forceinline__ int popcount32(int v) { return __popc((unsigned)v); }
// index of the n-th (0-based) set bit of v
__device__ __forceinline__ int nth_set_bit(int v, int n) { return (int)__fns((unsigned)v, 0u, n + 1); }
}  // namespace ur3e
#else
#define UR3E_LANE 0
#define WARP_FOR(i, n) for (int i = 0; i < (n); ++i)
#define WARP_SYNC() ((void)0)
#define BLOCK_SYNC() ((void)0)
#define IF_LANE0
#define UR3E_LDG(x) (x)
#define LANE_ARRAY(T, name, N) T name[32][N]
#define LA(name, lane) name[lane]
namespace ur3e {
template <typename T> inline T warp_sum(T v) { return v; }
template <typename T> inline T warp_max(T v) { return v; }
inline int warp_or(int v) { return v; }
inline int popcount32(int v) { return __builtin_popcount((unsigned)v); }
inline int nth_set_bit(int v, int n) { for (int b = 0; b < 32; ++b) if ((v >> b) & 1) { if (n == 0) return b; --n; } return 32; }
}  // namespace ur3e
#endif

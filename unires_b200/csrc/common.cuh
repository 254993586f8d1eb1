// Shared helpers for the unires_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/unires_b200.h"

namespace ur {

void set_error(const char *fmt, ...);
int sm_count();
void count_launch();  // bumps the process-wide kernel launch counter (ur_launch_count)
unsigned long long launches();

#define UR_CUDA_CHECK(expr)                                                         \
  do {                                                                              \
    cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      ur::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                    cudaGetErrorString(e_));                                        \
      return UR_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)
#define UR_LAUNCH_CHECK()                                                           \
  do {                                                                              \
    ur::count_launch();                                                             \
    UR_CUDA_CHECK(cudaGetLastError());                                              \
  } while (0)
#define UR_REQUIRE(cond, ...)                                                       \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ur::set_error(__VA_ARGS__);                                                   \
      return UR_ERR_ARG;                                                            \
    }                                                                               \
  } while (0)

struct Dim3i {
  int x, y, z;
  __host__ __device__ size_t numel() const { return (size_t)x * y * z; }
};
inline Dim3i make_dim(const int32_t d[3]) { return Dim3i{d[0], d[1], d[2]}; }

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------
// L2 eviction priorities (createpolicy + .L2::cache_hint): vectors that are read or written
// once per launch are marked evict_first so that the one vector that is touched again soon
// (the CG residual: written by the residual update, read by the next matvec and the next
// residual update; 67 MB at 256^3 against 126 MB of L2) survives in the cache.
// ---------------------------------------------------------------------------
enum { L2_NORMAL = 0, L2_FIRST = 1, L2_LAST = 2 };
__device__ __forceinline__ uint64_t l2_policy(int kind) {
  uint64_t p;
  if (kind == L2_FIRST)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == L2_LAST)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint4(const float *ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_hint4(float *ptr, const float4 &v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}

// ---------------------------------------------------------------------------
// deterministic float64 reductions
// ---------------------------------------------------------------------------
constexpr int kMaxWarps = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in thread 0.  `sh` holds >= kMaxWarps doubles.
__device__ __forceinline__ double block_sum(double v, double *sh) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, wid = tid >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect sh from a previous use
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = (nthr + 31) >> 5;
  if (wid == 0) {
    v = lane < nw ? sh[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// Two-stage grid reduction with a fixed summation order: every block writes
// its partial to partials[block]; the last block to arrive (atomic ticket)
// re-reads all partials in index order and returns the total in thread 0
// (is_last = true for every thread of that block).  The ticket counter is
// reset by the last block, so the same buffers serve the next launch.
struct GridReduce {
  double *partials;       // >= number of blocks
  unsigned int *counter;  // zero-initialised once
};

__device__ __forceinline__ bool grid_sum(double v, const GridReduce &gr, double *sh,
                                         double *total) {
  __shared__ bool s_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
  double s = block_sum(v, sh);
  if (tid == 0) {
    gr.partials[bid] = s;
    __threadfence();
    unsigned t = atomicAdd(gr.counter, 1u);
    s_last = (t == nblk - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc = 0.0;
  for (unsigned i = tid; i < nblk; i += nthr) acc += __ldcg(gr.partials + i);
  acc = block_sum(acc, sh);
  if (tid == 0) {
    *total = acc;
    *gr.counter = 0u;
  }
  return true;
}

}  // namespace ur

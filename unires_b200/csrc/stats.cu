// Intensity statistics for the hyper-parameter estimate (unires/_core.py:96-142): range and
// histogram of an observed volume, with the selections the reference makes by boolean-mask
// indexing (dat[dat >= 0], dat != 0, dat != max) folded into the passes.  HBM-bound, 4 B/voxel.
#include <math.h>

#include "common.cuh"

namespace ur {

// order-preserving map float -> uint32 (so that integer atomicMin / atomicMax order floats)
__device__ __forceinline__ unsigned enc(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float dec(unsigned e) {
  const unsigned u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// The reference's selections (nitorch.tools.img_statistics.estimate_noise after UniRes'
// dat[dat >= 0]): non-finite values count as 0; optionally v >= 0 only; optionally v != 0 and
// v != mask_value (the global maximum).
__device__ __forceinline__ bool select(float &v, int drop_negative, int mask, float mask_value) {
  if (isnan(v)) {
    if (drop_negative) return false;  // NaN >= 0 is false
    v = 0.f;
  } else if (isinf(v)) {
    if (drop_negative && v < 0.f) return false;
    v = 0.f;
  }
  if (drop_negative && !(v >= 0.f)) return false;
  if (mask && (v == 0.f || v == mask_value)) return false;
  return true;
}

__global__ void __launch_bounds__(256)
    range_kernel(const float *__restrict__ dat, size_t n, int drop_negative, int mask,
                 float mask_value, unsigned *out) {
  unsigned lo = 0xffffffffu, hi = 0u;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = dat[i];
    if (!select(v, drop_negative, mask, mask_value)) continue;
    const unsigned e = enc(v);
    lo = min(lo, e);
    hi = max(hi, e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (lo != 0xffffffffu) atomicMin(out, lo);
    if (hi != 0u) atomicMax(out + 1, hi);
  }
}

constexpr int kMaxBins = 4096;

__global__ void __launch_bounds__(256)
    histc_kernel(const float *__restrict__ dat, size_t n, int drop_negative, int mask,
                 float mask_value, double mn, double mx, int bins, unsigned long long *counts) {
  extern __shared__ unsigned s_hist[];
  for (int b = threadIdx.x; b < bins; b += blockDim.x) s_hist[b] = 0u;
  __syncthreads();
  const double width = mx - mn;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = dat[i];
    if (!select(v, drop_negative, mask, mask_value)) continue;
    const double d = (double)v;
    if (d < mn || d > mx) continue;  // torch.histc ignores out-of-range elements
    // torch.histc: bin = (int)((v - min) * bins / (max - min)), v == max in the last bin
    int b = (int)((d - mn) * (double)bins / width);  // same roundings as torch.histc
    if (b >= bins) b = bins - 1;
    atomicAdd(&s_hist[b], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x)
    if (s_hist[b]) atomicAdd(&counts[b], (unsigned long long)s_hist[b]);
}

}  // namespace ur

using namespace ur;

extern "C" int ur_intensity_range(const float *d_dat, size_t n, int drop_negative, int mask,
                                  float mask_value, float h_out[2], int32_t *h_any,
                                  ur_stream stream) {
  UR_REQUIRE(d_dat && h_out && n > 0, "ur_intensity_range: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned *d_enc = nullptr;
  UR_CUDA_CHECK(cudaMallocAsync((void **)&d_enc, 2 * sizeof(unsigned), st));
  const unsigned init[2] = {0xffffffffu, 0u};
  UR_CUDA_CHECK(cudaMemcpyAsync(d_enc, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const unsigned blocks = (unsigned)min((size_t)sm_count() * 8, (n + 255) / 256);
  range_kernel<<<blocks, 256, 0, st>>>(d_dat, n, drop_negative, mask, mask_value, d_enc);
  UR_LAUNCH_CHECK();
  unsigned h_enc[2];
  UR_CUDA_CHECK(cudaMemcpyAsync(h_enc, d_enc, sizeof(h_enc), cudaMemcpyDeviceToHost, st));
  UR_CUDA_CHECK(cudaStreamSynchronize(st));
  UR_CUDA_CHECK(cudaFreeAsync(d_enc, st));
  const bool any = h_enc[0] != 0xffffffffu;
  if (h_any) *h_any = any ? 1 : 0;
  h_out[0] = any ? dec(h_enc[0]) : 0.f;
  h_out[1] = any ? dec(h_enc[1]) : 0.f;
  return UR_OK;
}

extern "C" int ur_histc(const float *d_dat, size_t n, int drop_negative, int mask,
                        float mask_value, double mn, double mx, int bins,
                        unsigned long long *d_counts, ur_stream stream) {
  UR_REQUIRE(d_dat && d_counts && n > 0, "ur_histc: bad args");
  UR_REQUIRE(bins > 0 && bins <= kMaxBins, "ur_histc: bins must be in 1..%d", kMaxBins);
  UR_REQUIRE(mx > mn, "ur_histc: empty range [%g, %g]", mn, mx);
  cudaStream_t st = (cudaStream_t)stream;
  UR_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, (size_t)bins * sizeof(unsigned long long), st));
  const unsigned blocks = (unsigned)min((size_t)sm_count() * 8, (n + 255) / 256);
  histc_kernel<<<blocks, 256, (size_t)bins * sizeof(unsigned), st>>>(
      d_dat, n, drop_negative, mask, mask_value, mn, mx, bins, d_counts);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

// Slice-profile convolution along one axis (the separable factors of
// F.conv3d / F.conv_transpose3d with stride = ratio, unires/_project.py:153-154)
// and the even/odd slice scaling (_apply_scaling, unires/_project.py:9-24).
#include "common.cuh"

namespace ur {

struct Taps {
  float k[UR_MAX_TAPS];
};

// out[j] = sum_t k[t] * in[j*stride + t]              (valid cross-correlation)
// TRANSPOSE: out[i] = sum_{j, t = i - j*stride in [0,K)} k[t] * in[j]
template <bool TRANSPOSE>
__global__ void conv_axis_kernel(const float *__restrict__ in, float *__restrict__ out,
                                 Dim3i idim, Dim3i odim, int axis, Taps taps, int K, int stride) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (z >= odim.z || y >= odim.y) return;
  const size_t isy = idim.z, isx = (size_t)idim.y * idim.z;
  const int pos = axis == 0 ? x : (axis == 1 ? y : z);         // output index along axis
  const size_t step = axis == 0 ? isx : (axis == 1 ? isy : 1);  // input stride along axis
  const int n_in = axis == 0 ? idim.x : (axis == 1 ? idim.y : idim.z);
  // input offset with the axis coordinate zeroed
  const size_t base = (axis == 0 ? 0 : x * isx) + (axis == 1 ? 0 : y * isy) + (axis == 2 ? 0 : z);
  float acc = 0.f;
  if (!TRANSPOSE) {
    const float *p = in + base + (size_t)pos * stride * step;
    for (int t = 0; t < K; ++t) acc = fmaf(taps.k[t], __ldg(p + t * step), acc);
  } else {
    // j range: 0 <= pos - j*stride < K
    int j_hi = pos / stride;
    if (j_hi > n_in - 1) j_hi = n_in - 1;
    int j_lo = (pos - K + stride) / stride;  // ceil((pos-K+1)/stride) for pos-K+1 >= 0
    if (pos - K + 1 <= 0) j_lo = 0;
    for (int j = j_lo; j <= j_hi; ++j)
      acc = fmaf(taps.k[pos - j * stride], __ldg(in + base + (size_t)j * step), acc);
  }
  out[((size_t)x * odim.y + y) * odim.z + z] = acc;
}

// Same arithmetic for a conv axis other than z, marching: a thread owns one (z, other-axis)
// line and produces a chunk of consecutive outputs along the conv axis, so the K (or ~K/stride)
// inputs an output needs are re-read by the SAME thread from L1 instead of by K different
// blocks from L2; z stays the coalesced direction.
constexpr int kConvChunk = 32;

template <bool TRANSPOSE>
__global__ void __launch_bounds__(256)
    conv_march_kernel(const float *__restrict__ in, float *__restrict__ out, Dim3i idim,
                      Dim3i odim, int axis, Taps taps, int K, int stride) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int oth = blockIdx.y * blockDim.y + threadIdx.y;  // y (axis 0) or x (axis 1)
  const int n_oth = axis == 0 ? odim.y : odim.x;
  if (z >= odim.z || oth >= n_oth) return;
  const int n_out = axis == 0 ? odim.x : odim.y;
  const int n_in = axis == 0 ? idim.x : idim.y;
  const size_t istep = axis == 0 ? (size_t)idim.y * idim.z : (size_t)idim.z;
  const size_t ostep = axis == 0 ? (size_t)odim.y * odim.z : (size_t)odim.z;
  const size_t ibase = (axis == 0 ? (size_t)oth * idim.z : (size_t)oth * idim.y * idim.z) + z;
  const size_t obase = (axis == 0 ? (size_t)oth * odim.z : (size_t)oth * odim.y * odim.z) + z;
  const int p0 = blockIdx.z * kConvChunk;
  const int p1 = p0 + kConvChunk < n_out ? p0 + kConvChunk : n_out;
  for (int pos = p0; pos < p1; ++pos) {
    float acc = 0.f;
    if (!TRANSPOSE) {
      const float *p = in + ibase + (size_t)pos * stride * istep;
      for (int t = 0; t < K; ++t) acc = fmaf(taps.k[t], __ldg(p + t * istep), acc);
    } else {
      int j_hi = pos / stride;
      if (j_hi > n_in - 1) j_hi = n_in - 1;
      int j_lo = (pos - K + stride) / stride;
      if (pos - K + 1 <= 0) j_lo = 0;
      for (int j = j_lo; j <= j_hi; ++j)
        acc = fmaf(taps.k[pos - j * stride], __ldg(in + ibase + (size_t)j * istep), acc);
    }
    out[obase + (size_t)pos * ostep] = acc;
  }
}

__global__ void scaling_kernel(const float *__restrict__ in, float *__restrict__ out, Dim3i d,
                               float even, float odd, int axis) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (z >= d.z || y >= d.y) return;
  const size_t i = ((size_t)x * d.y + y) * d.z + z;
  const int pos = axis == 0 ? x : (axis == 1 ? y : z);
  out[i] = ((pos & 1) ? odd : even) * in[i];
}

int conv_axis(const float *in, Dim3i idim, float *out, int axis, const float *ker, int K,
              int stride, bool transpose, cudaStream_t st, Dim3i *odim_out) {
  Dim3i od = idim;
  int *n = axis == 0 ? &od.x : (axis == 1 ? &od.y : &od.z);
  if (!transpose) {
    if (*n < K) {
      set_error("conv_axis: extent %d shorter than kernel %d", *n, K);
      return UR_ERR_ARG;
    }
    *n = (*n - K) / stride + 1;
  } else {
    *n = (*n - 1) * stride + K;
  }
  if (odim_out) *odim_out = od;
  Taps taps;
  for (int t = 0; t < UR_MAX_TAPS; ++t) taps.k[t] = t < K ? ker[t] : 0.f;
  if (axis != 2) {
    const int n_oth = axis == 0 ? od.y : od.x, n_out = axis == 0 ? od.x : od.y;
    dim3 block(64, 4, 1), grid(div_up(od.z, 64), div_up(n_oth, 4), div_up(n_out, kConvChunk));
    if (transpose)
      conv_march_kernel<true><<<grid, block, 0, st>>>(in, out, idim, od, axis, taps, K, stride);
    else
      conv_march_kernel<false><<<grid, block, 0, st>>>(in, out, idim, od, axis, taps, K, stride);
    UR_LAUNCH_CHECK();
    return UR_OK;
  }
  dim3 block(64, 4, 1), grid(div_up(od.z, 64), div_up(od.y, 4), od.x);
  if (transpose)
    conv_axis_kernel<true><<<grid, block, 0, st>>>(in, out, idim, od, axis, taps, K, stride);
  else
    conv_axis_kernel<false><<<grid, block, 0, st>>>(in, out, idim, od, axis, taps, K, stride);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int apply_scaling(const float *in, float *out, Dim3i d, float scl, int axis, cudaStream_t st) {
  dim3 block(64, 4, 1), grid(div_up(d.z, 64), div_up(d.y, 4), d.x);
  scaling_kernel<<<grid, block, 0, st>>>(in, out, d, expf(scl), expf(-scl), axis);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

using namespace ur;

extern "C" int ur_conv_axis(const float *d_in, const int32_t idim[3], float *d_out, int axis,
                            const float *ker, int K, int stride, int transpose,
                            ur_stream stream) {
  UR_REQUIRE(d_in && d_out && idim && ker, "ur_conv_axis: null argument");
  UR_REQUIRE(axis >= 0 && axis < 3, "ur_conv_axis: axis must be 0, 1 or 2");
  UR_REQUIRE(K >= 1 && K <= UR_MAX_TAPS, "ur_conv_axis: kernel length %d not in [1,%d]", K,
             UR_MAX_TAPS);
  UR_REQUIRE(stride >= 1, "ur_conv_axis: stride must be >= 1");
  return conv_axis(d_in, make_dim(idim), d_out, axis, ker, K, stride, transpose != 0,
                   (cudaStream_t)stream, nullptr);
}

extern "C" int ur_scale_slices(const float *d_in, float *d_out, const int32_t dim[3], float f_even,
                               float f_odd, int axis, ur_stream stream) {
  UR_REQUIRE(d_in && d_out && dim, "ur_scale_slices: null argument");
  UR_REQUIRE(axis >= 0 && axis < 3, "ur_scale_slices: axis must be 0, 1 or 2");
  const Dim3i d = make_dim(dim);
  dim3 block(64, 4, 1), grid(div_up(d.z, 64), div_up(d.y, 4), d.x);
  scaling_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_in, d_out, d, f_even, f_odd, axis);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_apply_scaling(const float *d_in, float *d_out, const int32_t dim[3], float scl,
                                int axis, ur_stream stream) {
  UR_REQUIRE(d_in && d_out && dim, "ur_apply_scaling: null argument");
  UR_REQUIRE(axis >= 0 && axis < 3, "ur_apply_scaling: axis must be 0, 1 or 2");
  return apply_scaling(d_in, d_out, make_dim(dim), scl, axis, (cudaStream_t)stream);
}

// Library plumbing (error string, device info, scratch) and the small vector /
// reduction kernels: dot, axpy, the objective's data term and sqrt-sum
// (_compute_nll, unires/_update.py:396-427).
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "solver.cuh"

namespace ur {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Scratch for the stand-alone reductions (ur_dot, ur_cg_update_xr, ur_nll_data, ur_sqrt_sum,
// ur_scaling_sums, ur_rigid_sums): one ticket counter + partials buffer per (device, stream),
// so reductions in flight on different streams (sett.channel_streams) never share a counter.
constexpr size_t kScratchPartials = 1 << 16;
int scratch_reduce(GridReduce *gr, cudaStream_t st) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, void *> bufs;
  int dev = 0;
  UR_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  void *&slot = bufs[std::make_pair(dev, st)];
  if (!slot) {
    void *p = nullptr;
    UR_CUDA_CHECK(cudaMalloc(&p, 256 + kScratchPartials * sizeof(double)));
    UR_CUDA_CHECK(cudaMemset(p, 0, 256));
    slot = p;
  }
  gr->counter = (unsigned *)slot;
  gr->partials = (double *)((char *)slot + 256);
  return UR_OK;
}

static unsigned red_blocks(size_t n) {
  size_t want = (n + 1023) / 1024;
  const size_t cap = (size_t)sm_count() * 8;
  if (want > cap) want = cap;
  return (unsigned)(want ? want : 1);
}

// kind 0: sum a*b ; 1: 0.5*tau*sum_{a != 0} (a-b)^2 ; 2: sum sqrt(a)
template <int KIND>
__global__ void __launch_bounds__(256)
    reduce_kernel(const float *__restrict__ a, const float *__restrict__ b, size_t n, float tau,
                  int accumulate, GridReduce gr, double *out) {
  __shared__ double s_red[kMaxWarps];
  double part = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (KIND == 0) {
      part += (double)__fmul_rn(a[i], b[i]);
    } else if (KIND == 1) {
      const float av = a[i];
      if (av != 0.f) {
        const float d = __fsub_rn(av, b[i]);
        part += (double)__fmul_rn(d, d);
      }
    } else {
      part += (double)sqrtf(a[i]);
    }
  }
  double total;
  if (grid_sum(part, gr, s_red, &total) && threadIdx.x == 0) {
    if (KIND == 1) total = 0.5 * (double)tau * total;
    *out = accumulate ? *out + total : total;
  }
}

__global__ void axpy_kernel(float *__restrict__ y, const float *__restrict__ x, float a, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __fadd_rn(y[i], __fmul_rn(a, x[i]));
}

}  // namespace ur

using namespace ur;

extern "C" const char *ur_last_error(void) { return g_err; }
extern "C" int ur_version(void) { return 100; }
extern "C" uint64_t ur_launch_count(void) { return launches(); }

extern "C" int ur_device_info(int *sm, int *major, int *minor) {
  int dev = 0;
  UR_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  UR_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (sm) *sm = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return UR_OK;
}

extern "C" int ur_dot(const float *d_a, const float *d_b, size_t n, double *d_out,
                      ur_stream stream) {
  UR_REQUIRE(d_a && d_b && d_out && n > 0, "ur_dot: bad args");
  GridReduce gr;
  int rc = scratch_reduce(&gr, (cudaStream_t)stream);
  if (rc) return rc;
  reduce_kernel<0><<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, n, 0.f, 0, gr, d_out);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_nll_data(const float *d_x, const float *d_Ay, size_t n, float tau,
                           double *d_out, int accumulate, ur_stream stream) {
  UR_REQUIRE(d_x && d_Ay && d_out && n > 0, "ur_nll_data: bad args");
  GridReduce gr;
  int rc = scratch_reduce(&gr, (cudaStream_t)stream);
  if (rc) return rc;
  reduce_kernel<1><<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_x, d_Ay, n, tau, accumulate,
                                                                    gr, d_out);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

// Masked even/odd sums of the slice-scaling update (unires/_update.py:303-335): over the
// voxels with x != 0, with p = parity of the index along `axis`,
//   out[0] = sum (x - y)^2, out[1 + p] = sum y (x - y), out[3 + p] = sum y^2
// float32 products accumulated in float64 like torch.sum(..., dtype=float64); deterministic
// two-stage reduction (per-block partials, the last block sums them in index order).
__global__ void __launch_bounds__(256)
    scaling_sums_kernel(const float *__restrict__ x, const float *__restrict__ y, Dim3i d,
                        int axis, GridReduce gr, double *out) {
  __shared__ double s_red[5][kMaxWarps];
  __shared__ bool s_last;
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const size_t n = d.numel(), stride = (size_t)gridDim.x * blockDim.x;
  const size_t sx = (size_t)d.y * d.z;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float xv = x[i];
    if (xv == 0.f) continue;
    const float yv = y[i], df = __fsub_rn(xv, yv);
    const int pos = axis == 0 ? (int)(i / sx) : (axis == 1 ? (int)((i / d.z) % d.y) : (int)(i % d.z));
    const int p = pos & 1;
    part[0] += (double)__fmul_rn(df, df);
    part[1 + p] += (double)__fmul_rn(yv, df);
    part[3 + p] += (double)__fmul_rn(yv, yv);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double v = warp_sum(part[k]);
    if (lane == 0) s_red[k][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s_red[threadIdx.x][w];
    gr.partials[blockIdx.x * 5 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(gr.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += __ldcg(gr.partials + b * 5 + threadIdx.x);
    out[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *gr.counter = 0u;
}

extern "C" int ur_scaling_sums(const float *d_x, const float *d_y, const int32_t dim[3], int axis,
                               double *d_out, ur_stream stream) {
  UR_REQUIRE(d_x && d_y && d_out && dim && dim[0] > 0 && dim[1] > 0 && dim[2] > 0,
             "ur_scaling_sums: bad args");
  UR_REQUIRE(axis >= 0 && axis < 3, "ur_scaling_sums: axis must be 0, 1 or 2");
  GridReduce gr;
  int rc = scratch_reduce(&gr, (cudaStream_t)stream);
  if (rc) return rc;
  const Dim3i d = make_dim(dim);
  scaling_sums_kernel<<<red_blocks(d.numel()), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, d, axis,
                                                                               gr, d_out);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

// Gradient and Gauss-Newton Hessian of the rigid matching term w.r.t. the 6 Lie parameters
// (unires/_update.py:616-640).  Per voxel of the intermediate grid: g = spatial gradient of the
// warped recon (3), res = C'(A y - x) (masked), ctc = C'C 1; gr_m = g res, Hes_m = (g_a g_b) ctc.
// With dA[i][d] = ((m0 ix + m1 iy) + m2 iz) + m3, the float32 image of d(coordinate d)/d q_i
// (dm = 6 x 3 x 4 floats, rows of mat_y \ dR_i mat):
//   out[i]            = sum_d sum_vox gr_m[d] dA[i][d]                          (i < 6)
//   out[6 + tri(i,j)] = sum_{d1,d2} sum_vox (Hes_m[d1,d2] dA[i][d1]) dA[j][d2]   (i <= j)
// float32 products summed in float64, deterministic two-stage reduction.
struct RigidDm {
  float m[6][3][4];
};

__global__ void __launch_bounds__(256)
    rigid_sums_kernel(const float *__restrict__ g, const float *__restrict__ res,
                      const float *__restrict__ ctc, Dim3i d, RigidDm dm, GridReduce gr,
                      double *out) {
  constexpr int NS = 27;
  __shared__ double s_red[NS][kMaxWarps];
  __shared__ bool s_last;
  double part[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) part[k] = 0.0;
  const size_t n = d.numel(), stride = (size_t)gridDim.x * blockDim.x;
  const size_t sx = (size_t)d.y * d.z;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < n; v += stride) {
    const float fx = (float)(v / sx), fy = (float)((v / d.z) % d.y), fz = (float)(v % d.z);
    const float gv[3] = {g[3 * v], g[3 * v + 1], g[3 * v + 2]};
    const float rv = res[v], cv = ctc ? ctc[v] : 1.f;
    float dA[6][3];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int a = 0; a < 3; ++a)
        dA[i][a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dm.m[i][a][0], fx),
                                                 __fmul_rn(dm.m[i][a][1], fy)),
                                       __fmul_rn(dm.m[i][a][2], fz)),
                             dm.m[i][a][3]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gm = __fmul_rn(gv[a], rv);
#pragma unroll
      for (int i = 0; i < 6; ++i) part[i] += (double)__fmul_rn(gm, dA[i][a]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float h = __fmul_rn(gv[a < b ? a : b], gv[a < b ? b : a]);
        if (ctc) h = __fmul_rn(h, cv);
        int t = 6;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float left = __fmul_rn(h, dA[i][a]);
#pragma unroll
          for (int j = i; j < 6; ++j) part[t++] += (double)__fmul_rn(left, dA[j][b]);
        }
      }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const double w = warp_sum(part[k]);
    if (lane == 0) s_red[k][wid] = w;
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double w = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) w += s_red[threadIdx.x][q];
    gr.partials[blockIdx.x * NS + threadIdx.x] = w;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(gr.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < NS) {
    double w = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) w += __ldcg(gr.partials + b * NS + threadIdx.x);
    out[threadIdx.x] = w;
  }
  if (threadIdx.x == 0) *gr.counter = 0u;
}

extern "C" int ur_rigid_sums(const float *d_grad, const float *d_res, const float *d_ctc,
                             const int32_t dim[3], const float dm[72], double *d_out,
                             ur_stream stream) {
  UR_REQUIRE(d_grad && d_res && dm && d_out && dim && dim[0] > 0 && dim[1] > 0 && dim[2] > 0,
             "ur_rigid_sums: bad args");
  GridReduce gr;
  int rc = scratch_reduce(&gr, (cudaStream_t)stream);
  if (rc) return rc;
  RigidDm m;
  for (int k = 0; k < 72; ++k) (&m.m[0][0][0])[k] = dm[k];
  const Dim3i d = make_dim(dim);
  unsigned nb = red_blocks(d.numel());
  if (nb > 2048) nb = 2048;  // 27 partials per block in the shared scratch
  rigid_sums_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(d_grad, d_res, d_ctc, d, m, gr, d_out);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_sqrt_sum(const float *d_e, size_t n, double *d_out, ur_stream stream) {
  UR_REQUIRE(d_e && d_out && n > 0, "ur_sqrt_sum: bad args");
  GridReduce gr;
  int rc = scratch_reduce(&gr, (cudaStream_t)stream);
  if (rc) return rc;
  reduce_kernel<2><<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_e, nullptr, n, 0.f, 0, gr,
                                                                    d_out);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_axpy(float *d_y, const float *d_x, float a, size_t n, ur_stream stream) {
  UR_REQUIRE(d_y && d_x && n > 0, "ur_axpy: bad args");
  axpy_kernel<<<red_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_y, d_x, a, n);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

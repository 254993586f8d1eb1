// Finite-difference operators: im_gradient, im_divergence, DtD, ADMM RHS,
// nll prior energy.  Forward differences, zero bound:
//   grad_a[i] = (d[i+e_a] - d[i]) / vx_a,  d = 0 past the high edge
//   div[i]    = sum_a (v_a[i-e_a] - v_a[i]) / vx_a,  v_a = 0 before the low edge
// Divisions by vx are done as multiplications by 1/vx (exact for the
// power-of-two voxel sizes of every BASELINE config).
#include "common.cuh"

namespace ur {

struct DiffGeom {
  int nx, ny, nz;
  float ivx, ivy, ivz;
};

static inline DiffGeom make_geom(const int32_t dim[3], const float vx[3]) {
  return DiffGeom{dim[0], dim[1], dim[2], 1.0f / vx[0], 1.0f / vx[1], 1.0f / vx[2]};
}

static inline void launch_shape(const DiffGeom &g, dim3 &grid, dim3 &block) {
  block = dim3(64, 4, 1);
  grid = dim3(div_up(g.nz, 64), div_up(g.ny, 4), g.nx);
}

#define UR_VOXEL_PROLOGUE                                            \
  const int z = blockIdx.x * blockDim.x + threadIdx.x;               \
  const int y = blockIdx.y * blockDim.y + threadIdx.y;               \
  const int x = blockIdx.z;                                          \
  if (z >= g.nz || y >= g.ny) return;                                \
  const size_t sy = g.nz, sx = (size_t)g.ny * g.nz;                  \
  const size_t i = x * sx + y * sy + z;

__global__ void gradient_kernel(const float *__restrict__ d, float *__restrict__ out,
                                DiffGeom g) {
  UR_VOXEL_PROLOGUE
  const size_t n = sx * g.nx;
  const float c = d[i];
  const float xp = x + 1 < g.nx ? d[i + sx] : 0.f;
  const float yp = y + 1 < g.ny ? d[i + sy] : 0.f;
  const float zp = z + 1 < g.nz ? d[i + 1] : 0.f;
  out[i] = (xp - c) * g.ivx;
  out[n + i] = (yp - c) * g.ivy;
  out[2 * n + i] = (zp - c) * g.ivz;
}

__device__ __forceinline__ float div_at(const float *__restrict__ v, size_t n, size_t i,
                                        int x, int y, int z, size_t sx, size_t sy,
                                        const DiffGeom &g) {
  const float *v0 = v, *v1 = v + n, *v2 = v + 2 * n;
  const float t0 = ((x > 0 ? v0[i - sx] : 0.f) - v0[i]) * g.ivx;
  const float t1 = ((y > 0 ? v1[i - sy] : 0.f) - v1[i]) * g.ivy;
  const float t2 = ((z > 0 ? v2[i - 1] : 0.f) - v2[i]) * g.ivz;
  return (t0 + t1) + t2;
}

__global__ void divergence_kernel(const float *__restrict__ v, float *__restrict__ out,
                                  DiffGeom g) {
  UR_VOXEL_PROLOGUE
  out[i] = div_at(v, sx * g.nx, i, x, y, z, sx, sy, g);
}

// DtD in one pass: per axis  g[i] = (d[i+1]-d[i])/vx (d past the end = 0),
// t = (g[i-1] (0 if i == 0) - g[i]) / vx
__device__ __forceinline__ float dtd_axis(float lo, float c, float hi, bool has_lo,
                                          float iv) {
  const float gi = (hi - c) * iv;
  const float gm = has_lo ? (c - lo) * iv : 0.f;
  return (gm - gi) * iv;
}

__global__ void dtd_kernel(const float *__restrict__ d, float *__restrict__ out, DiffGeom g) {
  UR_VOXEL_PROLOGUE
  const float c = d[i];
  const float xm = x > 0 ? d[i - sx] : 0.f, xp = x + 1 < g.nx ? d[i + sx] : 0.f;
  const float ym = y > 0 ? d[i - sy] : 0.f, yp = y + 1 < g.ny ? d[i + sy] : 0.f;
  const float zm = z > 0 ? d[i - 1] : 0.f, zp = z + 1 < g.nz ? d[i + 1] : 0.f;
  const float t0 = dtd_axis(xm, c, xp, x > 0, g.ivx);
  const float t1 = dtd_axis(ym, c, yp, y > 0, g.ivy);
  const float t2 = dtd_axis(zm, c, zp, z > 0, g.ivz);
  out[i] = (t0 + t1) + t2;
}

// b -= lam * div(w - rho z)      (unires/_update.py:131-133)
__global__ void admm_rhs_kernel(float *__restrict__ b, const float *__restrict__ w,
                                const float *__restrict__ zz, DiffGeom g, float lam,
                                float rho) {
  UR_VOXEL_PROLOGUE
  const size_t n = sx * g.nx;
  auto q = [&](int a, size_t j) { return w[a * n + j] - rho * zz[a * n + j]; };
  const float t0 = ((x > 0 ? q(0, i - sx) : 0.f) - q(0, i)) * g.ivx;
  const float t1 = ((y > 0 ? q(1, i - sy) : 0.f) - q(1, i)) * g.ivy;
  const float t2 = ((z > 0 ? q(2, i - 1) : 0.f) - q(2, i)) * g.ivz;
  b[i] -= lam * ((t0 + t1) + t2);
}

}  // namespace ur

using namespace ur;

extern "C" int ur_im_gradient(const float *d_dat, float *d_grad, const int32_t dim[3],
                              const float vx[3], ur_stream stream) {
  UR_REQUIRE(d_dat && d_grad && dim[0] > 0 && dim[1] > 0 && dim[2] > 0, "ur_im_gradient: bad args");
  DiffGeom g = make_geom(dim, vx);
  dim3 grid, block;
  launch_shape(g, grid, block);
  gradient_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_dat, d_grad, g);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_im_divergence(const float *d_vec, float *d_div, const int32_t dim[3],
                                const float vx[3], ur_stream stream) {
  UR_REQUIRE(d_vec && d_div && dim[0] > 0 && dim[1] > 0 && dim[2] > 0, "ur_im_divergence: bad args");
  DiffGeom g = make_geom(dim, vx);
  dim3 grid, block;
  launch_shape(g, grid, block);
  divergence_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_vec, d_div, g);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_dtd(const float *d_dat, float *d_out, const int32_t dim[3],
                      const float vx[3], ur_stream stream) {
  UR_REQUIRE(d_dat && d_out && dim[0] > 0 && dim[1] > 0 && dim[2] > 0, "ur_dtd: bad args");
  DiffGeom g = make_geom(dim, vx);
  dim3 grid, block;
  launch_shape(g, grid, block);
  dtd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_dat, d_out, g);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_admm_rhs(float *d_b, const float *d_w, const float *d_z,
                           const int32_t dim[3], const float vx[3], float lam, float rho,
                           ur_stream stream) {
  UR_REQUIRE(d_b && d_w && d_z && dim[0] > 0 && dim[1] > 0 && dim[2] > 0, "ur_admm_rhs: bad args");
  DiffGeom g = make_geom(dim, vx);
  dim3 grid, block;
  launch_shape(g, grid, block);
  admm_rhs_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_b, d_w, d_z, g, lam, rho);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

// Trilinear / nearest resampling with zero bound: nitorch grid_pull and its
// exact transpose grid_push.  Coordinates come
// either from a dense (ox,oy,oz,3) grid or are evaluated in-kernel from a 3x4
// affine matrix in float32, the same arithmetic type the reference uses when
// it materialises affine_grid (unires/_project.py:159).
#include "common.cuh"

namespace ur {

constexpr float kFovTol = 5e-2f;  // nitorch's in-FOV tolerance when extrapolate=False

struct Affine {
  float m[12];
};

struct CoordDense {
  const float *grid;
  __device__ __forceinline__ void get(int i, int j, int k, size_t lin, float &cx, float &cy,
                                      float &cz) const {
    cx = grid[3 * lin + 0];
    cy = grid[3 * lin + 1];
    cz = grid[3 * lin + 2];
  }
};

struct CoordAffine {
  Affine a;
  __device__ __forceinline__ void get(int i, int j, int k, size_t lin, float &cx, float &cy,
                                      float &cz) const {
    const float fi = (float)i, fj = (float)j, fk = (float)k;
    cx = fmaf(a.m[2], fk, fmaf(a.m[1], fj, a.m[0] * fi)) + a.m[3];
    cy = fmaf(a.m[6], fk, fmaf(a.m[5], fj, a.m[4] * fi)) + a.m[7];
    cz = fmaf(a.m[10], fk, fmaf(a.m[9], fj, a.m[8] * fi)) + a.m[11];
  }
};

__device__ __forceinline__ bool in_fov(float cx, float cy, float cz, const Dim3i &s) {
  return cx > -kFovTol && cx < (float)(s.x - 1) + kFovTol && cy > -kFovTol &&
         cy < (float)(s.y - 1) + kFovTol && cz > -kFovTol && cz < (float)(s.z - 1) + kFovTol;
}

// PUSH = false: out[o] = sum_corners w * src[corner]
// PUSH = true : dst[corner] += scale * w * in[o]        (atomic)
template <class Coord, bool PUSH>
__global__ void resample_kernel(const float *__restrict__ in, float *__restrict__ out, Dim3i s,
                                Dim3i o, Coord coord, int order, int extrapolate, float scale) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= o.z || j >= o.y) return;
  const size_t lin = ((size_t)i * o.y + j) * o.z + k;
  float cx, cy, cz;
  coord.get(i, j, k, lin, cx, cy, cz);
  const bool ok = extrapolate || in_fov(cx, cy, cz, s);
  float val = 0.f;
  if (PUSH) val = scale * in[lin];
  if (!ok || (PUSH && val == 0.f)) {
    if (!PUSH) out[lin] = 0.f;
    return;
  }
  const size_t sy = s.z, sx = (size_t)s.y * s.z;
  if (order == 0) {
    const int ix = (int)floorf(cx + 0.5f), iy = (int)floorf(cy + 0.5f),
              iz = (int)floorf(cz + 0.5f);
    const bool inside = ix >= 0 && ix < s.x && iy >= 0 && iy < s.y && iz >= 0 && iz < s.z;
    if (PUSH) {
      if (inside) atomicAdd(out + ix * sx + iy * sy + iz, val);
    } else {
      out[lin] = inside ? __ldg(in + ix * sx + iy * sy + iz) : 0.f;
    }
    return;
  }
  const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const float wx1 = cx - fx, wy1 = cy - fy, wz1 = cz - fz;
  const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int bx = (c >> 2) & 1, by = (c >> 1) & 1, bz = c & 1;
    const int px = ix + bx, py = iy + by, pz = iz + bz;
    if (px < 0 || px >= s.x || py < 0 || py >= s.y || pz < 0 || pz >= s.z) continue;
    // same association as the oracle: (wx * wy) * wz
    const float wgt = ((bx ? wx1 : wx0) * (by ? wy1 : wy0)) * (bz ? wz1 : wz0);
    const size_t q = px * sx + py * sy + pz;
    if (PUSH)
      atomicAdd(out + q, val * wgt);
    else
      acc += __ldg(in + q) * wgt;
  }
  if (!PUSH) out[lin] = acc;
}

__global__ void affine_grid_kernel(float *__restrict__ grid, Dim3i o, CoordAffine coord) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= o.z || j >= o.y) return;
  const size_t lin = ((size_t)i * o.y + j) * o.z + k;
  float cx, cy, cz;
  coord.get(i, j, k, lin, cx, cy, cz);
  grid[3 * lin + 0] = cx;
  grid[3 * lin + 1] = cy;
  grid[3 * lin + 2] = cz;
}

// Gather form of the trilinear push for an AFFINE grid (deterministic, no atomics): a thread
// owns one TARGET voxel q and enumerates the source voxels p whose mapped coordinate
// c(p) = A p + t falls strictly within one voxel of q on every axis -- the preimage of that
// cube is contained in the box  A^-1 (q - t) +- sum_b |A^-1[a][b]|.  Coordinates, FOV test and
// corner weights are evaluated exactly like the scatter form (same float32 expressions), so
// the two differ only in summation order.  dst[q] += scale * sum_p w(p, q) in[p].
struct AffineInv {
  float m[12];  // rows of [A^-1 | -A^-1 t]
  float h[3];   // half extents of the candidate box per source axis
};

__global__ void affine_push_gather_kernel(const float *__restrict__ in, float *__restrict__ out,
                                          Dim3i s, Dim3i o, CoordAffine coord, AffineInv inv,
                                          int extrapolate, float scale) {
  const int qz = blockIdx.x * blockDim.x + threadIdx.x;
  const int qy = blockIdx.y * blockDim.y + threadIdx.y;
  const int qx = blockIdx.z;
  if (qz >= s.z || qy >= s.y) return;
  const float fqx = (float)qx, fqy = (float)qy, fqz = (float)qz;
  int lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pc = inv.m[4 * a + 0] * fqx + inv.m[4 * a + 1] * fqy + inv.m[4 * a + 2] * fqz +
                     inv.m[4 * a + 3];
    const int n = a == 0 ? o.x : (a == 1 ? o.y : o.z);
    lo[a] = max(0, (int)ceilf(pc - inv.h[a]));
    hi[a] = min(n - 1, (int)floorf(pc + inv.h[a]));
  }
  float acc = 0.f;
  for (int i = lo[0]; i <= hi[0]; ++i)
    for (int j = lo[1]; j <= hi[1]; ++j)
      for (int k = lo[2]; k <= hi[2]; ++k) {
        const size_t lin = ((size_t)i * o.y + j) * o.z + k;
        float cx, cy, cz;
        coord.get(i, j, k, lin, cx, cy, cz);
        if (!(extrapolate || in_fov(cx, cy, cz, s))) continue;
        const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
        const int bx = qx - (int)fx, by = qy - (int)fy, bz = qz - (int)fz;
        if ((unsigned)bx > 1u || (unsigned)by > 1u || (unsigned)bz > 1u) continue;
        const float wx1 = cx - fx, wy1 = cy - fy, wz1 = cz - fz;
        const float wgt = ((bx ? wx1 : 1.f - wx1) * (by ? wy1 : 1.f - wy1)) * (bz ? wz1 : 1.f - wz1);
        acc += (scale * __ldg(in + lin)) * wgt;
      }
  out[((size_t)qx * s.y + qy) * s.z + qz] += acc;
}

// Lattice-aligned operators (identity rotation, integer translation): pull is a shifted crop,
// push a shifted zero-pad embed -- one corner of weight exactly 1, a one-to-one mapping, so
// the push needs no atomics.  o = grid of the coordinates, s = the volume they point into.
template <bool PUSH>
__global__ void lattice_shift_kernel(const float *__restrict__ in, float *__restrict__ out,
                                     Dim3i s, Dim3i o, int tx, int ty, int tz, float scale) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= o.z || j >= o.y) return;
  const size_t lin = ((size_t)i * o.y + j) * o.z + k;
  const int px = i + tx, py = j + ty, pz = k + tz;
  const bool inside = px >= 0 && px < s.x && py >= 0 && py < s.y && pz >= 0 && pz < s.z;
  const size_t q = ((size_t)px * s.y + py) * s.z + pz;
  if (PUSH) {
    if (inside) out[q] += scale * in[lin];
  } else {
    out[lin] = inside ? __ldg(in + q) : 0.f;
  }
}

// Spatial gradient of the trilinearly interpolated volume w.r.t. the sampling coordinates
// (nitorch grid_grad, unires/_update.py:505): along axis a the corner weights (1 - t, t) become
// (-1, +1).  out is (ox, oy, oz, 3); zero bound, FOV tolerance as the pull.
__global__ void affine_grad_kernel(const float *__restrict__ in, float *__restrict__ out, Dim3i s,
                                   Dim3i o, CoordAffine coord, int extrapolate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= o.z || j >= o.y) return;
  const size_t lin = ((size_t)i * o.y + j) * o.z + k;
  float cx, cy, cz;
  coord.get(i, j, k, lin, cx, cy, cz);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (extrapolate || in_fov(cx, cy, cz, s)) {
    const size_t sy = s.z, sx = (size_t)s.y * s.z;
    const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const float wx1 = cx - fx, wy1 = cy - fy, wz1 = cz - fz;
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int bx = (c >> 2) & 1, by = (c >> 1) & 1, bz = c & 1;
      const int px = ix + bx, py = iy + by, pz = iz + bz;
      if (px < 0 || px >= s.x || py < 0 || py >= s.y || pz < 0 || pz >= s.z) continue;
      const float v = __ldg(in + px * sx + py * sy + pz);
      const float wx = bx ? wx1 : wx0, wy = by ? wy1 : wy0, wz = bz ? wz1 : wz0;
      const float sgx = bx ? 1.f : -1.f, sgy = by ? 1.f : -1.f, sgz = bz ? 1.f : -1.f;
      // same association as the oracle: (w_x * w_y) * w_z with the differentiated weight +-1
      g0 += v * ((sgx * wy) * wz);
      g1 += v * ((wx * sgy) * wz);
      g2 += v * ((wx * wy) * sgz);
    }
  }
  out[3 * lin + 0] = g0;
  out[3 * lin + 1] = g1;
  out[3 * lin + 2] = g2;
}

static inline void shape_for(const Dim3i &o, dim3 &grid, dim3 &block) {
  block = dim3(64, 4, 1);
  grid = dim3(div_up(o.z, 64), div_up(o.y, 4), o.x);
}

static inline CoordAffine make_affine(const float mat[12]) {
  CoordAffine c;
  for (int i = 0; i < 12; ++i) c.a.m[i] = mat[i];
  return c;
}

int affine_pull(const float *src, Dim3i s, const float mat[12], float *out, Dim3i o, int order,
                int extrapolate, cudaStream_t st) {
  dim3 grid, block;
  shape_for(o, grid, block);
  resample_kernel<CoordAffine, false>
      <<<grid, block, 0, st>>>(src, out, s, o, make_affine(mat), order, extrapolate, 1.f);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int affine_push(const float *in, Dim3i o, const float mat[12], float *out, Dim3i s, int order,
                int extrapolate, float scale, cudaStream_t st) {
  dim3 grid, block;
  shape_for(o, grid, block);
  resample_kernel<CoordAffine, true>
      <<<grid, block, 0, st>>>(in, out, s, o, make_affine(mat), order, extrapolate, scale);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

// Deterministic push for an affine grid; returns UR_ERR_UNSUPPORTED when the linear part is
// (numerically) singular or the candidate box would be unreasonably large.
int affine_push_gather(const float *in, Dim3i o, const float mat[12], float *out, Dim3i s,
                       int extrapolate, float scale, cudaStream_t st) {
  const double a[3][3] = {{mat[0], mat[1], mat[2]}, {mat[4], mat[5], mat[6]}, {mat[8], mat[9], mat[10]}};
  const double t[3] = {mat[3], mat[7], mat[11]};
  const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
                     a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                     a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
  if (!(fabs(det) > 1e-9)) return UR_ERR_UNSUPPORTED;
  double iv[3][3];
  iv[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det;
  iv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
  iv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det;
  iv[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det;
  iv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det;
  iv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
  iv[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det;
  iv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
  iv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
  AffineInv inv;
  double cand = 1.0;
  for (int r = 0; r < 3; ++r) {
    double h = 0.0, off = 0.0;
    for (int c = 0; c < 3; ++c) {
      inv.m[4 * r + c] = (float)iv[r][c];
      h += fabs(iv[r][c]);
      off -= iv[r][c] * t[c];
    }
    inv.m[4 * r + 3] = (float)off;
    inv.h[r] = (float)(h * 1.0005 + 1e-2);  // float32 slack on coordinates up to ~1e3
    cand *= 2.0 * inv.h[r] + 1.0;
  }
  if (cand > 4096.0) return UR_ERR_UNSUPPORTED;
  dim3 grid, block;
  shape_for(s, grid, block);
  affine_push_gather_kernel<<<grid, block, 0, st>>>(in, out, s, o, make_affine(mat), inv,
                                                    extrapolate, scale);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int lattice_pull(const float *src, Dim3i s, const float mat[12], float *out, Dim3i o,
                 cudaStream_t st) {
  dim3 grid, block;
  shape_for(o, grid, block);
  lattice_shift_kernel<false><<<grid, block, 0, st>>>(src, out, s, o, (int)lrintf(mat[3]),
                                                      (int)lrintf(mat[7]), (int)lrintf(mat[11]),
                                                      1.f);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int lattice_push(const float *in, Dim3i o, const float mat[12], float *out, Dim3i s, float scale,
                 cudaStream_t st) {
  dim3 grid, block;
  shape_for(o, grid, block);
  lattice_shift_kernel<true><<<grid, block, 0, st>>>(in, out, s, o, (int)lrintf(mat[3]),
                                                     (int)lrintf(mat[7]), (int)lrintf(mat[11]),
                                                     scale);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

using namespace ur;

static bool dims_ok(const int32_t d[3]) { return d && d[0] > 0 && d[1] > 0 && d[2] > 0; }

extern "C" int ur_grid_pull(const float *d_src, const int32_t sdim[3], const float *d_grid,
                            float *d_out, const int32_t odim[3], int order, int extrapolate,
                            ur_stream stream) {
  UR_REQUIRE(d_src && d_grid && d_out && dims_ok(sdim) && dims_ok(odim), "ur_grid_pull: bad args");
  UR_REQUIRE(order == 0 || order == 1, "ur_grid_pull: interpolation order must be 0 or 1");
  Dim3i s = make_dim(sdim), o = make_dim(odim);
  dim3 grid, block;
  shape_for(o, grid, block);
  resample_kernel<CoordDense, false><<<grid, block, 0, (cudaStream_t)stream>>>(
      d_src, d_out, s, o, CoordDense{d_grid}, order, extrapolate, 1.f);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_grid_push(const float *d_in, const int32_t idim[3], const float *d_grid,
                            float *d_out, const int32_t sdim[3], int order, int extrapolate,
                            float scale, ur_stream stream) {
  UR_REQUIRE(d_in && d_grid && d_out && dims_ok(sdim) && dims_ok(idim), "ur_grid_push: bad args");
  UR_REQUIRE(order == 0 || order == 1, "ur_grid_push: interpolation order must be 0 or 1");
  Dim3i s = make_dim(sdim), o = make_dim(idim);
  dim3 grid, block;
  shape_for(o, grid, block);
  resample_kernel<CoordDense, true><<<grid, block, 0, (cudaStream_t)stream>>>(
      d_in, d_out, s, o, CoordDense{d_grid}, order, extrapolate, scale);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_affine_pull(const float *d_src, const int32_t sdim[3], const float mat[12],
                              float *d_out, const int32_t odim[3], int order, int extrapolate,
                              ur_stream stream) {
  UR_REQUIRE(d_src && mat && d_out && dims_ok(sdim) && dims_ok(odim), "ur_affine_pull: bad args");
  UR_REQUIRE(order == 0 || order == 1, "ur_affine_pull: interpolation order must be 0 or 1");
  return affine_pull(d_src, make_dim(sdim), mat, d_out, make_dim(odim), order, extrapolate,
                     (cudaStream_t)stream);
}

extern "C" int ur_affine_push(const float *d_in, const int32_t idim[3], const float mat[12],
                              float *d_out, const int32_t sdim[3], int order, int extrapolate,
                              float scale, ur_stream stream) {
  UR_REQUIRE(d_in && mat && d_out && dims_ok(sdim) && dims_ok(idim), "ur_affine_push: bad args");
  UR_REQUIRE(order == 0 || order == 1, "ur_affine_push: interpolation order must be 0 or 1");
  return affine_push(d_in, make_dim(idim), mat, d_out, make_dim(sdim), order, extrapolate, scale,
                     (cudaStream_t)stream);
}

extern "C" int ur_affine_grad(const float *d_src, const int32_t sdim[3], const float mat[12],
                              float *d_out, const int32_t odim[3], int extrapolate,
                              ur_stream stream) {
  UR_REQUIRE(d_src && mat && d_out && dims_ok(sdim) && dims_ok(odim), "ur_affine_grad: bad args");
  Dim3i s = make_dim(sdim), o = make_dim(odim);
  dim3 grid, block;
  shape_for(o, grid, block);
  affine_grad_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_src, d_out, s, o,
                                                               make_affine(mat), extrapolate);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

extern "C" int ur_affine_grid(const float mat[12], float *d_grid, const int32_t odim[3],
                              ur_stream stream) {
  UR_REQUIRE(mat && d_grid && dims_ok(odim), "ur_affine_grid: bad args");
  Dim3i o = make_dim(odim);
  dim3 grid, block;
  shape_for(o, grid, block);
  affine_grid_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_grid, o, make_affine(mat));
  UR_LAUNCH_CHECK();
  return UR_OK;
}

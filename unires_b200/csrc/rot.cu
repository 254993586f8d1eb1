// Rotated observations (rot.cuh): the in-tile forward kernel  v -> C P v -> (S / C'S^2) and the
// host side that maps a ur_proj onto it.  Reference: unires/_project.py:147-179.
#include <math.h>
#include <string.h>

#include "rot.cuh"

namespace ur {

int g_rot_fused = 1;  // ur_tune("rot_fused")
int g_rot_cell = 1;   // ur_tune("rot_cell"): adjoint through per-cell corner coefficients
                      // (0 = per-voxel gather, 8 = always eight colour passes: test hook)

constexpr int kRotThreads = 256;

struct RotTile {
  int e[3];    // owned block of the OUTPUT index space (A: low-res rows; AtA: intermediate)
  int pe[3];   // pulled tile extents
  int nj_max;  // low-res rows per tile along the profile axis (AtA)
};

// trilinear pull of v at the intermediate voxel (i, j, k): same float32 expressions, FOV
// tolerance and corner order as resample_kernel (resample.cu), i.e. as the oracle
__device__ __forceinline__ float rot_pull(const float *__restrict__ v, const RotFwd &F, int i,
                                          int j, int k) {
  const float fi = (float)i, fj = (float)j, fk = (float)k;
  const float cx = fmaf(F.m[2], fk, fmaf(F.m[1], fj, F.m[0] * fi)) + F.m[3];
  const float cy = fmaf(F.m[6], fk, fmaf(F.m[5], fj, F.m[4] * fi)) + F.m[7];
  const float cz = fmaf(F.m[10], fk, fmaf(F.m[9], fj, F.m[8] * fi)) + F.m[11];
  if (!(cx > -kRotFovTol && cx < (float)(F.s[0] - 1) + kRotFovTol && cy > -kRotFovTol &&
        cy < (float)(F.s[1] - 1) + kRotFovTol && cz > -kRotFovTol &&
        cz < (float)(F.s[2] - 1) + kRotFovTol))
    return 0.f;
  const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const float wx1 = cx - fx, wy1 = cy - fy, wz1 = cz - fz;
  const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
  const int sy = F.s[2], sx = F.s[1] * F.s[2];
  const bool vx0 = ix >= 0, vx1 = ix + 1 < F.s[0];  // ix <= s-1 and ix+1 >= 0 by the FOV test
  const bool vy0 = iy >= 0, vy1 = iy + 1 < F.s[1];
  const bool vz0 = iz >= 0, vz1 = iz + 1 < F.s[2];
  const float *b = v + ((long long)ix * sx + (long long)iy * sy + iz);
  const float w00 = wx0 * wy0, w01 = wx0 * wy1, w10 = wx1 * wy0, w11 = wx1 * wy1;
  float acc = 0.f;
  if (vx0 && vx1 && vy0 && vy1 && vz0 && vz1) {
    // interior sample (almost all of them): eight unconditional loads in flight at once, summed
    // in the same order as below
    const float c000 = __ldg(b), c001 = __ldg(b + 1), c010 = __ldg(b + sy),
                c011 = __ldg(b + sy + 1), c100 = __ldg(b + sx), c101 = __ldg(b + sx + 1),
                c110 = __ldg(b + sx + sy), c111 = __ldg(b + sx + sy + 1);
    acc += c000 * (w00 * wz0);
    acc += c001 * (w00 * wz1);
    acc += c010 * (w01 * wz0);
    acc += c011 * (w01 * wz1);
    acc += c100 * (w10 * wz0);
    acc += c101 * (w10 * wz1);
    acc += c110 * (w11 * wz0);
    acc += c111 * (w11 * wz1);
    return acc;
  }
  if (vx0 && vy0 && vz0) acc += __ldg(b) * (w00 * wz0);
  if (vx0 && vy0 && vz1) acc += __ldg(b + 1) * (w00 * wz1);
  if (vx0 && vy1 && vz0) acc += __ldg(b + sy) * (w01 * wz0);
  if (vx0 && vy1 && vz1) acc += __ldg(b + sy + 1) * (w01 * wz1);
  if (vx1 && vy0 && vz0) acc += __ldg(b + sx) * (w10 * wz0);
  if (vx1 && vy0 && vz1) acc += __ldg(b + sx + 1) * (w10 * wz1);
  if (vx1 && vy1 && vz0) acc += __ldg(b + sx + sy) * (w11 * wz0);
  if (vx1 && vy1 && vz1) acc += __ldg(b + sx + sy + 1) * (w11 * wz1);
  return acc;
}

__device__ __forceinline__ int floordiv_i(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

// ATA = false: out (dim_x)  = weight * S C P v          block owns e[] low-res voxels
// ATA = true : out (dim_yx) = weight * C' S C P v        block owns e[] intermediate voxels
//              (S already holds the squared factors)
// The tile loops below run warps over the (c0, c1) rows and lanes along c2 -- no per-element
// integer division (the first version spent more instructions on idx / n, idx % n than on the
// trilinear pull: 258 warp instructions per intermediate voxel, ncu).
template <bool ATA>
__global__ void __launch_bounds__(kRotThreads)
    rot_forward_kernel(const float *__restrict__ v, float *__restrict__ out, const RotFwd F,
                       const RotTile T, const int *done) {
  extern __shared__ float sm[];
  __shared__ float s_ker[UR_MAX_TAPS];  // taps by LDS: the expansion indexes them per lane
  if (done && *done) return;
  const int ax = F.axis;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < UR_MAX_TAPS) s_ker[tid] = tid < F.K ? F.ker[tid] : 0.f;
  constexpr int NW = kRotThreads / 32;
  const int b3[3] = {(int)blockIdx.z, (int)blockIdx.y, (int)blockIdx.x};
  int o0[3];  // first owned output index per axis
#pragma unroll
  for (int a = 0; a < 3; ++a) o0[a] = b3[a] * T.e[a];
  const int o0a = ax == 0 ? o0[0] : (ax == 1 ? o0[1] : o0[2]);
  const int nyxa = ax == 0 ? F.nyx[0] : (ax == 1 ? F.nyx[1] : F.nyx[2]);
  const int ea = ax == 0 ? T.e[0] : (ax == 1 ? T.e[1] : T.e[2]);
  // low-res rows of this tile along the profile axis
  int j_min, nj_t;
  if (ATA) {
    const int u0 = o0a, u1 = min(u0 + ea, nyxa) - 1;
    j_min = -floordiv_i(-(u0 - F.k0 - F.K + 1), F.r);  // ceil
    const int j_max = floordiv_i(u1 - F.k0, F.r);
    nj_t = j_max - j_min + 1;
  } else {
    j_min = o0a;
    nj_t = min(ea, F.nj - o0a);
  }
  if (nj_t < 1) nj_t = 0;
  // pulled tile: intermediate indices [ps, ps + pn) per axis
  int ps0 = o0[0], ps1 = o0[1], ps2 = o0[2];
  int pn0 = min(T.e[0], (ATA ? F.nyx[0] : F.nlr[0]) - o0[0]);
  int pn1 = min(T.e[1], (ATA ? F.nyx[1] : F.nlr[1]) - o0[1]);
  int pn2 = min(T.e[2], (ATA ? F.nyx[2] : F.nlr[2]) - o0[2]);
  const int en0 = pn0, en1 = pn1, en2 = pn2;  // owned extents (ATA: intermediate voxels)
  {
    const int s_ = j_min * F.r + F.k0, n_ = nj_t > 0 ? (nj_t - 1) * F.r + F.K : 0;
    if (ax == 0) ps0 = s_, pn0 = n_;
    else if (ax == 1) ps1 = s_, pn1 = n_;
    else ps2 = s_, pn2 = n_;
  }
  float *pulled = sm;                                  // pitches from T.pe
  float *lres = sm + T.pe[0] * T.pe[1] * T.pe[2];      // low-res rows, same pitches
  const int pp1 = T.pe[2], pp0 = T.pe[1] * T.pe[2];
  const int astep = ax == 0 ? pp0 : (ax == 1 ? pp1 : 1);
  // ---- phase 1: pull ----
  if (pn1 > 0) {
    int c0 = warp / pn1, c1 = warp - c0 * pn1;
    while (c0 < pn0) {
      const int i = ps0 + c0, j = ps1 + c1;
      for (int c2 = lane; c2 < pn2; c2 += 32) {
        const int k = ps2 + c2;
        const int pa = ax == 0 ? i : (ax == 1 ? j : k);
        float val = 0.f;
        if (pa >= 0 && pa < nyxa) val = rot_pull(v, F, i, j, k);
        pulled[c0 * pp0 + c1 * pp1 + c2] = val;
      }
      c1 += NW;
      while (c1 >= pn1) {
        c1 -= pn1;
        ++c0;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: low-res rows  x[j] = s_j sum_t ker[t] yx[j r + k0 + t] ----
  {
    const int ln0 = ax == 0 ? nj_t : pn0, ln1 = ax == 1 ? nj_t : pn1, ln2 = ax == 2 ? nj_t : pn2;
    if (ln1 > 0) {
      int c0 = warp / ln1, c1 = warp - c0 * ln1;
      while (c0 < ln0) {
        for (int c2 = lane; c2 < ln2; c2 += 32) {
          const int jl = ax == 0 ? c0 : (ax == 1 ? c1 : c2);  // tile-local row
          const int j = j_min + jl;                            // global low-res row
          // pulled index: the axis coordinate jl becomes jl * r
          const int base = c0 * pp0 + c1 * pp1 + c2 + (jl * F.r - jl) * astep;
          float acc = 0.f;
          const bool valid = j >= 0 && j < F.nj;
          if (valid) {
            for (int tt = 0; tt < F.K; ++tt) acc = fmaf(s_ker[tt], pulled[base + tt * astep], acc);
            if (F.scl_axis >= 0) {
              const int g0 = ps0 + c0, g1 = ps1 + c1, g2 = ps2 + c2;
              const int js = F.scl_axis == ax ? j : (F.scl_axis == 0 ? g0 : (F.scl_axis == 1 ? g1 : g2));
              acc *= (js & 1) ? F.s_odd : F.s_even;
            }
          }
          if (ATA) {
            lres[c0 * pp0 + c1 * pp1 + c2] = acc;
          } else if (valid) {
            const int g0 = ax == 0 ? j : ps0 + c0, g1 = ax == 1 ? j : ps1 + c1,
                      g2 = ax == 2 ? j : ps2 + c2;
            out[((size_t)g0 * F.nlr[1] + g1) * F.nlr[2] + g2] = F.weight * acc;
          }
        }
        c1 += NW;
        while (c1 >= ln1) {
          c1 -= ln1;
          ++c0;
        }
      }
    }
  }
  if (!ATA) return;
  __syncthreads();
  // ---- phase 3: expand  u[p] = weight sum_j ker[p - k0 - j r] x[j] ----
  const float inv_r = 1.f / (float)F.r;
  if (en1 > 0) {
    int c0 = warp / en1, c1 = warp - c0 * en1;
    while (c0 < en0) {
      for (int c2 = lane; c2 < en2; c2 += 32) {
        const int ca = ax == 0 ? c0 : (ax == 1 ? c1 : c2);
        const int u = o0a + ca - F.k0;  // position relative to the first tap of row 0
        // floor(u / r) and ceil((u - K + 1) / r) through float (exact for |u| < 2^22; an integer
        // division per voxel costs more than the taps)
        const int fl_u = (int)floorf(((float)u + 0.5f) * inv_r);
        int jl_hi = fl_u - j_min;
        int jl_lo = (int)floorf(((float)(u - F.K + F.r) + 0.5f) * inv_r) - j_min;
        if (jl_lo < 0) jl_lo = 0;
        if (jl_hi > nj_t - 1) jl_hi = nj_t - 1;
        const int base = c0 * pp0 + c1 * pp1 + c2 - ca * astep;
        float acc = 0.f;
        for (int jl = jl_lo; jl <= jl_hi; ++jl)
          acc = fmaf(s_ker[u - (j_min + jl) * F.r], lres[base + jl * astep], acc);
        out[((size_t)(o0[0] + c0) * F.nyx[1] + (o0[1] + c1)) * F.nyx[2] + (o0[2] + c2)] =
            F.weight * acc;
      }
      c1 += NW;
      while (c1 >= en1) {
        c1 -= en1;
        ++c0;
      }
    }
  }
}

// u = weight * C' S x  (At): thread per intermediate voxel, x read from global (low-res)
__global__ void __launch_bounds__(256)
    rot_expand_kernel(const float *__restrict__ x, float *__restrict__ u, const RotFwd F) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.z;
  if (k >= F.nyx[2] || j >= F.nyx[1]) return;
  int g[3] = {i, j, k};
  const int ax = F.axis;
  const int uu = g[ax] - F.k0;
  int j_lo = -floordiv_i(-(uu - F.K + 1), F.r), j_hi = floordiv_i(uu, F.r);
  if (j_lo < 0) j_lo = 0;
  if (j_hi > F.nj - 1) j_hi = F.nj - 1;
  const size_t st = ax == 0 ? (size_t)F.nlr[1] * F.nlr[2] : (ax == 1 ? (size_t)F.nlr[2] : 1);
  int gl[3] = {i, j, k};
  gl[ax] = 0;
  const float *base = x + ((size_t)gl[0] * F.nlr[1] + gl[1]) * F.nlr[2] + gl[2];
  float acc = 0.f;
  for (int jj = j_lo; jj <= j_hi; ++jj) {
    float val = __ldg(base + (size_t)jj * st);
    if (F.scl_axis >= 0) {
      const int js = F.scl_axis == ax ? jj : g[F.scl_axis];
      val *= (js & 1) ? F.s_odd : F.s_even;
    }
    acc = fmaf(F.ker[uu - jj * F.r], val, acc);
  }
  u[((size_t)i * F.nyx[1] + j) * F.nyx[2] + k] = F.weight * acc;
}

__global__ void __launch_bounds__(256)
    rot_adjoint_kernel(const RotTerm T, float *__restrict__ out, int nx, int ny, int nz,
                       int accumulate) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (z >= nz || y >= ny) return;
  const float val = rot_gather(T, x, y, z, nx, ny, nz);
  const size_t i = ((size_t)x * ny + y) * nz + z;
  out[i] = accumulate ? out[i] + val : val;
}

// one thread per quad of z-consecutive voxels (nz % 4 == 0, 16-byte aligned `out`)
__global__ void __launch_bounds__(256)
    rot_adjoint4_kernel(const RotTerm T, float *__restrict__ out, int nx, int ny, int nz,
                        int accumulate) {
  const int z = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (z >= nz || y >= ny) return;
  float g[4];
  rot_gather4(T, x, y, z, nx, ny, nz, g);
  float4 *o = reinterpret_cast<float4 *>(out + ((size_t)x * ny + y) * nz + z);
  float4 q = make_float4(g[0], g[1], g[2], g[3]);
  if (accumulate) {
    const float4 p = *o;
    q.x += p.x, q.y += p.y, q.z += p.z, q.w += p.w;
  }
  *o = q;
}


// ---------------------------------------------------------------------------------------------
// Adjoint pull P'u through per-CELL corner coefficients (deterministic, no atomics).
//
// The pull samples v at c(p) = M p + t for every intermediate voxel p; its transpose hands
// u[p] w_abc(p) to the eight corners (a, b, c) of the recon cell floor(c(p)).  A CTA owns a
// TX x TY x TZ tile of recon voxels, i.e. the (TX+1)(TY+1)(TZ+1) cells that have a corner in
// it, and keeps eight coefficients per cell in shared memory:
//   phase 1 (SCATTER over the intermediate voxels of the tile's pre-image): p is visited once,
//            c(p), the FOV test and the corner weights with the pull's own float32
//            expressions; coef[abc][cell] += u[p] w_abc.  Two intermediate voxels can share a
//            cell only if they differ by a lattice vector whose image fits a unit cell; the
//            host picks a colouring of the lattice (2 colours: parity of i+j+k; 8: parities of
//            i, j, k) under which same-coloured voxels never do, and the colours run as
//            barrier-separated passes -- plain read-modify-write, fixed order, bit-reproducible;
//   phase 2 (GATHER): a recon voxel sums coefficient (a,b,c) of the cell at voxel - (a,b,c).
// ~7 warp instructions per recon voxel instead of ~18 for the per-voxel candidate search of
// rot_gather4 (which visits 18-27 candidates for 8 contributors).
// Reference: nitorch grid_push as called at unires/_project.py:172,179 (= transpose of :164).
constexpr int kCellTX = 8, kCellTY = 8, kCellTZ = 28;
constexpr int kCellMaxRows = 1000;
constexpr int kCellThreads = 512;

// k range (clipped into [k0, k1]) where lo <= m k + s < hi can hold; widened against rounding
__device__ __forceinline__ void cell_krange(float m, float s, float lo, float hi, int &k0, int &k1) {
  if (fabsf(m) < 1e-6f) return;  // no usable bound along k: the candidate test decides
  const float r = 1.f / m;
  const float e0 = (lo - s) * r, e1 = (hi - s) * r;
  const float slack = 2e-2f + 4e-4f * fabsf(r);
  const float a = fminf(e0, e1) - slack, b = fmaxf(e0, e1) + slack;
  // compare in float first: the bounds may exceed the int range
  if (a > (float)k0) k0 = a > (float)k1 ? k1 + 1 : (int)ceilf(a);
  if (b < (float)k1) k1 = b < (float)k0 ? k0 - 1 : (int)floorf(b);
}

template <int TX, int TY, int TZ>
__global__ void __launch_bounds__(kCellThreads, 2)
    rot_adjoint_cell_kernel(const RotTerm T, float *__restrict__ out, int nx, int ny, int nz,
                            int accumulate, const int *done) {
  constexpr int NW = kCellThreads / 32;
  static_assert(TX == 8 && NW % 8 == 0 && TY % (NW / 8) == 0, "a warp per x row and y slab");
  static_assert(TZ <= 32, "one lane per z voxel of the tile");
  constexpr int CX = TX + 1, CY = TY + 1, CZ = TZ + 1;
  constexpr int NCELL = CX * CY * CZ;
  constexpr int PL = (NCELL + 3) / 4 * 4;  // floats per coefficient plane
  extern __shared__ __align__(16) float cell_sm[];
  float *coef = cell_sm;  // [8][PL]
  int4 *rows = reinterpret_cast<int4 *>(cell_sm + 8 * PL);  // (i, j, k0, k1) of the live rows
  __shared__ int s_nrows, s_box[6];
  if (done && *done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.z * TX, y0 = blockIdx.y * TY, z0 = blockIdx.x * TZ;
  // sample positions of the owned cells: [x0 - 1, x0 + TX) x [y0 - 1, y0 + TY) x [z0 - 1, z0 + TZ)
  const float bl[3] = {(float)(x0 - 1), (float)(y0 - 1), (float)(z0 - 1)};
  const float bh[3] = {(float)(x0 + TX), (float)(y0 + TY), (float)(z0 + TZ)};
  if (tid < 3) {  // bounding box of the pre-image in intermediate indices, one axis per thread
    const int a = tid;
    const float base = T.inv[4 * a + 0] * bl[0] + T.inv[4 * a + 1] * bl[1] +
                       T.inv[4 * a + 2] * bl[2] + T.inv[4 * a + 3];
    const float dx = T.inv[4 * a + 0] * (float)(TX + 1), dy = T.inv[4 * a + 1] * (float)(TY + 1),
                dz = T.inv[4 * a + 2] * (float)(TZ + 1);
    const float lo = base + fminf(dx, 0.f) + fminf(dy, 0.f) + fminf(dz, 0.f) - 0.06f;
    const float hi = base + fmaxf(dx, 0.f) + fmaxf(dy, 0.f) + fmaxf(dz, 0.f) + 0.06f;
    s_box[a] = lo > 0.f ? (lo > 1e9f ? 1 << 30 : (int)ceilf(lo)) : 0;
    s_box[3 + a] = hi < (float)(T.n[a] - 1) ? (hi < -1e9f ? -2 : (int)floorf(hi)) : T.n[a] - 1;
    if (a == 0) s_nrows = 0;
  }
  __syncthreads();
  const int ilo[3] = {s_box[0], s_box[1], s_box[2]}, ihi[3] = {s_box[3], s_box[4], s_box[5]};
  const int ni = ihi[0] - ilo[0] + 1, nj = ihi[1] - ilo[1] + 1;
  const bool any = ni > 0 && nj > 0 && ihi[2] >= ilo[2];
  const int nbox = any ? ni * nj : 0;
  if (nbox > kCellMaxRows) asm volatile("trap;");  // the host checks the operator: never taken
  constexpr int YS = TY / (NW / 8);  // y rows per warp in the output phase
  const int xr = warp & 7, ys = (warp >> 3) * YS;
  const int x = x0 + xr, z = z0 + lane;
  if (nbox == 0) {  // nothing maps into this tile (outside the observation's field of view)
    if (!accumulate && x < nx && lane < TZ && z < nz) {
      float *o = out + ((size_t)x * ny + (y0 + ys)) * nz + z;
      for (int yy = 0; yy < YS && y0 + ys + yy < ny; ++yy, o += nz) *o = 0.f;
    }
    return;
  }

  // ---- phase 0: zero the coefficients; list of the rows (i, j) that can hit the tile with
  // their k range (list order is irrelevant: rows of one pass never share a cell) ----
  for (int q = tid; q < 8 * PL / 4; q += kCellThreads)
    reinterpret_cast<float4 *>(coef)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = tid; r < nbox; r += kCellThreads) {
    const int ii = r / nj, i = ilo[0] + ii, j = ilo[1] + (r - ii * nj);
    const float fi = (float)i, fj = (float)j;
    const float bx = fmaf(T.m[1], fj, T.m[0] * fi) + T.m[3];
    const float by = fmaf(T.m[5], fj, T.m[4] * fi) + T.m[7];
    const float bz = fmaf(T.m[9], fj, T.m[8] * fi) + T.m[11];
    int k0 = ilo[2], k1 = ihi[2];
    cell_krange(T.m[2], bx, bl[0], bh[0], k0, k1);
    cell_krange(T.m[6], by, bl[1], bh[1], k0, k1);
    cell_krange(T.m[10], bz, bl[2], bh[2], k0, k1);
    if (k0 <= k1) rows[atomicAdd(&s_nrows, 1)] = make_int4(i, j, k0, k1);
  }
  __syncthreads();
  const int nrows = s_nrows;

  // ---- phase 1: colour passes over the intermediate voxels ----
  const float fmax_x = (float)(nx - 1) + kRotFovTol, fmax_y = (float)(ny - 1) + kRotFovTol,
              fmax_z = (float)(nz - 1) + kRotFovTol;
  const int sub = lane >> 4, l16 = lane & 15;  // a half warp per candidate row
  const int ncol = T.cell_ncol;
  const int n12 = T.n[1] * T.n[2];
  // the pull's FOV test can only fail in cells next to a face of the recon grid
  const bool face = x0 == 0 || y0 == 0 || z0 == 0 || x0 + TX >= nx || y0 + TY >= ny ||
                    z0 + TZ >= nz;
  for (int c = 0; c < ncol && nrows > 0; ++c) {
    for (int rp = 2 * warp + sub; rp < nrows; rp += 2 * NW) {
      const int4 e = rows[rp];
      const int i = e.x, j = e.y;
      int kpar = (c + i + j) & 1;
      if (ncol == 8) {  // rows with the pass's parities of i and j only
        if ((((i & 1) << 1) | (j & 1)) != (c >> 1)) continue;
        kpar = c & 1;
      }
      const float fi = (float)i, fj = (float)j;
      const float bx = fmaf(T.m[1], fj, T.m[0] * fi);
      const float by = fmaf(T.m[5], fj, T.m[4] * fi);
      const float bz = fmaf(T.m[9], fj, T.m[8] * fi);
      const float *urow = T.u + (i * n12 + j * T.n[2]);
      for (int k = e.z + ((e.z ^ kpar) & 1) + 2 * l16; k <= e.w; k += 32) {
        const float uv = __ldg(urow + k);  // requested before the coordinates: latency overlaps
        const float fk = (float)k;
        const float cx = fmaf(T.m[2], fk, bx) + T.m[3];
        const float cy = fmaf(T.m[6], fk, by) + T.m[7];
        const float cz = fmaf(T.m[10], fk, bz) + T.m[11];
        bool ok = true;
        if (face)
          ok = cx > -kRotFovTol && cx < fmax_x && cy > -kRotFovTol && cy < fmax_y &&
               cz > -kRotFovTol && cz < fmax_z;
        const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
        const int lx = (int)fx - (x0 - 1), ly = (int)fy - (y0 - 1), lz = (int)fz - (z0 - 1);
        if (ok && (unsigned)lx < (unsigned)CX && (unsigned)ly < (unsigned)CY &&
            (unsigned)lz < (unsigned)CZ) {
          const float wx1 = cx - fx, wy1 = cy - fy, wz1 = cz - fz;
          const float wx0 = 1.f - wx1, wy0 = 1.f - wy1, wz0 = 1.f - wz1;
          const float w00 = wx0 * wy0, w01 = wx0 * wy1, w10 = wx1 * wy0, w11 = wx1 * wy1;
          float *cp = coef + (lx * CY + ly) * CZ + lz;
          cp[0 * PL] += uv * (w00 * wz0);
          cp[1 * PL] += uv * (w00 * wz1);
          cp[2 * PL] += uv * (w01 * wz0);
          cp[3 * PL] += uv * (w01 * wz1);
          cp[4 * PL] += uv * (w10 * wz0);
          cp[5 * PL] += uv * (w10 * wz1);
          cp[6 * PL] += uv * (w11 * wz0);
          cp[7 * PL] += uv * (w11 * wz1);
        }
      }
    }
    __syncthreads();
  }

  // ---- phase 2: voxel (x, y, z) <- coefficient (a, b, c) of the cell at (x - a, y - b, z - c) ----
  if (x < nx && lane < TZ && z < nz) {
    // the voxel's own cell (its corner 0,0,0) at y = y0 + ys
    const float *cc = coef + ((xr + 1) * CY + ys + 1) * CZ + lane + 1;
    float *o = out + ((size_t)x * ny + (y0 + ys)) * nz + z;
#pragma unroll 2
    for (int yy = 0; yy < YS; ++yy, cc += CZ, o += nz) {
      if (y0 + ys + yy >= ny) break;
      float acc = cc[0 * PL];
      acc += cc[1 * PL - 1];
      acc += cc[2 * PL - CZ];
      acc += cc[3 * PL - CZ - 1];
      acc += cc[4 * PL - CY * CZ];
      acc += cc[5 * PL - CY * CZ - 1];
      acc += cc[6 * PL - CY * CZ - CZ];
      acc += cc[7 * PL - CY * CZ - CZ - 1];
      *o = accumulate ? *o + acc : acc;
    }
  }
}

// Colour passes under which two intermediate voxels of one colour never fall into the same
// recon cell: 2 (parity of i + j + k) when no even-sum lattice vector maps into a unit cell, 8
// (parities of i, j, k) when no all-even vector does, 0 = neither (strongly anisotropic map).
static int rot_cell_colours(const float *mat) {
  bool ok2 = true, ok8 = true;
  for (int a = -4; a <= 4; ++a)
    for (int b = -4; b <= 4; ++b)
      for (int c = -4; c <= 4; ++c) {
        if (!a && !b && !c) continue;
        double mx = 0.0;
        for (int r = 0; r < 3; ++r) {
          const double v = fabs((double)mat[4 * r] * a + (double)mat[4 * r + 1] * b +
                                (double)mat[4 * r + 2] * c);
          if (v > mx) mx = v;
        }
        const bool fits = mx < 1.0 + 5e-4;  // float32 slack on coordinates up to ~1e3
        if (!fits) continue;
        if (((a + b + c) & 1) == 0) ok2 = false;
        if (!(a & 1) && !(b & 1) && !(c & 1)) ok8 = false;
      }
  return ok2 ? 2 : (ok8 ? 8 : 0);
}

bool rot_cell_enabled(const RotTerm &T) { return g_rot_cell && T.cell_ncol > 0; }


static bool dirac_axis(const ::ur_proj *po, int a) {
  return po->ksize[a] == 1 && po->ratio[a] == 1;
}

bool rot_describe(const ::ur_proj *po, int op, float tau, RotFwd *F, RotTerm *T) {
  const bool sr = po->method == UR_SUPERRES;
  const int32_t *src = sr ? po->dim_yx : po->dim_x;
  RotFwd f;
  memset(&f, 0, sizeof(f));
  f.axis = -1;
  float w = 1.f;
  for (int a = 0; a < 3; ++a) {
    f.s[a] = po->dim_y[a];
    f.nyx[a] = src[a];
    f.nlr[a] = po->dim_x[a];
    if (sr && !dirac_axis(po, a)) {
      if (f.axis >= 0) return false;  // several decimated axes: general path
      f.axis = a;
    } else if (sr) {
      w *= po->ker[a][0];
      if (po->dim_yx[a] != po->dim_x[a]) return false;
    }
  }
  for (int k = 0; k < 12; ++k) f.m[k] = po->mat[k];
  if (f.axis >= 0) {
    const int a = f.axis;
    int k0 = 0, k1 = po->ksize[a];
    while (k1 - k0 > 1 && po->ker[a][k0] == 0.f) ++k0;
    while (k1 - k0 > 1 && po->ker[a][k1 - 1] == 0.f) --k1;
    f.K = k1 - k0;
    f.k0 = k0;
    f.r = po->ratio[a];
    for (int t = 0; t < f.K; ++t) f.ker[t] = po->ker[a][k0 + t];
  } else {  // no profile: a 1-tap kernel along z
    f.axis = 2;
    f.K = 1;
    f.k0 = 0;
    f.r = 1;
    f.ker[0] = 1.f;
  }
  f.nj = f.nlr[f.axis];
  f.scl_axis = -1;
  f.s_even = f.s_odd = 1.f;
  if (sr && po->scl != 0.f) {
    f.scl_axis = po->dim_thick;
    const float e = op == UR_OP_ATA ? 2.f * po->scl : po->scl;
    f.s_even = expf(e);
    f.s_odd = expf(-e);
  }
  // A: in-plane 1-tap factors once; At: once; AtA: twice (and tau)
  f.weight = op == UR_OP_ATA ? tau * (w * w) : (op == UR_OP_AT ? tau * w : w);
  if (F) *F = f;
  if (T) {
    RotTerm t;
    memset(&t, 0, sizeof(t));
    const float *mat = po->mat;
    const double a[3][3] = {{mat[0], mat[1], mat[2]}, {mat[4], mat[5], mat[6]},
                            {mat[8], mat[9], mat[10]}};
    const double tr[3] = {mat[3], mat[7], mat[11]};
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) -
                       a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (!(fabs(det) > 1e-9)) return false;
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
    double cand = 1.0;
    for (int r = 0; r < 3; ++r) {
      double hh = 0.0, off = 0.0;
      for (int c = 0; c < 3; ++c) {
        t.inv[4 * r + c] = (float)iv[r][c];
        hh += fabs(iv[r][c]);
        off -= iv[r][c] * tr[c];
      }
      t.inv[4 * r + 3] = (float)off;
      t.h[r] = (float)(hh * 1.0005 + 1e-2);  // float32 slack on coordinates up to ~1e3
      cand *= 2.0 * t.h[r] + 1.0;
      t.n[r] = f.nyx[r];
    }
    if (cand > 4096.0) return false;
    for (int k = 0; k < 12; ++k) t.m[k] = mat[k];
    t.rz = fabsf(mat[10]) >= 0.5f ? 1.f / mat[10] : 0.f;
    // cell-coefficient adjoint: a colouring exists and the candidate rows of a tile fit the table
    {
      const double ext[3] = {kCellTX + 1.0, kCellTY + 1.0, kCellTZ + 1.0};
      double rows = 1.0;
      for (int r = 0; r < 2; ++r) {
        double e = 2.2;
        for (int c = 0; c < 3; ++c) e += fabs(iv[r][c]) * ext[c];
        rows *= e;
      }
      const double nvox = (double)f.nyx[0] * f.nyx[1] * f.nyx[2];  // 32-bit offsets in the kernel
      t.cell_ncol = (rows <= (double)kCellMaxRows && nvox < 2.0e9) ? rot_cell_colours(mat) : 0;
    }
    *T = t;
  }
  return true;
}

static void rot_tile(const RotFwd &F, bool ata, RotTile *T) {
  // owned block: ~64 voxels along z (two lane trips), a few rows; along a non-z profile axis
  // 32 intermediate voxels (AtA) or their low-res rows (A)
  int e[3] = {2, 2, 64};
  auto rows_of = [&](int ext) { return (ext + F.K - 2) / F.r + 1; };  // low-res rows touching ext
  if (F.axis == 2) {
    e[0] = 4;
    e[1] = 8;
    if (ata) {  // the pulled row should fit two lane trips (64)
      int ext = 64 / F.r * F.r;
      while (ext > F.r && (rows_of(ext) - 1) * F.r + F.K > 64) ext -= F.r;
      e[2] = ext;
    } else {
      int nj = 64 / F.r > 1 ? 64 / F.r : 1;
      while (nj > 1 && (nj - 1) * F.r + F.K > 64) --nj;
      e[2] = nj;
    }
  } else {
    e[F.axis] = ata ? 32 : (32 / F.r > 2 ? 32 / F.r : 2);
    e[1 - F.axis] = 2;
  }
  for (int a = 0; a < 3; ++a) {
    T->e[a] = e[a];
    T->pe[a] = e[a];
  }
  const int njm = ata ? rows_of(e[F.axis]) + 1 : e[F.axis];
  T->nj_max = njm;
  T->pe[F.axis] = (njm - 1) * F.r + F.K;
  if (F.axis == 2) T->pe[2] |= 1;  // odd z pitch: conflict-free strided reads of phase 2
}

int rot_forward_launch(int op, const RotFwd &F, const float *v, float *out, cudaStream_t st,
                       const int *done) {
  const bool ata = op == UR_OP_ATA;
  RotTile T;
  rot_tile(F, ata, &T);
  const int *on = ata ? F.nyx : F.nlr;
  dim3 grid(div_up(on[2], T.e[2]), div_up(on[1], T.e[1]), div_up(on[0], T.e[0]));
  const size_t tile = (size_t)T.pe[0] * T.pe[1] * T.pe[2];
  const size_t smem = (ata ? 2 * tile : tile) * sizeof(float);
  UR_REQUIRE(smem <= 96 * 1024, "rot_forward: tile does not fit shared memory (K=%d r=%d)", F.K,
             F.r);
  if (ata) {
    if (smem > 48 * 1024)
      UR_CUDA_CHECK(cudaFuncSetAttribute((const void *)rot_forward_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    rot_forward_kernel<true><<<grid, kRotThreads, smem, st>>>(v, out, F, T, done);
  } else {
    if (smem > 48 * 1024)
      UR_CUDA_CHECK(cudaFuncSetAttribute((const void *)rot_forward_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    rot_forward_kernel<false><<<grid, kRotThreads, smem, st>>>(v, out, F, T, done);
  }
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int rot_expand_launch(const RotFwd &F, const float *x, float *u, cudaStream_t st) {
  dim3 block(64, 4, 1), grid(div_up(F.nyx[2], 64), div_up(F.nyx[1], 4), F.nyx[0]);
  rot_expand_kernel<<<grid, block, 0, st>>>(x, u, F);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int rot_adjoint_launch(const RotTerm &T, const int dim_y[3], float *out, int accumulate,
                       cudaStream_t st, const int *done) {
  if (rot_cell_enabled(T)) {
    constexpr int CELLS = (kCellTX + 1) * (kCellTY + 1) * (kCellTZ + 1);
    constexpr size_t smem = (size_t)8 * ((CELLS + 3) / 4 * 4) * sizeof(float) +
                            (size_t)kCellMaxRows * sizeof(int4);
    static bool attr_set[64] = {false};  // the opt-in to > 48 KB is per device
    int dev = 0;
    UR_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      UR_CUDA_CHECK(cudaFuncSetAttribute(
          (const void *)rot_adjoint_cell_kernel<kCellTX, kCellTY, kCellTZ>,
          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[dev] = true;
    }
    dim3 grid(div_up(dim_y[2], kCellTZ), div_up(dim_y[1], kCellTY), div_up(dim_y[0], kCellTX));
    RotTerm Tc = T;
    if (g_rot_cell == 8) Tc.cell_ncol = 8;  // valid whenever two colours are
    rot_adjoint_cell_kernel<kCellTX, kCellTY, kCellTZ><<<grid, kCellThreads, smem, st>>>(
        Tc, out, dim_y[0], dim_y[1], dim_y[2], accumulate, done);
    UR_LAUNCH_CHECK();
    return UR_OK;
  }
  if (dim_y[2] % 4 == 0 && ((uintptr_t)out & 15u) == 0) {
    dim3 block(32, 8, 1), grid(div_up(dim_y[2], 128), div_up(dim_y[1], 8), dim_y[0]);
    rot_adjoint4_kernel<<<grid, block, 0, st>>>(T, out, dim_y[0], dim_y[1], dim_y[2], accumulate);
  } else {
    dim3 block(64, 4, 1), grid(div_up(dim_y[2], 64), div_up(dim_y[1], 4), dim_y[0]);
    rot_adjoint_kernel<<<grid, block, 0, st>>>(T, out, dim_y[0], dim_y[1], dim_y[2], accumulate);
  }
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

extern "C" int ur_rot_cell_colours(const float mat[12]) {
  if (!mat) return 0;
  return ur::rot_cell_colours(mat);
}

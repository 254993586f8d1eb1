// Rotated (rigidly mis-aligned) observations: A = S.C.P with P a trilinear pull under a general
// affine map (unires/_project.py:147-179: solve -> grid -> pull -> conv -> scale -> conv' -> push).
//
// Two kernels replace the reference's chain for the CG matvec:
//   rot_forward_kernel  v -> u = tau C' S^2 C P v  on the intermediate grid (dim_yx): a CTA
//                       pulls a tile of the intermediate grid into shared memory (each voxel
//                       pulled once, + the profile's halo along the thick axis), forms the
//                       low-resolution rows, scales them and expands them again in-tile;
//   the adjoint P' u    as a GATHER (no atomics, deterministic) evaluated inside the lhs kernel
//                       that also does D'D and the CG epilogue (lhs_direct_kernel, solver.cu).
// The same forward kernel produces A v (the low-resolution image) for the objective, the
// scaling and the rigid updates.
#pragma once
#include "common.cuh"

namespace ur {

constexpr float kRotFovTol = 5e-2f;  // nitorch's in-FOV tolerance (extrapolate=False)

// Adjoint side: everything a thread needs to gather P' u at one recon voxel.
struct RotTerm {
  float m[12];    // intermediate index -> recon voxel coordinates (float32, as the reference)
  float inv[12];  // recon voxel -> real intermediate index: rows of [A^-1 | -A^-1 t]
  float h[3];     // half extents of the candidate box per intermediate axis
  int n[3];       // dim_yx
  float rz;       // 1 / m[10] when the z rows are resolved exactly (|m[10]| >= 0.5), else 0
  int cell_ncol;  // cell-coefficient adjoint (rot_adjoint_cell_kernel): colour passes, 0 = n/a
  const float *u; // (dim_yx) volume written by rot_forward_kernel
};

// sum over the intermediate voxels p whose image c(p) = M p + t lies within one voxel of
// q = (x, y, z) on every axis and inside the FOV of:   u[p] * prod_a (1 - |c_a(p) - q_a|).
// The weights are bit-identical to the trilinear corner weights of the pull (c - q is exact).
__device__ __forceinline__ float rot_gather(const RotTerm &T, int x, int y, int z, int nx, int ny,
                                            int nz) {
  const float fq[3] = {(float)x, (float)y, (float)z};
  int lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pc = T.inv[4 * a + 0] * fq[0] + T.inv[4 * a + 1] * fq[1] +
                     T.inv[4 * a + 2] * fq[2] + T.inv[4 * a + 3];
    lo[a] = max(0, (int)ceilf(pc - T.h[a]));
    hi[a] = min(T.n[a] - 1, (int)floorf(pc + T.h[a]));
  }
  // the FOV test of the pull, -tol < c_a < n_a - 1 + tol, can only fail next to a face of the
  // recon grid (|c - q| < 1 otherwise keeps c inside); x / y faces are warp-uniform
  const bool face_xy = x == 0 || x == nx - 1 || y == 0 || y == ny - 1;
  const float fmax_x = (float)(nx - 1) + kRotFovTol, fmax_y = (float)(ny - 1) + kRotFovTol,
              fmax_z = (float)(nz - 1) + kRotFovTol;
  float acc = 0.f;
  for (int i = lo[0]; i <= hi[0]; ++i) {
    const float fi = (float)i;
    const float ax = T.m[0] * fi, ay = T.m[4] * fi, az = T.m[8] * fi;
    for (int j = lo[1]; j <= hi[1]; ++j) {
      const float fj = (float)j;
      const float bx = fmaf(T.m[1], fj, ax), by = fmaf(T.m[5], fj, ay), bz = fmaf(T.m[9], fj, az);
      int k0 = lo[2], k1 = hi[2];
      if (T.rz != 0.f) {  // -1 < m10 k + (bz + t_z - q_z) < 1, widened against rounding
        const float s = (bz + T.m[11]) - fq[2];
        const float e0 = (-1.f - s) * T.rz, e1 = (1.f - s) * T.rz;
        k0 = max(k0, (int)ceilf(fminf(e0, e1) - 1e-3f));
        k1 = min(k1, (int)floorf(fmaxf(e0, e1) + 1e-3f));
      }
      const float *row = T.u + ((size_t)i * T.n[1] + j) * T.n[2];
      for (int k = k0; k <= k1; ++k) {
        const float fk = (float)k;
        const float cx = fmaf(T.m[2], fk, bx) + T.m[3];
        const float cy = fmaf(T.m[6], fk, by) + T.m[7];
        const float cz = fmaf(T.m[10], fk, bz) + T.m[11];
        float wx = 1.f - fabsf(cx - fq[0]);
        float wy = 1.f - fabsf(cy - fq[1]);
        float wz = 1.f - fabsf(cz - fq[2]);
        bool ok = cz > -kRotFovTol && cz < fmax_z;
        if (face_xy) ok = ok && cx > -kRotFovTol && cx < fmax_x && cy > -kRotFovTol && cy < fmax_y;
        wx = fmaxf(wx, 0.f);
        wy = fmaxf(wy, 0.f);
        wz = ok ? fmaxf(wz, 0.f) : 0.f;
        acc = fmaf(__ldg(row + k), (wx * wy) * wz, acc);
      }
    }
  }
  return acc;
}

// The same gather for the four recon voxels (x, y, z .. z+3) of one thread: the x / y weights of
// an intermediate voxel are shared by the z neighbours it contributes to, so a quad costs about
// half the instructions of four single gathers (the gather is instruction-issue bound).
template <bool FACE_XY>
__device__ __forceinline__ void rot_gather4_impl(const RotTerm &T, int x, int y, int z, int nx,
                                                 int ny, int nz, float (&out)[4]) {
  const float fx = (float)x, fy = (float)y, fz = (float)z;
  int lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float p0 = T.inv[4 * a + 0] * fx + T.inv[4 * a + 1] * fy + T.inv[4 * a + 2] * fz +
                     T.inv[4 * a + 3];
    const float p3 = p0 + 3.f * T.inv[4 * a + 2];
    lo[a] = max(0, (int)ceilf(fminf(p0, p3) - T.h[a]));
    hi[a] = min(T.n[a] - 1, (int)floorf(fmaxf(p0, p3) + T.h[a]));
  }
  const float fmax_x = (float)(nx - 1) + kRotFovTol, fmax_y = (float)(ny - 1) + kRotFovTol,
              fmax_z = (float)(nz - 1) + kRotFovTol;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int i = lo[0]; i <= hi[0]; ++i) {
    const float fi = (float)i;
    const float ax = T.m[0] * fi, ay = T.m[4] * fi, az = T.m[8] * fi;
    for (int j = lo[1]; j <= hi[1]; ++j) {
      const float fj = (float)j;
      const float bx = fmaf(T.m[1], fj, ax), by = fmaf(T.m[5], fj, ay), bz = fmaf(T.m[9], fj, az);
      int k0 = lo[2], k1 = hi[2];
      if (T.rz != 0.f) {  // z - 1 < m10 k + (bz + t_z) < z + 4, widened against rounding
        const float s = bz + T.m[11];
        const float e0 = ((fz - 1.f) - s) * T.rz, e1 = ((fz + 4.f) - s) * T.rz;
        k0 = max(k0, (int)ceilf(fminf(e0, e1) - 1e-3f));
        k1 = min(k1, (int)floorf(fmaxf(e0, e1) + 1e-3f));
      }
      const float *row = T.u + ((size_t)i * T.n[1] + j) * T.n[2] + k0;
      float fk = (float)k0;
      // branch-free body: the load is unconditional (k is inside the grid), the FOV test only
      // zeroes the weight
      for (int k = k0; k <= k1; ++k, fk += 1.f, ++row) {
        const float uv = __ldg(row);
        const float cx = fmaf(T.m[2], fk, bx) + T.m[3];
        const float cy = fmaf(T.m[6], fk, by) + T.m[7];
        const float cz = fmaf(T.m[10], fk, bz) + T.m[11];
        const float wx = fmaxf(1.f - fabsf(cx - fx), 0.f);
        const float wy = fmaxf(1.f - fabsf(cy - fy), 0.f);
        float w = wx * wy;
        w = (cz > -kRotFovTol && cz < fmax_z) ? w : 0.f;
        if (FACE_XY)  // compiled out for the interior rows (12 % of the instructions, ncu)
          w = (cx > -kRotFovTol && cx < fmax_x && cy > -kRotFovTol && cy < fmax_y) ? w : 0.f;
        const float vt = uv * w;
        const float dz = cz - fz;
        a0 = fmaf(vt, fmaxf(1.f - fabsf(dz), 0.f), a0);
        a1 = fmaf(vt, fmaxf(1.f - fabsf(dz - 1.f), 0.f), a1);
        a2 = fmaf(vt, fmaxf(1.f - fabsf(dz - 2.f), 0.f), a2);
        a3 = fmaf(vt, fmaxf(1.f - fabsf(dz - 3.f), 0.f), a3);
      }
    }
  }
  out[0] = a0;
  out[1] = a1;
  out[2] = a2;
  out[3] = a3;
}

// x / y are warp-uniform in the kernels that call this (a warp owns one row): a real branch
__device__ __forceinline__ void rot_gather4(const RotTerm &T, int x, int y, int z, int nx, int ny,
                                            int nz, float (&out)[4]) {
  if (x == 0 || x == nx - 1 || y == 0 || y == ny - 1)
    rot_gather4_impl<true>(T, x, y, z, nx, ny, nz, out);
  else
    rot_gather4_impl<false>(T, x, y, z, nx, ny, nz, out);
}

// Host description of the forward kernel's operator (at most ONE decimated axis).
struct RotFwd {
  int s[3];    // recon grid (source of the pull)
  int nyx[3];  // intermediate grid
  int nlr[3];  // low-resolution grid (dim_x)
  float m[12];
  int axis;    // profile axis (a 1-tap profile along z when the operator has none)
  int K, r, k0;  // trimmed taps: x[j] = sum_{t<K} ker[t] yx[j r + k0 + t]
  int nj;
  float ker[UR_MAX_TAPS];
  int scl_axis;  // even/odd scaling along this low-res axis, -1 = none
  float s_even, s_odd;
  float weight;  // multiplies the output (tau and the in-plane 1-tap factors)
};

// po -> RotFwd / RotTerm; false when the operator does not fit (several decimated axes, ...)
bool rot_describe(const ::ur_proj *po, int op, float tau, RotFwd *F, RotTerm *T);
// op = UR_OP_A: out = S C P v on dim_x;  UR_OP_ATA: out = tau C' S^2 C P v on dim_yx
int rot_forward_launch(int op, const RotFwd &F, const float *v, float *out, cudaStream_t st,
                       const int *done = nullptr);
// stand-alone adjoint: out (dim_y) (+)= P' u  (accumulate = 0 overwrites): the cell-coefficient
// kernel when the operator allows it (T.cell_ncol > 0 and ur_tune("rot_cell") != 0), else the
// per-voxel gather
int rot_adjoint_launch(const RotTerm &T, const int dim_y[3], float *out, int accumulate,
                       cudaStream_t st, const int *done = nullptr);
bool rot_cell_enabled(const RotTerm &T);
// C' S x -> u on dim_yx for At (x on dim_x), scaled by F.weight
int rot_expand_launch(const RotFwd &F, const float *x, float *u, cudaStream_t st);

}  // namespace ur

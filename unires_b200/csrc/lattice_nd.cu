// Lattice-aligned observations decimated along SEVERAL axes (0.5 mm reconstruction of 1 mm
// isotropic data: ratio 2 on every axis, rect-5 x gauss-9 x gauss-9 profile; BASELINE configs[4]).
//
//   A  = B_x B_y B_z (x) crop      x[j] = sum_t ker_a[t] v[j r_a + t + off_a]   per axis
//   A'A v = A' (A v)
//
// Round 1 evaluated A'A as a chain of three full-resolution single-axis passes plus a final
// stencil pass: 36 bytes per voxel for 8 algorithmic.  Here the LOW-RESOLUTION image A v (1/8
// of the voxels at ratio 2) is the only intermediate that touches HBM:
//   nd_down_kernel   v -> A v        a CTA stages the input box of a low-res tile in shared
//                                    memory and decimates it axis by axis (z, y, x);
//   nd_up_kernel     A v -> tau A'(A v) + rho lam^2 D'D v + CG epilogue: a CTA expands the
//                                    low-res box of an output tile axis by axis (x, y, then z on
//                                    the fly per quad) and finishes with the 7-point stencil.
// 4.5 + 8.5 = 13 bytes per voxel.  The same kernels serve A (objective) and A' (right-hand
// side).  Reference: unires/_project.py:147-179 with an identity rotation (pull = crop, push =
// zero-pad embed, F.conv3d / F.conv_transpose3d with stride = ratio).
#include <string.h>

#include "lattice_nd.cuh"

namespace ur {

constexpr int kNdThreads = 256;
constexpr int kNdWarps = kNdThreads / 32;

struct NdDownTile {
  int L[3];   // low-res tile
  int I[3];   // input box extents (L-1) r + K
  int nt[3];  // tiles per axis
};

struct NdUpTile {
  int E[3];    // output tile
  int NJ[3];   // max low-res rows per axis touching a tile
  int nt[3];
};

// 4-byte asynchronous global -> shared copy with zero fill (src-size 0 reads nothing): every
// thread fires all the copies of a box back to back and waits once, instead of one exposed
// global-load round trip per row (the first version of these kernels ran at 1.8 ms per pass)
__device__ __forceinline__ void nd_cp_async4(float *dst, const float *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void nd_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// warps over rows (c0, c1), lanes along c2 -- no per-element division
#define ND_ROWS_BEGIN(n0, n1)                   \
  if ((n1) > 0) {                               \
    int c0 = warp / (n1), c1 = warp - c0 * (n1); \
    while (c0 < (n0)) {
#define ND_ROWS_END(n1)       \
      c1 += kNdWarps;         \
      while (c1 >= (n1)) {    \
        c1 -= (n1);           \
        ++c0;                 \
      }                       \
    }                         \
  }

// out (dim_x) = scale * A v
__global__ void __launch_bounds__(kNdThreads)
    nd_down_kernel(const float *__restrict__ v, float *__restrict__ out, const NdOp op,
                   const NdDownTile T, float scale, const int *done) {
  extern __shared__ float sm[];
  if (done && *done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int I0 = T.I[0], I1 = T.I[1], I2 = T.I[2];
  const int L1 = T.L[1], L2 = T.L[2];
  const int pI2 = I2 | 1;               // odd pitch: the stride-r reads of the z pass spread over banks
  float *in = sm;                       // [I0][I1][pI2]
  float *t1 = sm + I0 * I1 * pI2;       // [I0][I1][L2]
  float *t2 = sm;                       // [I0][L1][L2]  (aliases `in`, dead after the z pass)
  const size_t sy = op.n[2], sx = (size_t)op.n[1] * op.n[2];
  const int ntiles = T.nt[0] * T.nt[1] * T.nt[2];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b2 = tile % T.nt[2], tq = tile / T.nt[2];
    const int b1 = tq % T.nt[1], b0 = tq / T.nt[1];
    const int j0 = b0 * T.L[0], j1 = b1 * T.L[1], j2 = b2 * T.L[2];
    const int l0n = min(T.L[0], op.ax[0].nj - j0), l1n = min(T.L[1], op.ax[1].nj - j1),
              l2n = min(T.L[2], op.ax[2].nj - j2);
    const int s0 = j0 * op.ax[0].r + op.ax[0].off, s1 = j1 * op.ax[1].r + op.ax[1].off,
              s2 = j2 * op.ax[2].r + op.ax[2].off;
    const int i0n = (l0n - 1) * op.ax[0].r + op.ax[0].K, i1n = (l1n - 1) * op.ax[1].r + op.ax[1].K,
              i2n = (l2n - 1) * op.ax[2].r + op.ax[2].K;
    __syncthreads();  // previous tile's x pass is done with t2
    // ---- load the input box (zero outside the recon grid: bound = 'zero') ----
    ND_ROWS_BEGIN(i0n, i1n)
      const int gx = s0 + c0, gy = s1 + c1;
      const bool row_in = gx >= 0 && gx < op.n[0] && gy >= 0 && gy < op.n[1];
      const float *src = row_in ? v + (size_t)gx * sx + (size_t)gy * sy : v;
      for (int c2 = lane; c2 < i2n; c2 += 32) {
        const int gz = s2 + c2;
        const bool ok = row_in && gz >= 0 && gz < op.n[2];
        nd_cp_async4(in + (c0 * I1 + c1) * pI2 + c2, ok ? src + gz : v, ok);
      }
    ND_ROWS_END(i1n)
    nd_cp_async_wait_all();
    __syncthreads();
    // ---- z pass ----
    ND_ROWS_BEGIN(i0n, i1n)
      for (int l2 = lane; l2 < l2n; l2 += 32) {
        const float *p = in + (c0 * I1 + c1) * pI2 + l2 * op.ax[2].r;
        float acc = 0.f;
        for (int t = 0; t < op.ax[2].K; ++t) acc = fmaf(op.ax[2].ker[t], p[t], acc);
        t1[(c0 * I1 + c1) * L2 + l2] = acc;
      }
    ND_ROWS_END(i1n)
    __syncthreads();
    // ---- y pass ----
    ND_ROWS_BEGIN(i0n, l1n)
      for (int l2 = lane; l2 < l2n; l2 += 32) {
        const float *p = t1 + (c0 * I1 + c1 * op.ax[1].r) * L2 + l2;
        float acc = 0.f;
        for (int t = 0; t < op.ax[1].K; ++t) acc = fmaf(op.ax[1].ker[t], p[t * L2], acc);
        t2[(c0 * L1 + c1) * L2 + l2] = acc;
      }
    ND_ROWS_END(l1n)
    __syncthreads();
    // ---- x pass -> global ----
    ND_ROWS_BEGIN(l0n, l1n)
      float *dst = out + ((size_t)(j0 + c0) * op.ax[1].nj + (j1 + c1)) * op.ax[2].nj + j2;
      for (int l2 = lane; l2 < l2n; l2 += 32) {
        const float *p = t2 + (c0 * op.ax[0].r * L1 + c1) * L2 + l2;
        float acc = 0.f;
        for (int t = 0; t < op.ax[0].K; ++t) acc = fmaf(op.ax[0].ker[t], p[t * L1 * L2], acc);
        dst[l2] = scale * acc;
      }
    ND_ROWS_END(l1n)
  }
}

// MODE = LHS_TERM : out = (acc +) scale * A' xl                      (right-hand side)
// else            : out = scale * A' xl + w_ident v + acc + rho lam^2 D'D v, CG epilogue
// Requires nz % 4 == 0 and 16-byte aligned volumes.
template <int MODE>
__global__ void __launch_bounds__(kNdThreads, 2)
    nd_up_kernel(const float *__restrict__ xl, const NdOp op, const NdUpTile T, float scale,
                 const LhsArgs a) {
  extern __shared__ float sm[];
  __shared__ double s_red[kMaxWarps];
  __shared__ float s_kz[UR_MAX_TAPS];  // z taps: the per-lane tap index diverges (LDS, not LDC)
  if (a.done && *a.done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < UR_MAX_TAPS) s_kz[tid] = tid < op.ax[2].K ? op.ax[2].ker[tid] : 0.f;
  const int NJ1 = T.NJ[1], NJ2 = T.NJ[2];
  const int E0 = T.E[0], E1 = T.E[1];
  float *lr = sm;                         // [NJ0][NJ1][NJ2]
  float *u1 = lr + T.NJ[0] * NJ1 * NJ2;   // [E0][NJ1][NJ2]
  float *u2 = u1 + E0 * NJ1 * NJ2;        // [E0][E1][NJ2]
  const size_t ly = op.ax[2].nj, lx = (size_t)op.ax[1].nj * op.ax[2].nj;
  const int ntiles = T.nt[0] * T.nt[1] * T.nt[2];
  double part = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b2 = tile % T.nt[2], tq = tile / T.nt[2];
    const int b1 = tq % T.nt[1], b0 = tq / T.nt[1];
    const int o0 = b0 * T.E[0], o1 = b1 * T.E[1], o2 = b2 * T.E[2];
    const int e0n = min(T.E[0], op.n[0] - o0), e1n = min(T.E[1], op.n[1] - o1),
              e2n = min(T.E[2], op.n[2] - o2);
    // low-res rows whose support meets the tile, per axis
    int jb[3], jn[3];
    {
      const int o[3] = {o0, o1, o2}, en[3] = {e0n, e1n, e2n};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        jb[d] = nd_ceildiv(o[d] - op.ax[d].off - op.ax[d].K + 1, op.ax[d].r);
        jn[d] = nd_floordiv(o[d] + en[d] - 1 - op.ax[d].off, op.ax[d].r) - jb[d] + 1;
        if (jn[d] < 0) jn[d] = 0;
      }
    }
    __syncthreads();  // previous tile is done with u2
    // ---- load the low-res box (rows outside [0, nj) do not exist: zero) ----
    ND_ROWS_BEGIN(jn[0], jn[1])
      const int g0 = jb[0] + c0, g1 = jb[1] + c1;
      const bool row_in = g0 >= 0 && g0 < op.ax[0].nj && g1 >= 0 && g1 < op.ax[1].nj;
      const float *src = row_in ? xl + (size_t)g0 * lx + (size_t)g1 * ly : xl;
      for (int c2 = lane; c2 < jn[2]; c2 += 32) {
        const int g2 = jb[2] + c2;
        const bool ok = row_in && g2 >= 0 && g2 < op.ax[2].nj;
        nd_cp_async4(lr + (c0 * NJ1 + c1) * NJ2 + c2, ok ? src + g2 : xl, ok);
      }
    ND_ROWS_END(jn[1])
    nd_cp_async_wait_all();
    __syncthreads();
    // ---- x pass: u1[e0][c1][c2] = sum_j kx[(o0 + e0 - off) - j r] lr[j - jb][c1][c2] ----
    ND_ROWS_BEGIN(e0n, jn[1])
      const int u = o0 + c0 - op.ax[0].off;
      int ja = nd_ceildiv(u - op.ax[0].K + 1, op.ax[0].r) - jb[0];
      int jz = nd_floordiv(u, op.ax[0].r) - jb[0];
      if (ja < 0) ja = 0;
      if (jz > jn[0] - 1) jz = jn[0] - 1;
      for (int c2 = lane; c2 < jn[2]; c2 += 32) {
        float acc = 0.f;
        for (int jl = ja; jl <= jz; ++jl)
          acc = fmaf(op.ax[0].ker[u - (jb[0] + jl) * op.ax[0].r], lr[(jl * NJ1 + c1) * NJ2 + c2], acc);
        u1[(c0 * NJ1 + c1) * NJ2 + c2] = acc;
      }
    ND_ROWS_END(jn[1])
    __syncthreads();
    // ---- y pass: u2[e0][e1][c2] ----
    ND_ROWS_BEGIN(e0n, e1n)
      const int u = o1 + c1 - op.ax[1].off;
      int ja = nd_ceildiv(u - op.ax[1].K + 1, op.ax[1].r) - jb[1];
      int jz = nd_floordiv(u, op.ax[1].r) - jb[1];
      if (ja < 0) ja = 0;
      if (jz > jn[1] - 1) jz = jn[1] - 1;
      for (int c2 = lane; c2 < jn[2]; c2 += 32) {
        float acc = 0.f;
        for (int jl = ja; jl <= jz; ++jl)
          acc = fmaf(op.ax[1].ker[u - (jb[1] + jl) * op.ax[1].r], u1[(c0 * NJ1 + jl) * NJ2 + c2], acc);
        u2[(c0 * E1 + c1) * NJ2 + c2] = acc;
      }
    ND_ROWS_END(e1n)
    __syncthreads();
    // ---- z pass on the fly per quad, stencil, epilogue ----
    // a lane keeps its quad (z = o2 + 4 lane) for the whole tile: the low-res window and the
    // first tap of each of its four voxels are per-tile constants (no division per voxel)
    if (4 * lane < e2n) {
      const int z = o2 + 4 * lane;
      int wa[4], wn[4], wt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int u = z + k - op.ax[2].off;
        int ja = nd_ceildiv(u - op.ax[2].K + 1, op.ax[2].r) - jb[2];
        int jz = nd_floordiv(u, op.ax[2].r) - jb[2];
        if (ja < 0) ja = 0;
        if (jz > jn[2] - 1) jz = jn[2] - 1;
        wa[k] = ja;
        wn[k] = jz - ja + 1;
        wt[k] = u - (jb[2] + ja) * op.ax[2].r;  // tap of the first row of the window
      }
      const int rz = op.ax[2].r;
      if (e1n > 0) {
        int c0 = warp / e1n, c1 = warp - c0 * e1n;
        NdQuadIn cur, nxt;
        if (c0 < e0n) nd_quad_load<MODE>(a, o0 + c0, o1 + c1, z, cur);
        while (c0 < e0n) {
          int n0 = c0, n1 = c1 + kNdWarps;
          while (n1 >= e1n) {
            n1 -= e1n;
            ++n0;
          }
          if (n0 < e0n) nd_quad_load<MODE>(a, o0 + n0, o1 + n1, z, nxt);  // one row ahead
          const float *row = u2 + (c0 * E1 + c1) * NJ2;
          float data[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float acc = 0.f;
            for (int n = 0; n < wn[k]; ++n) acc = fmaf(s_kz[wt[k] - n * rz], row[wa[k] + n], acc);
            data[k] = scale * acc;
          }
          nd_quad_finish<MODE>(a, o0 + c0, o1 + c1, z,
                               ((size_t)(o0 + c0) * a.ny + (o1 + c1)) * a.nz + z, cur, data, part);
          cur = nxt;
          c0 = n0;
          c1 = n1;
        }
      }
    }
  }
  if (MODE == LHS_TERM) return;
  double total;
  if (grid_sum(part, a.gr, s_red, &total) && tid == 0) finalize(a.fin, total);
}

// ---------------------------------------------------------------------------
// Data term of the objective in ONE pass for lattice operators with at most one decimated
// axis (_compute_nll, unires/_update.py:411-417):  0.5 tau sum_{x != 0} (x - A y)^2.
// The reference materialises A y (pull, conv3d, scaling) and then compacts it with a boolean
// mask; here a thread forms (A y)[j] for its low-resolution voxel from the K taps it needs and
// accumulates the masked squared residual in float64: y is read once, nothing is written.
// Same per-element arithmetic as the unfused route (conv taps in order, then the in-plane
// 1-tap factors, then exp(+-scl); __fsub_rn / __fmul_rn on the residual).
// ---------------------------------------------------------------------------
struct NllNdArgs {
  NdOp op;
  int ca;          // decimated axis, -1 = none
  float inplane;   // product of the 1-tap factors of the other axes
  int scl_axis;    // even/odd scaling along this low-res axis, -1 = none
  float s_even, s_odd;
  float tau;
};

__global__ void __launch_bounds__(256)
    nll_nd_kernel(const float *__restrict__ y, const float *__restrict__ x, const NllNdArgs A,
                  int accumulate, GridReduce gr, double *out) {
  __shared__ double s_red[kMaxWarps];
  const NdOp &op = A.op;
  const int n0 = op.ax[0].nj, n1 = op.ax[1].nj, n2 = op.ax[2].nj;
  const size_t sy = op.n[2], sx = (size_t)op.n[1] * op.n[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  double part = 0.0;
  // warps over the (j0, j1) rows of the low-res grid, lanes along j2
  for (long long row = (long long)blockIdx.x * wpb + warp; row < (long long)n0 * n1;
       row += (long long)gridDim.x * wpb) {
    const int j0 = (int)(row / n1), j1 = (int)(row - (long long)j0 * n1);
    const int g0 = j0 * op.ax[0].r + op.ax[0].off, g1 = j1 * op.ax[1].r + op.ax[1].off;
    for (int j2 = lane; j2 < n2; j2 += 32) {
      const float xv = __ldg(x + ((size_t)j0 * n1 + j1) * n2 + j2);
      if (xv == 0.f) continue;
      const int g2 = j2 * op.ax[2].r + op.ax[2].off;
      int g[3] = {g0, g1, g2};
      float acc = 0.f;
      const int K = A.ca >= 0 ? op.ax[A.ca].K : 1;
      bool inside = true;  // the non-decimated axes: a crop (zero outside the recon grid)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (d != A.ca) inside = inside && g[d] >= 0 && g[d] < op.n[d];
      if (inside) {
        const size_t step = A.ca == 0 ? sx : (A.ca == 1 ? sy : 1);
        const int gc = A.ca >= 0 ? g[A.ca] : 0, nc = A.ca >= 0 ? op.n[A.ca] : 1;
        const float *base = y + (size_t)(A.ca == 0 ? 0 : g[0]) * sx +
                            (size_t)(A.ca == 1 ? 0 : g[1]) * sy + (A.ca == 2 ? 0 : g[2]);
        if (A.ca < 0) {
          acc = __ldg(base);
        } else {
          for (int t = 0; t < K; ++t) {
            const int q = gc + t;
            const float v = (q >= 0 && q < nc) ? __ldg(base + (size_t)q * step) : 0.f;
            acc = fmaf(op.ax[A.ca].ker[t], v, acc);
          }
        }
      }
      if (A.inplane != 1.f) acc *= A.inplane;
      if (A.scl_axis >= 0) {
        const int js = A.scl_axis == 0 ? j0 : (A.scl_axis == 1 ? j1 : j2);
        acc *= (js & 1) ? A.s_odd : A.s_even;
      }
      const float dlt = __fsub_rn(xv, acc);
      part += (double)__fmul_rn(dlt, dlt);
    }
  }
  double total;
  if (grid_sum(part, gr, s_red, &total) && threadIdx.x == 0) {
    total = 0.5 * (double)A.tau * total;
    *out = accumulate ? *out + total : total;
  }
}

int scratch_reduce(GridReduce *gr, cudaStream_t st);  // vecops.cu

// UR_ERR_UNSUPPORTED (nothing launched) unless the operator is lattice aligned with at most
// one decimated axis
int nll_nd_launch(const ::ur_proj *po, const float *y, const float *x, float tau, double *out,
                  int accumulate, cudaStream_t st) {
  ::ur_proj p0 = *po;
  p0.scl = 0.f;
  NllNdArgs A;
  memset(&A, 0, sizeof(A));
  if (!nd_describe(&p0, tau, &A.op) || nd_conv_axes(A.op) > 1) return UR_ERR_UNSUPPORTED;
  A.ca = -1;
  A.inplane = 1.f;
  for (int a = 0; a < 3; ++a) {
    if (A.op.ax[a].K > 1 || A.op.ax[a].r > 1)
      A.ca = a;
    else
      A.inplane *= A.op.ax[a].ker[0];
  }
  A.scl_axis = -1;
  A.s_even = A.s_odd = 1.f;
  if (po->method == UR_SUPERRES && po->scl != 0.f) {
    A.scl_axis = po->dim_thick;
    A.s_even = expf(po->scl);
    A.s_odd = expf(-po->scl);
  }
  A.tau = tau;
  GridReduce gr;
  int rc = scratch_reduce(&gr, st);
  if (rc) return rc;
  const long long rows = (long long)A.op.ax[0].nj * A.op.ax[1].nj;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  nll_nd_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, x, A, accumulate, gr, out);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int g_nd_fused = 1;  // ur_tune("nd_fused"): 0 = chained single-axis passes / general path

// po (lattice aligned, no even/odd scaling) -> NdOp
bool nd_describe(const ::ur_proj *po, float tau, NdOp *op) {
  if (!ur_proj_is_lattice(po) || po->scl != 0.f) return false;
  const bool sr = po->method == UR_SUPERRES;
  memset(op, 0, sizeof(*op));
  op->tau = tau;
  for (int a = 0; a < 3; ++a) {
    op->n[a] = po->dim_y[a];
    NdAxis &A = op->ax[a];
    const int shift = (int)lrintf(po->mat[4 * a + 3]);
    A.nj = po->dim_x[a];
    if (sr && (po->ksize[a] > 1 || po->ratio[a] > 1)) {
      int k0 = 0, k1 = po->ksize[a];
      while (k1 - k0 > 1 && po->ker[a][k0] == 0.f) ++k0;
      while (k1 - k0 > 1 && po->ker[a][k1 - 1] == 0.f) --k1;
      A.K = k1 - k0;
      A.r = po->ratio[a];
      A.off = shift + k0;
      for (int t = 0; t < A.K; ++t) A.ker[t] = po->ker[a][k0 + t];
    } else {
      A.K = 1;
      A.r = 1;
      A.off = shift;
      A.ker[0] = sr ? po->ker[a][0] : 1.f;
    }
  }
  return true;
}

int nd_conv_axes(const NdOp &op) {
  int n = 0;
  for (int a = 0; a < 3; ++a) n += op.ax[a].K > 1 || op.ax[a].r > 1;
  return n;
}

static int nd_grid(int ntiles, const void *kernel, size_t smem) {
  int resident = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, kNdThreads, smem) !=
          cudaSuccess ||
      resident < 1)
    resident = 1;
  const long long cap = (long long)resident * sm_count();
  return (int)(ntiles < cap ? ntiles : cap);
}

static int nd_opt_in(const void *kernel, size_t smem) {
  if (smem > 48 * 1024)
    UR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
  return UR_OK;
}

int nd_down_launch(const NdOp &op, const float *v, float *out, float scale, const int *done,
                   cudaStream_t st) {
  {
    const int rc = nd_down_spec_launch(op, v, out, scale, done, st);
    if (rc != UR_ERR_UNSUPPORTED) return rc;
  }
  NdDownTile T;
  const int want[3] = {4, 8, 32};
  for (int a = 0; a < 3; ++a) {
    int L = want[a];
    // undecimated axes move r = 1 voxel per row: a larger tile costs nothing in halo
    if (op.ax[a].r == 1 && op.ax[a].K == 1 && a < 2) L = a == 0 ? 4 : 8;
    if (L > op.ax[a].nj) L = op.ax[a].nj;
    T.L[a] = L;
    T.I[a] = (L - 1) * op.ax[a].r + op.ax[a].K;
    T.nt[a] = (op.ax[a].nj + L - 1) / L;
  }
  const size_t pI2 = (size_t)(T.I[2] | 1);
  const size_t n_in = (size_t)T.I[0] * T.I[1] * pI2;
  const size_t n_t1 = (size_t)T.I[0] * T.I[1] * T.L[2];
  const size_t n_t2 = (size_t)T.I[0] * T.L[1] * T.L[2];
  const size_t smem = ((n_in > n_t2 ? n_in : n_t2) + n_t1) * sizeof(float);
  if (smem > 200 * 1024) return UR_ERR_UNSUPPORTED;
  int rc = nd_opt_in((const void *)nd_down_kernel, smem);
  if (rc) return rc;
  const int ntiles = T.nt[0] * T.nt[1] * T.nt[2];
  const int grid = nd_grid(ntiles, (const void *)nd_down_kernel, smem);
  nd_down_kernel<<<grid, kNdThreads, smem, st>>>(v, out, op, T, scale, done);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

static bool nd_a16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// mode: LHS_PLAIN / LHS_RESID / LHS_ENERGY / LHS_TERM.  `A` carries the volumes.
int nd_up_launch(int mode, const NdOp &op, const float *xl, float scale, const LhsArgs &A,
                 cudaStream_t st) {
  if (op.n[2] % 4 != 0) return UR_ERR_UNSUPPORTED;
  if (!nd_a16(A.v) || !nd_a16(A.out) || !nd_a16(A.b) || !nd_a16(A.r) || !nd_a16(A.p) ||
      !nd_a16(A.acc))
    return UR_ERR_UNSUPPORTED;
  {
    const int rc = nd_up_spec_launch(mode, op, xl, scale, A, st);
    if (rc != UR_ERR_UNSUPPORTED) return rc;
  }
  NdUpTile T;
  const int want[3] = {8, 16, 128};  // 32 quads per row: one per lane
  for (int a = 0; a < 3; ++a) {
    int E = want[a];
    if (E > op.n[a]) E = a == 2 ? (op.n[a] + 3) / 4 * 4 : op.n[a];
    T.E[a] = E;
    T.NJ[a] = (E + op.ax[a].K - 2) / op.ax[a].r + 2;
    T.nt[a] = (op.n[a] + E - 1) / E;
  }
  const size_t smem = ((size_t)T.NJ[0] * T.NJ[1] * T.NJ[2] + (size_t)T.E[0] * T.NJ[1] * T.NJ[2] +
                       (size_t)T.E[0] * T.E[1] * T.NJ[2]) *
                      sizeof(float);
  if (smem > 200 * 1024) return UR_ERR_UNSUPPORTED;
  const void *kernel = mode == LHS_PLAIN   ? (const void *)nd_up_kernel<LHS_PLAIN>
                       : mode == LHS_RESID ? (const void *)nd_up_kernel<LHS_RESID>
                       : mode == LHS_TERM  ? (const void *)nd_up_kernel<LHS_TERM>
                                           : (const void *)nd_up_kernel<LHS_ENERGY>;
  int rc = nd_opt_in(kernel, smem);
  if (rc) return rc;
  const int ntiles = T.nt[0] * T.nt[1] * T.nt[2];
  const int grid = nd_grid(ntiles, kernel, smem);
  switch (mode) {
    case LHS_PLAIN:
      nd_up_kernel<LHS_PLAIN><<<grid, kNdThreads, smem, st>>>(xl, op, T, scale, A);
      break;
    case LHS_RESID:
      nd_up_kernel<LHS_RESID><<<grid, kNdThreads, smem, st>>>(xl, op, T, scale, A);
      break;
    case LHS_TERM:
      nd_up_kernel<LHS_TERM><<<grid, kNdThreads, smem, st>>>(xl, op, T, scale, A);
      break;
    default:
      nd_up_kernel<LHS_ENERGY><<<grid, kNdThreads, smem, st>>>(xl, op, T, scale, A);
      break;
  }
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

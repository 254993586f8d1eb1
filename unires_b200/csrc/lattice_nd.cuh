// Lattice-aligned observations decimated along several axes: host interface of lattice_nd.cu.
#pragma once
#include "solver.cuh"

namespace ur {

// one axis of the separable operator: x[j] = sum_t ker[t] v[j r + t + off], 0 <= j < nj
struct NdAxis {
  int K, r, off, nj;
  float ker[UR_MAX_TAPS];
};

struct NdOp {
  int n[3];   // recon grid
  NdAxis ax[3];
  float tau;
};

// po (lattice aligned, no even/odd scaling) -> NdOp; false when it does not apply
bool nd_describe(const ::ur_proj *po, float tau, NdOp *op);
int nd_conv_axes(const NdOp &op);  // number of axes with a real profile / decimation
// out (dim_x) = scale * A v
int nd_down_launch(const NdOp &op, const float *v, float *out, float scale, const int *done,
                   cudaStream_t st);
// LHS_TERM: A.out = (A.acc +) scale * A' xl;  LHS_PLAIN / RESID / ENERGY: the CG left-hand side
// scale * A' xl + w_ident v + acc + rho lam^2 D'D v with the epilogue of lhs_direct_kernel.
// UR_ERR_UNSUPPORTED (nothing launched) when nz % 4 != 0 or a volume is not 16-byte aligned.
int nd_up_launch(int mode, const NdOp &op, const float *xl, float scale, const LhsArgs &A,
                 cudaStream_t st);
extern int g_nd_fused;
// 0.5 tau sum_{x != 0} (x - A y)^2 in one pass (lattice operators, <= 1 decimated axis)
int nll_nd_launch(const ::ur_proj *po, const float *y, const float *x, float tau, double *out,
                  int accumulate, cudaStream_t st);
// compile-time specialised TMA variants (lattice_nd_spec.cu); UR_ERR_UNSUPPORTED when the
// operator has no instantiation
int nd_down_spec_launch(const NdOp &op, const float *v, float *out, float scale, const int *done,
                        cudaStream_t st);
int nd_up_spec_launch(int mode, const NdOp &op, const float *xl, float scale, const LhsArgs &A,
                      cudaStream_t st);

__device__ __forceinline__ int nd_floordiv(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
__device__ __forceinline__ int nd_ceildiv(int a, int b) { return -nd_floordiv(-a, b); }

// Everything a quad reads from HBM; loaded one row ahead of its use (software pipeline: the
// row loop would otherwise expose one global round trip per row).
struct NdQuadIn {
  float4 c, xm, xp, ym, yp, acc, b;
  float zl, zr;
};

template <int MODE>
__device__ __forceinline__ void nd_quad_load(const LhsArgs &a, int x, int y, int z, NdQuadIn &q) {
  const size_t sy = a.nz, sx = (size_t)a.ny * a.nz;
  const size_t i = x * sx + y * sy + z;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  q.acc = a.acc ? *reinterpret_cast<const float4 *>(a.acc + i) : zero4;
  if (MODE == LHS_TERM) return;
  const float *__restrict__ v = a.v;
  q.c = *reinterpret_cast<const float4 *>(v + i);
  q.xm = x > 0 ? *reinterpret_cast<const float4 *>(v + i - sx) : zero4;
  q.xp = x + 1 < a.nx ? *reinterpret_cast<const float4 *>(v + i + sx) : zero4;
  q.ym = y > 0 ? *reinterpret_cast<const float4 *>(v + i - sy) : zero4;
  q.yp = y + 1 < a.ny ? *reinterpret_cast<const float4 *>(v + i + sy) : zero4;
  q.zl = z > 0 ? __ldg(v + i - 1) : 0.f;
  q.zr = z + 4 < a.nz ? __ldg(v + i + 4) : 0.f;
  if (MODE == LHS_RESID || MODE == LHS_ENERGY) q.b = *reinterpret_cast<const float4 *>(a.b + i);
}

// The operands of a quad other than the x-column of v (which a marching thread keeps in
// registers): y / z neighbours, accumulator, right-hand side.  `i` = linear index of the quad.
template <int MODE>
__device__ __forceinline__ void nd_quad_load_side(const LhsArgs &a, size_t i, int y, int z,
                                                  NdQuadIn &q) {
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  q.acc = a.acc ? *reinterpret_cast<const float4 *>(a.acc + i) : zero4;
  if (MODE == LHS_TERM) return;
  const float *__restrict__ v = a.v;
  const size_t sy = a.nz;
  q.ym = y > 0 ? *reinterpret_cast<const float4 *>(v + i - sy) : zero4;
  q.yp = y + 1 < a.ny ? *reinterpret_cast<const float4 *>(v + i + sy) : zero4;
  q.zl = z > 0 ? __ldg(v + i - 1) : 0.f;
  q.zr = z + 4 < a.nz ? __ldg(v + i + 4) : 0.f;
  if (MODE == LHS_RESID || MODE == LHS_ENERGY) q.b = *reinterpret_cast<const float4 *>(a.b + i);
}

// the 7-point stencil + CG epilogue of one quad (same arithmetic as lhs_direct_kernel)
template <int MODE>
__device__ __forceinline__ void nd_quad_finish(const LhsArgs &a, int x, int y, int z, size_t i,
                                               const NdQuadIn &q, float (&data)[4], double &part) {
  data[0] += q.acc.x, data[1] += q.acc.y, data[2] += q.acc.z, data[3] += q.acc.w;
  if (MODE == LHS_TERM) {
    *reinterpret_cast<float4 *>(a.out + i) = make_float4(data[0], data[1], data[2], data[3]);
    return;
  }
  const float cc[4] = {q.c.x, q.c.y, q.c.z, q.c.w};
  const float xm[4] = {q.xm.x, q.xm.y, q.xm.z, q.xm.w}, xp[4] = {q.xp.x, q.xp.y, q.xp.z, q.xp.w};
  const float ym[4] = {q.ym.x, q.ym.y, q.ym.z, q.ym.w}, yp[4] = {q.yp.x, q.yp.y, q.yp.z, q.yp.w};
  float val[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float c = cc[k];
    const float lft = k == 0 ? q.zl : cc[k > 0 ? k - 1 : 0];
    const float rgt = k == 3 ? q.zr : cc[k < 3 ? k + 1 : 3];
    const float t0 = ((x > 0 ? (c - xm[k]) * a.ivx : 0.f) - (xp[k] - c) * a.ivx) * a.ivx;
    const float t1 = ((y > 0 ? (c - ym[k]) * a.ivy : 0.f) - (yp[k] - c) * a.ivy) * a.ivy;
    const float t2 = ((z + k > 0 ? (c - lft) * a.ivz : 0.f) - (rgt - c) * a.ivz) * a.ivz;
    val[k] = (a.w_ident * c + data[k]) + a.rl2 * ((t0 + t1) + t2);
  }
  if (MODE == LHS_PLAIN) {
    *reinterpret_cast<float4 *>(a.out + i) = make_float4(val[0], val[1], val[2], val[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) part += (double)__fmul_rn(cc[k], val[k]);
  } else if (MODE == LHS_RESID) {
    const float bb[4] = {q.b.x, q.b.y, q.b.z, q.b.w};
    float rr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      rr[k] = __fsub_rn(bb[k], val[k]);
      part += (double)__fmul_rn(rr[k], rr[k]);
    }
    const float4 r4 = make_float4(rr[0], rr[1], rr[2], rr[3]);
    *reinterpret_cast<float4 *>(a.r + i) = r4;
    *reinterpret_cast<float4 *>(a.p + i) = r4;
  } else {  // LHS_ENERGY
    const float bb[4] = {q.b.x, q.b.y, q.b.z, q.b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) part += (double)__fmul_rn(__fsub_rn(val[k], 2.f * bb[k]), cc[k]);
    if (a.update_p) {
      const float beta = (float)a.fin.st->beta;
      const float4 p4 = *reinterpret_cast<const float4 *>(a.p + i);
      const float4 r4 = *reinterpret_cast<const float4 *>(a.r + i);
      float4 pn;
      pn.x = __fadd_rn(__fmul_rn(beta, p4.x), r4.x);
      pn.y = __fadd_rn(__fmul_rn(beta, p4.y), r4.y);
      pn.z = __fadd_rn(__fmul_rn(beta, p4.z), r4.z);
      pn.w = __fadd_rn(__fmul_rn(beta, p4.w), r4.w);
      *reinterpret_cast<float4 *>(a.p + i) = pn;
    }
  }
}


}  // namespace ur

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

}  // namespace ur

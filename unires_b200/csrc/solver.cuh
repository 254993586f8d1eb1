// Device-side CG bookkeeping shared by the solver kernels.
#pragma once
#include "common.cuh"
#include "rot.cuh"

namespace ur {

// Lives in device memory (head of the CG workspace).  Dynamic values only; the
// loop bounds / stop rule are kernel arguments.
struct CgState {
  double rz, rz0, pAp, alpha, beta;
  double obj_min, obj_max;
  int n_iter;  // completed iterations
  int done;    // set on the device when |gain| < tolerance
  int done_iter;  // iteration whose stop test set `done`
  int p_cur;      // which of the two direction buffers holds the current p (fused update)
  int x_cur;      // which of the two x buffers holds the current iterate (fused energy rule)
  double obj[UR_CG_MAX_ITER + 1];
};

// What the last block of a reducing kernel does with the grid total.
enum Finalize {
  FIN_NONE = 0,    // *dot_out = total (stand-alone lhs / dot)
  FIN_INIT_RZ,     // r = b - A x has been formed: rz = total, reset the state
  FIN_ALPHA,       // total = p.Ap          -> alpha = rz / total
  FIN_BETA,        // total = r.r (iter n)  -> beta = rz_new / rz_old, residual stop rule
  FIN_ENERGY       // total = (Ax - 2b).x   -> obj[n] = total / 2, energy stop rule
};

struct FinalizeArgs {
  int kind;
  int iter;       // CG iteration this launch belongs to (0 = initialisation)
  int stop_rule;  // UR_STOP_*
  double tol;
  CgState *st;
  double *dot_out;
  int aux;  // FIN_ALPHA: index of the direction buffer this matvec produced; FIN_ENERGY: index
            // of the x buffer it produced (-1: x was not moved)
};

__device__ __forceinline__ void record_objective(CgState *st, int n, double o, double tol) {
  st->obj[n] = o;
  if (n == 0) {
    st->obj_min = o;
    st->obj_max = o;
    return;
  }
  // nitorch get_gain(obj[:n+1], 'decreasing'): (obj[n-1] - obj[n]) / (max - min)
  const double mn = fmin(st->obj_min, o), mx = fmax(st->obj_max, o);
  st->obj_min = mn;
  st->obj_max = mx;
  const double gain = (st->obj[n - 1] - o) / (mx - mn);
  if (fabs(gain) < tol) {
    st->done = 1;
    st->done_iter = n;
  }
}

// ---------------------------------------------------------------------------
// lhs kernel arguments (shared by the direct and the streaming kernels)
// ---------------------------------------------------------------------------
constexpr int kMaxFused = 4;
constexpr int kMaxRot = 2;  // rotated observations gathered inside the direct lhs kernel

// One "lattice" observation (identity rotation, integer shift): along `axis`
//   x[j] = s_j * sum_t ker[t] * v[j*r + t + off],  0 <= j < nj   (v = 0 outside the grid)
// and a plain crop [lo, hi) on the other two axes.
struct LatticeTerm {
  float tau;
  int axis;  // correlation axis, -1 = pure crop
  int r, K, off, nj;
  int lo[3], hi[3];
  int scl_axis;  // -1 = no even/odd scaling
  int scl_off;
  float s_even, s_odd;
  float ker[UR_MAX_TAPS];
};

// LHS_COMBINE (streaming kernel only): the direction update is folded into the matvec's
// load stage:  p = beta p_old + r  (tile + halo, on-chip),  x += alpha_prev p_old,
// out = A p, p.Ap  -- 24 instead of 8 + 20 bytes per voxel for the two sweeps it replaces.
// LHS_ECOMBINE (lean kernel only): the x update is folded into the energy matvec's load stage:
//   x_new = x_old + alpha p (tile + halo, on-chip; written to a second buffer),
//   0.5 (A x_new - 2 b).x_new  -- with LHS_COMBINE (without its x update) and the residual
// update this is the energy-rule iteration in 44 instead of 52 bytes per voxel.
// LHS_TERM (lean kernel only): out = (acc +) tau A'S^2A v of ONE lattice term -- no D'D, no
// identity term, no dot product: the passes of a multi-view / multi-axis evaluation.
enum LhsMode {
  LHS_PLAIN = 0,
  LHS_RESID = 1,
  LHS_ENERGY = 2,
  LHS_COMBINE = 3,
  LHS_ECOMBINE = 4,
  LHS_TERM = 5
};

struct LhsArgs {
  int nx, ny, nz;
  int pitch;  // elements between z rows (0 / nz: dense).  > nz only with the lean TMA kernel:
              // volumes whose nz is not a multiple of 4 are solved in a zero-padded copy
  float ivx, ivy, ivz;
  float rl2;      // rho * lam^2
  float w_ident;  // sum of tau over identity observations (do_proj = 0)
  const float *acc;
  int nterm;
  LatticeTerm term[kMaxFused];
  int nrot;  // rotated observations: P' u gathered in-kernel, u from rot_forward_kernel (rot.cuh)
  RotTerm rot[kMaxRot];
  const float *v;
  float *out;      // PLAIN: A v
  const float *b;  // RESID / ENERGY
  float *r;        // RESID: r = b - A v ; ENERGY (p update): read
  float *p;        // RESID: p = r       ; ENERGY (p update): p = beta p + r
  int update_p;    // ENERGY only
  // COMBINE: v = p_old (read through TMA), rres = residual, p_out = new direction (a second
  // buffer: neighbouring CTAs still read p_old halos), xup = x (updated in place)
  const float *rres;
  float *p_out;
  float *xup;
  const int *done;
  GridReduce gr;
  FinalizeArgs fin;
};

// Executed by ONE thread of the last block.
__device__ __forceinline__ void finalize(const FinalizeArgs &f, double total) {
  CgState *st = f.st;
  switch (f.kind) {
    case FIN_NONE:
      if (f.dot_out) *f.dot_out = total;
      break;
    case FIN_INIT_RZ:
      st->rz = total;
      st->rz0 = total;
      st->n_iter = 0;
      st->done = 0;
      if (f.stop_rule == UR_STOP_RESIDUAL) record_objective(st, 0, sqrt(total), f.tol);
      break;
    case FIN_ALPHA:
      st->pAp = total;
      st->alpha = st->rz / total;
      st->p_cur = f.aux;
      break;
    case FIN_BETA:
      st->rz0 = st->rz;
      st->rz = total;
      st->beta = total / st->rz0;
      st->n_iter = f.iter;
      if (f.stop_rule == UR_STOP_RESIDUAL) record_objective(st, f.iter, sqrt(total), f.tol);
      break;
    case FIN_ENERGY:
      record_objective(st, f.iter, 0.5 * total, f.tol);
      if (f.aux >= 0) st->x_cur = f.aux;
      break;
  }
}

}  // namespace ur

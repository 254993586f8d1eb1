// Projection operator A / At / AtA  (_proj_apply, unires/_project.py:99-190):
//   super-resolution:  A = S . C . P      At = P' . C' . S      AtA = P' C' S^2 C P
//   denoising:         A = P              At = P'               AtA = P' P
// P = trilinear pull on the intermediate grid (affine coordinates evaluated
// in-kernel, no dense grid), C = separable strided slice-profile correlation,
// S = even/odd slice scaling.  This file is the general path (any rigid
// transform); the lattice-aligned fused kernels live in solver.cu.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "rot.cuh"
#include "lattice_nd.cuh"

namespace ur {

int affine_pull(const float *, Dim3i, const float[12], float *, Dim3i, int, int, cudaStream_t);
int affine_push(const float *, Dim3i, const float[12], float *, Dim3i, int, int, float,
                cudaStream_t);
int affine_push_gather(const float *, Dim3i, const float[12], float *, Dim3i, int, float,
                       cudaStream_t);
int lattice_pull(const float *, Dim3i, const float[12], float *, Dim3i, cudaStream_t);
int lattice_push(const float *, Dim3i, const float[12], float *, Dim3i, float, cudaStream_t);
int conv_axis(const float *, Dim3i, float *, int, const float *, int, int, bool, cudaStream_t,
              Dim3i *);
int apply_scaling(const float *, float *, Dim3i, float, int, cudaStream_t);
extern int g_rot_fused;  // rot.cu

static bool identity_pass(const ur_proj *po, int a) {
  return po->ksize[a] == 1 && po->ratio[a] == 1 && po->ker[a][0] == 1.0f;
}

int validate_proj(const ur_proj *po) {
  UR_REQUIRE(po != nullptr, "projection operator is NULL");
  UR_REQUIRE(po->method == UR_SUPERRES || po->method == UR_DENOISE, "Undefined method");
  for (int a = 0; a < 3; ++a) {
    UR_REQUIRE(po->dim_y[a] > 0 && po->dim_x[a] > 0, "projection operator: non-positive dims");
    if (po->method == UR_SUPERRES) {
      UR_REQUIRE(po->ksize[a] >= 1 && po->ksize[a] <= UR_MAX_TAPS,
                 "slice-profile factor length %d not in [1,%d]", po->ksize[a], UR_MAX_TAPS);
      UR_REQUIRE(po->ratio[a] >= 1, "ratio must be >= 1");
      UR_REQUIRE((po->dim_yx[a] - po->ksize[a]) / po->ratio[a] + 1 == po->dim_x[a] &&
                     po->dim_yx[a] >= po->ksize[a],
                 "dim_yx/ksize/ratio inconsistent with dim_x on axis %d", a);
    }
  }
  UR_REQUIRE(po->dim_thick >= 0 && po->dim_thick < 3, "dim_thick must be 0, 1 or 2");
  return UR_OK;
}

static size_t src_numel(const ur_proj *po) {
  const int32_t *d = po->method == UR_SUPERRES ? po->dim_yx : po->dim_x;
  return (size_t)d[0] * d[1] * d[2];
}

size_t proj_workspace_bytes(const ur_proj *po) {
  // two ping-pong volumes of the intermediate grid
  return 2 * ((src_numel(po) * sizeof(float) + 255) / 256 * 256);
}

// C: correlate+decimate along the three axes, most-decimating axis first.
// Returns the buffer that holds the result (bufs[0], bufs[1] or `last` if given).
static int conv_down(const ur_proj *po, const float *in, Dim3i d, float *bufs[2], float *last,
                     cudaStream_t st, const float **result, Dim3i *rdim) {
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (po->ratio[order[j]] > po->ratio[order[i]]) {
        int t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
  int npass = 0;
  for (int a = 0; a < 3; ++a) npass += !identity_pass(po, a);
  const float *cur = in;
  int done = 0, flip = (in == bufs[0]) ? 1 : 0;
  for (int i = 0; i < 3; ++i) {
    const int a = order[i];
    if (identity_pass(po, a)) continue;
    ++done;
    float *dst = (done == npass && last) ? last : bufs[flip];
    Dim3i od;
    int rc = conv_axis(cur, d, dst, a, po->ker[a], po->ksize[a], po->ratio[a], false, st, &od);
    if (rc) return rc;
    cur = dst;
    d = od;
    flip ^= 1;
  }
  *result = cur;
  *rdim = d;
  return UR_OK;
}

// C': transpose passes in exactly the reverse order of conv_down (the z pass, whose accesses
// are strided along the contiguous axis, then runs on the smallest volume in both directions).
static int conv_up(const ur_proj *po, const float *in, Dim3i d, float *bufs[2],
                   cudaStream_t st, const float **result, Dim3i *rdim) {
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (po->ratio[order[j]] > po->ratio[order[i]]) {
        int t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
  {
    const int t = order[0];
    order[0] = order[2];
    order[2] = t;
  }
  const float *cur = in;
  int flip = (in == bufs[0]) ? 1 : 0;
  for (int i = 0; i < 3; ++i) {
    const int a = order[i];
    if (identity_pass(po, a)) continue;
    float *dst = bufs[flip];
    Dim3i od;
    int rc = conv_axis(cur, d, dst, a, po->ker[a], po->ksize[a], po->ratio[a], true, st, &od);
    if (rc) return rc;
    cur = dst;
    d = od;
    flip ^= 1;
  }
  *result = cur;
  *rdim = d;
  return UR_OK;
}

// General-path operator.  For At / AtA the result is ACCUMULATED into d_out as
// out += scale * (...) (push is a scatter); the caller zeroes d_out first when a
// plain result is wanted.  For A, d_out is overwritten (scale ignored).
int proj_apply_general(int op, const ur_proj *po, const float *d_in, float *d_out, float scale,
                       void *d_ws, size_t ws_bytes, cudaStream_t st) {
  int rc = validate_proj(po);
  if (rc) return rc;
  UR_REQUIRE(op == UR_OP_A || op == UR_OP_AT || op == UR_OP_ATA, "Undefined operator");
  UR_REQUIRE(ws_bytes >= proj_workspace_bytes(po) && d_ws, "projection workspace too small");
  const size_t half = proj_workspace_bytes(po) / 2;
  float *bufs[2] = {(float *)d_ws, (float *)((char *)d_ws + half)};
  const Dim3i dy = make_dim(po->dim_y), dx = make_dim(po->dim_x);
  const bool sr = po->method == UR_SUPERRES;
  const Dim3i dsrc = sr ? make_dim(po->dim_yx) : dx;
  const float *cur;
  Dim3i d;
  // identity rotation + integer shift: crop / zero-pad embed instead of gather / atomic scatter
  const bool lat = ur_proj_is_lattice(po) != 0;
  // rotated operator with at most one decimated axis: in-tile forward kernel + gather adjoint
  // (rot.cuh) instead of pull / conv / scale / conv' / push through HBM
  if (!lat && g_rot_fused) {
    RotFwd F;
    RotTerm T;
    if (rot_describe(po, op, scale, &F, &T)) {
      if (op == UR_OP_A) return rot_forward_launch(UR_OP_A, F, d_in, d_out, st);
      if (op == UR_OP_AT)
        rc = rot_expand_launch(F, d_in, bufs[0], st);
      else
        rc = rot_forward_launch(UR_OP_ATA, F, d_in, bufs[0], st);
      if (rc) return rc;
      T.u = bufs[0];
      return rot_adjoint_launch(T, po->dim_y, d_out, 1, st);
    }
  }
  // lattice-aligned operator decimated along several axes: through the low-resolution image
  // (lattice_nd.cu) instead of crop / three conv passes / three conv' passes / embed
  if (lat && g_nd_fused) {
    NdOp nd;
    if (nd_describe(po, scale, &nd) && nd_conv_axes(nd) >= 2) {
      if (op == UR_OP_A) return nd_down_launch(nd, d_in, d_out, 1.f, nullptr, st);
      LhsArgs T;
      memset(&T, 0, sizeof(T));
      T.nx = po->dim_y[0];
      T.ny = po->dim_y[1];
      T.nz = po->dim_y[2];
      T.out = d_out;
      T.acc = d_out;
      const float *xl = d_in;
      if (op == UR_OP_ATA) {
        rc = nd_down_launch(nd, d_in, bufs[0], 1.f, nullptr, st);
        if (rc) return rc;
        xl = bufs[0];
      }
      rc = nd_up_launch(LHS_TERM, nd, xl, scale, T, st);
      if (rc != UR_ERR_UNSUPPORTED) return rc;
    }
  }
  auto pull = [&](const float *src, float *dst) {
    return lat ? lattice_pull(src, dy, po->mat, dst, dsrc, st)
               : affine_pull(src, dy, po->mat, dst, dsrc, 1, 0, st);
  };

  if (op == UR_OP_A) {
    bool any_pass = false;
    for (int a = 0; a < 3 && sr; ++a) any_pass |= !identity_pass(po, a);
    const bool need_scale = sr && po->scl != 0.f;
    if (!any_pass && !need_scale) return pull(d_in, d_out);
    rc = pull(d_in, bufs[0]);
    if (rc) return rc;
    cur = bufs[0];
    d = dsrc;
    if (any_pass) {
      rc = conv_down(po, cur, d, bufs, need_scale ? nullptr : d_out, st, &cur, &d);
      if (rc) return rc;
    }
    if (need_scale) return apply_scaling(cur, d_out, dx, po->scl, po->dim_thick, st);
    return UR_OK;
  }

  if (op == UR_OP_AT) {
    cur = d_in;
    d = dx;
    if (sr && po->scl != 0.f) {
      rc = apply_scaling(cur, bufs[0], dx, po->scl, po->dim_thick, st);
      if (rc) return rc;
      cur = bufs[0];
    }
  } else {  // AtA
    rc = pull(d_in, bufs[0]);
    if (rc) return rc;
    cur = bufs[0];
    d = dsrc;
    if (sr) {
      rc = conv_down(po, cur, d, bufs, nullptr, st, &cur, &d);
      if (rc) return rc;
      if (po->scl != 0.f) {
        float *dst = (cur == bufs[0]) ? bufs[1] : bufs[0];
        rc = apply_scaling(cur, dst, d, 2.f * po->scl, po->dim_thick, st);
        if (rc) return rc;
        cur = dst;
      }
    }
  }
  if (sr) {
    rc = conv_up(po, cur, d, bufs, st, &cur, &d);
    if (rc) return rc;
  }
  if (lat) return lattice_push(cur, dsrc, po->mat, d_out, dy, scale, st);
  // rotated operators: deterministic gather-form push; the atomic scatter only as a fallback
  rc = affine_push_gather(cur, dsrc, po->mat, d_out, dy, 0, scale, st);
  if (rc != UR_ERR_UNSUPPORTED) return rc;
  return affine_push(cur, dsrc, po->mat, d_out, dy, 1, 0, scale, st);
}

}  // namespace ur

using namespace ur;

extern "C" size_t ur_proj_workspace_bytes(const ur_proj *po) {
  return po ? proj_workspace_bytes(po) : 0;
}

extern "C" int ur_proj_is_lattice(const ur_proj *po) {
  if (!po) return 0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      const float want = r == c ? 1.f : 0.f;
      if (fabsf(po->mat[4 * r + c] - want) > 2e-7f) return 0;
    }
    const float t = po->mat[4 * r + 3];
    if (fabsf(t - rintf(t)) > 2e-5f) return 0;
  }
  return 1;
}

extern "C" int ur_proj_apply(int op, const ur_proj *po, const float *d_in, float *d_out,
                             void *d_ws, size_t ws_bytes, ur_stream stream) {
  UR_REQUIRE(op == UR_OP_A || op == UR_OP_AT || op == UR_OP_ATA, "Undefined operator");
  UR_REQUIRE(d_in && d_out, "ur_proj_apply: null data pointer");
  int rc = validate_proj(po);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (op != UR_OP_A) {
    const size_t n = (size_t)po->dim_y[0] * po->dim_y[1] * po->dim_y[2];
    UR_CUDA_CHECK(cudaMemsetAsync(d_out, 0, n * sizeof(float), st));
  }
  return proj_apply_general(op, po, d_in, d_out, 1.f, d_ws, ws_bytes, st);
}

extern "C" int ur_proj_accumulate(int op, const ur_proj *po, const float *d_in, float *d_out,
                                  float scale, void *d_ws, size_t ws_bytes, ur_stream stream) {
  UR_REQUIRE(op == UR_OP_AT || op == UR_OP_ATA, "ur_proj_accumulate: operator must be At or AtA");
  UR_REQUIRE(d_in && d_out, "ur_proj_accumulate: null data pointer");
  return proj_apply_general(op, po, d_in, d_out, scale, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int ur_nll_data_proj(const ur_proj *po, const float *d_y, const float *d_x, float tau,
                                double *d_out, int accumulate, void *d_ws, size_t ws_bytes,
                                ur_stream stream) {
  UR_REQUIRE(d_y && d_x && d_out, "ur_nll_data_proj: null pointer");
  int rc = validate_proj(po);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (ur_proj_is_lattice(po) && g_nd_fused) {
    rc = nll_nd_launch(po, d_y, d_x, tau, d_out, accumulate, st);
    if (rc != UR_ERR_UNSUPPORTED) return rc;
  }
  // general route: A y into the workspace (after the operator's own scratch), then the reduction
  const size_t nx = (size_t)po->dim_x[0] * po->dim_x[1] * po->dim_x[2];
  const size_t pw = proj_workspace_bytes(po);
  UR_REQUIRE(d_ws && ws_bytes >= pw + nx * sizeof(float), "ur_nll_data_proj: workspace too small");
  float *Ay = (float *)((char *)d_ws + pw);
  rc = proj_apply_general(UR_OP_A, po, d_y, Ay, 1.f, d_ws, pw, st);
  if (rc) return rc;
  return ur_nll_data(d_x, Ay, nx, tau, d_out, accumulate, stream);
}

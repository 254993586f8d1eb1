/* unires_b200.h -- C ABI of the B200-native UniRes ADMM/CG hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer named
 * d_* / documented "device" is a CUDA device pointer; volumes are float32,
 * C-contiguous (X, Y, Z) with Z fastest, exactly the layout UniRes keeps
 * (unires/_update.py:29 allocates (C,3,X,Y,Z); unires/_project.py:78 adds
 * the two leading singleton dims).  All calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*) unless stated otherwise.  Return value:
 * UR_OK or an error code; ur_last_error() gives the message.
 *
 * "Reference" citations are relative to the upstream repo brudfors/UniRes
 * (file:line); "nitorch" is its pinned third-party dependency
 * (setup.py:11), whose functions these entry points replace at the call
 * sites listed.
 */
#ifndef UNIRES_B200_H
#define UNIRES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UR_MAX_TAPS 32     /* max length of one separable slice-profile factor */
#define UR_MAX_OBS 8       /* max observations (repeats) of one channel        */
#define UR_MAX_CHANNELS 16 /* max channels in one JTV-prox launch              */
#define UR_CG_MAX_ITER 256 /* capacity of the on-device CG objective trace     */

typedef void *ur_stream;

enum {
  UR_OK = 0,
  UR_ERR_ARG = 1,         /* bad argument (the reference raises ValueError)   */
  UR_ERR_CUDA = 2,        /* CUDA runtime failure                             */
  UR_ERR_UNSUPPORTED = 3  /* valid in nitorch, not implemented here           */
};

enum { UR_SUPERRES = 0, UR_DENOISE = 1 };        /* unires/_project.py:125 */
enum { UR_OP_A = 0, UR_OP_AT = 1, UR_OP_ATA = 2 }; /* unires/_project.py:123 */
/* nitorch cg `stop` (Appendix A.7 / Q1 of SURVEY.md): NONE = tolerance 0,
 * RESIDUAL = sqrt(r.z) ('e'), ENERGY = 0.5 x'Ax - b'x (what 'max_gain' selects) */
enum { UR_STOP_NONE = 0, UR_STOP_RESIDUAL = 1, UR_STOP_ENERGY = 2 };

const char *ur_last_error(void);
int ur_version(void);
/* device in use: SM count and compute capability */
int ur_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* number of CUDA kernels this library has launched in this process */
uint64_t ur_launch_count(void);
/* Instrumentation for bench.py's roofline: while enabled, every CG matvec
 * launch (the lhs kernel producing A p) is bracketed by CUDA events on its own
 * stream.  ur_profile_matvec_read synchronises, returns the summed duration,
 * the number of launches since the last read and the sum over those launches
 * of their algorithmic HBM bytes per voxel (8 for A p; 24 when the direction
 * and x updates are fused into the matvec), and resets.                   */
int ur_profile_matvec(int enable);
int ur_profile_matvec_read(double *total_ms, int32_t *count,
                           double *bytes_per_voxel_sum);
/* Tuning / test knobs (process-wide).  "lhs_variant": 0 automatic (lean
 * specialised TMA kernel -> generic TMA streaming kernel -> direct kernel),
 * 1 force the direct (non-TMA) lhs kernel, 2 skip the lean kernel;
 * "stream_mc" / "stream_rpt" / "stream_pf": planes per CTA chunk, rows per
 * thread (0 automatic, 1 = 8-row tiles, 2 = 16-row tiles) and extra prefetch
 * slots of the generic streaming kernel; "fast_q" / "fast_rpt" / "fast_depth":
 * work units per CTA, rows per thread and ring prefetch depth (plane pairs)
 * of the lean kernel; "cg_fuse": fold the direction update into the matvec;
 * "cg_graph": 0 disables the CUDA-graph replay of repeated solves; "nd_fused": 0
 * sends multi-axis lattice operators through chained single-axis passes;
 * "rot_fused": 0 routes rotated operators through the unfused pull / conv /
 * conv' / push chain instead of the in-tile kernels (A/B and tests);
 * "rot_cell": adjoint pull of rotated operators through per-cell corner
 * coefficients (1, default), with eight colour passes forced (8, test hook),
 * or by the per-voxel candidate gather (0); "fast_to" / "fast_segs": output
 * rows per tile (0 automatic, <= 8 x rows per thread) and lock-step segments
 * per column (0 automatic) of the lean kernel; "l2_hints": L2 eviction
 * priorities of the fused CG iteration (bit 0: stores and x loads are
 * evict_first; bit 1: the residual is evict_last; bit 2: the TMA tiles of v
 * are evict_first; default 0 -- measured a loss with concurrent channels);
 * "vol_skew": bytes added to every workspace volume (placement experiment,
 * default 0).
 * Unknown names return UR_ERR_ARG.                                          */
int ur_tune(const char *name, int value);
/* Which kernel served the most recent lhs launch of this process:
 * 0 direct, 1 generic TMA streaming kernel, 2 lean specialised TMA kernel,
 * 3 rotated-operator kernels (forward tile kernel + quad gather adjoint),
 * 4 multi-axis lattice kernels (nd_down + nd_up through the low-res image),
 * 5 rotated-operator kernels (forward tile kernel + cell-coefficient adjoint
 *   into the accumulator) followed by the lean TMA kernel.                  */
int ur_last_lhs_path(void);

/* ---------------------------------------------------------------- finite
 * differences: nitorch.spatial.im_gradient / im_divergence, 'forward',
 * bound 'zero' (unires/_project.py:314-315; unires/_update.py:132,168,176,
 * 188,419).  grad/vec are (3, X, Y, Z).                                    */
int ur_im_gradient(const float *d_dat, float *d_grad, const int32_t dim[3],
                   const float vx[3], ur_stream stream);
int ur_im_divergence(const float *d_vec, float *d_div, const int32_t dim[3],
                     const float vx[3], ur_stream stream);
/* _DtD (unires/_project.py:300-317): out = div(grad(dat)) in one pass */
int ur_dtd(const float *d_dat, float *d_out, const int32_t dim[3],
           const float vx[3], ur_stream stream);

/* ---------------------------------------------------------------- resampling:
 * nitorch.spatial.grid_pull / grid_push, bound 'zero' (unires/_project.py:
 * 164,172,174,179,183,185,187,188).  order 0|1; extrapolate 0 = the
 * reference's setting.  `grid` is the dense (ox,oy,oz,3) voxel-coordinate
 * field of nitorch.affine_grid; the *_affine variants evaluate the same
 * coordinates in-kernel from the 3x4 matrix (row major) and never touch a
 * dense grid (replaces unires/_project.py:159).  push ACCUMULATES
 * scale*value into d_out (zero it first for a plain push).                 */
int ur_grid_pull(const float *d_src, const int32_t sdim[3], const float *d_grid,
                 float *d_out, const int32_t odim[3], int order, int extrapolate,
                 ur_stream stream);
int ur_grid_push(const float *d_in, const int32_t idim[3], const float *d_grid,
                 float *d_out, const int32_t sdim[3], int order, int extrapolate,
                 float scale, ur_stream stream);
int ur_affine_pull(const float *d_src, const int32_t sdim[3], const float mat[12],
                   float *d_out, const int32_t odim[3], int order, int extrapolate,
                   ur_stream stream);
int ur_affine_push(const float *d_in, const int32_t idim[3], const float mat[12],
                   float *d_out, const int32_t sdim[3], int order, int extrapolate,
                   float scale, ur_stream stream);
/* nitorch.spatial.affine_grid materialised (only for callers that index the
 * grid, unires/run.py:169-174): d_grid is (ox,oy,oz,3).                    */
int ur_affine_grid(const float mat[12], float *d_grid, const int32_t odim[3],
                   ur_stream stream);

/* ---------------------------------------------------------------- slice profile:
 * F.conv3d / F.conv_transpose3d with a 1->1 channel separable kernel and
 * stride = ratio (unires/_project.py:153-154); valid padding.  One axis per
 * call: odim[axis] = (idim[axis]-K)/stride+1 (conv) or (idim[axis]-1)*stride+K
 * (transpose); the other two extents are unchanged.                        */
int ur_conv_axis(const float *d_in, const int32_t idim[3], float *d_out, int axis,
                 const float *ker, int K, int stride, int transpose,
                 ur_stream stream);
/* _apply_scaling (unires/_project.py:9-24): out = exp(+-scl) * in, sign
 * alternating along `axis` (even slices +).                                */
int ur_apply_scaling(const float *d_in, float *d_out, const int32_t dim[3],
                     float scl, int axis, ur_stream stream);

/* ---------------------------------------------------------------- projection
 * operator.  Plain-data image of unires.struct._proj_op (unires/struct.py:
 * 36-54) after _proj_info (unires/_project.py:193-297), with smo_ker split
 * into its three 1-D factors (it is an outer product, _project.py:277) and
 * `mat` = mat_y^-1 . rigid . mat_yx (super-resolution) or mat_y^-1 . rigid .
 * mat_x (denoising) solved ONCE in float64 and cast to float32 exactly as
 * unires/_project.py:147,150,159 does on every call.                       */
typedef struct ur_proj {
  int32_t method; /* UR_SUPERRES | UR_DENOISE */
  int32_t dim_y[3];
  int32_t dim_x[3];
  int32_t dim_yx[3];
  int32_t ratio[3];
  int32_t ksize[3];
  float ker[3][UR_MAX_TAPS];
  float mat[12];
  float scl;         /* even/odd log-scaling, 0 = off (po.scl)   */
  int32_t dim_thick; /* axis the scaling alternates along        */
} ur_proj;

/* 1 if the operator is "lattice aligned" (mat = identity + integer shift):
 * pull/push degenerate to crop/zero-pad and the fused kernels apply.       */
int ur_proj_is_lattice(const ur_proj *po);
size_t ur_proj_workspace_bytes(const ur_proj *po);
/* Rotated operators, adjoint pull (nitorch grid_push at unires/_project.py:172,
 * 179) through per-cell corner coefficients: number of colour passes under
 * which two intermediate voxels of one colour never fall into the same unit
 * cell of the recon grid for the 3x4 row-major map `mat` (intermediate index
 * -> recon voxel): 2 = parity of i+j+k, 8 = parities of i, j, k, 0 = neither
 * (the per-voxel gather is used).  Host-only; no GPU needed.                */
int ur_rot_cell_colours(const float mat[12]);
/* _proj_apply (unires/_project.py:99-190): op in {A, At, AtA}.
 * A: in = y-space, out = x-space; At: reverse; AtA: y -> y.  out = result
 * (overwritten).                                                           */
int ur_proj_apply(int op, const ur_proj *po, const float *d_in, float *d_out,
                  void *d_ws, size_t ws_bytes, ur_stream stream);
/* out += scale * op(in) for op in {At, AtA}: the accumulation of
 * unires/_update.py:125-128 (tmp += tau * At x) without a temporary.      */
int ur_proj_accumulate(int op, const ur_proj *po, const float *d_in, float *d_out,
                       float scale, void *d_ws, size_t ws_bytes, ur_stream stream);

/* ---------------------------------------------------------------- CG left-hand
 * side of one channel: v -> sum_n tau_n An'An v + rho lam^2 D'D v
 * (unires/_project.py:73-87 with operator 'AtA'; do_proj = 0 is the
 * operator 'none' branch, :76-77, A = identity).                          */
typedef struct ur_lhs {
  int32_t dim_y[3];
  float vx[3];
  float rho_lam2; /* rho * lam^2, formed in float32 by the host as the reference does */
  int32_t do_proj;
  int32_t n_obs;
  float tau[UR_MAX_OBS];
  ur_proj obs[UR_MAX_OBS];
} ur_lhs;

size_t ur_lhs_workspace_bytes(const ur_lhs *lhs);
/* out = lhs(v).  If d_dot != NULL also writes sum(v*out) (float64). */
int ur_lhs_apply(const ur_lhs *lhs, const float *d_v, float *d_out, double *d_dot,
                 void *d_ws, size_t ws_bytes, ur_stream stream);

/* ---------------------------------------------------------------- CG solve:
 * nitorch.core.optim.cg as called at unires/_update.py:142-148 (in place on
 * x, identity preconditioner, float64 dot products), run entirely on the
 * device: no host synchronisation between iterations; the stop test
 * |gain| < tolerance is evaluated on the device and later launches of the
 * same solve early-out.  A solve repeated with the same operator, options,
 * buffers, stream and knobs (what an ADMM run does every outer iteration) is
 * stream-captured on its second call and replayed as one CUDA graph afterwards
 * (ur_tune("cg_graph", 0) disables this); results are bitwise identical.     */
typedef struct ur_cg_opts {
  int32_t max_iter;  /* sett.cgs_max_iter (unires/struct.py:65), <= UR_CG_MAX_ITER */
  int32_t stop_rule; /* UR_STOP_* */
  double tolerance;  /* sett.cgs_tol (unires/struct.py:66) */
  int32_t variant;   /* kernel selection: 0 = auto (for A/B measurements) */
} ur_cg_opts;

size_t ur_cg_workspace_bytes(const ur_lhs *lhs);
int ur_cg_solve(const ur_lhs *lhs, const float *d_b, float *d_x, void *d_ws,
                size_t ws_bytes, const ur_cg_opts *opts, ur_stream stream);
/* Synchronises `stream`, then returns the iterate count and the objective
 * trace obj[0..n_iter] (as many as fit in n_obj).                          */
int ur_cg_fetch(const void *d_ws, int32_t *n_iter, double *obj, int32_t n_obj,
                ur_stream stream);

/* building blocks for cg() with an arbitrary host callable as A */
int ur_dot(const float *d_a, const float *d_b, size_t n, double *d_out,
           ur_stream stream);
/* x += alpha p ; r -= alpha Ap ; *d_rr = sum(r*r).  alpha read from device */
int ur_cg_update_xr(float *d_x, float *d_r, const float *d_p, const float *d_Ap,
                    size_t n, const double *d_alpha, double *d_rr,
                    ur_stream stream);
/* p = beta p + r, beta read from device */
int ur_cg_update_p(float *d_p, const float *d_r, size_t n, const double *d_beta,
                   ur_stream stream);

/* ---------------------------------------------------------------- ADMM pieces
 * (unires/_update.py:105-195).                                             */
/* RHS of the y-update (:124-133): b = acc - lam * div(w_c - rho z_c), where
 * acc = sum_n tau_n An' x_n has already been accumulated into d_b by the
 * caller (ur_proj_apply + ur_axpy) or is taken as tau*x when do_proj = 0.  */
int ur_admm_rhs(float *d_b, const float *d_w, const float *d_z,
                const int32_t dim[3], const float vx[3], float lam, float rho,
                ur_stream stream);
/* The whole right-hand side in ONE pass when every observation of the channel
 * is lattice aligned (else UR_ERR_UNSUPPORTED, nothing launched):
 *   b = sum_n tau_n An' x_n - lam * div(w_c - rho z_c)
 * d_x: host array of lhs->n_obs device pointers (the observations x_n).     */
int ur_admm_rhs_fused(const ur_lhs *lhs, const float *const *d_x, float *d_b,
                      const float *d_w, const float *d_z, float lam, float rho,
                      ur_stream stream);
/* Back-projection of the observations of one channel in ONE pass, lattice
 * aligned observations only (else UR_ERR_UNSUPPORTED, nothing launched):
 *   out = scale .* sum_n An' x_n      (no tau_n; d_scale may be NULL)
 * With scale = 1 / (A' 1) this is the normalised back-projection used as the
 * device-side initial estimate of a freshly uploaded subject (the role of
 * _init_y_dat, unires/_core.py:371-399, in the host pipeline: the estimate
 * is formed from the uploaded observations instead of being uploaded).      */
int ur_backproject(const ur_lhs *lhs, const float *const *d_x, float *d_out,
                   const float *d_scale, ur_stream stream);
/* y += a * x (float32) */
int ur_axpy(float *d_y, const float *d_x, float a, size_t n, ur_stream stream);

/* z- and w-update in ONE pass over all channels (:160-193):
 *   g_c = lam_c grad(y_c)  [alpha != 1: g_c = alpha g_c + (1-alpha) z_c_old]
 *   u_c = w_c/rho + g_c; s = sqrt(sum_c |u_c|^2);
 *   f = max(s - 1/rho, 0)/(s + 1e-7); z_c = f u_c; w_c += rho (g_c - z_c)
 * d_y: array of C device pointers (host array), d_z,d_w: (C,3,X,Y,Z),
 * d_jtv: (X,Y,Z) receives f.                                               */
int ur_jtv_prox(const float *const *d_y, float *d_z, float *d_w, float *d_jtv,
                int n_channels, const float *lam, const int32_t dim[3],
                const float vx[3], float rho, float alpha, ur_stream stream);
/* channel-sharded variant (multi-GPU): (1) accumulate this rank's
 * sum_c |u_c|^2 into d_nrm2; (all-reduce d_nrm2 across ranks); (2) apply. */
int ur_jtv_norm2(const float *const *d_y, const float *d_z, const float *d_w,
                 float *d_nrm2, int n_channels, const float *lam,
                 const int32_t dim[3], const float vx[3], float rho, float alpha,
                 int accumulate, ur_stream stream);
int ur_jtv_apply(const float *const *d_y, float *d_z, float *d_w,
                 const float *d_nrm2, float *d_jtv, int n_channels,
                 const float *lam, const int32_t dim[3], const float vx[3],
                 float rho, float alpha, ur_stream stream);

/* objective (_compute_nll, unires/_update.py:396-427), float64 sums.
 * data term of one observation: *d_out (+)= 0.5 tau sum_{x!=0} (x - Ay)^2,
 * d_Ay already projected.  prior: accumulate sum_c sum_d (lam_c grad y_c)_d^2
 * into d_e (X,Y,Z), then ur_sqrt_sum.                                      */
int ur_nll_data(const float *d_x, const float *d_Ay, size_t n, float tau,
                double *d_out, int accumulate, ur_stream stream);
/* The data term of one observation WITHOUT materialising A y when the operator is lattice
 * aligned with at most one decimated axis (one pass over y; the reference's pull / conv3d /
 * scaling / boolean-mask compaction of unires/_update.py:411-417 folded into the reduction):
 * *d_out (+)= 0.5 tau sum_{x != 0} (x - A y)^2.  Any other operator: A y is formed in the
 * workspace first (ws_bytes >= ur_proj_workspace_bytes(po) + 4 numel(dim_x)).              */
int ur_nll_data_proj(const ur_proj *po, const float *d_y, const float *d_x, float tau,
                     double *d_out, int accumulate, void *d_ws, size_t ws_bytes,
                     ur_stream stream);
int ur_nll_prior_energy(const float *const *d_y, float *d_e, int n_channels,
                        const float *lam, const int32_t dim[3], const float vx[3],
                        int accumulate, ur_stream stream);
int ur_sqrt_sum(const float *d_e, size_t n, double *d_out, ur_stream stream);

/* ---------------------------------------------------------------- even/odd slice
 * scaling update, _update_scaling (unires/_update.py:270-393).
 * ur_scaling_sums: the five masked float64 sums of one Gauss-Newton step over
 * the voxels with x != 0 (d_x observed, d_y = A y with the current scaling;
 * p = parity of the slice index along `axis`):
 *   d_out[0] = sum (x-y)^2   d_out[1+p] = sum y (x-y)   d_out[3+p] = sum y^2
 * (unires/_update.py:321 log-likelihood, :334-339 gradient and Hessian; the
 * reference's "odd" slices are ::2, i.e. p = 0).  ur_scale_slices: out = f_even
 * * in on slices ::2 and f_odd * in on slices 1::2 (_apply_scaling with the
 * factors exp(+-s) formed by the caller, unires/_update.py:361); in place ok. */
int ur_scaling_sums(const float *d_x, const float *d_y, const int32_t dim[3], int axis,
                    double *d_out, ur_stream stream);
int ur_scale_slices(const float *d_in, float *d_out, const int32_t dim[3], float f_even,
                    float f_odd, int axis, ur_stream stream);

/* ---------------------------------------------------------------- rigid
 * Gauss-Newton update, _update_rigid / _rigid_match / _update_rigid_channel
 * (unires/_update.py:198-267, 448-710).
 * ur_affine_grad: nitorch.spatial.grid_grad (unires/_update.py:505) with the
 * coordinates evaluated from the 3x4 matrix: d_out (ox,oy,oz,3) = gradient of
 * the trilinearly interpolated d_src w.r.t. the sampling coordinates, zero
 * bound.  ur_rigid_sums: the chain-rule reductions of unires/_update.py:
 * 616-640 in one pass: with g = d_grad, res = d_res (C'(Ay - x), masked),
 * ctc = d_ctc (C'C 1, may be NULL) and dA[i][d] the float32 field
 * ((dm[i][d][0] ix + dm[i][d][1] iy) + dm[i][d][2] iz) + dm[i][d][3],
 *   d_out[i]            = sum_d sum_vox (g_d res) dA[i][d]            i < 6
 *   d_out[6 + tri(i,j)] = sum_{d1,d2} sum_vox ((g_d1 g_d2 ctc) dA[i][d1]) dA[j][d2]
 * for i <= j in row-major upper-triangular order (21 values), float64.      */
int ur_affine_grad(const float *d_src, const int32_t sdim[3], const float mat[12],
                   float *d_out, const int32_t odim[3], int extrapolate,
                   ur_stream stream);
int ur_rigid_sums(const float *d_grad, const float *d_res, const float *d_ctc,
                  const int32_t dim[3], const float dm[72], double *d_out,
                  ur_stream stream);

/* ---------------------------------------------------------------- intensity
 * statistics for the hyper-parameter estimate, _estimate_hyperpar
 * (unires/_core.py:96-142).  The reference compacts the volume by
 * boolean masks (dat[dat >= 0] at _core.py:118, then nitorch's estimate_noise
 * drops zeros and the maximum) and bins it with torch.histc; here the
 * selections are folded into two streaming passes.  Selection of a voxel v:
 * non-finite -> 0; drop_negative: only v >= 0; mask: v != 0 and v != mask_value.
 * ur_intensity_range: h_out[0] = min, h_out[1] = max of the selected voxels
 * (HOST floats; synchronises the stream), *h_any = 0 if nothing was selected.
 * ur_histc: d_counts[b] (device, `bins` uint64, zeroed here) = number of
 * selected voxels with mn <= v <= mx in bin (int)((v - mn) * bins / (mx - mn))
 * (v == mx in the last bin), computed in float64 like torch.histc on the
 * reference's float64 copy.                                                   */
int ur_intensity_range(const float *d_dat, size_t n, int drop_negative, int mask,
                       float mask_value, float h_out[2], int32_t *h_any,
                       ur_stream stream);
int ur_histc(const float *d_dat, size_t n, int drop_negative, int mask,
             float mask_value, double mn, double mx, int bins,
             unsigned long long *d_counts, ur_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIRES_B200_H */

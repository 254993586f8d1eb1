// Host side of the lean streaming lhs kernel (lhs_fast.cuh): eligibility, specialisation
// lookup, ring sizing, balanced one-wave work split in row-aligned units, launch.
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "lhs_fast.cuh"

namespace ur {

// lhs_stream.cu: cached cuTensorMapEncodeTiled of an (X, Y, Z) float volume with a box of one
// plane tile (sz floats along z, `rows` rows, 1 plane along the march axis)
bool stream_tensor_map(const float *v, int nx, int ny, int nz, int sz, int march_y, int rows,
                       CUtensorMap *out, int pitch);

int fast_rpt = 0;     // rows per thread: 0 automatic, 1 | 2
int fast_depth = 1;   // prefetch depth of the ring in plane pairs (>= 1)
int fast_q_units = 0; // units per CTA override (0 automatic)
int fast_pfd = 0;     // planes prefetched into L2 ahead of the ring
int fast_lock = 1;    // cut every column into the same segments (neighbours march in step)
int fast_diag_residue = 1;  // carry the rounding residue of the stencil diagonal (debug knob)
int fast_to = 0;      // output rows per tile (0 automatic; <= 8 rpt)
int fast_segs = 0;    // segments per column override (0 automatic)
int g_l2_hints = 0;   // ur_tune("l2_hints"): L2 eviction priorities of the fused CG iteration -- bit 0:
                      // stores and x loads evict_first, bit 1: residual evict_last, bit 2: TMA tiles of v
                      // evict_first.  -2.6 % single stream at 256^3, but a LOSS with three channel
                      // streams and at 384^3 (profiles/r02_l2_eviction_hints.txt): off by default;
                      // solver.cu applies the same to the residual update

static bool a16(const void *p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; }

// dry_run: eligibility check only, nothing is launched
int lhs_fast_launch(int mode, const LhsArgs &A, bool dry_run, cudaStream_t st) {
  using namespace fast;
  if (A.nterm > 1) return UR_ERR_UNSUPPORTED;  // solver.cu splits several terms into passes
  if (A.nrot > 0) return UR_ERR_UNSUPPORTED;   // rotated terms are gathered in the direct kernel
  if (!a16(A.acc)) return UR_ERR_UNSUPPORTED;
  const int pitch = A.pitch > 0 ? A.pitch : A.nz;
  if (pitch % 4 != 0 || pitch < A.nz || pitch - A.nz >= 4 || A.nz < 4) return UR_ERR_UNSUPPORTED;
  if ((long long)A.nx * A.ny * pitch + 64ll * A.ny * A.nz + 64ll * A.nx * A.nz > 0x7fffffffll)
    return UR_ERR_UNSUPPORTED;
  if (!a16(A.v) || !a16(A.out) || !a16(A.b) || !a16(A.r) || !a16(A.p)) return UR_ERR_UNSUPPORTED;
  if (!a16(A.rres) || !a16(A.p_out) || !a16(A.xup)) return UR_ERR_UNSUPPORTED;
  const bool combine = mode == LHS_COMBINE || mode == LHS_ECOMBINE;
  const bool x_fused = mode == LHS_COMBINE && A.xup != nullptr;

  FastArgs S;
  memset(&S, 0, sizeof(S));
  int kind = FK_NONE, kp = 1, r = 1, e = 0, march = 0;
  S.lo_m = S.lo_o = S.lo_z = 0;
  S.s_even = S.s_odd = 1.f;
  if (A.nterm == 1) {
    const LatticeTerm &T = A.term[0];
    if (T.axis == 1) march = 1;
    const int ax_m = march, ax_o = 1 - march;
    if (T.scl_axis >= 0 && T.scl_axis != T.axis) return UR_ERR_UNSUPPORTED;  // thin-axis scaling
    S.tau = T.tau;
    S.off = T.off;
    S.nj = T.nj;
    S.lo_m = T.lo[ax_m];
    S.hi_m = T.hi[ax_m];
    S.lo_o = T.lo[ax_o];
    S.hi_o = T.hi[ax_o];
    S.lo_z = T.lo[2];
    S.hi_z = T.hi[2];
    S.scl_conv = T.scl_axis >= 0;
    S.s_even = T.s_even;
    S.s_odd = T.s_odd;
    if (T.axis < 0) {
      kind = FK_POINT;
    } else {
      kp = T.K;
      r = T.r;
      if (kp > kTaps) return UR_ERR_UNSUPPORTED;
      for (int t = 0; t < kp; ++t) {
        S.ker[t] = T.ker[t];
        S.kerT[t] = T.tau * T.ker[t];
      }
      e = ((T.off % r) + r) % r;
      if (T.axis == 2) {
        // register scheme when the ratio divides the quad, per-lane windows otherwise
        kind = ((kp == 5 && r == 4) || (kp == 3 && r == 2) || (kp == 9 && r == 2)) ? FK_THICK_Z
                                                                                    : FK_THICK_ZG;
        S.lo_z = 0;
        S.hi_z = A.nz;
      } else {
        kind = FK_THICK_M;
        S.lo_m = 0;
        S.hi_m = march ? A.ny : A.nx;
      }
    }
  }
  S.march_y = march;
  S.nm = march ? A.ny : A.nx;
  S.no = march ? A.nx : A.ny;
  S.nz = A.nz;
  S.gs_m = march ? pitch : A.ny * pitch;
  S.gs_o = march ? A.ny * pitch : pitch;
  const float iv_m = march ? A.ivy * A.ivy : A.ivx * A.ivx;
  const float iv_o = march ? A.ivx * A.ivx : A.ivy * A.ivy;
  const float iv_z = A.ivz * A.ivz;
  S.a_m = A.rl2 * iv_m;
  S.a_o = A.rl2 * iv_o;
  S.a_z = A.rl2 * iv_z;
  // diag - sum(neighbour weights) must cancel exactly: carry the float rounding residue of the
  // diagonal separately (6e-8 relative to d0, but coherent over the volume: it would act as a
  // spurious identity term on the large near-constant part of an image)
  const double d0_exact = (double)A.w_ident + 2.0 * ((double)S.a_m + (double)S.a_o + (double)S.a_z);
  S.d0 = (float)d0_exact;
  S.nd0l = fast_diag_residue ? (float)((double)S.d0 - d0_exact) : 0.f;

  int rpt = kind == FK_THICK_M ? 1 : 2;  // measured at 256^3: 8-row tiles win with a deep ring
  if (fast_rpt == 1 || fast_rpt == 2) rpt = fast_rpt;
  if (S.no <= NWARP) rpt = 1;

  FastKernel kernel = nullptr;
  const int ez = kind == FK_THICK_Z ? e : 0;
  switch (mode) {
    case LHS_PLAIN:
      kernel = fast_lookup_plain(kind, kp, r, ez, rpt);
      break;
    case LHS_RESID:
      kernel = fast_lookup_resid(kind, kp, r, ez, rpt);
      break;
    case LHS_ENERGY:
      kernel = fast_lookup_energy(kind, kp, r, ez, rpt);
      break;
    case LHS_ECOMBINE:
      kernel = fast_lookup_ecombine(kind, kp, r, ez, rpt);
      break;
    case LHS_TERM:
      kernel = fast_lookup_term(kind, kp, r, ez, rpt);
      break;
    default:
      kernel = fast_lookup_combine(kind, kp, r, ez, rpt);
      break;
  }
  if (!kernel && rpt == 2) {  // some specialisations exist for 8-row tiles only
    rpt = 1;
    switch (mode) {
      case LHS_PLAIN: kernel = fast_lookup_plain(kind, kp, r, ez, rpt); break;
      case LHS_RESID: kernel = fast_lookup_resid(kind, kp, r, ez, rpt); break;
      case LHS_ENERGY: kernel = fast_lookup_energy(kind, kp, r, ez, rpt); break;
      case LHS_ECOMBINE: kernel = fast_lookup_ecombine(kind, kp, r, ez, rpt); break;
      case LHS_TERM: kernel = fast_lookup_term(kind, kp, r, ez, rpt); break;
      default: kernel = fast_lookup_combine(kind, kp, r, ez, rpt); break;
    }
  }
  if (!kernel) return UR_ERR_UNSUPPORTED;
  const int hz = fast_hz(kind, kp), sz = TZ + 2 * hz;

  const int to_max = NWARP * rpt;  // rows of the tile in shared memory (compile time)
  int to = to_max;                  // rows that are output
  if (fast_to > 0 && fast_to <= to_max) to = fast_to;
  const int L = kind == FK_THICK_M ? kp - 1 : 1;
  const int B = kind == FK_THICK_M ? kp - 1 : 0;
  const int depth = fast_depth < 1 ? 1 : fast_depth;
  S.ns = L + 3 + 2 * depth;
  S.nrs = combine ? S.ns - L - 1 : 0;
  if (S.ns > kMaxSlots) return UR_ERR_UNSUPPORTED;
  const size_t plane_b = ((size_t)(to_max + 2) * sz * 4 + 127) / 128 * 128;
  const size_t smem = (size_t)(S.ns + S.nrs) * plane_b + 128;
  if (smem > 200u * 1024u) return UR_ERR_UNSUPPORTED;

  int resident = 0;
  if (!dry_run) {
    static std::unordered_map<size_t, int> occ_cache;
    static std::mutex occ_mu;
    // the opt-in to > 48 KB of dynamic shared memory and the occupancy are per DEVICE
    int dev = 0;
    UR_CUDA_CHECK(cudaGetDevice(&dev));
    const size_t key = (size_t)kernel ^ (smem * 0x9E3779B97F4A7C15ull) ^
                       ((size_t)(dev + 1) * 0xC2B2AE3D27D4EB4Full);
    std::lock_guard<std::mutex> lock(occ_mu);
    auto it = occ_cache.find(key);
    if (it != occ_cache.end()) {
      resident = it->second;
    } else {
      UR_CUDA_CHECK(cudaFuncSetAttribute((const void *)kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident,
                                                                     (const void *)kernel, NTHR,
                                                                     smem);
      if (oe != cudaSuccess || resident < 1) resident = 1;
      occ_cache[key] = resident;
    }
  }

  // work units: groups of unit_r planes of one column, cut at low-res row starts
  S.unit_r = kind == FK_THICK_M ? r : 1;
  S.unit_e2 = kind == FK_THICK_M ? (e > 0 ? e : r) : 1;
  S.units_per_col = 1 + (S.nm > S.unit_e2 ? (S.nm - S.unit_e2 + S.unit_r - 1) / S.unit_r : 0);
  const long long slots = (long long)resident * sm_count();
  const long long q_min = (4 * (B + L + 1) + S.unit_r - 1) / S.unit_r;
  const long long max_segs = S.units_per_col / q_min > 0 ? S.units_per_col / q_min : 1;
  const unsigned gx = div_up(S.nz, TZ);
  // CTAs of the lock-step split with `rows` output rows per tile
  auto lock_ctas = [&](int rows) {
    const long long ncol = (long long)gx * div_up(S.no, rows);
    if (ncol > slots) return 0ll;
    const long long segs = slots / ncol < max_segs ? slots / ncol : max_segs;
    return segs * ncol;
  };
  // Output rows per tile: a 7-row tile (the eighth warp idles) when that fills more of the
  // CTA slots -- 256^3: 37 x 2 columns x 6 segments = 444 = 3 x 148 instead of 384 (-2 %)
  if (!dry_run && rpt == 1 && fast_to == 0 && fast_lock && fast_q_units == 0 && fast_segs == 0 &&
      lock_ctas(to_max - 1) > lock_ctas(to_max))
    to = to_max - 1;

  CUtensorMap map_v, map_r, map_x;
  if (!stream_tensor_map(A.v, A.nx, A.ny, A.nz, sz, march, to + 2, &map_v, pitch))
    return UR_ERR_UNSUPPORTED;
  map_r = map_v;
  if (combine && !stream_tensor_map(A.rres, A.nx, A.ny, A.nz, sz, march, to + 2, &map_r, pitch))
    return UR_ERR_UNSUPPORTED;
  map_x = map_v;
  if (x_fused && !stream_tensor_map(A.xup, A.nx, A.ny, A.nz, TZ, march, to, &map_x, pitch))
    return UR_ERR_UNSUPPORTED;
  if (dry_run) return UR_OK;

  const unsigned gy = div_up(S.no, to);
  S.gx = (int)gx;
  S.ncol = (int)(gx * gy);
  const long long total = (long long)S.ncol * S.units_per_col;
  long long q = (total + slots - 1) / slots;
  if (q < q_min) q = q_min;
  if (fast_q_units > 0) q = fast_q_units;
  if (q > total) q = total;
  S.q_units = (int)q;
  unsigned n_cta = (unsigned)((total + q - 1) / q);
  // Lock-step split: the same cuts in every column, one CTA per segment.  Neighbouring tiles
  // are then read at (nearly) the same time and the halo re-reads hit L2 instead of HBM.
  S.segs = 0;
  if (fast_lock && fast_q_units == 0 && S.ncol <= slots) {
    long long segs = slots / S.ncol;
    if (segs > max_segs) segs = max_segs;
    if (fast_segs > 0) segs = fast_segs < max_segs ? fast_segs : max_segs;
    if (segs * S.ncol * 4 >= slots * 3 || fast_segs > 0) {  // keep >= 75 % of the CTA slots busy
      S.segs = (int)segs;
      n_cta = (unsigned)(segs * S.ncol);
    }
  }
  S.to = to;
  // only the fused CG iteration has a vector worth keeping (the residual); elsewhere: normal
  S.l2_stream = (combine && (g_l2_hints & 1)) ? L2_FIRST : L2_NORMAL;
  S.l2_keep = (combine && (g_l2_hints & 2)) ? L2_LAST : L2_NORMAL;
  S.l2_v = (combine && (g_l2_hints & 4)) ? L2_FIRST : L2_NORMAL;
  S.pfd = fast_pfd < 0 ? 0 : fast_pfd;

  S.v = A.v;
  S.out = A.out;
  S.b = A.b;
  S.r = A.r;
  S.p = A.p;
  S.update_p = A.update_p;
  S.p_out = A.p_out;
  S.xup = A.xup;
  S.acc = A.acc;
  S.done = A.done;
  S.gr = A.gr;
  S.fin = A.fin;

  kernel<<<dim3(n_cta), dim3(NTHR), smem, st>>>(map_v, map_r, map_x, S);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

// TMA-staged streaming kernel for the CG left-hand side (sm_100a)
//
//   out = w_ident * v  +  tau * A'S^2A v  +  rho lam^2 * D'D v        (+ CG epilogue)
//
// for the lattice-aligned operators of solver.cuh (at most one observation term).
// A CTA owns a (TO x TZ) column of the volume in the two non-marching axes (TO = 8 or 16
// rows, TZ = 128: 8 warps, RPT rows per warp, one float4 per lane and row) and marches
// along the third axis ("m": x, or y when the slices are thick along y) over a contiguous
// range of planes.  Each plane tile (with a 1-row / hz-column halo) is brought into a
// shared-memory ring by ONE cp.async.bulk.tensor (TMA) per plane, completing on an
// mbarrier; out-of-volume elements are zero-filled by the TMA unit, which is exactly the
// reference's bound='zero'.  Every thread keeps the m-1 / m / m+1 values of its own
// z-quads in registers and takes its z neighbours from the adjacent lanes by shuffle, so
// per plane it reads only the next plane's quads and the two row-halo quads from shared
// memory for the 7-point D'D stencil.  The slice-profile term is evaluated through the
// decimated grid, as the reference does (pull -> conv -> scale -> conv' -> push), but
// entirely on-chip:
//   thick along m: each low-res row j is formed ONCE per thread when the march reaches
//                  its first plane (K taps over the look-ahead planes of the ring) and
//                  parked in a thread-private shared-memory slot until its last plane;
//   thick along z: the low-res row segment of the NEXT plane is formed cooperatively
//                  into a double-buffered shared array while the current plane is output.
// The (column, plane) tiles are cut into equal contiguous ranges, one per resident CTA
// slot, so the grid is exactly one balanced wave.  HBM traffic is the algorithmic 8 bytes
// per voxel (+ halo re-reads served by L2).
#include <cuda.h>
#include <math.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "solver.cuh"

namespace ur {

constexpr int TZ = 128;  // z extent of the tile (32 lanes x float4)
constexpr int NTHR = 256;
constexpr int NWARP = NTHR / 32;
constexpr int kMaxSlots = 24;
constexpr int kLrzPitch = TZ / 2 + 8;

enum { SK_NONE = 0, SK_CROP = 1, SK_THICK_M = 2, SK_THICK_Z = 3 };
enum { SC_NONE = 0, SC_CONV = 1, SC_M = 2, SC_O = 3, SC_Z = 4 };

struct StreamTerm {
  int kind;
  float tau;
  int r, K, off, nj;
  int lo_m, hi_m, lo_o, hi_o, lo_z, hi_z;
  int scl_kind, scl_off;
  float s_even, s_odd;
  float ker[UR_MAX_TAPS];
};

struct StreamArgs {
  int nm, no, nz;
  int gs_m, gs_o;  // element strides of the marching / row axis (volumes < 2^31 voxels)
  int march_y;
  float iv_m, iv_o, iv_z, rl2, w_ident;  // iv_* = 1 / vx^2
  StreamTerm T;
  int q;     // plane-tiles per CTA: contiguous range in (column, plane) order
  int ncol;  // number of (o, z) columns
  int gx;    // columns along z
  int L, B, ns, hz, sz, plane_floats, nlr;
  const float *v;
  float *out;
  const float *b;
  float *r;
  float *p;
  int update_p;
  const float *rres;  // COMBINE: residual r
  float *p_out;       // COMBINE: new direction p = beta p_old + r
  float *xup;         // COMBINE: x += alpha_prev p_old
  const int *done;
  GridReduce gr;
  FinalizeArgs fin;
};

// ------------------------------------------------------------------ PTX helpers
// Shared memory is addressed through 32-bit shared-window addresses and explicit
// ld.shared / st.shared: the ring base is computed at run time (128-byte aligned), which
// would otherwise demote every access to a generic load.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 &v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                              int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ float comp(const float4 &q, int k) {
  return k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w));
}

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

// Ring geometry in shared-window addresses.
struct Ring {
  uint32_t base, end, plane_b;  // planes
  uint32_t bar;                 // mbarriers (8 bytes each)
};

// One position of the ring: plane address, its mbarrier and the phase parity to wait for.
struct RingPos {
  uint32_t pa, ba, par;
  __device__ __forceinline__ void inc(const Ring &R) {
    pa += R.plane_b;
    ba += 8u;
    if (pa == R.end) {
      pa = R.base;
      ba = R.bar;
      par ^= 1u;
    }
  }
};

// MODE: LhsMode epilogue.  KIND: observation term (SK_*).  RPT: rows per thread (1 | 2).
template <int MODE, int KIND, int RPT>
__global__ void __launch_bounds__(NTHR, RPT == 1 ? 3 : 2)
    lhs_stream_kernel(const __grid_constant__ CUtensorMap tmap, const StreamArgs a) {
  constexpr int TO = NWARP * RPT;  // rows of the tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_red[kMaxWarps];
  __shared__ uint64_t s_bar[kMaxSlots];
  __shared__ float s_ker[UR_MAX_TAPS];
  if (a.done && *a.done) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = warp * RPT;  // first tile row of this thread
  const StreamTerm &T = a.T;
  const int ns = a.ns;

  Ring R;
  R.base = (smem_u32(smem_raw) + 127u) & ~127u;
  R.plane_b = (uint32_t)a.plane_floats * 4u;
  R.end = R.base + (uint32_t)ns * R.plane_b;
  R.bar = smem_u32(s_bar);
  const uint32_t lrm_a = R.end;  // [nlr][TO][TZ] floats (thick along m)
  const uint32_t lrm_stride = TO * TZ * 4u;
  const uint32_t lrm_end = lrm_a + (uint32_t)a.nlr * lrm_stride;
  const uint32_t lrz_a = lrm_end;  // [2][TO][kLrzPitch] floats (thick along z)

  const uint32_t plane_bytes = (uint32_t)(a.sz * (TO + 2) * sizeof(float));
  const uint32_t sz_b = (uint32_t)a.sz * 4u;
  // first own quad inside a plane (tile row row0 is plane row row0 + 1); next rows: + sz_b
  const uint32_t own_b = (uint32_t)((row0 + 1) * a.sz + a.hz + 4 * lane) * 4u;
  const uint32_t lr_own = (uint32_t)(row0 * TZ + 4 * lane) * 4u;  // next rows: + TZ * 4

  if (tid < UR_MAX_TAPS) s_ker[tid] = T.ker[tid];
  if (tid == 0) {
    for (int s = 0; s < ns; ++s) mbar_init(&s_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // contiguous range of plane-tiles (column-major: column, then plane) of this CTA
  const long long total = (long long)a.ncol * a.nm;
  const long long t_begin = (long long)blockIdx.x * a.q;
  const long long t_end = t_begin + a.q < total ? t_begin + a.q : total;

  RingPos ip{R.base, R.bar, 0u};  // producer position (thread 0)
  RingPos wp{R.base, R.bar, 0u};  // consumer position (every thread)
  double part = 0.0;

  for (long long t = t_begin; t < t_end;) {
    const int col = (int)(t / a.nm);
    const int m0 = (int)(t - (long long)col * a.nm);
    const long long left_tiles = t_end - t;
    const int m1 = (left_tiles < (long long)(a.nm - m0)) ? m0 + (int)left_tiles : a.nm;
    t += m1 - m0;
    const int cz = col % a.gx, co = col / a.gx;
    const int z0 = cz * TZ, o0 = co * TO;
    const int o_first = o0 + row0, z = z0 + 4 * lane;
    const bool z_ok = z < a.nz;

    const int u_begin = m0 - a.B;
    const int first = u_begin - 1;
    const int last = m1 - 1 + a.L;

    int iq = first;  // next plane to issue
    auto issue_next = [&]() {
      // COMBINE rewrote the slot with generic-proxy stores: order them before the TMA write
      if (MODE == LHS_COMBINE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx_a(ip.ba, plane_bytes);
      if (a.march_y)
        tma_load_3d_a(ip.pa, &tmap, ip.ba, z0 - a.hz, iq, o0 - 1);
      else
        tma_load_3d_a(ip.pa, &tmap, ip.ba, z0 - a.hz, o0 - 1, iq);
      ip.inc(R);
      ++iq;
    };
    if (tid == 0)
      for (int n = 0; n < ns && iq <= last; ++n) issue_next();

    // ---- per-thread constants of the observation term ----
    bool active[RPT], o_in[RPT];
    float4 thin[RPT];  // tau x (scaling that alternates along a thin axis): per-voxel factor
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int o = o_first + i;
      active[i] = z_ok && o < a.no;
      o_in[i] = KIND != SK_NONE && o >= T.lo_o && o < T.hi_o;
      thin[i] = make_float4(T.tau, T.tau, T.tau, T.tau);
      if (KIND != SK_NONE) {
        if (T.scl_kind == SC_O) {
          const float f = T.tau * (((o - T.scl_off) & 1) ? T.s_odd : T.s_even);
          thin[i] = make_float4(f, f, f, f);
        } else if (T.scl_kind == SC_Z) {
          const float f0 = T.tau * (((z - T.scl_off) & 1) ? T.s_odd : T.s_even);
          const float f1 = T.tau * (((z - T.scl_off) & 1) ? T.s_even : T.s_odd);
          thin[i] = make_float4(f0, f1, f0, f1);
        }
      }
    }
    float4 zmask = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KIND == SK_CROP || KIND == SK_THICK_M) {
      zmask.x = (z + 0 >= T.lo_z && z + 0 < T.hi_z) ? 1.f : 0.f;
      zmask.y = (z + 1 >= T.lo_z && z + 1 < T.hi_z) ? 1.f : 0.f;
      zmask.z = (z + 2 >= T.lo_z && z + 2 < T.hi_z) ? 1.f : 0.f;
      zmask.w = (z + 3 >= T.lo_z && z + 3 < T.hi_z) ? 1.f : 0.f;
    }
    // thick along z: low-res rows touching this tile and, per component of this thread's
    // quad, the local index of its highest row and the tap that row contributes
    int jz_lo = 0, njt = 0;
    int zj0[4] = {0, 0, 0, 0}, zt0[4] = {UR_MAX_TAPS, UR_MAX_TAPS, UR_MAX_TAPS, UR_MAX_TAPS};
    uint32_t lrz_src_off = 0;
    if (KIND == SK_THICK_Z) {
      const int a0 = z0 - T.off - T.K + 1;
      jz_lo = a0 <= 0 ? 0 : (a0 + T.r - 1) / T.r;
      int jz_hi = (z0 + TZ - 1 - T.off) >= 0 ? (z0 + TZ - 1 - T.off) / T.r : -1;
      if (jz_hi > T.nj - 1) jz_hi = T.nj - 1;
      njt = jz_hi - jz_lo + 1;
      if (njt < 0) njt = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int up = z + k - T.off;
        if (up >= 0) {
          int j_hi = up / T.r;
          if (j_hi > T.nj - 1) j_hi = T.nj - 1;
          zj0[k] = j_hi - jz_lo;
          zt0[k] = up - j_hi * T.r;
        }
      }
      lrz_src_off = (uint32_t)((row0 + 1) * a.sz + a.hz + (jz_lo * T.r + T.off - z0)) * 4u;
    }
    auto build_lrz = [&](uint32_t plane_a, int buf) {  // low-res z-rows, RPT rows per warp
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const uint32_t dst = lrz_a + (uint32_t)((buf * TO + row0 + i) * kLrzPitch) * 4u;
        for (int jj = lane; jj < njt; jj += 32) {
          const uint32_t src = plane_a + lrz_src_off + i * sz_b + (uint32_t)(jj * T.r) * 4u;
          float acc = 0.f;
          for (int tt = 0; tt < T.K; ++tt) acc = fmaf(s_ker[tt], lds32(src + 4u * tt), acc);
          if (T.scl_kind == SC_CONV) acc *= ((jz_lo + jj) & 1) ? T.s_odd : T.s_even;
          sts32(dst + 4u * jj, acc);
        }
      }
    };

    // ---- COMBINE: p = beta p_old + r on the whole staged tile (own quads + halo) ----
    // Every thread owns its RPT interior quads and (threads < halo_n) one halo quad of the
    // tile.  r (and x for the delayed x += alpha p_old) come straight from global memory,
    // software-pipelined one plane ahead; p_old is what the TMA brought into the ring slot
    // and is overwritten in place by the new direction.
    float beta_c = 0.f, alpha_c = 0.f;
    uint32_t h_off = 0;
    int h_g = 0;
    bool h_has = false, h_in = false;
    int g_own = 0;
    if (MODE == LHS_COMBINE) {
      beta_c = (float)a.fin.st->beta;
      alpha_c = (float)a.fin.st->alpha;
      const int sz4 = a.sz >> 2, hz4 = a.hz >> 2;
      const int halo_n = 2 * sz4 + TO * 2 * hz4;
      h_has = tid < halo_n;
      int h_row = 0, h_c4 = 0;
      if (tid < sz4) {
        h_row = 0;
        h_c4 = tid;
      } else if (tid < 2 * sz4) {
        h_row = TO + 1;
        h_c4 = tid - sz4;
      } else {
        const int k = tid - 2 * sz4;
        h_row = 1 + k / (2 * hz4);
        const int s = k - (h_row - 1) * (2 * hz4);
        h_c4 = s < hz4 ? s : sz4 - 2 * hz4 + s;
      }
      h_off = (uint32_t)(h_row * a.sz + 4 * h_c4) * 4u;
      const int h_o = o0 - 1 + h_row, h_z = z0 - a.hz + 4 * h_c4;
      h_in = h_has && h_o >= 0 && h_o < a.no && h_z >= 0 && h_z < a.nz;
      h_g = h_in ? h_o * a.gs_o + h_z : 0;
      g_own = o_first * a.gs_o + z;
    }
    float4 cr_own[RPT], cx_own[RPT], cr_h;  // prefetched r / x quads of the plane to combine
    // gq = q * gs_m as a signed element offset, advanced incrementally by the callers (it is
    // only dereferenced for planes inside the volume)
    auto fetch_rx = [&](int q, int gq) {
      const bool q_in = q >= 0 && q < a.nm;
      const bool q_own = q >= m0 && q < m1;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int gi = gq + g_own + i * a.gs_o;
        cr_own[i] = (q_in && active[i]) ? *reinterpret_cast<const float4 *>(a.rres + gi)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        cx_own[i] = (q_own && active[i]) ? *reinterpret_cast<const float4 *>(a.xup + gi)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cr_h = (q_in && h_in) ? *reinterpret_cast<const float4 *>(a.rres + (gq + h_g))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto combine = [&](int q, uint32_t aq, int gq) {
      const bool q_own = q >= m0 && q < m1;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const uint32_t sa = aq + own_b + i * sz_b;
        const float4 po = lds128(sa);
        float4 pn;  // torch: p *= beta; p += r  (two roundings)
        pn.x = __fadd_rn(__fmul_rn(beta_c, po.x), cr_own[i].x);
        pn.y = __fadd_rn(__fmul_rn(beta_c, po.y), cr_own[i].y);
        pn.z = __fadd_rn(__fmul_rn(beta_c, po.z), cr_own[i].z);
        pn.w = __fadd_rn(__fmul_rn(beta_c, po.w), cr_own[i].w);
        sts128(sa, pn);
        if (q_own && active[i]) {
          const int gi = gq + g_own + i * a.gs_o;
          *reinterpret_cast<float4 *>(a.p_out + gi) = pn;
          float4 xn;  // the previous iteration's x += alpha p, done now that p_old is at hand
          xn.x = __fadd_rn(cx_own[i].x, __fmul_rn(alpha_c, po.x));
          xn.y = __fadd_rn(cx_own[i].y, __fmul_rn(alpha_c, po.y));
          xn.z = __fadd_rn(cx_own[i].z, __fmul_rn(alpha_c, po.z));
          xn.w = __fadd_rn(cx_own[i].w, __fmul_rn(alpha_c, po.w));
          *reinterpret_cast<float4 *>(a.xup + gi) = xn;
        }
      }
      if (h_has) {
        const float4 po = lds128(aq + h_off);
        float4 pn;
        pn.x = __fadd_rn(__fmul_rn(beta_c, po.x), cr_h.x);
        pn.y = __fadd_rn(__fmul_rn(beta_c, po.y), cr_h.y);
        pn.z = __fadd_rn(__fmul_rn(beta_c, po.z), cr_h.z);
        pn.w = __fadd_rn(__fmul_rn(beta_c, po.w), cr_h.w);
        sts128(aq + h_off, pn);
      }
    };

    // ---- prime the pipeline: planes first .. u_begin + L - 1 (COMBINE: one more, each
    //      combined as it lands) ----
    const uint32_t a_first = wp.pa;
    int gq_c = first * a.gs_m;  // COMBINE: element offset of the plane to combine
    for (int n = 0; n < a.L + 1 + (MODE == LHS_COMBINE ? 1 : 0); ++n) {
      const uint32_t aq = wp.pa;
      mbar_wait_a(wp.ba, wp.par);
      wp.inc(R);
      if (MODE == LHS_COMBINE) {
        fetch_rx(first + n, gq_c);
        combine(first + n, aq, gq_c);
        gq_c += a.gs_m;
      }
    }
    if (MODE == LHS_COMBINE) {
      __syncthreads();  // halo quads of the primed planes were written by other threads
      fetch_rx(u_begin + a.L + 1, gq_c);
    }
    uint32_t au = a_first + R.plane_b;  // plane u
    if (au == R.end) au = R.base;
    float4 prev[RPT], cur[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      prev[i] = lds128(a_first + own_b + i * sz_b);
      cur[i] = lds128(au + own_b + i * sz_b);
    }

    int ph = 0, jrow = 0;    // thick along m: phase inside the stride / current low-res row
    uint32_t jaddr = lrm_a;  // parking slot of row jrow
    if (KIND == SK_THICK_M) {
      const int t0 = u_begin - T.off;
      jrow = floordiv(t0, T.r);
      ph = t0 - jrow * T.r;
      jaddr = lrm_a + (uint32_t)(jrow - floordiv(jrow, a.nlr) * a.nlr) * lrm_stride;
    }
    if (KIND == SK_THICK_Z) {
      build_lrz(au, u_begin & 1);
      __syncthreads();
    }
    const bool o_is0 = o_first == 0, z_is0 = z == 0;
    // global element offset of this thread's first quad in plane u (advanced per output plane)
    int goff = m0 * a.gs_m + o_first * a.gs_o + z;
    int pend = 0;  // planes consumed since the ring was last refilled

    for (int u = u_begin; u < m1; ++u) {
      if (MODE == LHS_COMBINE) {
        // plane u + L + 1 lands now and is combined one iteration before its first use (the
        // barrier at the end of this iteration publishes the halo quads)
        const int qc = u + a.L + 1;
        if (qc <= last) {
          const uint32_t aq = wp.pa;
          mbar_wait_a(wp.ba, wp.par);
          wp.inc(R);
          combine(qc, aq, gq_c);
          gq_c += a.gs_m;
          if (qc + 1 <= last) fetch_rx(qc + 1, gq_c);
        }
      } else {
        mbar_wait_a(wp.ba, wp.par);  // plane u + L has landed
        wp.inc(R);
      }
      uint32_t an = au + R.plane_b;  // plane u + 1
      if (an == R.end) an = R.base;
      float4 next[RPT];
      float zl[RPT], zr[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        next[i] = lds128(an + own_b + i * sz_b);
        // z neighbours of the quad come from the adjacent lanes; only the tile edges read smem
        zl[i] = __shfl_up_sync(0xffffffffu, cur[i].w, 1);
        zr[i] = __shfl_down_sync(0xffffffffu, cur[i].x, 1);
        if (lane == 0) zl[i] = lds32(au + own_b + i * sz_b - 4u);
        if (lane == 31) zr[i] = lds32(au + own_b + i * sz_b + 16u);
      }

      if (KIND == SK_THICK_M) {
        if (ph == 0 && jrow >= 0 && jrow < T.nj) {
          // low-res row jrow starts at this plane: form it once and park it
          float sc = 1.f;
          if (T.scl_kind == SC_CONV) sc = (jrow & 1) ? T.s_odd : T.s_even;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            if (o_in[i] && active[i]) {
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              uint32_t s = au;
              for (int tt = 0; tt < T.K; ++tt) {
                const float4 q =
                    tt == 0 ? cur[i] : (tt == 1 ? next[i] : lds128(s + own_b + i * sz_b));
                const float k = s_ker[tt];
                acc.x = fmaf(k, q.x, acc.x);
                acc.y = fmaf(k, q.y, acc.y);
                acc.z = fmaf(k, q.z, acc.z);
                acc.w = fmaf(k, q.w, acc.w);
                s += R.plane_b;
                if (s == R.end) s = R.base;
              }
              acc.x *= sc * zmask.x;
              acc.y *= sc * zmask.y;
              acc.z *= sc * zmask.z;
              acc.w *= sc * zmask.w;
              sts128(jaddr + lr_own + i * (TZ * 4u), acc);
            }
          }
        }
      } else if (KIND == SK_THICK_Z) {
        if (u + 1 < m1) build_lrz(an, (u + 1) & 1);
      }

      if (u >= m0) {
        // row halos: above the first own row, below the last own row
        float4 om_edge = lds128(au + own_b - sz_b);
        const float4 op_edge = lds128(au + own_b + RPT * sz_b);
        if (o_is0) om_edge = cur[0];
        float mfac = 1.f;
        if (KIND != SK_NONE && T.scl_kind == SC_M)
          mfac = ((u - T.scl_off) & 1) ? T.s_odd : T.s_even;
        const bool m_in = KIND != SK_NONE && u >= T.lo_m && u < T.hi_m;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          if (!active[i]) continue;
          const int gi = goff + i * a.gs_o;
          float4 bq = make_float4(0.f, 0.f, 0.f, 0.f), rq = bq, pq = bq;
          if (MODE == LHS_RESID || MODE == LHS_ENERGY)
            bq = *reinterpret_cast<const float4 *>(a.b + gi);
          if (MODE == LHS_ENERGY && a.update_p) {
            rq = *reinterpret_cast<const float4 *>(a.r + gi);
            pq = *reinterpret_cast<const float4 *>(a.p + gi);
          }
          // ---- observation term through the decimated grid ----
          float4 dat = make_float4(0.f, 0.f, 0.f, 0.f);
          if (KIND == SK_CROP) {
            if (o_in[i] && m_in) {
              dat.x = cur[i].x * zmask.x;
              dat.y = cur[i].y * zmask.y;
              dat.z = cur[i].z * zmask.z;
              dat.w = cur[i].w * zmask.w;
            }
          } else if (KIND == SK_THICK_M) {
            if (o_in[i]) {
              int jr = jrow;
              uint32_t js = jaddr;
              for (int tap = ph; tap < T.K; tap += T.r) {
                if (jr >= 0 && jr < T.nj) {
                  const float4 lr = lds128(js + lr_own + i * (TZ * 4u));
                  const float k = s_ker[tap];
                  dat.x = fmaf(k, lr.x, dat.x);
                  dat.y = fmaf(k, lr.y, dat.y);
                  dat.z = fmaf(k, lr.z, dat.z);
                  dat.w = fmaf(k, lr.w, dat.w);
                }
                --jr;
                js = (js == lrm_a ? lrm_end : js) - lrm_stride;
              }
            }
          } else if (KIND == SK_THICK_Z) {
            if (o_in[i] && m_in) {
              const uint32_t lr =
                  lrz_a + (uint32_t)(((u & 1) * TO + row0 + i) * kLrzPitch) * 4u;
              float d[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float acc = 0.f;
                int jl = zj0[k];
                for (int tap = zt0[k]; tap < T.K && jl >= 0; tap += T.r, --jl)
                  acc = fmaf(s_ker[tap], lds32(lr + 4u * jl), acc);
                d[k] = acc;
              }
              dat = make_float4(d[0], d[1], d[2], d[3]);
            }
          }
          // ---- D'D: per axis (2c - lo - hi) / vx^2 with the "lo" term dropped on the low
          //      edge; values past the high edge are the TMA's zero fill (bound = zero) ----
          const float4 pv = u > 0 ? prev[i] : cur[i];
          const float4 om = i == 0 ? om_edge : cur[i > 0 ? i - 1 : 0];
          const float4 op = i == RPT - 1 ? op_edge : cur[i < RPT - 1 ? i + 1 : 0];
          const float zlv = z_is0 ? cur[i].x : zl[i];
          float val[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float c = comp(cur[i], k);
            const float lft = k == 0 ? zlv : comp(cur[i], k - 1);
            const float rgt = k == 3 ? zr[i] : comp(cur[i], k + 1);
            // (2c - lo) - hi per axis: one FFMA + one FADD; on a low edge lo := c makes the
            // first term exactly c
            const float d_m = fmaf(2.f, c, -comp(pv, k)) - comp(next[i], k);
            const float d_o = fmaf(2.f, c, -comp(om, k)) - comp(op, k);
            const float d_z = fmaf(2.f, c, -lft) - rgt;
            const float dtd = fmaf(d_z, a.iv_z, fmaf(d_o, a.iv_o, d_m * a.iv_m));
            float data = a.w_ident * c;
            if (KIND != SK_NONE) data = fmaf(mfac * comp(thin[i], k), comp(dat, k), data);
            val[k] = fmaf(a.rl2, dtd, data);
          }
          if (MODE == LHS_PLAIN || MODE == LHS_COMBINE) {
            *reinterpret_cast<float4 *>(a.out + gi) = make_float4(val[0], val[1], val[2], val[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) part += (double)__fmul_rn(comp(cur[i], k), val[k]);
          } else if (MODE == LHS_RESID) {
            float4 rr;
            rr.x = __fsub_rn(bq.x, val[0]);
            rr.y = __fsub_rn(bq.y, val[1]);
            rr.z = __fsub_rn(bq.z, val[2]);
            rr.w = __fsub_rn(bq.w, val[3]);
            *reinterpret_cast<float4 *>(a.r + gi) = rr;
            *reinterpret_cast<float4 *>(a.p + gi) = rr;
            part += (double)__fmul_rn(rr.x, rr.x) + (double)__fmul_rn(rr.y, rr.y) +
                    (double)__fmul_rn(rr.z, rr.z) + (double)__fmul_rn(rr.w, rr.w);
          } else if (MODE == LHS_ENERGY) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              part += (double)__fmul_rn(__fsub_rn(val[k], 2.f * comp(bq, k)), comp(cur[i], k));
            if (a.update_p) {
              const float beta = (float)a.fin.st->beta;
              float4 pn;
              pn.x = __fadd_rn(__fmul_rn(beta, pq.x), rq.x);
              pn.y = __fadd_rn(__fmul_rn(beta, pq.y), rq.y);
              pn.z = __fadd_rn(__fmul_rn(beta, pq.z), rq.z);
              pn.w = __fadd_rn(__fmul_rn(beta, pq.w), rq.w);
              *reinterpret_cast<float4 *>(a.p + gi) = pn;
            }
          }
        }
      }
      // ---- advance ----
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        prev[i] = cur[i];
        cur[i] = next[i];
      }
      au = an;
      if (KIND == SK_THICK_M) {
        if (++ph == T.r) {
          ph = 0;
          ++jrow;
          jaddr += lrm_stride;
          if (jaddr == lrm_end) jaddr = lrm_a;
        }
      }
      if (u >= m0) goff += a.gs_m;
      // Refill the ring every kRefill planes (every plane for the z-thick kernel, whose lrz
      // double buffer needs the barrier anyway): after the barrier every warp is done with
      // the planes up to u - 1, so that many slots are free again.
      // (COMBINE: the quads other warps rewrote in plane u + L + 1 are first read L + 1 >= 2
      // iterations later, so a barrier every second plane also publishes them in time)
      constexpr int kRefill = KIND == SK_THICK_Z ? 1 : 2;
      if (++pend == kRefill || u == m1 - 1) {
        __syncthreads();
        if (tid == 0)
          for (int k = 0; k < pend && iq <= last; ++k) issue_next();
        pend = 0;
      }
    }
  }
  double total_sum;
  if (grid_sum(part, a.gr, s_red, &total_sum) && tid == 0) finalize(a.fin, total_sum);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault,
                                         &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct MapKey {
  const void *ptr;
  int nx, ny, nz, sz, march_y, rows, pitch;
  bool operator==(const MapKey &o) const {
    return ptr == o.ptr && nx == o.nx && ny == o.ny && nz == o.nz && sz == o.sz &&
           march_y == o.march_y && rows == o.rows && pitch == o.pitch;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey &k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1315423911u + k.nx;
    h = h * 1315423911u + k.ny;
    h = h * 1315423911u + k.nz;
    h = h * 1315423911u + k.sz * 2 + k.march_y;
    h = h * 1315423911u + k.rows;
    h = h * 1315423911u + k.pitch;
    return h;
  }
};

// 3-D tensor map over the (X, Y, Z) float volume with a box of one plane tile:
// sz floats along z, `rows` rows along the non-marching axis, 1 plane along the march.
// `pitch`: elements between consecutive z rows in memory (>= nz, multiple of 4).
static bool get_tensor_map(const float *v, int nx, int ny, int nz, int sz, int march_y, int rows,
                           CUtensorMap *out, int pitch = 0) {
  if (pitch <= 0) pitch = nz;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{v, nx, ny, nz, sz, march_y, rows, pitch};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx};
  cuuint64_t gstr[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * ny * 4};
  cuuint32_t box[3] = {(cuuint32_t)sz, march_y ? 1u : (cuuint32_t)rows,
                       march_y ? (cuuint32_t)rows : 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult rc = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)v, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

// 3-D tensor map over an (n0, n1, n2) float volume (n2 contiguous) with an arbitrary box
// (b0, b1, b2): the input boxes of the multi-axis lattice kernels (lattice_nd.cu).  Needs
// n2 % 4 == 0, b2 % 4 == 0, box extents <= 256 and a 16-byte aligned base.
bool box_tensor_map(const float *v, int n0, int n1, int n2, int b0, int b1, int b2,
                    CUtensorMap *out) {
  if (n2 % 4 != 0 || b2 % 4 != 0 || b0 > 256 || b1 > 256 || b2 > 256 || b0 < 1 || b1 < 1 ||
      b2 < 1 || ((uintptr_t)v & 15u) != 0)
    return false;
  // same cache as the plane maps: (sz, rows, pitch) carry the box, march_y = 2 marks the kind
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{v, n0, n1, n2, b2, 2, b1, b0};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
  cuuint64_t gstr[2] = {(cuuint64_t)n2 * 4, (cuuint64_t)n2 * n1 * 4};
  cuuint32_t box[3] = {(cuuint32_t)b2, (cuuint32_t)b1, (cuuint32_t)b0};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult rc = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)v, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

bool stream_tensor_map(const float *v, int nx, int ny, int nz, int sz, int march_y, int rows,
                       CUtensorMap *out, int pitch) {
  return get_tensor_map(v, nx, ny, nz, sz, march_y, rows, out, pitch);
}

static bool a16(const void *p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; }

int stream_mc_override = 0;  // test / tuning hook (plane-tiles per CTA), 0 = automatic
int stream_rpt = 0;          // rows per thread: 0 automatic, 1 (8-row tiles), 2 (16-row tiles)
int stream_pf = 0;           // extra prefetch slots of the ring on top of the default
typedef void (*StreamKernel)(const CUtensorMap, const StreamArgs);

// variant < 0: dry run (eligibility check only, nothing is launched)
int lhs_stream_launch(int mode, const LhsArgs &A, int variant, cudaStream_t st) {
  const bool dry_run = variant < 0;
  if (A.acc != nullptr || A.nterm > 1 || A.nrot > 0) return UR_ERR_UNSUPPORTED;
  if (A.nz % 4 != 0 || A.nz < 4) return UR_ERR_UNSUPPORTED;
  // 32-bit element offsets inside the kernel (incl. the look-ahead planes past the end)
  if ((long long)A.nx * A.ny * A.nz + 64ll * A.ny * A.nz + 64ll * A.nx * A.nz > 0x7fffffffll)
    return UR_ERR_UNSUPPORTED;
  if (!a16(A.v) || !a16(A.out) || !a16(A.b) || !a16(A.r) || !a16(A.p)) return UR_ERR_UNSUPPORTED;
  if (!a16(A.rres) || !a16(A.p_out) || !a16(A.xup)) return UR_ERR_UNSUPPORTED;
  const bool combine = mode == LHS_COMBINE;

  StreamArgs S;
  memset(&S, 0, sizeof(S));
  StreamTerm &T = S.T;
  T.kind = SK_NONE;
  T.r = T.K = 1;
  T.s_even = T.s_odd = 1.f;
  int march = 0;  // 0: x, 1: y
  int hz = 4;
  if (A.nterm == 1) {
    const LatticeTerm &L = A.term[0];
    if (L.axis == 1) march = 1;
    const int ax_m = march, ax_o = 1 - march;
    T.tau = L.tau;
    T.r = L.r;
    T.K = L.K;
    T.off = L.off;
    T.nj = L.nj;
    T.lo_m = L.lo[ax_m];
    T.hi_m = L.hi[ax_m];
    T.lo_o = L.lo[ax_o];
    T.hi_o = L.hi[ax_o];
    T.lo_z = L.lo[2];
    T.hi_z = L.hi[2];
    for (int t = 0; t < UR_MAX_TAPS; ++t) T.ker[t] = L.ker[t];
    T.s_even = L.s_even;
    T.s_odd = L.s_odd;
    T.scl_off = L.scl_off;
    if (L.axis < 0) {
      T.kind = SK_CROP;
    } else if (L.axis == 2) {
      T.kind = SK_THICK_Z;
      if (T.r < 2 || T.K - 1 > 8) return UR_ERR_UNSUPPORTED;
      hz = (T.K - 1 + 3) / 4 * 4;
      if (hz < 4) hz = 4;
      T.lo_z = 0;
      T.hi_z = A.nz;
    } else {
      T.kind = SK_THICK_M;
      if (T.K - 1 > 12) return UR_ERR_UNSUPPORTED;
      T.lo_m = 0;
      T.hi_m = march ? A.ny : A.nx;
    }
    if (L.scl_axis < 0)
      T.scl_kind = SC_NONE;
    else if (L.scl_axis == L.axis)
      T.scl_kind = SC_CONV;
    else if (L.scl_axis == 2)
      T.scl_kind = SC_Z;
    else
      T.scl_kind = (L.scl_axis == ax_m) ? SC_M : SC_O;
  }
  S.march_y = march;
  S.nm = march ? A.ny : A.nx;
  S.no = march ? A.nx : A.ny;
  S.nz = A.nz;
  S.gs_m = march ? A.nz : A.ny * A.nz;
  S.gs_o = march ? A.ny * A.nz : A.nz;
  S.iv_m = march ? A.ivy * A.ivy : A.ivx * A.ivx;  // 1 / vx^2 per axis
  S.iv_o = march ? A.ivx * A.ivx : A.ivy * A.ivy;
  S.iv_z = A.ivz * A.ivz;
  S.rl2 = A.rl2;
  S.w_ident = A.w_ident;
  S.L = 1;
  S.B = 0;
  S.nlr = 0;
  if (T.kind == SK_THICK_M) {
    S.L = T.K - 1 > 1 ? T.K - 1 : 1;
    S.B = T.K - 1;
    S.nlr = (T.K + T.r - 1) / T.r + 1;
    if (S.nlr > 6) return UR_ERR_UNSUPPORTED;
  }
  // rows per thread: 16-row tiles halve the per-plane bookkeeping per voxel; fall back to
  // 8-row tiles when the volume has few rows or the ring would not fit twice per SM
  // (measured at 256^3: 16-row tiles win without a thick term, 8-row tiles with one --
  // the thick kernels need 3 resident CTAs per SM to hide the ring latency)
  int rpt = (T.kind == SK_THICK_M || T.kind == SK_THICK_Z) ? 1 : 2;
  if (stream_rpt == 1 || stream_rpt == 2) rpt = stream_rpt;
  if (S.no <= NWARP) rpt = 1;
  size_t smem = 0;
  for (;; rpt = 1) {
    const int to = NWARP * rpt;
    // window [u-1, u+L] + prefetch + one slot of slack for the every-other-plane refill
    // (COMBINE looks one plane further ahead and refills every plane)
    S.ns = S.L + 2 + (rpt == 1 ? 3 : 2) + stream_pf + (combine ? 1 : 0) +
           (T.kind == SK_THICK_Z ? 0 : 1);
    if (combine && 2 * ((TZ + 2 * hz) / 4) + to * 2 * (hz / 4) > NTHR) return UR_ERR_UNSUPPORTED;
    if (S.ns > kMaxSlots) return UR_ERR_UNSUPPORTED;
    S.hz = hz;
    S.sz = TZ + 2 * hz;
    S.plane_floats = (S.sz * (to + 2) + 31) / 32 * 32;
    smem = ((size_t)S.ns * S.plane_floats + (size_t)S.nlr * to * TZ +
            (T.kind == SK_THICK_Z ? 2 * to * kLrzPitch : 0)) *
               sizeof(float) +
           128;  // slack for the 128-byte alignment of the ring
    if (smem <= (rpt == 1 ? 200u : 112u) * 1024u) break;
    if (rpt == 1) return UR_ERR_UNSUPPORTED;
  }
  const int to = NWARP * rpt;

#define UR_SK_ROW(M)                                                                        \
  {                                                                                         \
    {lhs_stream_kernel<M, SK_NONE, 1>, lhs_stream_kernel<M, SK_NONE, 2>},                   \
        {lhs_stream_kernel<M, SK_CROP, 1>, lhs_stream_kernel<M, SK_CROP, 2>},               \
        {lhs_stream_kernel<M, SK_THICK_M, 1>, lhs_stream_kernel<M, SK_THICK_M, 2>},         \
        {lhs_stream_kernel<M, SK_THICK_Z, 1>, lhs_stream_kernel<M, SK_THICK_Z, 2>},         \
  }
  static StreamKernel table[4][4][2] = {UR_SK_ROW(LHS_PLAIN), UR_SK_ROW(LHS_RESID),
                                        UR_SK_ROW(LHS_ENERGY), UR_SK_ROW(LHS_COMBINE)};
#undef UR_SK_ROW
  if (!encode_fn()) return UR_ERR_UNSUPPORTED;  // no cuTensorMapEncodeTiled: no TMA path
  if (dry_run) return UR_OK;
  // per DEVICE: the opt-in to > 48 KB of dynamic shared memory is a device-side attribute
  static bool attr_set_dev[64] = {false};
  int dev = 0;
  UR_CUDA_CHECK(cudaGetDevice(&dev));
  UR_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  bool &attr_set = attr_set_dev[dev];
  if (!attr_set) {
    for (int i = 0; i < 4; ++i)
      for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 2; ++j)
          UR_CUDA_CHECK(cudaFuncSetAttribute((const void *)table[i][k][j],
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024));
    attr_set = true;
  }
  const int mi = mode == LHS_PLAIN ? 0 : (mode == LHS_RESID ? 1 : (mode == LHS_ENERGY ? 2 : 3));
  StreamKernel kernel = table[mi][T.kind][rpt - 1];
  // One wave of equally loaded CTAs: the (column, plane) tiles are cut into contiguous
  // ranges, one per resident CTA slot (occupancy x SM count).
  const unsigned gx = div_up(S.nz, TZ), gy = div_up(S.no, to);
  S.gx = (int)gx;
  S.ncol = (int)(gx * gy);
  const long long total = (long long)S.ncol * S.nm;
  int resident = 0;
  {
    // the occupancy query costs several microseconds of host time: remember the answer
    static std::unordered_map<size_t, int> occ_cache;
    static std::mutex occ_mu;
    const size_t key = (size_t)kernel ^ (smem * 0x9E3779B97F4A7C15ull) ^
                       ((size_t)(dev + 1) * 0xC2B2AE3D27D4EB4Full);
    std::lock_guard<std::mutex> lock(occ_mu);
    auto it = occ_cache.find(key);
    if (it != occ_cache.end()) {
      resident = it->second;
    } else {
      cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &resident, (const void *)kernel, NTHR, smem);
      if (oe != cudaSuccess || resident < 1) resident = 1;
      occ_cache[key] = resident;
    }
  }
  const long long slots = (long long)resident * sm_count();
  const long long q_min = 4 * (S.B + S.L + 1);  // keep the warm-up planes a small fraction
  long long q = (total + slots - 1) / slots;
  if (q < q_min) q = q_min;
  if (stream_mc_override > 0) q = stream_mc_override;
  if (q > total) q = total;
  S.q = (int)q;
  const unsigned n_cta = (unsigned)((total + q - 1) / q);

  CUtensorMap map;
  if (!get_tensor_map(A.v, A.nx, A.ny, A.nz, S.sz, march, to + 2, &map))
    return UR_ERR_UNSUPPORTED;

  S.v = A.v;
  S.out = A.out;
  S.b = A.b;
  S.r = A.r;
  S.p = A.p;
  S.update_p = A.update_p;
  S.rres = A.rres;
  S.p_out = A.p_out;
  S.xup = A.xup;
  S.done = A.done;
  S.gr = A.gr;
  S.fin = A.fin;

  kernel<<<dim3(n_cta), dim3(NTHR), smem, st>>>(map, S);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

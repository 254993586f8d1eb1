// Lean TMA streaming kernel for the CG left-hand side (sm_100a), compile-time specialised.
//
//   out = w_ident v + tau A'S^2A v + rho lam^2 D'D v      (+ CG epilogue, see LhsMode)
//
// Same operator and tiling idea as lhs_stream.cu (a CTA owns a (TO x 128) column and marches
// along the third axis through a shared-memory ring of TMA-loaded plane tiles; the TMA zero
// fill is the reference's bound='zero'), rebuilt so that the per-plane work of a thread is a
// short straight-line sequence:
//   * the slice-profile taps (KP), the decimation ratio (R) and the tile shape (RPT rows per
//     thread) are template parameters; everything that depends on the thread's position is
//     folded into per-thread constants once per column (stencil diagonal incl. the low-edge
//     rule of D'D, FOV masks, even/odd scaling);
//   * D'D is evaluated as  diag*c - sum_axis a_axis (lo + hi): 8 instructions per voxel;
//   * thick slices along the march axis: the low-resolution rows live in REGISTERS (each
//     thread forms "its" quad of row j once, when the march reaches the row's first plane);
//     the column ranges are cut at row starts, so a cut costs KP-1 pre-roll planes + 1;
//   * thick slices along z: every lane forms the <= 8/R low-resolution values whose windows
//     start in the quad left of it or in its own quad from three LDS.128 -- no shuffles, no
//     shared staging, no barrier;
//   * fused direction update (LHS_COMBINE): the residual tile (with halo) arrives through a
//     second TMA ring on the same mbarrier; p = beta p_old + r is formed in place in shared
//     memory, x += alpha p_old uses x quads prefetched two planes ahead;
//   * planes are processed in pairs: [arrive (+combine) 2 planes | __syncthreads | refill the
//     ring | stencil 2 planes], one barrier per two planes.
// Reference semantics: unires/_project.py:73-87 (lhs), :99-190 (AtA), :300-317 (DtD).
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "solver.cuh"

namespace ur {
namespace fast {

constexpr int TZ = 128;          // z extent of a tile (32 lanes x float4)
// z halo of a tile row: one quad each side; two for the thick-z kinds with > 5 taps
constexpr int fast_hz(int kind, int kp) { return ((kind == 3 || kind == 4) && kp - 1 > 4) ? 8 : 4; }
constexpr int NTHR = 256;
constexpr int NWARP = NTHR / 32;
constexpr int kMaxSlots = 32;
constexpr int kTaps = 16;

// FK_THICK_Z: ratio divides 4 (register scheme); FK_THICK_ZG: any ratio (per-lane windows)
enum { FK_NONE = 0, FK_POINT = 1, FK_THICK_M = 2, FK_THICK_Z = 3, FK_THICK_ZG = 4 };

struct FastArgs {
  int nm, no, nz;
  int gs_m, gs_o;  // element strides of the march / row axis
  int march_y;
  float a_m, a_o, a_z;  // rho lam^2 / vx^2 per axis
  float d0;             // w_ident + 2 (a_m + a_o + a_z), rounded to float ...
  float nd0l;           // ... and minus the rounding residue: constants stay in the null space of
                        // D'D (an uncorrected residue e acts as a spurious e * identity term)
  // observation term (lattice aligned)
  float tau;
  int off, nj;
  int lo_m, hi_m, lo_o, hi_o, lo_z, hi_z;
  int scl_conv;
  float s_even, s_odd;
  float ker[kTaps];   // taps of the decimating correlation
  float kerT[kTaps];  // tau * taps (transpose side), zero padded
  // work split: a unit is a group of unit_r planes of one column starting at unit_e (mod unit_r)
  int unit_r, unit_e2, units_per_col, q_units, ncol, gx;
  int ns, nrs;  // ring slots: planes of v / planes of the residual (COMBINE)
  int segs;     // > 0: every column is cut into `segs` equal segments, one CTA each (neighbouring
                // columns then march in step and share their halos through L2); 0: contiguous
                // ranges of q_units units in (column, unit) order
  int pfd;      // planes prefetched into L2 ahead of the ring's issue front
  int l2_stream, l2_keep;  // L2 eviction priority (L2_*) of the vectors without reuse (x loads and
                           // every store) and of the residual ring's source
  int l2_v;                // ... and of the TMA tiles of v (their halos are re-read by neighbours)
  int to;       // rows of a tile that are OUTPUT (<= 8 RPT; the threads of the other rows idle): a
                // shorter tile makes more columns, so that columns x segments fills the CTA slots
  const float *v;
  float *out;
  const float *b;
  float *r;
  float *p;
  int update_p;
  float *p_out;
  float *xup;
  const float *acc;  // optional volume added to A v (other observations' terms, general path)
  const int *done;
  GridReduce gr;
  FinalizeArgs fin;
};

typedef void (*FastKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const FastArgs);

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
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
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// the same with an L2 eviction priority (l2_policy, common.cuh)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ float &cmp(float4 &q, int k) {
  return k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w));
}
__device__ __forceinline__ float cmpv(const float4 &q, int k) {
  return k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w));
}
__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

// MODE: LhsMode.  KIND: FK_*.  KP taps / R ratio of the thick axis (1/1 otherwise).
// E = off mod R (thick along z only).  RPT: rows per thread (tile = 8 RPT rows x 128 z).
template <int MODE, int KIND, int KP, int R, int E, int RPT>
__global__ void __launch_bounds__(NTHR, RPT == 1 ? 3 : 2)
    lhs_fast_kernel(const __grid_constant__ CUtensorMap tmap_v,
                    const __grid_constant__ CUtensorMap tmap_r,
                    const __grid_constant__ CUtensorMap tmap_x,
                    const __grid_constant__ FastArgs a) {
  constexpr int TO = NWARP * RPT;
  constexpr int HZ = fast_hz(KIND, KP);
  constexpr int SZ = TZ + 2 * HZ;  // floats per tile row in shared memory
  constexpr uint32_t ROWB = SZ * 4u;
  constexpr uint32_t PLANE_BYTES = (TO + 2) * ROWB;
  constexpr uint32_t PLANE_B = (PLANE_BYTES + 127u) / 128u * 128u;
  constexpr bool THICK_M = KIND == FK_THICK_M;
  constexpr bool THICK_Z = KIND == FK_THICK_Z;
  constexpr bool THICK_ZG = KIND == FK_THICK_ZG;
  constexpr int NW = THICK_ZG ? (KP + 2) / R + 1 : 1;  // candidate windows per quad (generic z)
  constexpr int L = THICK_M ? KP - 1 : 1;    // look-ahead planes of the march
  constexpr int B = THICK_M ? KP - 1 : 0;    // pre-roll planes before the first output plane
  constexpr int PRE = THICK_M ? 0 : 1;       // plane u_begin - 1 needed (D'D of the first plane)
  // COMBINE: both fused modes (an in-place update of the tile + halo from a second TMA ring)
  constexpr bool COMBINE = MODE == LHS_COMBINE || MODE == LHS_ECOMBINE;
  constexpr bool ECOMB = MODE == LHS_ECOMBINE;
  constexpr bool TERM = MODE == LHS_TERM;  // term only: no stencil, no dot product
  constexpr int NT = THICK_M ? (KP + R - 1) / R : 1;  // low-res rows alive at a plane (thick m)
  static_assert(NT <= 5, "thick-m: at most five live low-res rows");
  static_assert(!THICK_ZG || KP - 1 <= HZ, "thick-z: windows inside the z halo");
  static_assert(!THICK_M || (R - 1) + (NT - 1) * R < kTaps, "kerT zero padding");
  static_assert(!THICK_Z || (4 % R == 0 && KP - 1 <= HZ), "thick-z: ratio divides the quad");
  constexpr int NQ = HZ / 4;  // halo quads each side of the own quad
  // THICK_Z: candidate low-res rows per quad = windows starting in [z - 4 NQ, z + 3]
  constexpr int NRZ = THICK_Z ? 4 * (NQ + 1) / R : 1;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_red[kMaxWarps];
  __shared__ uint64_t s_bar[kMaxSlots];
  if (a.done && *a.done) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = warp * RPT;
  const uint32_t ring_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t ring_end = ring_base + (uint32_t)a.ns * PLANE_B;
  const uint32_t rring_base = ring_end;
  const uint32_t rring_end = rring_base + (uint32_t)a.nrs * PLANE_B;
  const uint32_t bar_base = smem_u32(s_bar);
  const uint32_t bar_end = bar_base + 8u * (uint32_t)a.ns;
  const uint32_t own_b = (uint32_t)((row0 + 1) * SZ + HZ + 4 * lane) * 4u;
  const uint32_t box_bytes = (uint32_t)(a.to + 2) * ROWB;  // what one TMA box delivers

  if (tid == 0) {
    for (int s = 0; s < a.ns; ++s) mbar_init(bar_base + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t pol_s = l2_policy(a.l2_stream), pol_k = l2_policy(a.l2_keep);
  const uint64_t pol_v = l2_policy(a.l2_v);
  float beta_c = 0.f, alpha_c = 0.f;
  if (COMBINE) {
    beta_c = (float)a.fin.st->beta;
    alpha_c = (float)a.fin.st->alpha;
  }
  const bool x_fused = MODE == LHS_COMBINE && a.xup != nullptr;  // x += alpha p_old rides along
  float beta_e = 0.f;
  if (MODE == LHS_ENERGY && a.update_p) beta_e = (float)a.fin.st->beta;

  // halo quad of this thread (COMBINE): rows 0 and TO+1 entirely, first/last quad of the others
  constexpr int SZ4 = SZ / 4, HQ = HZ / 4;
  constexpr int HALO_N = 2 * SZ4 + 2 * HQ * TO;
  static_assert(HALO_N <= NTHR, "one halo quad per thread");
  uint32_t h_off = 0;
  const bool h_has = COMBINE && tid < HALO_N;
  if (COMBINE) {
    int h_row, h_c4;
    if (tid < SZ4) {
      h_row = 0;
      h_c4 = tid;
    } else if (tid < 2 * SZ4) {
      h_row = TO + 1;
      h_c4 = tid - SZ4;
    } else {
      const int k = tid - 2 * SZ4;
      h_row = 1 + k / (2 * HQ);
      const int sq = k - (h_row - 1) * (2 * HQ);
      h_c4 = sq < HQ ? sq : SZ4 - 2 * HQ + sq;
    }
    h_off = (uint32_t)(h_row * SZ + 4 * h_c4) * 4u;
  }

  auto adv = [&](uint32_t s) {
    s += PLANE_B;
    return s == ring_end ? ring_base : s;
  };
  auto ustart = [&](int k) {
    if (k <= 0) return 0;
    const int m = a.unit_e2 + a.unit_r * (k - 1);
    return m < a.nm ? m : a.nm;
  };

  const long long total_units = (long long)a.ncol * a.units_per_col;
  long long t = (long long)blockIdx.x * a.q_units;
  long long t_end = t + a.q_units < total_units ? t + a.q_units : total_units;
  if (a.segs > 0) {  // one segment of one column
    const int c = blockIdx.x / a.segs, sg = blockIdx.x - c * a.segs;
    t = (long long)c * a.units_per_col + (long long)sg * a.units_per_col / a.segs;
    t_end = (long long)c * a.units_per_col + (long long)(sg + 1) * a.units_per_col / a.segs;
  }

  // ring positions persist across the column segments of this CTA
  uint32_t arr_pa = ring_base, arr_ba = bar_base, arr_par = 0u, arr_ra = rring_base;
  uint32_t ip_pa = ring_base, ip_ba = bar_base, ip_ra = rring_base;
  double part = 0.0;

  while (t < t_end) {
    const int col = (int)(t / a.units_per_col);
    const int k0 = (int)(t - (long long)col * a.units_per_col);
    const long long left = t_end - t;
    const int k1 = left < (long long)(a.units_per_col - k0) ? k0 + (int)left : a.units_per_col;
    t += k1 - k0;
    const int m0 = ustart(k0), m1 = ustart(k1);
    const int cz = col % a.gx, co = col / a.gx;
    const int z0 = cz * TZ, o0 = co * a.to;
    const int o_first = o0 + row0, z = z0 + 4 * lane;

    const int u_begin = m0 - B;
    const int first = u_begin - PRE;
    int last = m1 - 1 + L;
    // cut at a row start: the last row of the segment starts at m1 - R and ends at m1 - R + KP - 1
    if (THICK_M && m1 < a.nm) last = m1 + (KP - 1 > R ? KP - 1 - R : 0);
    const int u_start = first - L - 1;    // first (virtual) trip: arrives plane `first`

    // ---- per-thread constants of this column ----
    bool act[RPT];
    const bool ztail = z + 4 > a.nz;  // padded rows (nz % 4 != 0): keep the pad voxels zero
    float4 Dq[RPT];
    float4 tm[RPT];        // POINT: tau x FOV mask per voxel; THICK_M: FOV mask of the row quad
    float czr[RPT][NRZ];   // THICK_Z: validity x scaling of the candidate low-res rows
    // THICK_ZG: per-lane candidate windows (start offset in bytes from the own quad, weights of
    // the four output voxels, validity x scaling per row)
    int g_so[NW];
    float g_w[NW][4], g_c[RPT][NW];
    if (THICK_ZG) {
      const int lo = z - (KP - 1);
      const int s0 = lo + ((((a.off - lo) % R) + R) % R);
#pragma unroll
      for (int n = 0; n < NW; ++n) {
        const int sn = s0 + n * R;
        const int j = (sn - a.off) / R;  // exact: sn = off (mod R); may be negative
        const bool ok = sn <= z + 3 && sn - a.off >= 0 && j < a.nj;
        g_so[n] = (sn - z) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int tap = z + k - sn;
          g_w[n][k] = (ok && tap >= 0 && tap < KP) ? a.kerT[tap] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int o = o_first + i;
          const bool o_in = o >= a.lo_o && o < a.hi_o;
          g_c[i][n] = (ok && o_in) ? (a.scl_conv ? ((j & 1) ? a.s_odd : a.s_even) : 1.f) : 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int o = o_first + i;
      act[i] = z < a.nz && o < a.no && row0 + i < a.to;
      const float dz0 = a.d0 - (o == 0 ? a.a_o : 0.f);
      Dq[i] = make_float4(dz0 - (z == 0 ? a.a_z : 0.f), dz0, dz0, dz0);
      const bool o_in = o >= a.lo_o && o < a.hi_o;
      tm[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (KIND == FK_POINT || THICK_M) {
        const float f = o_in ? (KIND == FK_POINT ? a.tau : 1.f) : 0.f;
        tm[i].x = (z + 0 >= a.lo_z && z + 0 < a.hi_z) ? f : 0.f;
        tm[i].y = (z + 1 >= a.lo_z && z + 1 < a.hi_z) ? f : 0.f;
        tm[i].z = (z + 2 >= a.lo_z && z + 2 < a.hi_z) ? f : 0.f;
        tm[i].w = (z + 3 >= a.lo_z && z + 3 < a.hi_z) ? f : 0.f;
      }
#pragma unroll
      for (int n = 0; n < NRZ; ++n) {
        czr[i][n] = 0.f;
        if (THICK_Z) {
          // window n starts at z - 4 NQ + E + n R  ( = R j + off )
          const int s = z - 4 * NQ + E + n * R - a.off;
          const int j = s / R;  // exact when s >= 0
          if (o_in && s >= 0 && j < a.nj)
            czr[i][n] = a.scl_conv ? ((j & 1) ? a.s_odd : a.s_even) : 1.f;
        }
      }
    }

    // ---- producer ----
    __syncthreads();  // every thread is done with the previous segment's slots
    int iq = first, pq = first;
    auto issue = [&](int limit) {
      while (iq <= last && iq < limit) {
        if (COMBINE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(ip_ba, (COMBINE ? 2u : 1u) * box_bytes);
        if (a.march_y) {
          tma_load_3d(ip_pa, &tmap_v, ip_ba, z0 - HZ, iq, o0 - 1, pol_v);
          if (COMBINE) tma_load_3d(ip_ra, &tmap_r, ip_ba, z0 - HZ, iq, o0 - 1, pol_k);
        } else {
          tma_load_3d(ip_pa, &tmap_v, ip_ba, z0 - HZ, o0 - 1, iq, pol_v);
          if (COMBINE) tma_load_3d(ip_ra, &tmap_r, ip_ba, z0 - HZ, o0 - 1, iq, pol_k);
        }
        ip_pa += PLANE_B;
        ip_ba += 8u;
        if (ip_pa == ring_end) {
          ip_pa = ring_base;
          ip_ba = bar_base;
        }
        if (COMBINE) {
          ip_ra += PLANE_B;
          if (ip_ra == rring_end) ip_ra = rring_base;
        }
        ++iq;
      }
      // L2 prefetch of the planes the ring cannot hold yet
      if (pq < iq) pq = iq;
      while (pq <= last && pq < iq + a.pfd) {
        const int c1 = a.march_y ? pq : o0 - 1, c2 = a.march_y ? o0 - 1 : pq;
        tma_prefetch_3d(&tmap_v, z0 - HZ, c1, c2);
        if (COMBINE) {
          tma_prefetch_3d(&tmap_r, z0 - HZ, c1, c2);
          if (x_fused && pq >= m0 && pq < m1)
            tma_prefetch_3d(&tmap_x, z0, a.march_y ? pq : o0, a.march_y ? o0 : pq);
        }
        ++pq;
      }
    };
    if (tid == 0) issue(u_start + a.ns);

    // ---- consumer state ----
    uint32_t sa_cur = PRE ? adv(arr_pa) : arr_pa;  // slot of plane u_begin
    int goff_c = first * a.gs_m + o_first * a.gs_o + z;    // plane being combined
    int goff_u = u_begin * a.gs_m + o_first * a.gs_o + z;  // plane being output
    float4 prev[RPT], cur[RPT];
    float4 lr[NT][RPT];  // THICK_M: the live low-res rows, newest first (registers)
    float4 xr[2][RPT];          // COMBINE: x quads of the next two planes to combine
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      prev[i] = cur[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NT; ++n) lr[n][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      xr[0][i] = xr[1][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (x_fused) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = first + h;
        if (q >= m0 && q < m1) {
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            if (act[i])
              xr[h][i] = ldg_hint4(a.xup + (goff_c + h * a.gs_m + i * a.gs_o), pol_s);
        }
      }
    }
    int ph = 0, jrow = 0;
    if (THICK_M) {
      const int t0 = u_begin - a.off;
      jrow = floordiv(t0, R);
      ph = t0 - jrow * R;
    }

    for (int u = u_start; u < m1; u += 2) {
      // ---- arrive (and combine) planes u + L + 1, u + L + 2 ----
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = u + L + 1 + h;
        if (c <= last) {
          mbar_wait(arr_ba, arr_par);
          // the first two planes of the segment seed the register pipeline of the stencil
          const bool cap_prev = PRE && c == first, cap_cur = c == u_begin;
          if (!COMBINE && (cap_prev || cap_cur)) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const float4 q = lds128(arr_pa + own_b + i * ROWB);
              if (cap_prev) prev[i] = q;
              if (cap_cur) cur[i] = q;
            }
          }
          if (COMBINE) {
            const bool c_own = c >= m0 && c < m1;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              const uint32_t sa = arr_pa + own_b + i * ROWB;
              const float4 po = lds128(sa);
              const float4 rr = lds128(arr_ra + own_b + i * ROWB);
              float4 pn;
              if (ECOMB) {  // torch: x += alpha * p  (two roundings)
                pn.x = __fadd_rn(po.x, __fmul_rn(alpha_c, rr.x));
                pn.y = __fadd_rn(po.y, __fmul_rn(alpha_c, rr.y));
                pn.z = __fadd_rn(po.z, __fmul_rn(alpha_c, rr.z));
                pn.w = __fadd_rn(po.w, __fmul_rn(alpha_c, rr.w));
              } else {  // torch: p *= beta; p += r  (two roundings)
                pn.x = __fadd_rn(__fmul_rn(beta_c, po.x), rr.x);
                pn.y = __fadd_rn(__fmul_rn(beta_c, po.y), rr.y);
                pn.z = __fadd_rn(__fmul_rn(beta_c, po.z), rr.z);
                pn.w = __fadd_rn(__fmul_rn(beta_c, po.w), rr.w);
              }
              sts128(sa, pn);
              if (cap_prev) prev[i] = pn;
              if (cap_cur) cur[i] = pn;
              if (c_own && act[i]) {
                const int gi = goff_c + i * a.gs_o;
                stg_hint4(a.p_out + gi, pn, pol_s);
              }
              if (x_fused && c_own && act[i]) {
                const int gi = goff_c + i * a.gs_o;
                float4 xn;  // the previous iteration's x += alpha p
                xn.x = __fadd_rn(xr[h][i].x, __fmul_rn(alpha_c, po.x));
                xn.y = __fadd_rn(xr[h][i].y, __fmul_rn(alpha_c, po.y));
                xn.z = __fadd_rn(xr[h][i].z, __fmul_rn(alpha_c, po.z));
                xn.w = __fadd_rn(xr[h][i].w, __fmul_rn(alpha_c, po.w));
                stg_hint4(a.xup + gi, xn, pol_s);
              }
            }
            if (h_has) {
              const float4 po = lds128(arr_pa + h_off);
              const float4 rr = lds128(arr_ra + h_off);
              float4 pn;
              if (ECOMB) {
                pn.x = __fadd_rn(po.x, __fmul_rn(alpha_c, rr.x));
                pn.y = __fadd_rn(po.y, __fmul_rn(alpha_c, rr.y));
                pn.z = __fadd_rn(po.z, __fmul_rn(alpha_c, rr.z));
                pn.w = __fadd_rn(po.w, __fmul_rn(alpha_c, rr.w));
              } else {
                pn.x = __fadd_rn(__fmul_rn(beta_c, po.x), rr.x);
                pn.y = __fadd_rn(__fmul_rn(beta_c, po.y), rr.y);
                pn.z = __fadd_rn(__fmul_rn(beta_c, po.z), rr.z);
                pn.w = __fadd_rn(__fmul_rn(beta_c, po.w), rr.w);
              }
              sts128(arr_pa + h_off, pn);
            }
            // x quads of plane c + 2 (combined by the next trip)
            const int q2 = c + 2;
            const bool q_own = q2 >= m0 && q2 < m1;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              xr[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (x_fused && q_own && act[i])
                xr[h][i] = ldg_hint4(a.xup + (goff_c + 2 * a.gs_m + i * a.gs_o), pol_s);
            }
            arr_ra += PLANE_B;
            if (arr_ra == rring_end) arr_ra = rring_base;
          }
          arr_pa += PLANE_B;
          arr_ba += 8u;
          if (arr_pa == ring_end) {
            arr_pa = ring_base;
            arr_ba = bar_base;
            arr_par ^= 1u;
          }
          goff_c += a.gs_m;
        }
      }
      __syncthreads();
      // every thread is done with the planes before u: refill their slots
      if (tid == 0) issue(u + a.ns);

      // ---- stencil: planes u, u + 1 ----
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int uu = u + h;
        if (uu < u_begin || uu >= m1) continue;
        const uint32_t an = adv(sa_cur);
        float4 next[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) next[i] = lds128(an + own_b + i * ROWB);

        if (THICK_M) {
          if (ph == 0) {  // low-res row jrow starts at this plane: form this thread's quad of it
            const bool jv = jrow >= 0 && jrow < a.nj;
            float sc = jv ? 1.f : 0.f;
            if (a.scl_conv && jv) sc = (jrow & 1) ? a.s_odd : a.s_even;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
              float4 acc;
              acc.x = fmaf(a.ker[1], next[i].x, a.ker[0] * cur[i].x);
              acc.y = fmaf(a.ker[1], next[i].y, a.ker[0] * cur[i].y);
              acc.z = fmaf(a.ker[1], next[i].z, a.ker[0] * cur[i].z);
              acc.w = fmaf(a.ker[1], next[i].w, a.ker[0] * cur[i].w);
              uint32_t s = an;
#pragma unroll
              for (int tt = 2; tt < KP; ++tt) {
                s = adv(s);
                const float4 q = lds128(s + own_b + i * ROWB);
                acc.x = fmaf(a.ker[tt], q.x, acc.x);
                acc.y = fmaf(a.ker[tt], q.y, acc.y);
                acc.z = fmaf(a.ker[tt], q.z, acc.z);
                acc.w = fmaf(a.ker[tt], q.w, acc.w);
              }
#pragma unroll
              for (int n = NT - 1; n > 0; --n) lr[n][i] = lr[n - 1][i];
              lr[0][i].x = acc.x * (sc * tm[i].x);
              lr[0][i].y = acc.y * (sc * tm[i].y);
              lr[0][i].z = acc.z * (sc * tm[i].z);
              lr[0][i].w = acc.w * (sc * tm[i].w);
            }
          }
        }

        if (uu >= m0) {
          if (!TERM && uu == 0) {  // low edge of the march axis: (c - hi), i.e. lo := c
#pragma unroll
            for (int i = 0; i < RPT; ++i) prev[i] = cur[i];
          }
          float4 om_edge = make_float4(0.f, 0.f, 0.f, 0.f), op_edge = om_edge;
          if (!TERM) {
            om_edge = lds128(sa_cur + own_b - ROWB);
            op_edge = lds128(sa_cur + own_b + RPT * ROWB);
          }
          const bool m_in = uu >= a.lo_m && uu < a.hi_m;
          float wt[NT];  // tap of live row n at this plane: kerT[ph + n R] (zero padded)
#pragma unroll
          for (int n = 0; n < NT; ++n) wt[n] = THICK_M ? a.kerT[ph + n * R] : 0.f;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const uint32_t ca = sa_cur + own_b + i * ROWB;
            const float4 om = i == 0 ? om_edge : cur[i > 0 ? i - 1 : 0];
            const float4 op = i == RPT - 1 ? op_edge : cur[i < RPT - 1 ? i + 1 : 0];
            float zl, zr;
            float4 dat = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 Dk = Dq[i];
            if (THICK_Z) {
              const float4 ql = lds128(ca - 16u);
              const float4 qr = lds128(ca + 16u);
              zl = ql.w;
              zr = qr.x;
              if (m_in) {
                // the 2 NQ + 1 quads around the own one, in registers
                float V[4 * (2 * NQ + 1)];
                if constexpr (NQ == 2) {
                  const float4 ql2 = lds128(ca - 32u);
                  const float4 qr2 = lds128(ca + 32u);
                  V[0] = ql2.x, V[1] = ql2.y, V[2] = ql2.z, V[3] = ql2.w;
                  V[16] = qr2.x, V[17] = qr2.y, V[18] = qr2.z, V[19] = qr2.w;
                }
                constexpr int VB = 4 * (NQ - 1);
                V[VB + 0] = ql.x, V[VB + 1] = ql.y, V[VB + 2] = ql.z, V[VB + 3] = ql.w;
                V[VB + 4] = cur[i].x, V[VB + 5] = cur[i].y, V[VB + 6] = cur[i].z;
                V[VB + 7] = cur[i].w;
                V[VB + 8] = qr.x, V[VB + 9] = qr.y, V[VB + 10] = qr.z, V[VB + 11] = qr.w;
#pragma unroll
                for (int n = 0; n < NRZ; ++n) {
                  const int s0 = E + n * R;  // window start relative to z - 4 NQ
                  // cannot touch the own quad
                  if (s0 + KP - 1 < 4 * NQ || s0 > 4 * NQ + 3) continue;
                  float lr = 0.f;
#pragma unroll
                  for (int tt = 0; tt < KP; ++tt)
                    if (s0 + tt < 4 * (2 * NQ + 1)) lr = fmaf(a.ker[tt], V[s0 + tt], lr);
                  lr *= czr[i][n];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int tap = k + 4 * NQ - s0;
                    if (tap >= 0 && tap < KP) cmp(dat, k) = fmaf(a.kerT[tap], lr, cmpv(dat, k));
                  }
                }
              }
            } else {
              zl = zr = 0.f;
              if (!TERM) {
                zl = lds32(ca - 4u);
                zr = lds32(ca + 16u);
              }
              if (THICK_ZG) {
                if (m_in) {
#pragma unroll
                  for (int n = 0; n < NW; ++n) {
                    if (g_c[i][n] == 0.f) continue;
                    const uint32_t wa = ca + (uint32_t)g_so[n];
                    float lr = 0.f;
#pragma unroll
                    for (int tt = 0; tt < KP; ++tt) lr = fmaf(a.ker[tt], lds32(wa + 4u * tt), lr);
                    lr *= g_c[i][n];
                    dat.x = fmaf(g_w[n][0], lr, dat.x);
                    dat.y = fmaf(g_w[n][1], lr, dat.y);
                    dat.z = fmaf(g_w[n][2], lr, dat.z);
                    dat.w = fmaf(g_w[n][3], lr, dat.w);
                  }
                }
              } else if (THICK_M) {
                dat.x = wt[0] * lr[0][i].x;
                dat.y = wt[0] * lr[0][i].y;
                dat.z = wt[0] * lr[0][i].z;
                dat.w = wt[0] * lr[0][i].w;
#pragma unroll
                for (int n = 1; n < NT; ++n) {
                  dat.x = fmaf(wt[n], lr[n][i].x, dat.x);
                  dat.y = fmaf(wt[n], lr[n][i].y, dat.y);
                  dat.z = fmaf(wt[n], lr[n][i].z, dat.z);
                  dat.w = fmaf(wt[n], lr[n][i].w, dat.w);
                }
              } else if (KIND == FK_POINT) {
                if (m_in) {
                  Dk.x += tm[i].x;
                  Dk.y += tm[i].y;
                  Dk.z += tm[i].z;
                  Dk.w += tm[i].w;
                }
              }
            }
            float val[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (TERM) {
                val[k] = cmpv(dat, k);
                continue;
              }
              const float c = cmpv(cur[i], k);
              const float lft = k == 0 ? zl : cmpv(cur[i], k - 1);
              const float rgt = k == 3 ? zr : cmpv(cur[i], k + 1);
              const float s_m = cmpv(prev[i], k) + cmpv(next[i], k);
              const float s_o = cmpv(om, k) + cmpv(op, k);
              const float s_z = lft + rgt;
              const float S = fmaf(a.a_z, s_z, fmaf(a.a_o, s_o, fmaf(a.a_m, s_m, a.nd0l * c)));
              val[k] = fmaf(cmpv(Dk, k), c, cmpv(dat, k)) - S;
            }
            if (ztail) {
#pragma unroll
              for (int k = 1; k < 4; ++k)
                if (z + k >= a.nz) val[k] = 0.f;
            }
            if (act[i]) {
              const int gi = goff_u + i * a.gs_o;
              if (a.acc != nullptr) {  // may alias `out` (read before the store below)
                const float4 aq = *reinterpret_cast<const float4 *>(a.acc + gi);
                val[0] += aq.x;
                val[1] += aq.y;
                val[2] += aq.z;
                val[3] += aq.w;
              }
              if (TERM) {
                *reinterpret_cast<float4 *>(a.out + gi) =
                    make_float4(val[0], val[1], val[2], val[3]);
              } else if (MODE == LHS_PLAIN || MODE == LHS_COMBINE) {
                stg_hint4(a.out + gi, make_float4(val[0], val[1], val[2], val[3]), pol_s);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  part += (double)__fmul_rn(cmpv(cur[i], k), val[k]);
              } else if (MODE == LHS_RESID) {
                const float4 bq = *reinterpret_cast<const float4 *>(a.b + gi);
                float4 rr;
                rr.x = __fsub_rn(bq.x, val[0]);
                rr.y = __fsub_rn(bq.y, val[1]);
                rr.z = __fsub_rn(bq.z, val[2]);
                rr.w = __fsub_rn(bq.w, val[3]);
                *reinterpret_cast<float4 *>(a.r + gi) = rr;
                *reinterpret_cast<float4 *>(a.p + gi) = rr;
                part += (double)__fmul_rn(rr.x, rr.x) + (double)__fmul_rn(rr.y, rr.y) +
                        (double)__fmul_rn(rr.z, rr.z) + (double)__fmul_rn(rr.w, rr.w);
              } else {  // LHS_ENERGY / LHS_ECOMBINE
                const float4 bq = *reinterpret_cast<const float4 *>(a.b + gi);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  part += (double)__fmul_rn(__fsub_rn(val[k], 2.f * cmpv(bq, k)),
                                            cmpv(cur[i], k));
                if (MODE == LHS_ENERGY && a.update_p) {
                  const float4 rq = *reinterpret_cast<const float4 *>(a.r + gi);
                  const float4 pq = *reinterpret_cast<const float4 *>(a.p + gi);
                  float4 pn;
                  pn.x = __fadd_rn(__fmul_rn(beta_e, pq.x), rq.x);
                  pn.y = __fadd_rn(__fmul_rn(beta_e, pq.y), rq.y);
                  pn.z = __fadd_rn(__fmul_rn(beta_e, pq.z), rq.z);
                  pn.w = __fadd_rn(__fmul_rn(beta_e, pq.w), rq.w);
                  *reinterpret_cast<float4 *>(a.p + gi) = pn;
                }
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
        sa_cur = an;
        goff_u += a.gs_m;
        if (THICK_M) {
          if (++ph == R) {
            ph = 0;
            ++jrow;
          }
        }
      }
    }
  }
  if (TERM) return;
  double total_sum;
  if (grid_sum(part, a.gr, s_red, &total_sum) && tid == 0) finalize(a.fin, total_sum);
}

// kernel lookup, one translation unit per mode (lhs_fast_m*.cu)
FastKernel fast_lookup_plain(int kind, int kp, int r, int e, int rpt);
FastKernel fast_lookup_resid(int kind, int kp, int r, int e, int rpt);
FastKernel fast_lookup_energy(int kind, int kp, int r, int e, int rpt);
FastKernel fast_lookup_combine(int kind, int kp, int r, int e, int rpt);
FastKernel fast_lookup_ecombine(int kind, int kp, int r, int e, int rpt);
FastKernel fast_lookup_term(int kind, int kp, int r, int e, int rpt);

#define UR_FAST_LOOKUP_BODY(MODE)                                                          \
  {                                                                                        \
    const int R1 = rpt == 1;                                                               \
    switch (kind) {                                                                        \
      case FK_NONE:                                                                        \
        return R1 ? lhs_fast_kernel<MODE, FK_NONE, 1, 1, 0, 1>                             \
                  : lhs_fast_kernel<MODE, FK_NONE, 1, 1, 0, 2>;                            \
      case FK_POINT:                                                                       \
        return R1 ? lhs_fast_kernel<MODE, FK_POINT, 1, 1, 0, 1>                            \
                  : lhs_fast_kernel<MODE, FK_POINT, 1, 1, 0, 2>;                           \
      case FK_THICK_M:                                                                     \
        if (kp == 5 && r == 4)                                                             \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_M, 5, 4, 0, 1>                        \
                    : lhs_fast_kernel<MODE, FK_THICK_M, 5, 4, 0, 2>;                       \
        if (kp == 3 && r == 2)                                                             \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_M, 3, 2, 0, 1>                        \
                    : lhs_fast_kernel<MODE, FK_THICK_M, 3, 2, 0, 2>;                       \
        if (!R1) return nullptr;                                                           \
        if (kp == 5 && r == 3) return lhs_fast_kernel<MODE, FK_THICK_M, 5, 3, 0, 1>;       \
        if (kp == 7 && r == 5) return lhs_fast_kernel<MODE, FK_THICK_M, 7, 5, 0, 1>;       \
        if (kp == 7 && r == 6) return lhs_fast_kernel<MODE, FK_THICK_M, 7, 6, 0, 1>;       \
        if (kp == 9 && r == 8) return lhs_fast_kernel<MODE, FK_THICK_M, 9, 8, 0, 1>;       \
        if (kp == 9 && r == 2) return lhs_fast_kernel<MODE, FK_THICK_M, 9, 2, 0, 1>;       \
        return nullptr;                                                                    \
      case FK_THICK_ZG:                                                                    \
        if (kp == 5 && r == 3)                                                             \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_ZG, 5, 3, 0, 1>                       \
                    : lhs_fast_kernel<MODE, FK_THICK_ZG, 5, 3, 0, 2>;                      \
        if (kp == 7 && r == 5)                                                             \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_ZG, 7, 5, 0, 1>                       \
                    : lhs_fast_kernel<MODE, FK_THICK_ZG, 7, 5, 0, 2>;                      \
        if (kp == 7 && r == 6)                                                             \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_ZG, 7, 6, 0, 1>                       \
                    : lhs_fast_kernel<MODE, FK_THICK_ZG, 7, 6, 0, 2>;                      \
        if (kp == 9 && r == 8)                                                             \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_ZG, 9, 8, 0, 1>                       \
                    : lhs_fast_kernel<MODE, FK_THICK_ZG, 9, 8, 0, 2>;                      \
        return nullptr;                                                                    \
      case FK_THICK_Z:                                                                     \
        if (kp == 5 && r == 4) {                                                           \
          switch (e) {                                                                     \
            case 0:                                                                        \
              return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 0, 1>                    \
                        : lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 0, 2>;                   \
            case 1:                                                                        \
              return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 1, 1>                    \
                        : lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 1, 2>;                   \
            case 2:                                                                        \
              return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 2, 1>                    \
                        : lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 2, 2>;                   \
            default:                                                                       \
              return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 3, 1>                    \
                        : lhs_fast_kernel<MODE, FK_THICK_Z, 5, 4, 3, 2>;                   \
          }                                                                                \
        }                                                                                  \
        if (kp == 3 && r == 2) {                                                           \
          if (e == 0)                                                                      \
            return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 3, 2, 0, 1>                      \
                      : lhs_fast_kernel<MODE, FK_THICK_Z, 3, 2, 0, 2>;                     \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 3, 2, 1, 1>                        \
                    : lhs_fast_kernel<MODE, FK_THICK_Z, 3, 2, 1, 2>;                       \
        }                                                                                  \
        if (kp == 9 && r == 2) { /* two-quad halo, 5 quads in registers */                 \
          if (e == 0)                                                                      \
            return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 9, 2, 0, 1>                      \
                      : lhs_fast_kernel<MODE, FK_THICK_Z, 9, 2, 0, 2>;                     \
          return R1 ? lhs_fast_kernel<MODE, FK_THICK_Z, 9, 2, 1, 1>                        \
                    : lhs_fast_kernel<MODE, FK_THICK_Z, 9, 2, 1, 2>;                       \
        }                                                                                  \
        return nullptr;                                                                    \
    }                                                                                      \
    return nullptr;                                                                        \
  }

}  // namespace fast
}  // namespace ur

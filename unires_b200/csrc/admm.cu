// Joint-total-variation prox (z-update) and dual update (w-update) of the ADMM
// iteration, unires/_update.py:160-193, in one pass over all
// channels, plus the prior energy field of _compute_nll (unires/_update.py:
// 419-425).  Per voxel and channel c (forward differences, zero past the high
// edge, A.4):
//   g_cd = lam_c * (y_c[i+e_d] - y_c[i]) / vx_d
//          (alpha != 1, over/under-relaxation: g_cd = alpha g_cd + (1-alpha) z_cd_old)
//   u_cd = w_cd / rho + g_cd
//   s    = sqrt(sum_c sum_d u_cd^2);  f = max(s - 1/rho, 0) / (s + 1e-7)
//   z_cd = f * u_cd;  w_cd += rho * (g_cd - z_cd)
// w_cd / rho is evaluated as w_cd * (1 / rho) with the float32 reciprocal formed once on the
// host (within 1 ulp of the IEEE quotient, 6e-8 relative against the 1e-4 parity bar): the nine
// IEEE divisions per voxel were ~20 % of the instructions of a kernel that ncu shows
// latency / issue bound at 16 warps per SM (round 2), and with a one-FMA u the vector variant
// no longer has to keep u in registers.
#include "common.cuh"

namespace ur {

struct ChannelPtrs {
  const float *y[UR_MAX_CHANNELS];
  float lam[UR_MAX_CHANNELS];
};

struct JtvGeom {
  int nx, ny, nz;
  float ivx, ivy, ivz;
  float rho;
  float alpha;
  float irho;  // 1 / rho (float32)
};

__device__ __forceinline__ void scaled_grad(const float *__restrict__ y, float lam, size_t i,
                                            int x, int yy, int z, size_t sx, size_t sy,
                                            const JtvGeom &g, float (&o)[3]) {
  const float c = __ldg(y + i);
  const float xp = x + 1 < g.nx ? __ldg(y + i + sx) : 0.f;
  const float yp = yy + 1 < g.ny ? __ldg(y + i + sy) : 0.f;
  const float zp = z + 1 < g.nz ? __ldg(y + i + 1) : 0.f;
  o[0] = lam * ((xp - c) * g.ivx);
  o[1] = lam * ((yp - c) * g.ivy);
  o[2] = lam * ((zp - c) * g.ivz);
}

enum JtvMode { JTV_FUSED = 0, JTV_NORM2 = 1, JTV_APPLY = 2, JTV_PRIOR = 3 };

// CT > 0: channel count known at compile time, u kept in registers.
// CT = 0: run-time channel count, u recomputed in the second sweep.
template <int MODE, int CT>
__global__ void __launch_bounds__(256)
    jtv_kernel(ChannelPtrs ch, float *__restrict__ zz, float *__restrict__ w,
               float *__restrict__ nrm2, float *__restrict__ jtv, int n_channels, JtvGeom g,
               int accumulate) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (z >= g.nz || y >= g.ny) return;
  const size_t sy = g.nz, sx = (size_t)g.ny * g.nz, n = sx * g.nx;
  const size_t i = x * sx + y * sy + z;
  const int C = CT > 0 ? CT : n_channels;
  constexpr int KEEP = CT > 0 ? CT : 1;
  float gkeep[KEEP][3], ukeep[KEEP][3];

  float s2 = 0.f;
  if (MODE != JTV_APPLY) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float gr[3];
      scaled_grad(ch.y[c], ch.lam[c], i, x, y, z, sx, sy, g, gr);
      float e = 0.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (MODE != JTV_PRIOR && g.alpha != 1.f)
          gr[d] = __fadd_rn(__fmul_rn(g.alpha, gr[d]),
                            __fmul_rn(__fsub_rn(1.f, g.alpha), zz[((size_t)c * 3 + d) * n + i]));
        float u = gr[d];
        if (MODE != JTV_PRIOR) u = __fadd_rn(__fmul_rn(w[((size_t)c * 3 + d) * n + i], g.irho), u);
        if (CT > 0) {
          gkeep[c][d] = gr[d];
          ukeep[c][d] = u;
        }
        e = d == 0 ? __fmul_rn(u, u) : __fadd_rn(e, __fmul_rn(u, u));
      }
      s2 = __fadd_rn(s2, e);
    }
    if (MODE == JTV_NORM2 || MODE == JTV_PRIOR) {
      nrm2[i] = accumulate ? __fadd_rn(nrm2[i], s2) : s2;
      return;
    }
  } else {
    s2 = nrm2[i];
  }
  const float s = sqrtf(s2);
  const float f = __fdiv_rn(fmaxf(__fsub_rn(s, g.irho), 0.f), __fadd_rn(s, 1e-7f));
  if (jtv) jtv[i] = f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float gr[3];
    if (!(CT > 0 && MODE == JTV_FUSED)) scaled_grad(ch.y[c], ch.lam[c], i, x, y, z, sx, sy, g, gr);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t q = ((size_t)c * 3 + d) * n + i;
      float gd, u;
      const float wv = w[q];
      if (CT > 0 && MODE == JTV_FUSED) {
        gd = gkeep[c][d];
        u = ukeep[c][d];
      } else {
        gd = gr[d];
        if (g.alpha != 1.f)
          gd = __fadd_rn(__fmul_rn(g.alpha, gd), __fmul_rn(__fsub_rn(1.f, g.alpha), zz[q]));
        u = __fadd_rn(__fmul_rn(wv, g.irho), gd);
      }
      const float zv = __fmul_rn(f, u);
      zz[q] = zv;
      w[q] = __fadd_rn(wv, __fmul_rn(g.rho, __fsub_rn(gd, zv)));
    }
  }
}


// ---------------------------------------------------------------------------
// Vector variant: one thread owns V (4 or 2) consecutive z voxels of one row, all channels.
// Every access to w / z / jtv is a 128- / 64-bit load or store and all loads of the first
// sweep are in flight together; the arithmetic per voxel is the scalar kernel's, in the same
// order (bit-identical results).  V is chosen so that at least two blocks stay resident per
// SM (loads of one block overlap the stores of another).  Needs nz % V == 0 and aligned volumes.
// ---------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void ldv(float (&o)[V], const float *p) {
  if (V == 4) {
    const float4 q = *reinterpret_cast<const float4 *>(p);
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[V - 1] = q.w;
  } else {
    const float2 q = *reinterpret_cast<const float2 *>(p);
    o[0] = q.x; o[V - 1] = q.y;
  }
}
template <int V>
__device__ __forceinline__ void ldgv(float (&o)[V], const float *p) {
  if (V == 4) {
    const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[V - 1] = q.w;
  } else {
    const float2 q = __ldg(reinterpret_cast<const float2 *>(p));
    o[0] = q.x; o[V - 1] = q.y;
  }
}
template <int V>
__device__ __forceinline__ void stv(float *p, const float (&o)[V]) {
  if (V == 4)
    *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[V - 1]);
  else
    *reinterpret_cast<float2 *>(p) = make_float2(o[0], o[V - 1]);
}

template <int MODE, int CT, int V>
__global__ void __launch_bounds__(256, 2)
    jtv_kernel_vec(ChannelPtrs ch, float *__restrict__ zz, float *__restrict__ w,
                   float *__restrict__ nrm2, float *__restrict__ jtv, JtvGeom g, int accumulate) {
  const int z = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.z;
  if (z >= g.nz || y >= g.ny) return;
  const size_t sy = g.nz, sx = (size_t)g.ny * g.nz, n = sx * g.nx;
  const size_t i = x * sx + y * sy + z;
  float gk[CT][3][V], wk[CT][3][V];
  float s2[V];
#pragma unroll
  for (int k = 0; k < V; ++k) s2[k] = 0.f;
  {
    // issue every load of the sweep first
    float yc[CT][V], yx[CT][V], yy[CT][V], yz[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float *yp = ch.y[c] + i;
      ldgv<V>(yc[c], yp);
#pragma unroll
      for (int k = 0; k < V; ++k) yx[c][k] = yy[c][k] = 0.f;
      if (x + 1 < g.nx) ldgv<V>(yx[c], yp + sx);
      if (y + 1 < g.ny) ldgv<V>(yy[c], yp + sy);
      yz[c] = z + V < g.nz ? __ldg(yp + V) : 0.f;
      if (MODE != JTV_PRIOR) {
#pragma unroll
        for (int d = 0; d < 3; ++d) ldv<V>(wk[c][d], w + ((size_t)c * 3 + d) * n + i);
      }
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const float lam = ch.lam[c];
      float zo[3][V];
      if (MODE != JTV_PRIOR && g.alpha != 1.f) {
#pragma unroll
        for (int d = 0; d < 3; ++d) ldv<V>(zo[d], zz + ((size_t)c * 3 + d) * n + i);
      }
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float cv = yc[c][k];
        const float zp = k < V - 1 ? yc[c][k < V - 1 ? k + 1 : k] : yz[c];
        float gr[3];
        gr[0] = lam * ((yx[c][k] - cv) * g.ivx);
        gr[1] = lam * ((yy[c][k] - cv) * g.ivy);
        gr[2] = lam * ((zp - cv) * g.ivz);
        float e = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (MODE != JTV_PRIOR && g.alpha != 1.f)
            gr[d] = __fadd_rn(__fmul_rn(g.alpha, gr[d]),
                              __fmul_rn(__fsub_rn(1.f, g.alpha), zo[d][k]));
          float u = gr[d];
          if (MODE != JTV_PRIOR) u = __fadd_rn(__fmul_rn(wk[c][d][k], g.irho), u);
          gk[c][d][k] = gr[d];
          e = d == 0 ? __fmul_rn(u, u) : __fadd_rn(e, __fmul_rn(u, u));
        }
        s2[k] = __fadd_rn(s2[k], e);
      }
    }
  }
  if (MODE == JTV_NORM2 || MODE == JTV_PRIOR) {
    if (accumulate) {
      float o[V];
      ldv<V>(o, nrm2 + i);
#pragma unroll
      for (int k = 0; k < V; ++k) s2[k] = __fadd_rn(o[k], s2[k]);
    }
    stv<V>(nrm2 + i, s2);
    return;
  }
  if (MODE == JTV_APPLY) ldv<V>(s2, nrm2 + i);
  float f[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float s = sqrtf(s2[k]);
    f[k] = __fdiv_rn(fmaxf(__fsub_rn(s, g.irho), 0.f), __fadd_rn(s, 1e-7f));
  }
  if (jtv) stv<V>(jtv + i, f);
#pragma unroll
  for (int c = 0; c < CT; ++c) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const size_t q = ((size_t)c * 3 + d) * n + i;
      float zv[V], wn[V];
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float uu = __fadd_rn(__fmul_rn(wk[c][d][k], g.irho), gk[c][d][k]);
        zv[k] = __fmul_rn(f[k], uu);
        wn[k] = __fadd_rn(wk[c][d][k], __fmul_rn(g.rho, __fsub_rn(gk[c][d][k], zv[k])));
      }
      stv<V>(zz + q, zv);
      stv<V>(w + q, wn);
    }
  }
}


int jtv_wide = 1;  // ur_tune("jtv_wide"): 128-bit accesses for 3-4 channels too (measured at
                   // 3x256^3: 64-bit 359 us, 128-bit 333 us = 0.94 of the HBM roofline)
int jtv_block_rows = 4;  // measured at 3x256^3: 8 rows 502 us, 4 rows 478 us, 2 rows 484 us

static bool ptr16(const void *p) { return p == nullptr || ((uintptr_t)p & 15u) == 0; }

static int fill(ChannelPtrs *ch, const float *const *y, const float *lam, int C) {
  UR_REQUIRE(C >= 1 && C <= UR_MAX_CHANNELS, "channel count %d not in [1,%d]", C, UR_MAX_CHANNELS);
  UR_REQUIRE(y && lam, "null channel array");
  for (int c = 0; c < UR_MAX_CHANNELS; ++c) {
    ch->y[c] = c < C ? y[c] : nullptr;
    ch->lam[c] = c < C ? lam[c] : 0.f;
    if (c < C) UR_REQUIRE(y[c] != nullptr, "channel %d: null volume", c);
  }
  return UR_OK;
}

static JtvGeom geom(const int32_t dim[3], const float vx[3], float rho, float alpha) {
  return JtvGeom{dim[0], dim[1], dim[2], 1.f / vx[0], 1.f / vx[1], 1.f / vx[2], rho, alpha,
                 1.f / rho};
}

template <int MODE>
static int launch(const float *const *y, float *z, float *w, float *nrm2, float *jtv, int C,
                  const float *lam, const int32_t dim[3], const float vx[3], float rho,
                  float alpha, int accumulate, cudaStream_t st) {
  UR_REQUIRE(dim && vx && dim[0] > 0 && dim[1] > 0 && dim[2] > 0, "bad dims");
  ChannelPtrs ch;
  int rc = fill(&ch, y, lam, C);
  if (rc) return rc;
  JtvGeom g = geom(dim, vx, rho, alpha);
  // vector path: whole quads (pairs for 3-4 channels: register budget for two resident
  // blocks), aligned volumes, channel count known at compile time
  bool vec = g.nz % 4 == 0 && C <= 4 && ptr16(z) && ptr16(w) && ptr16(nrm2) && ptr16(jtv);
  for (int c = 0; c < C; ++c) vec = vec && ptr16(y[c]);
  if (vec) {
    const bool wide = C <= 2 || MODE == JTV_PRIOR || jtv_wide;
    const int V = wide ? 4 : 2;
    const int by = jtv_block_rows;  // rows per block (tuning knob "jtv_rows": 8 | 4 | 2)
    dim3 vblock(32, by, 1), vgrid(div_up(g.nz, 32 * V), div_up(g.ny, by), g.nx);
#define UR_JTV_V(CT, VV)                                                                    \
  jtv_kernel_vec<MODE, CT, VV><<<vgrid, vblock, 0, st>>>(ch, z, w, nrm2, jtv, g, accumulate)
    switch (C) {
      case 1: UR_JTV_V(1, 4); break;
      case 2: UR_JTV_V(2, 4); break;
      case 3: if (wide) UR_JTV_V(3, 4); else UR_JTV_V(3, 2); break;
      default: if (wide) UR_JTV_V(4, 4); else UR_JTV_V(4, 2); break;
    }
#undef UR_JTV_V
    UR_LAUNCH_CHECK();
    return UR_OK;
  }
  dim3 block(64, 4, 1), grid(div_up(g.nz, 64), div_up(g.ny, 4), g.nx);
#define UR_JTV_CASE(CT)                                                                     \
  jtv_kernel<MODE, CT><<<grid, block, 0, st>>>(ch, z, w, nrm2, jtv, C, g, accumulate)
  if (MODE == JTV_FUSED) {
    switch (C) {
      case 1: UR_JTV_CASE(1); break;
      case 2: UR_JTV_CASE(2); break;
      case 3: UR_JTV_CASE(3); break;
      case 4: UR_JTV_CASE(4); break;
      default: UR_JTV_CASE(0); break;
    }
  } else {
    UR_JTV_CASE(0);
  }
#undef UR_JTV_CASE
  UR_LAUNCH_CHECK();
  return UR_OK;
}

}  // namespace ur

using namespace ur;

extern "C" int ur_jtv_prox(const float *const *d_y, float *d_z, float *d_w, float *d_jtv,
                           int n_channels, const float *lam, const int32_t dim[3],
                           const float vx[3], float rho, float alpha, ur_stream stream) {
  UR_REQUIRE(d_z && d_w, "ur_jtv_prox: null z/w");
  return launch<JTV_FUSED>(d_y, d_z, d_w, nullptr, d_jtv, n_channels, lam, dim, vx, rho, alpha, 0,
                           (cudaStream_t)stream);
}

extern "C" int ur_jtv_norm2(const float *const *d_y, const float *d_z, const float *d_w,
                            float *d_nrm2, int n_channels, const float *lam,
                            const int32_t dim[3], const float vx[3], float rho, float alpha,
                            int accumulate, ur_stream stream) {
  UR_REQUIRE(d_w && d_nrm2, "ur_jtv_norm2: null w/nrm2");
  UR_REQUIRE(alpha == 1.f || d_z, "ur_jtv_norm2: relaxation needs z");
  return launch<JTV_NORM2>(d_y, const_cast<float *>(d_z), const_cast<float *>(d_w), d_nrm2,
                           nullptr, n_channels, lam, dim, vx, rho, alpha, accumulate,
                           (cudaStream_t)stream);
}

extern "C" int ur_jtv_apply(const float *const *d_y, float *d_z, float *d_w, const float *d_nrm2,
                            float *d_jtv, int n_channels, const float *lam, const int32_t dim[3],
                            const float vx[3], float rho, float alpha, ur_stream stream) {
  UR_REQUIRE(d_z && d_w && d_nrm2, "ur_jtv_apply: null z/w/nrm2");
  return launch<JTV_APPLY>(d_y, d_z, d_w, const_cast<float *>(d_nrm2), d_jtv, n_channels, lam,
                           dim, vx, rho, alpha, 0, (cudaStream_t)stream);
}

extern "C" int ur_nll_prior_energy(const float *const *d_y, float *d_e, int n_channels,
                                   const float *lam, const int32_t dim[3], const float vx[3],
                                   int accumulate, ur_stream stream) {
  UR_REQUIRE(d_e, "ur_nll_prior_energy: null output");
  return launch<JTV_PRIOR>(d_y, nullptr, nullptr, d_e, nullptr, n_channels, lam, dim, vx, 1.f,
                           1.f, accumulate, (cudaStream_t)stream);
}

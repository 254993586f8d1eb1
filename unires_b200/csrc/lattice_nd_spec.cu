// Compile-time specialised, TMA-fed variants of the multi-axis lattice kernels (lattice_nd.cu).
//
// The generic kernels turned out INSTRUCTION bound (ncu, 512^3: 190 warp instructions per
// voxel, 4.8 % of them FFMA -- run-time tap loops, per-element index arithmetic and 4-byte
// cp.async copies).  Here the taps, ratios and tile shapes are template parameters, a tile's
// input box arrives with ONE cp.async.bulk.tensor.3d (zero fill outside the volume = the
// reference's bound 'zero'; the next tile's box is in flight while this one is processed) and
// every pass is a straight-line register computation:
//   nd_down_spec:  x pass (a thread owns a column of the box) -> y pass (a thread slides along
//                  y) -> z pass (lanes along z, 64-bit LDS) -> low-res store
//   nd_up_spec:    x pass -> y pass -> per quad: z pass + 7-point stencil + CG epilogue
// Instantiated for BASELINE configs[4] (rect-3 / gauss-9 / gauss-9 at ratio 2); anything else
// runs the generic kernels.
#include <cuda.h>
#include <string.h>

#include "lattice_nd.cuh"
#include "lhs_fast.cuh"

namespace ur {

bool box_tensor_map(const float *v, int n0, int n1, int n2, int b0, int b1, int b2,
                    CUtensorMap *out);  // lhs_stream.cu

using fast::mbar_expect_tx;
using fast::mbar_init;
using fast::mbar_wait;
using fast::smem_u32;
using fast::tma_load_3d;

// tile index -> (b0, b1, b2) through float reciprocals (exact for tile < 2^22, see fdiv_small in
// solver.cu): the integer divisions of the decode were 14 % of nd_down's instructions (ncu)
struct TileDec {
  int nt1, nt2;
  float inv1, inv2;
};
__device__ __forceinline__ void tile_decode(const TileDec &d, int tile, int &b0, int &b1, int &b2) {
  const int tq = (int)floorf(((float)tile + 0.5f) * d.inv2);
  b2 = tile - tq * d.nt2;
  b0 = (int)floorf(((float)tq + 0.5f) * d.inv1);
  b1 = tq - b0 * d.nt1;
}

constexpr int kSpecThreads = 256;
constexpr int kSpecWarps = kSpecThreads / 32;

struct SpecTaps {
  float k0[UR_MAX_TAPS], k1[UR_MAX_TAPS], k2[UR_MAX_TAPS];
};

// ---------------------------------------------------------------- down: v -> scale * A v
template <int K0, int R0, int K1, int R1, int K2, int R2>
struct DownCfg {
  static constexpr int L0 = 2, L1 = 8, L2 = 32;  // low-res tile (86 KB of shared memory: 2 CTAs/SM)
  static constexpr int I0 = (L0 - 1) * R0 + K0, I1 = (L1 - 1) * R1 + K1,
                       I2 = (L2 - 1) * R2 + K2;                  // input box
  static constexpr int I2P = (I2 + 3) / 4 * 4;                   // TMA inner extent (16 B)
  static constexpr int BOX = I0 * I1 * I2P;                      // floats per box
  static constexpr int BOXP = (BOX + 31) / 32 * 32;              // TMA destinations: 128-byte aligned
  // passes run x -> y -> z: the first pass sees the whole (halo-amplified) box, so it should
  // be the cheap one -- 3 taps along x here against 9 along z (z first cost 2.4x the
  // instructions: 7.1 N against 4.9 N FMAs plus their loads)
  static constexpr int T1 = L0 * I1 * I2P, T2 = L0 * L1 * I2P;
  static constexpr size_t SMEM = (size_t)(2 * BOXP + T1 + T2) * 4 + 128;
};

template <int K0, int R0, int K1, int R1, int K2, int R2>
__global__ void __launch_bounds__(kSpecThreads, 2)
    nd_down_spec_kernel(const __grid_constant__ CUtensorMap map_v, float *__restrict__ out,
                        const SpecTaps taps, int off0, int off1, int off2, int nj0, int nj1,
                        int nj2, int nt0, int nt1, int nt2, float scale, const int *done) {
  using C = DownCfg<K0, R0, K1, R1, K2, R2>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_bar[2];
  if (done && *done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (no integer round trip on the pointer: it would demote every access below to a generic LD /
  // ST with 64-bit address arithmetic -- ncu: 25 % IMAD, 11 % LEA, 10 % LD in the first version)
  float *box0 = reinterpret_cast<float *>(smem_raw);
  float *t1 = box0 + 2 * C::BOXP;
  float *t2 = t1 + C::T1;
  const uint32_t bar = smem_u32(s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ntiles = nt0 * nt1 * nt2;
  const TileDec dec{nt1, nt2, 1.f / (float)nt1, 1.f / (float)nt2};
  // the descriptor must be addressed in PARAMETER space: take its address here, not through a
  // by-reference lambda capture (which may spill a copy to local memory -> illegal instruction)
  const CUtensorMap *const pmap = &map_v;
  const uint32_t box_u32 = smem_u32(box0);
  auto issue = [=](int tile, int buf) {
    int b0, b1, b2;
    tile_decode(dec, tile, b0, b1, b2);
    mbar_expect_tx(bar + 8u * buf, (uint32_t)C::BOX * 4u);
    tma_load_3d(box_u32 + (uint32_t)(buf * C::BOXP) * 4u, pmap, bar + 8u * buf,
                b2 * C::L2 * R2 + off2, b1 * C::L1 * R1 + off1, b0 * C::L0 * R0 + off0);
  };
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  uint32_t par[2] = {0u, 0u};
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int nxt = tile + gridDim.x;
    if (tid == 0 && nxt < ntiles) issue(nxt, buf ^ 1);  // that buffer was consumed two syncs ago
    mbar_wait(bar + 8u * buf, par[buf]);
    par[buf] ^= 1u;
    const float *box = box0 + buf * C::BOXP;
    // ---- x pass: thread per (c1, c2) column: I0 planes in registers -> L0 outputs ----
    {
      constexpr int COLS = C::I1 * C::I2P;
      for (int col = tid; col < COLS; col += kSpecThreads) {
        float in[C::I0];
#pragma unroll
        for (int t = 0; t < C::I0; ++t) in[t] = box[t * COLS + col];
#pragma unroll
        for (int m = 0; m < C::L0; ++m) {
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < K0; ++t) acc = fmaf(taps.k0[t], in[m * R0 + t], acc);
          t1[m * COLS + col] = acc;
        }
      }
    }
    __syncthreads();
    // ---- y pass: a thread slides along y for HALF of the L1 outputs of one (l0, c2) ----
    {
      constexpr int H = C::L1 / 2, NI = (H - 1) * R1 + K1;
      constexpr int TASKS = C::L0 * 2 * C::I2P;
      for (int task = tid; task < TASKS; task += kSpecThreads) {
        const int c2 = task % C::I2P, r = task / C::I2P;
        const int l0 = r >> 1, h = r & 1;
        const float *p = t1 + (l0 * C::I1 + h * H * R1) * C::I2P + c2;
        float in[NI];
#pragma unroll
        for (int t = 0; t < NI; ++t) in[t] = p[t * C::I2P];
#pragma unroll
        for (int m = 0; m < H; ++m) {
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < K1; ++t) acc = fmaf(taps.k1[t], in[m * R1 + t], acc);
          t2[(l0 * C::L1 + h * H + m) * C::I2P + c2] = acc;
        }
      }
    }
    __syncthreads();
    // ---- z pass -> global: lane = l2, rows (l0, l1) over the warps, 64-bit LDS ----
    {
      int b0, b1, b2;
      tile_decode(dec, tile, b0, b1, b2);
      const int j2 = b2 * C::L2 + lane;
      for (int row = warp; row < C::L0 * C::L1; row += kSpecWarps) {
        const int l0 = row / C::L1, l1 = row - l0 * C::L1;
        const int j0 = b0 * C::L0 + l0, j1 = b1 * C::L1 + l1;
        const float *p = t2 + row * C::I2P + R2 * lane;
        float acc = 0.f;
        if (R2 == 2) {
#pragma unroll
          for (int t = 0; t + 1 < K2; t += 2) {
            const float2 q = *reinterpret_cast<const float2 *>(p + t);
            acc = fmaf(taps.k2[t], q.x, acc);
            acc = fmaf(taps.k2[t + 1], q.y, acc);
          }
          if (K2 & 1) acc = fmaf(taps.k2[K2 - 1], p[K2 - 1], acc);
        } else {
#pragma unroll
          for (int t = 0; t < K2; ++t) acc = fmaf(taps.k2[t], p[t], acc);
        }
        if (j0 < nj0 && j1 < nj1 && j2 < nj2)
          out[((size_t)j0 * nj1 + j1) * nj2 + j2] = scale * acc;
      }
    }
    // t1 / t2 are rewritten only after the next tile's sync points; the box buffer `buf` is
    // re-armed by the issue at the top of the iteration after next (a sync lies in between)
    __syncthreads();
  }
}

// ---------------------------------------------------------------- up
template <int K0, int R0, int K1, int R1, int K2, int R2>
struct UpCfg {
  static constexpr int E0 = 8, E1 = 16, E2 = 128;  // output tile (112 KB of shared memory: 2 CTAs/SM;
                                                   // 4 x 16 x 128 at 3 CTAs/SM measured slower)
  static constexpr int N0 = (E0 + K0 - 2) / R0 + 2, N1 = (E1 + K1 - 2) / R1 + 2,
                       N2 = (E2 + K2 - 2) / R2 + 2;                // low-res box
  static constexpr int N2P = (N2 + 3) / 4 * 4;
  static constexpr int BOX = N0 * N1 * N2P;
  static constexpr int BOXP = (BOX + 31) / 32 * 32;  // TMA destinations: 128-byte aligned
  static constexpr int U1 = E0 * N1 * N2P, U2 = E0 * E1 * N2P;
  static constexpr size_t SMEM = (size_t)(2 * BOXP + U1 + U2) * 4 + 128;
  static constexpr int M0 = (K0 + R0 - 1) / R0, M1 = (K1 + R1 - 1) / R1,
                       M2 = (K2 + R2 - 1) / R2;                    // low-res rows per output
  static_assert(R0 == 2 && R1 == 2 && R2 == 2, "register-window expansion: ratio 2");
  static_assert(E0 / 2 + (K0 - 1) / 2 <= N0 && E1 / 2 + (K1 - 1) / 2 <= N1 &&
                    E2 / 2 + (K2 - 1) / 2 <= N2P,
                "window inside the box");
  // the zero-padded tap tables absorb indices down to -(R M - K + R - 1) >= -4
  static_assert(R0 * M0 - K0 + R0 - 1 <= 4 && R1 * M1 - K1 + R1 - 1 <= 4 &&
                    R2 * M2 - K2 + R2 - 1 <= 4,
                "tap table padding");
};

// NOUT consecutive outputs of a ratio-2 transposed profile (K odd taps) from low-res rows held
// in registers.  v[i] is box row (first row of the window) + i; P is the parity of the first
// output's tap index on that row (0: tap K - 1, 1: tap K - 2).  An output with an even tap
// index sums taps K-1, K-3, .., 0 over H + 1 rows, an odd one taps K-2, .., 1 over H rows
// (H = (K - 1) / 2); every index below is a compile-time constant after unrolling: one FFMA
// per tap, taps as constant-bank operands.
template <int K, int NOUT, int P, int NV>
__device__ __forceinline__ void expand_r2(const float (&kk)[UR_MAX_TAPS], const float (&v)[NV],
                                          float (&out)[NOUT]) {
  static_assert((K & 1) == 1, "odd tap count");
  constexpr int H = (K - 1) / 2;
  static_assert(NV >= NOUT / 2 + H, "window rows");
#pragma unroll
  for (int n = 0; n < NOUT; ++n) {
    float acc = 0.f;
    if (((n + P) & 1) == 0) {
      const int i0 = P == 0 ? n / 2 : (n - 1) / 2;
#pragma unroll
      for (int i = 0; i <= H; ++i) acc = fmaf(kk[K - 1 - 2 * i], v[i0 + i], acc);
    } else {
      const int i0 = P == 0 ? (n - 1) / 2 + 1 : n / 2;
#pragma unroll
      for (int i = 0; i < H; ++i) acc = fmaf(kk[K - 2 - 2 * i], v[i0 + i], acc);
    }
    out[n] = acc;
  }
}

// 4-byte asynchronous global -> shared copy with zero fill (src-size 0 reads nothing)
__device__ __forceinline__ void spec_cp_async4(float *dst, const float *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz)
               : "memory");
}

// The low-res box of a tile is small (22 KB against 64 KB of output): it is fetched with
// zero-filling cp.async copies, one tile ahead (a 3-D TMA box over the low-res volume faulted
// with "illegal instruction" on this driver; not worth more GPU minutes than the 10 % of
// instructions the copies cost).
template <int MODE, int K0, int R0, int K1, int R1, int K2, int R2>
__global__ void __launch_bounds__(kSpecThreads, 2)
    nd_up_spec_kernel(const float *__restrict__ xl, const SpecTaps taps, int off0, int off1,
                      int off2, int nj0, int nj1, int nj2, int nt0, int nt1, int nt2, float scale,
                      const LhsArgs a) {
  using C = UpCfg<K0, R0, K1, R1, K2, R2>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_red[kMaxWarps];
  __shared__ float s_k0[UR_MAX_TAPS + 8], s_k1[UR_MAX_TAPS + 8], s_k2[UR_MAX_TAPS + 8];
  if (a.done && *a.done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (no integer round trip on the pointer: it would demote every access below to a generic LD /
  // ST with 64-bit address arithmetic -- ncu: 25 % IMAD, 11 % LEA, 10 % LD in the first version)
  float *box0 = reinterpret_cast<float *>(smem_raw);
  float *u1 = box0 + 2 * C::BOXP;
  float *u2 = u1 + C::U1;
  // taps with zero padding on both sides: out-of-range tap indices read 0 (branch-free)
  if (tid < UR_MAX_TAPS + 8) {
    const int t = tid - 4;
    s_k0[tid] = (t >= 0 && t < K0) ? taps.k0[t] : 0.f;
    s_k1[tid] = (t >= 0 && t < K1) ? taps.k1[t] : 0.f;
    s_k2[tid] = (t >= 0 && t < K2) ? taps.k2[t] : 0.f;
  }
  const int ntiles = nt0 * nt1 * nt2;
  const TileDec dec{nt1, nt2, 1.f / (float)nt1, 1.f / (float)nt2};
  // first low-res row whose support can reach output index o:  ceil((o - off - K + 1) / R)
  auto jb_of = [](int o, int off, int K, int R) { return nd_ceildiv(o - off - K + 1, R); };
  const size_t ly = nj2, lx = (size_t)nj1 * nj2;
  auto issue = [&](int tile, int buf) {  // every thread copies its share of the box
    int b0, b1, b2;
    tile_decode(dec, tile, b0, b1, b2);
    const int g0 = jb_of(b0 * C::E0, off0, K0, R0), g1 = jb_of(b1 * C::E1, off1, K1, R1),
              g2 = jb_of(b2 * C::E2, off2, K2, R2);
    float *dst = box0 + buf * C::BOXP;
    for (int row = warp; row < C::N0 * C::N1; row += kSpecWarps) {
      const int c0 = row / C::N1, c1 = row - c0 * C::N1;
      const int j0 = g0 + c0, j1 = g1 + c1;
      const bool row_in = j0 >= 0 && j0 < nj0 && j1 >= 0 && j1 < nj1;
      const float *src = row_in ? xl + (size_t)j0 * lx + (size_t)j1 * ly : xl;
      for (int c2 = lane; c2 < C::N2P; c2 += 32) {
        const int j2 = g2 + c2;
        const bool ok = row_in && j2 >= 0 && j2 < nj2;
        spec_cp_async4(dst + row * C::N2P + c2, ok ? src + j2 : xl, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int buf = 0;
  double part = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int nxt = tile + gridDim.x;
    int b0, b1, b2;
    tile_decode(dec, tile, b0, b1, b2);
    const int o0 = b0 * C::E0, o1 = b1 * C::E1, o2 = b2 * C::E2;
    const int jb0 = jb_of(o0, off0, K0, R0), jb1 = jb_of(o1, off1, K1, R1),
              jb2 = jb_of(o2, off2, K2, R2);
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // this tile's box has landed
    __syncthreads();  // ... for every thread; the previous tile is done with u1 / u2 / buf ^ 1
    if (nxt < ntiles) issue(nxt, buf ^ 1);  // next tile's box, in flight during this tile
    // L2 prefetch of everything the finish stage of THIS tile reads from HBM (v with its one
    // voxel halo, acc, b): the stage prefetches registers only one row ahead, which leaves too
    // few bytes in flight at 16 warps per SM (ncu: long-scoreboard 4.5 warps per issue slot);
    // these fire-and-forget prefetches run under the x and y passes.
    {
      constexpr int PR = (C::E0 + 2) * (C::E1 + 2) * (C::E2 / 32);
      const size_t sy = a.nz, sx = (size_t)a.ny * a.nz;
      for (int r = tid; r < PR; r += kSpecThreads) {
        const int seg = r % (C::E2 / 32), row = r / (C::E2 / 32);
        const int px = o0 - 1 + row / (C::E1 + 2), py = o1 - 1 + row % (C::E1 + 2);
        const int pz = o2 + 32 * seg;
        if (px < 0 || px >= a.nx || py < 0 || py >= a.ny || pz >= a.nz) continue;
        const size_t i = px * sx + py * sy + pz;
        const bool inner = px >= o0 && px < o0 + C::E0 && py >= o1 && py < o1 + C::E1;
        if (MODE != LHS_TERM) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.v + i));
        if (inner && a.acc) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.acc + i));
        if (inner && (MODE == LHS_RESID || MODE == LHS_ENERGY))
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.b + i));
      }
    }
    const float *lr = box0 + buf * C::BOXP;  // [N0][N1][N2P], rows outside the low-res grid = 0
    // The box starts at the first row that reaches the tile, so relative to it the first output
    // of the tile has tap index K - 1 or K - 2 on box row 0: only its PARITY is a run-time
    // (uniform) quantity; the register-window expansions below are compiled for both parities.
    const int p0 = (o0 - off0 - R0 * jb0) & 1, p1 = (o1 - off1 - R1 * jb1) & 1,
              p2 = (o2 - off2 - R2 * jb2) & 1;
    // ---- x pass: a thread owns a (c1, c2) column: N0 rows in registers -> E0 outputs ----
    {
      constexpr int COLS = C::N1 * C::N2P, NR = C::E0 / 2 + (K0 - 1) / 2;
      for (int col = tid; col < COLS; col += kSpecThreads) {
        float v[NR], o[C::E0];
#pragma unroll
        for (int i = 0; i < NR; ++i) v[i] = lr[i * COLS + col];
        if (p0)
          expand_r2<K0, C::E0, 1>(taps.k0, v, o);
        else
          expand_r2<K0, C::E0, 0>(taps.k0, v, o);
#pragma unroll
        for (int e = 0; e < C::E0; ++e) u1[e * COLS + col] = o[e];
      }
    }
    __syncthreads();
    // ---- y pass: a thread owns an (e0, c2) column: N1 rows in registers -> E1 outputs ----
    {
      constexpr int COLS = C::E0 * C::N2P, NR = C::E1 / 2 + (K1 - 1) / 2;
      for (int col = tid; col < COLS; col += kSpecThreads) {
        const int e0 = col / C::N2P, c2 = col - e0 * C::N2P;
        const float *src = u1 + e0 * (C::N1 * C::N2P) + c2;
        float v[NR], o[C::E1];
#pragma unroll
        for (int i = 0; i < NR; ++i) v[i] = src[i * C::N2P];
        if (p1)
          expand_r2<K1, C::E1, 1>(taps.k1, v, o);
        else
          expand_r2<K1, C::E1, 0>(taps.k1, v, o);
        float *dst = u2 + e0 * (C::E1 * C::N2P) + c2;
#pragma unroll
        for (int e = 0; e < C::E1; ++e) dst[e * C::N2P] = o[e];
      }
    }
    __syncthreads();
    // ---- per quad: z pass, stencil, epilogue.  A warp owns E1 / 8 rows of the tile and MARCHES
    // along x in each: the x-column of v stays in registers (previous / current / next plane),
    // the other operands of the next plane are requested before this plane's arithmetic, and
    // the linear index advances by one plane stride (no per-quad index arithmetic). ----
    {
      constexpr int NR = 4 / 2 + (K2 - 1) / 2;
      const int z = o2 + 4 * lane;
      const size_t sy = a.nz, sx = (size_t)a.ny * a.nz;
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      auto ldv = [&](size_t idx) { return *reinterpret_cast<const float4 *>(a.v + idx); };
      for (int e1 = warp; e1 < C::E1; e1 += kSpecWarps) {
        const int y = o1 + e1;
        if (z >= a.nz || y >= a.ny || o0 >= a.nx) continue;
        size_t i = (size_t)o0 * sx + (size_t)y * sy + z;
        NdQuadIn q, qn;
        q.c = q.xm = q.xp = zero4;
        float4 c_next2 = zero4;
        if (MODE != LHS_TERM) {
          q.xm = o0 > 0 ? ldv(i - sx) : zero4;
          q.c = ldv(i);
          q.xp = o0 + 1 < a.nx ? ldv(i + sx) : zero4;
        }
        nd_quad_load_side<MODE>(a, i, y, z, q);
        const float2 *row = reinterpret_cast<const float2 *>(u2 + e1 * C::N2P) + lane;
#pragma unroll 1
        for (int e0 = 0; e0 < C::E0; ++e0, i += sx, row += C::E1 * C::N2P / 2) {
          const int x = o0 + e0;
          if (x >= a.nx) break;
          const bool more = e0 + 1 < C::E0 && x + 1 < a.nx;
          if (more) {  // operands of the next plane
            nd_quad_load_side<MODE>(a, i + sx, y, z, qn);
            c_next2 = (MODE != LHS_TERM && x + 2 < a.nx) ? ldv(i + 2 * sx) : zero4;
          }
          // the quad at lane l starts 4 l outputs = 2 l low-res rows after the tile's first
          float v[NR + (NR & 1)], data[4];
#pragma unroll
          for (int k = 0; k < NR; k += 2) {
            const float2 t = row[k / 2];
            v[k] = t.x;
            v[k + 1] = t.y;
          }
          if (p2)
            expand_r2<K2, 4, 1>(taps.k2, v, data);
          else
            expand_r2<K2, 4, 0>(taps.k2, v, data);
#pragma unroll
          for (int k = 0; k < 4; ++k) data[k] *= scale;
          nd_quad_finish<MODE>(a, x, y, z, i, q, data, part);
          if (more) {
            qn.xm = q.c;
            qn.c = q.xp;
            qn.xp = c_next2;
            q = qn;
          }
        }
      }
    }
    __syncthreads();  // u1 / u2 / the box buffer may be overwritten from here on
  }
  if (MODE == LHS_TERM) return;
  double total;
  if (grid_sum(part, a.gr, s_red, &total) && tid == 0) finalize(a.fin, total);
}

// ---------------------------------------------------------------- host side
static bool spec_match(const NdOp &op, int K0, int R0, int K1, int R1, int K2, int R2) {
  return op.ax[0].K == K0 && op.ax[0].r == R0 && op.ax[1].K == K1 && op.ax[1].r == R1 &&
         op.ax[2].K == K2 && op.ax[2].r == R2;
}

static SpecTaps spec_taps(const NdOp &op) {
  SpecTaps t;
  memset(&t, 0, sizeof(t));
  for (int k = 0; k < op.ax[0].K; ++k) t.k0[k] = op.ax[0].ker[k];
  for (int k = 0; k < op.ax[1].K; ++k) t.k1[k] = op.ax[1].ker[k];
  for (int k = 0; k < op.ax[2].K; ++k) t.k2[k] = op.ax[2].ker[k];
  return t;
}

static int spec_opt_in(const void *kernel, size_t smem) {
  static bool done_dev[64][8] = {};
  static const void *slots[8] = {};
  int dev = 0;
  UR_CUDA_CHECK(cudaGetDevice(&dev));
  UR_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  int s = 0;
  while (s < 8 && slots[s] && slots[s] != kernel) ++s;
  if (s == 8) s = 7;
  slots[s] = kernel;
  if (!done_dev[dev][s] || s == 7) {
    UR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    done_dev[dev][s] = true;
  }
  return UR_OK;
}

int nd_down_spec_launch(const NdOp &op, const float *v, float *out, float scale, const int *done,
                        cudaStream_t st) {
  if (!spec_match(op, 3, 2, 9, 2, 9, 2)) return UR_ERR_UNSUPPORTED;
  using C = DownCfg<3, 2, 9, 2, 9, 2>;
  CUtensorMap map;
  if (!box_tensor_map(v, op.n[0], op.n[1], op.n[2], C::I0, C::I1, C::I2P, &map))
    return UR_ERR_UNSUPPORTED;
  auto kernel = nd_down_spec_kernel<3, 2, 9, 2, 9, 2>;
  int rc = spec_opt_in((const void *)kernel, C::SMEM);
  if (rc) return rc;
  const int nt0 = (op.ax[0].nj + C::L0 - 1) / C::L0, nt1 = (op.ax[1].nj + C::L1 - 1) / C::L1,
            nt2 = (op.ax[2].nj + C::L2 - 1) / C::L2;
  const long long ntiles = (long long)nt0 * nt1 * nt2;
  const long long slots = 2ll * sm_count();
  const int grid = (int)(ntiles < slots ? ntiles : slots);
  kernel<<<grid, kSpecThreads, C::SMEM, st>>>(map, out, spec_taps(op), op.ax[0].off, op.ax[1].off,
                                             op.ax[2].off, op.ax[0].nj, op.ax[1].nj, op.ax[2].nj,
                                             nt0, nt1, nt2, scale, done);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

template <int MODE>
static int up_spec(const NdOp &op, const float *xl, float scale, const LhsArgs &A,
                   cudaStream_t st) {
  using C = UpCfg<3, 2, 9, 2, 9, 2>;
  auto kernel = nd_up_spec_kernel<MODE, 3, 2, 9, 2, 9, 2>;
  int rc = spec_opt_in((const void *)kernel, C::SMEM);
  if (rc) return rc;
  const int nt0 = (op.n[0] + C::E0 - 1) / C::E0, nt1 = (op.n[1] + C::E1 - 1) / C::E1,
            nt2 = (op.n[2] + C::E2 - 1) / C::E2;
  const long long ntiles = (long long)nt0 * nt1 * nt2;
  const long long slots = 2ll * sm_count();  // 112 KB of shared memory: two CTAs per SM
  const int grid = (int)(ntiles < slots ? ntiles : slots);
  kernel<<<grid, kSpecThreads, C::SMEM, st>>>(xl, spec_taps(op), op.ax[0].off, op.ax[1].off,
                                             op.ax[2].off, op.ax[0].nj, op.ax[1].nj, op.ax[2].nj,
                                             nt0, nt1, nt2, scale, A);
  UR_LAUNCH_CHECK();
  return UR_OK;
}

int nd_up_spec_launch(int mode, const NdOp &op, const float *xl, float scale, const LhsArgs &A,
                      cudaStream_t st) {
  if (!spec_match(op, 3, 2, 9, 2, 9, 2)) return UR_ERR_UNSUPPORTED;
  switch (mode) {
    case LHS_PLAIN: return up_spec<LHS_PLAIN>(op, xl, scale, A, st);
    case LHS_RESID: return up_spec<LHS_RESID>(op, xl, scale, A, st);
    case LHS_TERM: return up_spec<LHS_TERM>(op, xl, scale, A, st);
    default: return up_spec<LHS_ENERGY>(op, xl, scale, A, st);
  }
}

}  // namespace ur

// K1 (main): register-tiled z-marching form of the FDI regularisation normal operator for the rows
// 3 <= j <= ny-4, every column and every plane.  Same operator, same ghost-cell formulation and same
// reference lines as fdi_stencil.cuh (LI/_finite_difference_interpolator.py:643-692,884-1164); the
// six rows next to the y faces are left to k_fdi_normal (fdi_stencil.cuh) running in its y-shell mode,
// so that here every y coefficient is a launch constant.
//
// Why a second kernel: ncu of k_fdi_normal at 256^3 (profiles/r01b) shows 103-168 warp instructions
// and 40 fp64 operations per 32 nodes, two to five resident warps per scheduler and "wait" /
// "short scoreboard" as the top stalls: it is bound by instruction latency, not by HBM.  Here a
// thread owns a 2 (x) by YB (y) patch of the tile:
//   * every staged row is read once with three 16-byte shared loads and reused by the YB outputs
//     that touch it (3 LDS.128 per node pair instead of 11);
//   * the row sums p(i-2)+p(i+2) are formed once per staged row and serve both the in-row tap and the
//     four xy corner taps of the rows two above and below;
//   * 28 fp64 operations per node, eight independent accumulation chains per thread;
//   * the five output-plane accumulators rotate statically (five copies of the plane body);
//   * planes are staged by 16-byte cp.async (LDGSTS) into a ring of ST planes, ST-3 of them in flight;
//     the ring keeps planes kp-1 and kp-2 resident so that p(k) for the fused p.q reduction is
//     re-read from shared memory instead of being carried in registers.
// Planes within three of a z face take a general path (plane_general of fdi_stencil.cuh, row by row).
#pragma once
#include "fdi_stencil.cuh"

namespace lsb {

struct MainConsts {
  double cy1, cy2;   // y taps +-1 / +-2 of ayy T_y - 2 byz [|t|=2]   (masked by mx at use)
  double bxy, bxz, byz;
  double zt0, zt1;   // azz/6, -2 azz/3
};

template <int YB, int TYT, int ST>
struct MainCfg {
  static constexpr int XB = 2, TXT = 32;
  static constexpr int NT = TXT * TYT;
  static constexpr int TX = TXT * XB;
  static constexpr int TY = TYT * YB;
  static constexpr int SX = TX + 2 * HALO;
  static constexpr int SY = TY + 2 * HALO;
  static constexpr int PLANE = SX * SY;
  static constexpr int LD8 = (PLANE + NT - 1) / NT;
  static constexpr int LD16 = (PLANE / 2 + NT - 1) / NT;
  static constexpr int Y_FIRST = 3;  // first row handled here; rows 0..2 and ny-3..ny-1 are the y shell
};

__device__ __forceinline__ double2 lds2(const double* s) { return *reinterpret_cast<const double2*>(s); }

// One incoming regular plane (3 <= kp <= nz-2) for the thread's 2 x YB patch.  Sc points at the staged
// element (row j0-2, column i0-2).  A*[r][e] accumulate output planes kp-2..kp+2.
template <int YB, int SX>
__device__ __forceinline__ void main_plane_fast(const double* __restrict__ Sc, const double (&cx)[2][5],
                                                const double (&mx)[2], const MainConsts& K, double (&Am2)[YB][2],
                                                double (&Am1)[YB][2], double (&A0)[YB][2], double (&Ap1)[YB][2],
                                                double (&Ap2)[YB][2]) {
  constexpr int NR = YB + 4;
  double pc[NR][2], h2[NR][2];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const double2 a = lds2(Sc + rr * SX), b = lds2(Sc + rr * SX + 2), c = lds2(Sc + rr * SX + 4);
    h2[rr][0] = a.x + c.x;
    h2[rr][1] = a.y + c.y;
    pc[rr][0] = b.x;
    pc[rr][1] = b.y;
    if (rr >= 2 && rr < YB + 2) {
      // x line of output row rr-2 (folded coefficients, see k_fdi_main prologue), on top of A0
      double s0 = fma(cx[0][0], a.x, A0[rr - 2][0]);
      s0 = fma(cx[0][1], a.y, s0);
      s0 = fma(cx[0][2], b.x, s0);
      s0 = fma(cx[0][3], b.y, s0);
      s0 = fma(cx[0][4], c.x, s0);
      double s1 = fma(cx[1][0], a.y, A0[rr - 2][1]);
      s1 = fma(cx[1][1], b.x, s1);
      s1 = fma(cx[1][2], b.y, s1);
      s1 = fma(cx[1][3], c.x, s1);
      s1 = fma(cx[1][4], c.y, s1);
      A0[rr - 2][0] = s0;
      A0[rr - 2][1] = s1;
    }
  }
#pragma unroll
  for (int r = 0; r < YB; ++r) {
    const int rr = r + 2;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double p0 = pc[rr][e];
      const double sy1 = pc[rr - 1][e] + pc[rr + 1][e];
      const double sy2 = pc[rr - 2][e] + pc[rr + 2][e];
      const double cs = h2[rr - 2][e] + h2[rr + 2][e];
      double Y = K.cy1 * sy1;
      Y = fma(K.cy2, sy2, Y);
      double S = fma(mx[e], Y, A0[r][e]);
      S = fma(K.bxy, fma(-2.0, sy2, cs), S);
      const double pm = mx[e] * p0;
      const double uym = mx[e] * fma(2.0, p0, -sy2);
      const double u = fma(K.bxz, fma(2.0, p0, -h2[rr][e]), K.byz * uym);
      Am2[r][e] = fma(K.zt0, pm, Am2[r][e]) - u;
      Am1[r][e] = fma(K.zt1, pm, Am1[r][e]);
      A0[r][e] = S;
      Ap1[r][e] = fma(K.zt1, pm, Ap1[r][e]);
      Ap2[r][e] = fma(K.zt0, pm, Ap2[r][e]) - u;
    }
  }
}

template <int YB, int TYT, int ST, int MINB, bool VEC16>
__global__ void __launch_bounds__(32 * TYT, MINB)
    k_fdi_main(const GridDesc g, const StencilWeights w, const MainConsts K, const double* __restrict__ p,
               double* __restrict__ q, const int cz, double* __restrict__ partials, unsigned int* __restrict__ counter,
               CgScalars* __restrict__ sc, const int cg_mode) {
  using C = MainCfg<YB, TYT, ST>;
  static_assert(ST >= 5, "ring needs planes kp-2..kp plus at least two in flight");
  extern __shared__ __align__(16) double S[];
  __shared__ double red[C::NT / 32];

  if (cg_mode && sc->done) return;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + 32 * ty;
  const int x0 = blockIdx.x * C::TX, y0 = C::Y_FIRST + blockIdx.y * C::TY;
  const int ka = g.k0 + blockIdx.z * cz;
  const int kb = min(ka + cz, g.k1);
  const int i0 = x0 + 2 * tx;
  const int j0 = y0 + YB * ty;
  const int nrows = max(0, min(YB, (g.ny - 3) - j0));  // rows of this thread that belong to the main range

  // staging descriptors: staged slot -> in-plane offset of its (x ghost-mapped) source, -1 = zero fill
  constexpr int LD = VEC16 ? C::LD16 : C::LD8;
  constexpr int NSLOT = VEC16 ? C::PLANE / 2 : C::PLANE;
  int off[LD];
#pragma unroll
  for (int t = 0; t < LD; ++t) {
    const int idx = tid + t * C::NT;
    off[t] = -1;
    if (idx < NSLOT) {
      const int e0 = VEC16 ? 2 * idx : idx;
      const int sy = e0 / C::SX, sx = e0 - sy * C::SX;
      const int gi = ghost_map(x0 - HALO + sx, g.nx);
      const int gj = y0 - HALO + sy;  // >= 1: never a ghost row
      if (gi >= 0 && gj < g.ny) off[t] = gi + g.pitch * gj;
    }
  }
  auto issue_plane = [&](const int kp, double* dstS) {
    const int ks = ghost_map(kp, g.nz);
    const double* src = p + (long long)(ks - g.k0 + HALO) * g.plane;
#pragma unroll
    for (int t = 0; t < LD; ++t) {
      const int idx = tid + t * C::NT;
      if (idx < NSLOT) {
        const bool ok = off[t] >= 0;
        if (VEC16) cp_async_16(dstS + 2 * idx, src + (ok ? off[t] : 0), ok);
        else cp_async_8(dstS + idx, src + (ok ? off[t] : 0), ok);
      }
    }
  };

  // per-thread x coefficients (rows here have my = 1 and uniform y coefficients)
  double cx[2][5], mx[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = i0 + e;
    mx[e] = (i > 0 && i < g.nx - 1) ? 1.0 : 0.0;
#pragma unroll
    for (int s = 0; s < 5; ++s) cx[e][s] = w.axx * coef_T(i, g.nx, s - 2) + w.dx * coef_B(i, g.nx, s - 2);
    cx[e][0] -= 2.0 * (w.bxy + w.bxz);
    cx[e][4] -= 2.0 * (w.bxy + w.bxz);
    cx[e][2] += mx[e] * w.ayy + 4.0 * (w.bxy + w.bxz + w.byz * mx[e]) + w.azz * mx[e] + w.ridge2;
  }

  double A[5][YB][2];
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int r = 0; r < YB; ++r) A[d][r][0] = A[d][r][1] = 0.0;
  double dot = 0.0;

  const int kstart = ka - HALO, kend = kb + HALO;
  int kp = kstart;
  // prologue: ST-3 planes in flight
#pragma unroll
  for (int s = 0; s < ST - 3; ++s) {
    if (kp + s < kend) issue_plane(kp + s, S + s * C::PLANE);
    cp_async_commit();
  }
  int slot = 0;  // ring slot of plane kp
  const int my_off = (YB * ty) * C::SX + 2 * tx;

  // wait for plane kp; then refill the slot of plane kp-3 (no longer needed) with plane kp+ST-3
  auto advance_pipeline = [&]() {
    cp_async_wait<ST - 4>();
    __syncthreads();
    const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;
    if (kp + ST - 3 < kend) issue_plane(kp + ST - 3, S + fs * C::PLANE);
    cp_async_commit();
  };
  auto next_slot = [&]() { slot = (slot + 1 == ST) ? 0 : slot + 1; };
  auto back_slot = [&](const int n) { return slot >= n ? slot - n : slot + ST - n; };

  // store output plane kp-2 from accumulator `out` and add p(kp-2).out to the dot product
  auto store_out = [&](const double (&out)[YB][2]) {
    const int ko = kp - HALO;
    if (ko < ka) return;
    const double* Sp = S + back_slot(2) * C::PLANE + my_off + 2 * C::SX + 2;
    double* dst = q + (long long)(ko - g.k0 + HALO) * g.plane + (long long)g.pitch * j0 + i0;
#pragma unroll
    for (int r = 0; r < YB; ++r) {
      if (r < nrows) {
        const double2 pv = lds2(Sp + r * C::SX);
        if (VEC16) {
          if (i0 < g.nx) {
            *reinterpret_cast<double2*>(dst + (long long)r * g.pitch) = make_double2(out[r][0], out[r][1]);
            dot = fma(pv.x, out[r][0], dot);
            dot = fma(pv.y, out[r][1], dot);
          }
        } else {
          if (i0 < g.nx) {
            dst[(long long)r * g.pitch] = out[r][0];
            dot = fma(pv.x, out[r][0], dot);
          }
          if (i0 + 1 < g.nx) {
            dst[(long long)r * g.pitch + 1] = out[r][1];
            dot = fma(pv.y, out[r][1], dot);
          }
        }
      }
    }
  };

  // general planes (z ends): plane_general of fdi_stencil.cuh row by row, shifting accumulators
  auto slow_planes = [&](const int kstop) {
    for (; kp < kstop; ++kp) {
      advance_pipeline();
      const double* Sc = S + slot * C::PLANE + my_off;
      const bool h1 = kp - 1 >= kstart, h2ok = kp - 2 >= kstart;
      const double* S1 = S + back_slot(1) * C::PLANE + my_off + 2;
      const double* S2 = S + back_slot(2) * C::PLANE + my_off + 2;
#pragma unroll
      for (int r = 0; r < YB; ++r) {
        double rc[6], rm2[6], rp2[6], rm1[2], rp1[2], prev1[2] = {0.0, 0.0}, prev2[2] = {0.0, 0.0};
        load_rows<2, C::SX>(Sc + (r + 2) * C::SX, rc, rm2, rp2, rm1, rp1);
        if (h1) { const double2 v = lds2(S1 + (r + 2) * C::SX); prev1[0] = v.x; prev1[1] = v.y; }
        if (h2ok) { const double2 v = lds2(S2 + (r + 2) * C::SX); prev2[0] = v.x; prev2[1] = v.y; }
        plane_general<2>(g, w, kp, i0, j0 + r, rc, rm2, rp2, rm1, rp1, prev1, prev2, A[0][r], A[1][r], A[2][r], A[3][r],
                         A[4][r]);
      }
      store_out(A[0]);
#pragma unroll
      for (int r = 0; r < YB; ++r)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          A[0][r][e] = A[1][r][e]; A[1][r][e] = A[2][r][e]; A[2][r][e] = A[3][r][e]; A[3][r][e] = A[4][r][e];
          A[4][r][e] = 0.0;
        }
      next_slot();
    }
  };

  slow_planes(min(3, kend));

  const int kfast = min(g.nz - 1, kend);  // exclusive
  if (kp < kfast) {
    int phase = 0;
#define LSB_MAIN_STEP(PH)                                                                                     \
  {                                                                                                           \
    if (kp >= kfast) { phase = PH; break; }                                                                   \
    advance_pipeline();                                                                                       \
    main_plane_fast<YB, C::SX>(S + slot * C::PLANE + my_off, cx, mx, K, A[(PH + 0) % 5], A[(PH + 1) % 5],     \
                               A[(PH + 2) % 5], A[(PH + 3) % 5], A[(PH + 4) % 5]);                            \
    store_out(A[(PH + 0) % 5]);                                                                               \
    _Pragma("unroll") for (int r = 0; r < YB; ++r) A[(PH + 0) % 5][r][0] = A[(PH + 0) % 5][r][1] = 0.0;       \
    next_slot();                                                                                              \
    ++kp;                                                                                                     \
  }
    for (;;) {
      LSB_MAIN_STEP(0)
      LSB_MAIN_STEP(1)
      LSB_MAIN_STEP(2)
      LSB_MAIN_STEP(3)
      LSB_MAIN_STEP(4)
    }
#undef LSB_MAIN_STEP
    // back to canonical order for the trailing general planes
    double T[5][YB][2];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
      for (int r = 0; r < YB; ++r) T[d][r][0] = A[d][r][0], T[d][r][1] = A[d][r][1];
#pragma unroll
    for (int ph = 0; ph < 5; ++ph)
      if (phase == ph) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
          for (int r = 0; r < YB; ++r) A[d][r][0] = T[(ph + d) % 5][r][0], A[d][r][1] = T[(ph + d) % 5][r][1];
      }
  }

  slow_planes(kend);
  cp_async_wait<0>();

  const double bsum = block_sum<C::NT>(dot, red);
  const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  double total;
  if (grid_sum_last_block<C::NT>(bsum, partials, counter, nblocks, bid, red, &total)) sc->pq_part[0] = total;
}

}  // namespace lsb

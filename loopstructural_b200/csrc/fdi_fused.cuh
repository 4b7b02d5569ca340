// K1, large grids, ONE launch: q = (R^T R + ridge^2 I) p over every owned node of a handle, fused p.q.
//
// Round 1 ran three launches on three streams (k_fdi_box uniform, k_fdi_box<ZGEN> for the z faces, k_fdi_shell for the
// x/y ring).  The ring kernel -- one thread per column gathering from global memory -- is latency bound (39 us alone
// at 256^3 for 5 % of the nodes, profiles/r02c) and cannot co-reside with the register-hungry box CTAs, so ~28 us of it
// trailed every matvec; the z-face launch queued behind the box for the same reason.  Here every node is computed by a
// CTA of the same launch, staged through shared memory like the box:
//
//   BOX CTAs    tile of 64 x TY nodes over the columns 4 <= i < nx-4 and ALL rows 0 <= j < ny, a chunk of cz planes.
//               Every plane step is the register-tiled constant-coefficient update of k_fdi_box (uni_plane_fast, 20
//               fp64 operations per node); the steps whose incoming plane lies within 5 planes of a z face use the
//               z-general twin (zgen_plane_fast, coefficients of that plane from a 10-entry table in shared memory).
//               The rows next to a y face (j < 3 or j >= ny-3) are exact through the ghost-cell formulation of
//               fdi_stencil.cuh -- rows -1,-2 / ny,ny+1 of the staged tile hold the mirrored rows -- plus a fix-up
//               (y_edge_fix) that only the warps owning such rows execute (a warp owns whole rows: no divergence).
//   STRIP CTAs  the 4 columns next to an x face, NT rows, a chunk of planes; lanes run along y, a thread owns the 4
//               nodes of its row.  The right strip is staged mirrored (local column c = global nx-1-c), so both sides
//               run the same code with compile-time x coefficients (T_x end corrections, the dx border rows, C_x on
//               ghost columns); y and z coefficients are general.  Strips are 3 % of the nodes at 256^3; their CTAs
//               are ordinary members of the launch, sized by the planner so that the whole grid is one even wave.
//
// Operator (fdi_main.cuh / fdi_stencil.cuh): q = ridge^2 p + sum_a (a_aa m_b m_c T_a + d_a B_a) p
//   + sum_{a<b} b_ab m_c (C_a (x) C_b) p, masks m(r) = [1 <= r <= n-2]; reference: assemble_inner / assemble_borders
//   (LI/_finite_difference_interpolator.py:1052-1164), row normalisation (LI/_discrete_interpolator.py:361-366).
#pragma once
#include "fdi_main.cuh"

namespace lsb {

struct FusedPlan {
  int tiles_x, tiles_y;   // box tiles over columns [4, nx-4) x rows [0, ny)
  int cz, chunks;         // z chunks of the box CTAs over the owned planes
  int n_box;              // tiles_x * tiles_y * chunks
  int strips_y;           // strip CTAs per side and z chunk: ceil(ny / NT)
  int scz, schunks;       // z chunks of the strip CTAs
  int n_strip;            // 2 * strips_y * schunks
};

// z coefficients of one incoming plane kp: what it contributes to the output planes kp-2..kp+2.  l* = azz T_z + dz B_z
// (what an interior column uses), t* = azz T_z alone (the part the masks m_x m_y switch off at an x / y face).
struct ZFull {
  double mz;
  double l0, lm1, lp1, lm2, lp2;
  double t0, tm1, tp1, tm2, tp2;
  double c0, cm2, cp2;
};

__device__ __forceinline__ ZFull z_full(const GridDesc& g, const double* __restrict__ ztab, const double azz, const double dz,
                                        const int k) {
  auto row = [&](const int ko, const int col) {
    return (ko >= 0 && ko < g.nz) ? ztab[(long long)ko * SHELL_ZSTRIDE + col] : 0.0;
  };
  // tap s = k - ko of output row ko: T at column 2 + s, B at 7 + s, C_z taps -2,0,+2 at 10,11,12
  ZFull Z;
  Z.mz = row(k, 13);
  Z.t0 = azz * row(k, 2);
  Z.tm1 = azz * row(k - 1, 3);
  Z.tp1 = azz * row(k + 1, 1);
  Z.tm2 = azz * row(k - 2, 4);
  Z.tp2 = azz * row(k + 2, 0);
  Z.l0 = fma(dz, row(k, 7), Z.t0);
  Z.lm1 = fma(dz, row(k - 1, 8), Z.tm1);
  Z.lp1 = fma(dz, row(k + 1, 6), Z.tp1);
  Z.lm2 = fma(dz, row(k - 2, 9), Z.tm2);
  Z.lp2 = fma(dz, row(k + 2, 5), Z.tp2);
  Z.c0 = row(k, 11);
  Z.cm2 = row(k - 2, 12);
  Z.cp2 = row(k + 2, 10);
  return Z;
}

__device__ __forceinline__ ZFull z_uniform(const double azz) {
  ZFull Z;
  Z.mz = 1.0;
  Z.t0 = Z.l0 = azz;
  Z.tm1 = Z.tp1 = Z.lm1 = Z.lp1 = azz * (-2.0 / 3.0);
  Z.tm2 = Z.tp2 = Z.lm2 = Z.lp2 = azz * (1.0 / 6.0);
  Z.c0 = 2.0;
  Z.cm2 = Z.cp2 = -1.0;
  return Z;
}

// an incoming plane kp whose five output planes all have the constant interior z coefficients
__device__ __forceinline__ bool z_step_uniform(const int kp, const int nz) { return kp >= 5 && kp <= nz - 6; }

// index of a non-uniform plane in the 10-entry table of a box CTA: planes 0..4 and nz-5..nz-1 (nz >= 10 is not needed:
// an entry is recomputed per plane, the two ranges may overlap)
__device__ __forceinline__ int z_slot(const int kp, const int nz) { return kp < 5 ? kp : 5 + (kp - (nz - 5)); }

// end corrections of the ghost-cell 1-D operators at index i of n (fdi_stencil.cuh header): T differs from the constant
// 5-tap stencil by  q(0) -= p0 - p1,  q(1) += (p0 - p1)/3  (mirrored at the upper end); border rows B = d d^T / 6 on the
// rows 1 and n-2.  pm2..pp2 are p(i-2..i+2) on ghost data.
__device__ __forceinline__ double t_end_corr(const int i, const int n, const double pm1, const double p0, const double pp1) {
  double c = 0.0;
  if (i == 0) c = -(p0 - pp1);
  if (i == 1) c = (pm1 - p0) * (1.0 / 3.0);
  if (i == n - 1) c += -(p0 - pm1);
  if (i == n - 2) c += (pp1 - p0) * (1.0 / 3.0);
  return c;
}
__device__ __forceinline__ double b_rows(const int i, const int n, const double pm2, const double pm1, const double p0,
                                         const double pp1, const double pp2) {
  double c = 0.0;
  // lower border row, centred on index 1, touches 0,1,2; upper, centred on n-2, touches n-3,n-2,n-1
  if (i == 0) c += (p0 - 2.0 * pp1 + pp2);
  if (i == 1) c += -2.0 * (pm1 - 2.0 * p0 + pp1);
  if (i == 2) c += (pm2 - 2.0 * pm1 + p0);
  if (i == n - 1) c += (p0 - 2.0 * pm1 + pm2);
  if (i == n - 2) c += -2.0 * (pp1 - 2.0 * p0 + pm1);
  if (i == n - 3) c += (pp2 - 2.0 * pp1 + p0);
  return c * (1.0 / 6.0);
}

// Fix-up of the box update for the output rows of this thread that lie within 3 rows of a y face.  Sc as in
// uni_plane_fast (staged element (row j0-2, column i0-2) of the incoming plane); the uniform update has already been
// accumulated with m_y = 1 and the constant T_y, no dy rows.  Adds (true - uniform) for those rows:
//   dA0   = mz (-d axx X_u + ayy ycorr) + dy yb - d t0 p0 + c0 dG,   dG = -d bxz gx,  d = 1 - m_y(j)
//   dA+-1 = -d t+-1 p0,   dA+-2 = -d t+-2 p0 + c+-2 dG
template <int YB, int SX>
__device__ __forceinline__ void y_edge_fix(const double* __restrict__ Sc, const UniConsts& K, const double dyw, const ZFull& Z,
                                           const int j0, const int ny, double (&Am2)[YB][2], double (&Am1)[YB][2],
                                           double (&A0)[YB][2], double (&Ap1)[YB][2], double (&Ap2)[YB][2]) {
#pragma unroll
  for (int r = 0; r < YB; ++r) {
    const int j = j0 + r;
    if (j >= 3 && j < ny - 3) continue;
    if (j >= ny) continue;
    const double* R = Sc + (r + 2) * SX;  // row j, columns i0-2 .. i0+3
    const double d = (j == 0 || j == ny - 1) ? 1.0 : 0.0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double p0 = R[2 + e];
      const double ym2 = R[2 + e - 2 * SX], ym1 = R[2 + e - SX], yp1 = R[2 + e + SX], yp2 = R[2 + e + 2 * SX];
      const double ycorr = t_end_corr(j, ny, ym1, p0, yp1);
      const double yb = b_rows(j, ny, ym2, ym1, p0, yp1, yp2);
      const double h2 = R[e] + R[4 + e];
      const double xu = fma(1.0 / 6.0, h2, fma(-2.0 / 3.0, R[1 + e] + R[3 + e], p0));  // constant T_x on interior columns
      const double gx = 2.0 * p0 - h2;
      const double dG = -d * K.bxz * gx;
      const double axx = K.cx1 * (-1.5);  // cx1 = -2 axx / 3
      const double ayy = K.cy1 * (-1.5);
      const double dp = d * p0;
      A0[r][e] += Z.mz * fma(ayy, ycorr, -d * axx * xu) + dyw * yb - Z.t0 * dp + Z.c0 * dG;
      Am1[r][e] -= Z.tm1 * dp;
      Ap1[r][e] -= Z.tp1 * dp;
      Am2[r][e] += fma(Z.cm2, dG, -Z.tm2 * dp);
      Ap2[r][e] += fma(Z.cp2, dG, -Z.tp2 * dp);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// STRIP role: 4 columns next to an x face (side 0: columns 0..3, side 1: columns nx-1..nx-4 as local 0..3), rows
// [y0, y0 + NT), output planes [ka, kb).  Staged plane: NT + 4 rows (ghost-mapped in y) of 6 local columns.
constexpr int STRIP_SC = 6;  // staged local columns 0..5

template <int NT, int ST>
__device__ __forceinline__ double strip_cta(const GridDesc& g, const StencilWeights& w, const double* __restrict__ ztab,
                                            const double* __restrict__ p, double* __restrict__ q, const int side,
                                            const int y0, const int ka, const int kb, double* __restrict__ S, const int tid) {
  constexpr int ROWS = NT + 2 * HALO;
  constexpr int PLANE = ROWS * STRIP_SC;
  constexpr int LD = (PLANE + NT - 1) / NT;
  int off[LD];
#pragma unroll
  for (int t = 0; t < LD; ++t) {
    const int idx = tid + t * NT;
    off[t] = -1;
    if (idx < PLANE) {
      const int sy = idx / STRIP_SC, c = idx - sy * STRIP_SC;
      const int gj = ghost_map(y0 - HALO + sy, g.ny);
      const int gi = side ? g.nx - 1 - c : c;
      if (gj >= 0) off[t] = gi + g.pitch * gj;
    }
  }
  const int j = y0 + tid;
  const bool row_ok = j < g.ny;
  const double my = (j >= 1 && j <= g.ny - 2) ? 1.0 : 0.0;
  double A[5][4];
#pragma unroll
  for (int dd = 0; dd < 5; ++dd)
#pragma unroll
    for (int c = 0; c < 4; ++c) A[dd][c] = 0.0;
  double dot = 0.0;

  const int kstart = ka - HALO, kend = ka < kb ? kb + HALO : ka - HALO;
  const double* src = p + (long long)(kstart - g.k0 + HALO) * g.plane;
  int kstage = kstart;
  auto issue_plane = [&](double* dstS) {
    if (kstage >= 0 && kstage < g.nz) {
#pragma unroll
      for (int t = 0; t < LD; ++t) {
        const int idx = tid + t * NT;
        if (idx < PLANE) {
          const bool ok = off[t] >= 0;
          cp_async_8(dstS + idx, src + (ok ? off[t] : 0), ok);
        }
      }
    }
    src += g.plane;
    ++kstage;
  };
  int issued = kstart;
#pragma unroll
  for (int s = 0; s < ST - 3; ++s) {
    if (issued < kend) { issue_plane(S + s * PLANE); ++issued; }
    cp_async_commit();
  }
  int slot = 0;
  double* qrow = q + (long long)(ka - g.k0 + HALO) * g.plane + (long long)g.pitch * j;  // output plane ka, row j
  for (int kp = kstart; kp < kend; ++kp) {
    cp_async_wait<ST - 4>();
    __syncthreads();
    {
      const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;
      if (issued < kend) { issue_plane(S + fs * PLANE); ++issued; }
      cp_async_commit();
    }
    if (kp >= 0 && kp < g.nz) {
      const ZFull Z = z_full(g, ztab, w.azz, w.dz, kp);
      const double* V = S + slot * PLANE + tid * STRIP_SC;  // row j-2 of this thread's window
      double v[5][STRIP_SC];
#pragma unroll
      for (int rr = 0; rr < 5; ++rr)
#pragma unroll
        for (int c = 0; c < STRIP_SC; c += 2) {
          const double2 t2 = lds2(V + rr * STRIP_SC + c);
          v[rr][c] = t2.x;
          v[rr][c + 1] = t2.y;
        }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // ghost columns: p(-1) := p(1), p(-2) := p(0)
        auto xg = [&](const int rr, const int cc) { return cc >= 0 ? v[rr][cc] : (cc == -1 ? v[rr][1] : v[rr][0]); };
        const double p0 = v[2][c];
        const double xm1 = xg(2, c - 1), xm2 = xg(2, c - 2), xp1 = v[2][c + 1], xp2 = v[2][c + 2];
        double X = fma(1.0 / 6.0, xm2 + xp2, fma(-2.0 / 3.0, xm1 + xp1, p0));
        if (c == 0) X -= (v[2][0] - v[2][1]);
        if (c == 1) X += (v[2][0] - v[2][1]) * (1.0 / 3.0);
        const double e2 = v[2][0] - 2.0 * v[2][1] + v[2][2];
        const double xb = (c == 0 || c == 2) ? e2 * (1.0 / 6.0) : (c == 1 ? e2 * (-2.0 / 6.0) : 0.0);
        const double gx0 = 2.0 * v[0][c] - xg(0, c - 2) - v[0][c + 2];
        const double gx2 = 2.0 * p0 - xm2 - xp2;
        const double gx4 = 2.0 * v[4][c] - xg(4, c - 2) - v[4][c + 2];
        const double XY = 2.0 * gx2 - gx0 - gx4;
        const double ym2 = v[0][c], ym1 = v[1][c], yp1 = v[3][c], yp2 = v[4][c];
        const double Y = fma(1.0 / 6.0, ym2 + yp2, fma(-2.0 / 3.0, ym1 + yp1, p0)) + t_end_corr(j, g.ny, ym1, p0, yp1);
        const double yb = b_rows(j, g.ny, ym2, ym1, p0, yp1, yp2);
        const double gy = 2.0 * p0 - (ym2 + yp2);
        const double mx = c == 0 ? 0.0 : 1.0;
        const double same = Z.mz * (w.axx * my * X + w.ayy * mx * Y + w.bxy * XY) + w.ridge2 * p0 + w.dx * xb + w.dy * yb;
        const double G = w.bxz * my * gx2 + w.byz * mx * gy;
        const double mm = mx * my;
        // l = t + (dz B): the masks act on the azz T part only
        A[2][c] += same + fma(Z.t0, mm * p0, (Z.l0 - Z.t0) * p0) + Z.c0 * G;
        A[1][c] += fma(Z.tm1, mm * p0, (Z.lm1 - Z.tm1) * p0);
        A[3][c] += fma(Z.tp1, mm * p0, (Z.lp1 - Z.tp1) * p0);
        A[0][c] += fma(Z.tm2, mm * p0, (Z.lm2 - Z.tm2) * p0) + Z.cm2 * G;
        A[4][c] += fma(Z.tp2, mm * p0, (Z.lp2 - Z.tp2) * p0) + Z.cp2 * G;
      }
    }
    // output plane kp-2 is complete
    if (kp - HALO >= ka) {
      if (row_ok) {
        const int s2 = slot >= 2 ? slot - 2 : slot + ST - 2;
        const double* P = S + s2 * PLANE + (tid + HALO) * STRIP_SC;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int gi = side ? g.nx - 1 - c : c;
          qrow[gi] = A[0][c];
          dot = fma(P[c], A[0][c], dot);
        }
      }
      qrow += g.plane;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      A[0][c] = A[1][c]; A[1][c] = A[2][c]; A[2][c] = A[3][c]; A[3][c] = A[4][c]; A[4][c] = 0.0;
    }
    slot = (slot + 1 == ST) ? 0 : slot + 1;
  }
  cp_async_wait<0>();
  return dot;
}

// ---------------------------------------------------------------------------------------------------------------
template <int TYT, int ST, int MINB, bool VEC16>
__global__ void __launch_bounds__(32 * TYT, MINB)
    k_fdi_fused(const GridDesc g, const UniConsts U, const StencilWeights w, const FusedPlan pl, const double* __restrict__ p,
                double* __restrict__ q, const double* __restrict__ ztab, double* __restrict__ partials,
                unsigned int* __restrict__ counter, CgScalars* __restrict__ sc, const int cg_mode) {
  constexpr int YB = 4;
  using C = MainCfg<YB, TYT, ST>;
  static_assert(ST >= 5, "ring needs planes kp-2..kp plus at least two in flight");
  static_assert((C::NT + 2 * HALO) * STRIP_SC <= C::PLANE, "a strip plane must fit a box plane slot");
  LSB_DYN_SMEM(S);
  __shared__ double red[C::NT / 32];
  __shared__ ZStep zsm[10];
  __shared__ ZFull zfm[11];  // [10]: the constant interior coefficients

  if (cg_mode && sc->done) return;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + 32 * ty;
  const int bid = blockIdx.x;
  double dot = 0.0;

  if (bid < pl.n_strip) {
    // ---- strip role
    const int per_chunk = 2 * pl.strips_y;
    const int ch = bid / per_chunk, rem = bid - ch * per_chunk;
    const int side = rem / pl.strips_y, sy = rem - side * pl.strips_y;
    const int ka = g.k0 + ch * pl.scz, kb = min(ka + pl.scz, g.k1);
    dot = strip_cta<C::NT, ST>(g, w, ztab, p, q, side, sy * C::NT, ka, kb, S, tid);
  } else {
    // ---- box role
    const int b = bid - pl.n_strip;
    const int tiles = pl.tiles_x * pl.tiles_y;
    const int ch = b / tiles, trem = b - ch * tiles;
    const int by = trem / pl.tiles_x, bx = trem - by * pl.tiles_x;
    const int x0 = C::X_FIRST + bx * C::TX, y0 = by * C::TY;
    const int ka = g.k0 + ch * pl.cz, kb = min(ka + pl.cz, g.k1);
    const int i0 = x0 + 2 * tx;
    const int j0 = y0 + YB * ty;
    const int nrows = max(0, min(YB, g.ny - j0));
    const int xlim = g.nx - C::X_FIRST;
    const bool y_edge = nrows > 0 && (j0 < 3 || j0 + YB > g.ny - 3);  // warp uniform

    // Staging of one plane, row by row: warp ty copies the staged rows ty, ty + TYT, ...; a lane copies the column pair
    // (VEC16) or the columns lane, lane + 32, lane + 64 of a row.  Addresses are affine in (row, lane): no per-thread
    // offset table (round 1's table of 6 offsets was the first thing ptxas spilled once the edge code was added).
    const int kstart = ka - HALO, kend = ka < kb ? kb + HALO : ka - HALO;
    int issued = kstart;  // next plane to stage (global index)
    auto issue_plane = [&](double* dstS) {
      if (issued >= 0 && issued < g.nz) {
        const double* src = p + (long long)(issued - g.k0 + HALO) * g.plane + (x0 - HALO);
#pragma unroll
        for (int sy = ty; sy < C::SY; sy += TYT) {
          const int gj = ghost_map(y0 - HALO + sy, g.ny);  // rows -1,-2 / ny,ny+1 mirror the rows inside; -1: overhang
          const double* row = src + (long long)g.pitch * max(gj, 0);
          double* drow = dstS + sy * C::SX;
          if (VEC16) {
            // SX = 68 doubles = 34 pairs: every lane one pair, lanes 0 and 1 the last two
            const bool ok = gj >= 0 && (x0 - HALO + 2 * tx) < g.nx;
            cp_async_16(drow + 2 * tx, row + (ok ? 2 * tx : 0), ok);
            if (tx < 2) {
              const bool ok2 = gj >= 0 && (x0 - HALO + 64 + 2 * tx) < g.nx;
              cp_async_16(drow + 64 + 2 * tx, row + (ok2 ? 64 + 2 * tx : 0), ok2);
            }
          } else {
#pragma unroll
            for (int c = tx; c < C::SX; c += 32) {
              const bool ok = gj >= 0 && (x0 - HALO + c) < g.nx;
              cp_async_8(drow + c, row + (ok ? c : 0), ok);
            }
          }
        }
      }
      ++issued;
    };
    // z coefficients of the (at most ten) incoming planes next to a z face
    if (tid < 10) {
      const int k = tid < 5 ? tid : g.nz - 5 + (tid - 5);
      ZStep Zs{};
      ZFull Zf{};
      if (k >= 0 && k < g.nz) {
        Zf = z_full(g, ztab, U.azz, U.dz, k);
        Zs.mz = Zf.mz; Zs.l0 = Zf.l0; Zs.c0 = Zf.c0; Zs.lm1 = Zf.lm1; Zs.lp1 = Zf.lp1;
        Zs.lm2 = Zf.lm2; Zs.cm2 = Zf.cm2; Zs.lp2 = Zf.lp2; Zs.cp2 = Zf.cp2;
      }
      // planes 0..4 fill slots 0..4, planes nz-5..nz-1 slots 5..9 (z_slot); on grids of fewer than 10 planes a plane
      // may own two slots with the same content
      zsm[tid] = Zs;
      zfm[tid] = Zf;
    }
    if (tid == 10) zfm[10] = z_uniform(U.azz);

    double A[5][YB][2];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
      for (int r = 0; r < YB; ++r) A[d][r][0] = A[d][r][1] = 0.0;

    int kp = kstart;
#pragma unroll
    for (int s = 0; s < ST - 3; ++s) {
      if (issued < kend) issue_plane(S + s * C::PLANE);
      cp_async_commit();
    }
    int slot = 0;
    const int my_off = (YB * ty) * C::SX + 2 * tx;
    double* qdst = q + (long long)(ka - g.k0 + HALO) * g.plane + (long long)g.pitch * j0 + i0;  // output plane ka

    auto advance_pipeline = [&]() {
      cp_async_wait<ST - 4>();
      __syncthreads();
      const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;
      if (issued < kend) issue_plane(S + fs * C::PLANE);
      cp_async_commit();
    };
    auto store_out = [&](const double (&out)[YB][2]) {
      if (kp - HALO < ka) return;
      const int s2 = slot >= 2 ? slot - 2 : slot + ST - 2;
      const double* Sp = S + s2 * C::PLANE + my_off + 2 * C::SX + 2;
#pragma unroll
      for (int r = 0; r < YB; ++r) {
        if (r < nrows) {
          const double2 pv = lds2(Sp + r * C::SX);
          if (VEC16) {
            if (i0 < xlim) {
              *reinterpret_cast<double2*>(qdst + (long long)r * g.pitch) = make_double2(out[r][0], out[r][1]);
              dot = fma(pv.x, out[r][0], dot);
              dot = fma(pv.y, out[r][1], dot);
            }
          } else {
            if (i0 < xlim) {
              qdst[(long long)r * g.pitch] = out[r][0];
              dot = fma(pv.x, out[r][0], dot);
            }
            if (i0 + 1 < xlim) {
              qdst[(long long)r * g.pitch + 1] = out[r][1];
              dot = fma(pv.y, out[r][1], dot);
            }
          }
        }
      }
      qdst += g.plane;
    };
#define LSB_FUSED_STEP(PH)                                                                                          \
  {                                                                                                                 \
    if (kp >= kend) break;                                                                                          \
    advance_pipeline();                                                                                             \
    if (kp >= 0 && kp < g.nz) {                                                                                     \
      const double* Sc = S + slot * C::PLANE + my_off;                                                              \
      int zs = 10;                                                                                                  \
      if (z_step_uniform(kp, g.nz)) {                                                                               \
        uni_plane_fast<YB, C::SX>(Sc, U, A[(PH + 0) % 5], A[(PH + 1) % 5], A[(PH + 2) % 5], A[(PH + 3) % 5],        \
                                  A[(PH + 4) % 5]);                                                                 \
      } else {                                                                                                      \
        zs = z_slot(kp, g.nz);                                                                                      \
        const ZStep Z = zsm[zs];                                                                                    \
        zgen_plane_fast<YB, C::SX>(Sc, U, Z, A[(PH + 0) % 5], A[(PH + 1) % 5], A[(PH + 2) % 5], A[(PH + 3) % 5],    \
                                   A[(PH + 4) % 5]);                                                                \
      }                                                                                                             \
      if (y_edge)                                                                                                   \
        y_edge_fix<YB, C::SX>(Sc, U, w.dy, zfm[zs], j0, g.ny, A[(PH + 0) % 5], A[(PH + 1) % 5], A[(PH + 2) % 5],    \
                              A[(PH + 3) % 5], A[(PH + 4) % 5]);                                                    \
    }                                                                                                               \
    store_out(A[(PH + 0) % 5]);                                                                                     \
    _Pragma("unroll") for (int r = 0; r < YB; ++r) A[(PH + 0) % 5][r][0] = A[(PH + 0) % 5][r][1] = 0.0;             \
    slot = (slot + 1 == ST) ? 0 : slot + 1;                                                                         \
    ++kp;                                                                                                           \
  }
    if (ka < kb) {
      for (;;) {
        LSB_FUSED_STEP(0)
        LSB_FUSED_STEP(1)
        LSB_FUSED_STEP(2)
        LSB_FUSED_STEP(3)
        LSB_FUSED_STEP(4)
      }
    }
#undef LSB_FUSED_STEP
    cp_async_wait<0>();
  }

  const double bsum = block_sum<C::NT>(dot, red);
  double total;
  if (grid_sum_last_block<C::NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[0] = total;
}

}  // namespace lsb

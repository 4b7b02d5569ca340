// K1, large grids: the FDI regularisation normal operator q = (R^T R + ridge^2 I) p split by region.
//   k_fdi_box<.., ZGEN=false>  the interior columns 4 <= i <= nx-5 and rows 3 <= j <= ny-4 of the planes
//                3 <= k <= nz-4 (92 % of a 256^3 grid, 98 % of a 1024^3 one), where R^T R is the constant 25-tap
//                composite kernel of LI/_finite_difference_interpolator.py:884-1050;
//   k_fdi_box<.., ZGEN=true>   the same columns and rows of the (<= 8) planes next to a z face: same tiling and x/y
//                taps, z coefficients per plane from a table (true boundary operators along z);
//   k_fdi_shell  the ring of columns outside that x/y range, every plane: one thread per column marching along z,
//                gather form with the true boundary operators (row masks of LC/supports/_3d_structured_grid.py:
//                164-293, border families of LI/_finite_difference_interpolator.py:1052-1144).
// The three launches write disjoint nodes and run on three streams (fork/join by events); their p.q parts are
// summed by the CG update.  k_fdi_normal (fdi_stencil.cuh) stays the single-kernel path for grids too small to
// have a box.
//
// Why: ncu of k_fdi_normal at 256^3 (profiles/r01b) shows 103-168 warp instructions and 40 fp64 operations per
// 32 nodes, two to five resident warps per scheduler, "wait" and "short scoreboard" as the top stalls: bound by
// instruction latency, not by HBM.  In k_fdi_box a thread owns a 2 (x) by YB (y) patch of the tile:
//   * every staged row is read once with three 16-byte shared loads and reused by the YB outputs that touch it
//     (3.5 LDS.128 per node pair instead of 11);
//   * the row sums p(i-2)+p(i+2) are formed once per staged row and serve the in-row tap, the xz tap and the xy
//     corner taps of the rows two above and below;
//   * 20 fp64 operations per node (25 in the z-general variant), every coefficient a launch constant, 2 YB
//     independent chains per thread, 162-167 registers, no spills;
//   * the five output-plane accumulators rotate statically (five copies of the plane body);
//   * planes are staged by 16-byte cp.async (LDGSTS) into a ring of ST planes, ST-3 of them in flight; the ring
//     keeps planes kp-1 and kp-2 resident so that p(k) for the fused p.q reduction is re-read from shared
//     memory instead of being carried in registers.
#pragma once
#include "fdi_stencil.cuh"

namespace lsb {

// coefficients of the uniform interior operator (every mask 1, no end corrections, no border rows):
// the 25-tap composite kernel of LI/_finite_difference_interpolator.py:884-1050 split into its
// same-plane part and the parts scattered to planes kp+-1, kp+-2
struct UniConsts {
  double cc;         // centre: axx + ayy + azz + 4 (bxy + bxz + byz) + ridge^2
  double cx1, cx2;   // x taps +-1: -2 axx/3 ; +-2: axx/6 - 2 (bxy + bxz)
  double cy1, cy2;   // y taps +-1: -2 ayy/3 ; +-2: ayy/6 - 2 (bxy + byz)
  double bxy;        // the four (+-2, +-2, 0) corners
  double kz, bxz, byz;  // plane kp+-2: kz p + bxz (p(i-2)+p(i+2)) + byz (p(j-2)+p(j+2)),  kz = azz/6 - 2 (bxz + byz)
  double zt1;        // plane kp+-1: -2 azz/3
  // z-general variant (chunks that touch a z face): the x/y part alone and the raw weights
  double cxy;        // centre of the xy part: axx + ayy + 4 bxy
  double cx2xy, cy2xy;  // x / y taps +-2 of the xy part: axx/6 - 2 bxy, ayy/6 - 2 bxy
  double g0;         // 2 (bxz + byz): centre of G = bxz C_x p + byz C_y p
  double azz, dz, ridge2;
};

// per-plane z coefficients (shell_fill_z_table), 16 doubles per GLOBAL plane k: [0..4] T_z taps -2..2 (1/6 included),
// [5..9] B_z taps (1/6 included), [10..12] C_z taps -2,0,+2, [13] m_z(k)
constexpr int SHELL_ZSTRIDE = 16;

// z coefficients of one incoming plane kp in the z-general variant, from the per-plane table:
// what plane kp contributes to the output planes kp-2 .. kp+2 through the z line L = azz T_z + dz B_z and C_z
constexpr int BOX_ZGEN_MAX_STEPS = 160;  // a z-general chunk has at most 156 output planes (+ 4 halo steps)
struct ZStep {
  double mz;                      // m_z(kp): mask of the xy part of output plane kp
  double l0, c0;                  // L(kp, 0), C_z(kp, 0)
  double lm1, lp1;                // L(kp-1, +1), L(kp+1, -1)
  double lm2, cm2, lp2, cp2;      // L(kp-2, +2), C_z(kp-2, +2), L(kp+2, -2), C_z(kp+2, -2)
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
  static constexpr int X_FIRST = 4;  // uniform variant: first column; columns 0..3 and nx-4..nx-1 are the x shell
};

__device__ __forceinline__ double2 lds2(const double* s) { return *reinterpret_cast<const double2*>(s); }

// One incoming plane for the thread's 2 x YB patch.  Sc points at the staged element (row j0-2, column i0-2);
// A*[r][e] accumulate the output planes kp-2..kp+2.  Every coefficient is a launch constant: 20 fp64
// operations per node, no per-thread coefficient registers.
template <int YB, int SX>
__device__ __forceinline__ void uni_plane_fast(const double* __restrict__ Sc, const UniConsts& K, double (&Am2)[YB][2],
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
      double s0 = fma(K.cc, b.x, A0[rr - 2][0]);
      s0 = fma(K.cx1, a.y + b.y, s0);
      s0 = fma(K.cx2, h2[rr][0], s0);
      double s1 = fma(K.cc, b.y, A0[rr - 2][1]);
      s1 = fma(K.cx1, b.x + c.x, s1);
      s1 = fma(K.cx2, h2[rr][1], s1);
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
      double S = fma(K.cy1, sy1, A0[r][e]);
      S = fma(K.cy2, sy2, S);
      S = fma(K.bxy, cs, S);
      double wv = K.kz * p0;
      wv = fma(K.bxz, h2[rr][e], wv);
      wv = fma(K.byz, sy2, wv);
      Am2[r][e] += wv;
      Am1[r][e] = fma(K.zt1, p0, Am1[r][e]);
      A0[r][e] = S;
      Ap1[r][e] = fma(K.zt1, p0, Ap1[r][e]);
      Ap2[r][e] += wv;
    }
  }
}

// z-general twin of uni_plane_fast for the chunks next to a z face: same x/y taps (columns and rows are interior),
// z coefficients of this step in Z.  25 fp64 operations per node.
template <int YB, int SX>
__device__ __forceinline__ void zgen_plane_fast(const double* __restrict__ Sc, const UniConsts& K, const ZStep& Z,
                                                double (&Am2)[YB][2], double (&Am1)[YB][2], double (&A0)[YB][2],
                                                double (&Ap1)[YB][2], double (&Ap2)[YB][2]) {
  constexpr int NR = YB + 4;
  double pc[NR][2], h2[NR][2], xl[YB][2];
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) {
    const double2 a = lds2(Sc + rr * SX), b = lds2(Sc + rr * SX + 2), c = lds2(Sc + rr * SX + 4);
    h2[rr][0] = a.x + c.x;
    h2[rr][1] = a.y + c.y;
    pc[rr][0] = b.x;
    pc[rr][1] = b.y;
    if (rr >= 2 && rr < YB + 2) {
      xl[rr - 2][0] = fma(K.cx2xy, h2[rr][0], fma(K.cx1, a.y + b.y, K.cxy * b.x));
      xl[rr - 2][1] = fma(K.cx2xy, h2[rr][1], fma(K.cx1, b.x + c.x, K.cxy * b.y));
    }
  }
  const double cen = K.ridge2 + Z.l0;
#pragma unroll
  for (int r = 0; r < YB; ++r) {
    const int rr = r + 2;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double p0 = pc[rr][e];
      const double sy1 = pc[rr - 1][e] + pc[rr + 1][e];
      const double sy2 = pc[rr - 2][e] + pc[rr + 2][e];
      const double cs = h2[rr - 2][e] + h2[rr + 2][e];
      double xy = fma(K.cy1, sy1, xl[r][e]);
      xy = fma(K.cy2xy, sy2, xy);
      xy = fma(K.bxy, cs, xy);
      const double G = fma(-K.byz, sy2, fma(-K.bxz, h2[rr][e], K.g0 * p0));  // (bxz C_x + byz C_y) p
      double s = fma(Z.mz, xy, A0[r][e]);
      s = fma(cen, p0, s);
      A0[r][e] = fma(Z.c0, G, s);
      Am1[r][e] = fma(Z.lm1, p0, Am1[r][e]);
      Ap1[r][e] = fma(Z.lp1, p0, Ap1[r][e]);
      Am2[r][e] = fma(Z.cm2, G, fma(Z.lm2, p0, Am2[r][e]));
      Ap2[r][e] = fma(Z.cp2, G, fma(Z.lp2, p0, Ap2[r][e]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_fdi_shell: every owned node OUTSIDE the interior box, in gather form with the true boundary operators
//   q = ridge^2 p + sum_a (a_aa m_b m_c T_a + d_a B_a) p + sum_{a<b} b_ab m_c (C_a (x) C_b) p
// (no ghost cells): T = D^T diag(m) D / 6, C = G^T diag(m) G, B = sum_{r in {1,n-2}} d_r d_r^T / 6,
// m(r) = [1 <= r <= n-2] -- the row masks of StructuredGrid.neighbour_global_indexes / FDI.assemble_borders
// (LC/supports/_3d_structured_grid.py:164-293, LI/_finite_difference_interpolator.py:1052-1144) after row
// normalisation (LI/_discrete_interpolator.py:361-366).
//
// A thread owns one xy column and marches along z: its x and y coefficient sets are computed once, the z sets
// come from a per-plane table (ShellZ), and of the 25 taps of an output plane only the 13 in-plane ones are
// loaded -- the z neighbours are the per-plane scalars (p, C_x p, C_y p) kept in a five-deep register history.
// ~100 instructions per node instead of the ~450 of the one-thread-per-node form it replaced (profiles/r01m).
// Work items: the ring columns (outside [ilo,ihi) x [jlo,jhi)) over z chunks of [klo,khi) -- every owned plane
// in the current launch plan; the kernel can also take every column of the planes outside [klo,khi) ("caps").
struct ShellBox {
  int ilo, ihi, jlo, jhi, klo, khi;  // the box that is NOT computed here (klo/khi clipped to the owned planes)
  int ring_cols;                     // columns outside [ilo,ihi) x [jlo,jhi)
  int ring_blocks;                   // CTAs per z chunk of the ring: ceil(ring_cols / NT)
  int ring_cz, ring_chunks;          // z chunking of [klo,khi)
  int cap_blocks;                    // CTAs per cap: ceil(nx*ny / NT)
  int capA0, capA1, capB0, capB1;    // owned planes below klo / from khi up (either may be empty)
  int ring_stagger;                  // k_fdi_box<RING>: 1 = CTAs of odd parity do their ring share BEFORE the box work
};

// per-plane z coefficients, 16 doubles per GLOBAL plane k: [0..4] T_z taps -2..2 (1/6 included), [5..9] B_z taps
// (1/6 included), [10..12] C_z taps -2,0,+2, [13] m_z(k)

__device__ __forceinline__ double vrow(int r, int n) { return (r >= 1 && r <= n - 2) ? 1.0 : 0.0; }
// coefficient of p(i+s) in B = sum_{r in {1,n-2}} d_r d_r^T / 6 (both rows are counted when n = 3, as the
// reference's lower and upper border families both exist then)
__host__ __device__ inline double bline(int i, int n, int s) {
  if (n < 3) return 0.0;
  double c = 0.0;
  const int t = i + s;
  if (i <= 2 && t >= 0 && t <= 2) c += (i == 1 ? -2.0 : 1.0) * (t == 1 ? -2.0 : 1.0);
  if (i >= n - 3 && t >= n - 3 && t <= n - 1) c += (i == n - 2 ? -2.0 : 1.0) * (t == n - 2 ? -2.0 : 1.0);
  return c * (1.0 / 6.0);
}
// coefficient of p(i+s) in (T p)(i), T = D^T diag(m) D / 6
__host__ __device__ inline double tline(int i, int n, int s) {
  auto v = [n](int r) { return (r >= 1 && r <= n - 2) ? 1.0 : 0.0; };
  const double a = v(i - 1), b = v(i), c = v(i + 1);
  switch (s) {
    case -2: return a * (1.0 / 6.0);
    case -1: return (a + b) * (-2.0 / 6.0);
    case 0: return (a + 4.0 * b + c) * (1.0 / 6.0);
    case 1: return (b + c) * (-2.0 / 6.0);
    default: return c * (1.0 / 6.0);
  }
}
// host: fill the per-plane table (nz * SHELL_ZSTRIDE doubles)
inline void shell_fill_z_table(int nz, double* tab) {
  for (int k = 0; k < nz; ++k) {
    double* t = tab + (size_t)k * SHELL_ZSTRIDE;
    auto v = [nz](int r) { return (r >= 1 && r <= nz - 2) ? 1.0 : 0.0; };
    for (int s = -2; s <= 2; ++s) {
      t[s + 2] = tline(k, nz, s);
      t[5 + s + 2] = bline(k, nz, s);
    }
    t[10] = -v(k - 1);
    t[11] = v(k - 1) + v(k + 1);
    t[12] = -v(k + 1);
    t[13] = v(k);
    t[14] = t[15] = 0.0;
  }
}

// column number r of the ring (outside [ilo,ihi) x [jlo,jhi)): first the full rows below / above the box, then the
// left / right column stubs of the rows in between
__device__ __forceinline__ void ring_column(const GridDesc& g, const ShellBox& sb, const int r, int* ci, int* cj) {
  const int ny_out = g.ny - (sb.jhi - sb.jlo);
  if (r < ny_out * g.nx) {
    const int jj = r / g.nx;
    *ci = r - jj * g.nx;
    *cj = jj < sb.jlo ? jj : sb.jhi + (jj - sb.jlo);
  } else {
    const int nx_out = max(g.nx - (sb.ihi - sb.ilo), 1);
    const int r2 = r - ny_out * g.nx;
    const int jr = r2 / nx_out, ii = r2 - jr * nx_out;
    *cj = sb.jlo + jr;
    *ci = ii < sb.ilo ? ii : sb.ihi + (ii - sb.ilo);
  }
}

// One thread, one xy column (ci, cj), output planes [ka, kb): the march described above.  Returns its part of p.q.
__device__ __forceinline__ double shell_march(const GridDesc& g, const StencilWeights& w, const double* __restrict__ ztab,
                                              const double* __restrict__ p, double* __restrict__ q, const int ci,
                                              const int cj, const int ka, const int kb) {
  double dot = 0.0;
  // ---- x and y coefficient sets of the column
  const double vax = vrow(ci - 1, g.nx), vbx = vrow(ci, g.nx), vcx = vrow(ci + 1, g.nx);
  const double vay = vrow(cj - 1, g.ny), vby = vrow(cj, g.ny), vcy = vrow(cj + 1, g.ny);
  const double ax6 = w.axx * vby * (1.0 / 6.0), ay6 = w.ayy * vbx * (1.0 / 6.0);
  // T lines (to be multiplied by m_z of the plane) and border lines (unmasked); tap order -2,-1,0,+1,+2
  const double tx0 = ax6 * vax, tx1 = -2.0 * ax6 * (vax + vbx), tx2 = ax6 * (vax + 4.0 * vbx + vcx),
               tx3 = -2.0 * ax6 * (vbx + vcx), tx4 = ax6 * vcx;
  const double ty0 = ay6 * vay, ty1 = -2.0 * ay6 * (vay + vby), ty2 = ay6 * (vay + 4.0 * vby + vcy),
               ty3 = -2.0 * ay6 * (vby + vcy), ty4 = ay6 * vcy;
  const bool bdx = (ci <= 2 || ci >= g.nx - 3), bdy = (cj <= 2 || cj >= g.ny - 3);
  double bx[5] = {0, 0, 0, 0, 0}, by[5] = {0, 0, 0, 0, 0};
  if (bdx)
#pragma unroll
    for (int s = 0; s < 5; ++s) bx[s] = w.dx * bline(ci, g.nx, s - 2);
  if (bdy)
#pragma unroll
    for (int s = 0; s < 5; ++s) by[s] = w.dy * bline(cj, g.ny, s - 2);
  const double cx0 = -vax, cx1 = vax + vcx, cx2 = -vcx;  // C_x taps -2, 0, +2
  const double cy0 = -vay, cy1 = vay + vcy, cy2 = -vcy;
  const double azm = w.azz * vbx * vby, bxzm = w.bxz * vby, byzm = w.byz * vbx;
  // ---- march: history of (S, c, gx, gy) for planes kp-4 .. kp
  double hS[5], hc[5], hgx[5], hgy[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) hS[t] = hc[t] = hgx[t] = hgy[t] = 0.0;
  const double* pc = p + ((long long)(ka - 2 - g.k0 + HALO) * g.plane + cj * g.pitch + ci);
  double* qc = q + ((long long)(ka - g.k0 + HALO) * g.plane + cj * g.pitch + ci);
  const int P = g.pitch;
  // clamped in-plane offsets are not needed: a tap outside the grid has coefficient 0 and the address stays
  // inside the vector's allocation (two halo planes on either side), whose contents are finite
  for (int kp = ka - 2; kp < kb + 2; ++kp) {
    double S = 0.0, c = 0.0, gx = 0.0, gy = 0.0;
    if (kp >= 0 && kp < g.nz) {
      c = pc[0];
      const double xm2 = pc[-2], xm1 = pc[-1], xp1 = pc[1], xp2 = pc[2];
      const double ym2 = pc[-2 * P], ym1 = pc[-P], yp1 = pc[P], yp2 = pc[2 * P];
      gx = fma(cx2, xp2, fma(cx0, xm2, cx1 * c));
      gy = fma(cy2, yp2, fma(cy0, ym2, cy1 * c));
      if (kp >= ka && kp < kb) {  // in-plane part of an output plane
        const double cmm = pc[-2 * P - 2], cpm = pc[-2 * P + 2], cmp = pc[2 * P - 2], cpp = pc[2 * P + 2];
        const double mz = ztab[(long long)kp * SHELL_ZSTRIDE + 13];
        double s1 = tx2 * c;
        s1 = fma(tx0, xm2, s1); s1 = fma(tx1, xm1, s1); s1 = fma(tx3, xp1, s1); s1 = fma(tx4, xp2, s1);
        s1 = fma(ty2, c, s1);
        s1 = fma(ty0, ym2, s1); s1 = fma(ty1, ym1, s1); s1 = fma(ty3, yp1, s1); s1 = fma(ty4, yp2, s1);
        // (C_x (x) C_y) p: C_y over the rows j-2, j, j+2 of (C_x p)
        const double rm = fma(cx2, cpm, fma(cx0, cmm, cx1 * ym2));
        const double rp = fma(cx2, cpp, fma(cx0, cmp, cx1 * yp2));
        s1 = fma(w.bxy, fma(cy2, rp, fma(cy0, rm, cy1 * gx)), s1);
        S = fma(mz, s1, w.ridge2 * c);
        if (bdx) S += bx[0] * xm2 + bx[1] * xm1 + bx[2] * c + bx[3] * xp1 + bx[4] * xp2;
        if (bdy) S += by[0] * ym2 + by[1] * ym1 + by[2] * c + by[3] * yp1 + by[4] * yp2;
      }
    }
    // shift the history (registers) and append plane kp
#pragma unroll
    for (int t = 0; t < 4; ++t) { hS[t] = hS[t + 1]; hc[t] = hc[t + 1]; hgx[t] = hgx[t + 1]; hgy[t] = hgy[t + 1]; }
    hS[4] = S; hc[4] = c; hgx[4] = gx; hgy[4] = gy;
    const int ko = kp - 2;  // output plane completed by plane kp: history index 2 is plane ko
    if (ko >= ka) {
      const double* z = ztab + (long long)ko * SHELL_ZSTRIDE;
      double acc = hS[2];
      // z line: a_zz m_x m_y T_z + d_z B_z
      double lz = 0.0;
#pragma unroll
      for (int t = 0; t < 5; ++t) lz = fma(fma(azm, z[t], w.dz * z[5 + t]), hc[t], lz);
      acc += lz;
      // (C_x (x) C_z) p m_y and (C_y (x) C_z) p m_x
      acc = fma(bxzm, fma(z[12], hgx[4], fma(z[10], hgx[0], z[11] * hgx[2])), acc);
      acc = fma(byzm, fma(z[12], hgy[4], fma(z[10], hgy[0], z[11] * hgy[2])), acc);
      qc[0] = acc;
      dot = fma(hc[2], acc, dot);
      qc += g.plane;
    }
    pc += g.plane;
  }
  return dot;
}

template <int NT>
__global__ void __launch_bounds__(NT, 4)
    k_fdi_shell(const GridDesc g, const StencilWeights w, const ShellBox sb, const double* __restrict__ ztab,
                const double* __restrict__ p, double* __restrict__ q, double* __restrict__ partials,
                unsigned int* __restrict__ counter, CgScalars* __restrict__ sc, const int cg_mode) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  // ---- which column and which output planes
  int ci = 0, cj = 0, ka, kb;
  bool active;
  {
    const int b = blockIdx.x;
    const int nring = sb.ring_blocks * sb.ring_chunks;
    if (b < nring) {
      const int ch = b / sb.ring_blocks;
      const int r = (b - ch * sb.ring_blocks) * NT + threadIdx.x;
      active = r < sb.ring_cols;
      ka = sb.klo + ch * sb.ring_cz;
      kb = min(ka + sb.ring_cz, sb.khi);
      if (active) ring_column(g, sb, r, &ci, &cj);
    } else {
      const int b2 = b - nring;
      const int cap = b2 / sb.cap_blocks;
      const int r = (b2 - cap * sb.cap_blocks) * NT + threadIdx.x;
      cj = r / g.nx;
      ci = r - cj * g.nx;
      active = cj < g.ny;
      ka = cap == 0 ? sb.capA0 : sb.capB0;
      kb = cap == 0 ? sb.capA1 : sb.capB1;
    }
    active = active && kb > ka;
  }
  double dot = 0.0;
  if (active) dot = shell_march(g, w, ztab, p, q, ci, cj, ka, kb);
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[2] = total;
}

// ---------------------------------------------------------------------------------------------
// k_fdi_box: the interior box 4 <= i <= nx-5, 3 <= j <= ny-4, 3 <= k <= nz-4 only.  Every input plane is a
// real plane, every coefficient a launch constant: no ghost mapping, no masks, no general path.  Everything
// outside the box is k_fdi_shell's.  Output planes [zlo, zhi) of this handle, z-chunks of cz planes.
// (A persistent-CTA variant with an equal split of the (tile, plane) list was measured and dropped: it loses the
// L2 sharing of the tile halos between CTAs that march through the same planes together -- 573 us instead
// of 375 us at 512^3, profiles/r01h.)
// ZGEN = true: the launch for the (at most two) chunks that touch a z face: blockIdx.z = 0 -> output planes [zlo,zhi),
// 1 -> [zlo2,zhi2); z coefficients per plane from ztab, planes outside the grid are skipped.  Its p.q goes to
// pq_part[3], its partial sums to its own slot.
// RING = true (LSB200_RING=1): no k_fdi_shell launch.  After its box work every CTA also computes an equal share of
// the ring columns over its own planes [ka,kb) with shell_march -- the columns of a z chunk are dealt out over the
// tiles of that chunk, each column's planes split into sub-ranges so that all threads of the CTA march at once.
// The march gathers from global memory (L2: the planes were just streamed by this chunk's CTAs), needs no shared
// memory and fewer registers than the box part, and its latency hides behind the other CTAs resident on the SM --
// which a separate shell kernel cannot do, because the box CTAs leave it no registers to become resident
// (profiles/r01r).  Its p.q goes into this launch's sum.
template <int YB, int TYT, int ST, int MINB, bool VEC16, bool ZGEN, bool RING = false>
__global__ void __launch_bounds__(32 * TYT, MINB)
    k_fdi_box(const GridDesc g, const UniConsts U, const double* __restrict__ p, double* __restrict__ q, const int cz,
              const int zlo, const int zhi, const int zlo2, const int zhi2, const double* __restrict__ ztab,
              double* __restrict__ partials, unsigned int* __restrict__ counter, CgScalars* __restrict__ sc,
              const int cg_mode, const StencilWeights w, const ShellBox sb) {
  using C = MainCfg<YB, TYT, ST>;
  static_assert(ST >= 5, "ring needs planes kp-2..kp plus at least two in flight");
  LSB_DYN_SMEM(S);
  __shared__ double red[C::NT / 32];
  __shared__ ZStep zsm[ZGEN ? BOX_ZGEN_MAX_STEPS : 1];  // z coefficients of every plane step of this CTA (ZGEN)

  if (cg_mode && sc->done) return;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + 32 * ty;
  const int x0 = C::X_FIRST + blockIdx.x * C::TX, y0 = C::Y_FIRST + blockIdx.y * C::TY;
  const int ka = ZGEN ? (blockIdx.z == 0 ? zlo : zlo2) : zlo + blockIdx.z * cz;
  const int kb = ZGEN ? (blockIdx.z == 0 ? zhi : zhi2) : min(ka + cz, zhi);
  const int i0 = x0 + 2 * tx;
  const int j0 = y0 + YB * ty;
  const int nrows = max(0, min(YB, (g.ny - 3) - j0));
  const int xlim = g.nx - C::X_FIRST;

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
      const int gi = x0 - HALO + sx, gj = y0 - HALO + sy;  // >= 2, >= 1: real nodes or tile overhang
      if (gi < g.nx && gj < g.ny) off[t] = gi + g.pitch * gj;
    }
  }
  // next plane to stage (input planes ka-2 .. kb+1 of this handle's storage)
  const double* src = p + (long long)(ka - HALO - g.k0 + HALO) * g.plane;
  int kstage = ka - HALO;  // global index of that plane
  auto issue_plane = [&](double* dstS) {
    const bool exists = !ZGEN || (kstage >= 0 && kstage < g.nz);
#pragma unroll
    for (int t = 0; t < LD && exists; ++t) {
      const int idx = tid + t * C::NT;
      if (idx < NSLOT) {
        const bool ok = off[t] >= 0;
        if (VEC16) cp_async_16(dstS + 2 * idx, src + (ok ? off[t] : 0), ok);
        else cp_async_8(dstS + idx, src + (ok ? off[t] : 0), ok);
      }
    }
    src += g.plane;
    ++kstage;
  };

  double A[5][YB][2];
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int r = 0; r < YB; ++r) A[d][r][0] = A[d][r][1] = 0.0;
  double dot = 0.0;

  // this CTA's share of the ring columns over its planes [ka,kb) (RING only)
  auto ring_phase = [&]() {
    const int ntiles = gridDim.x * gridDim.y;
    const int tile = blockIdx.x + gridDim.x * blockIdx.y;
    const int per = (sb.ring_cols + ntiles - 1) / ntiles;
    const int c0 = tile * per, ncol = min(c0 + per, sb.ring_cols) - c0;
    const int nplanes = kb - ka;
    if (ncol > 0 && nplanes > 0) {
      // sub-ranges of at least RING_MIN_PLANES planes (each re-reads 4 halo planes), enough of them to occupy the CTA
      constexpr int RING_MIN_PLANES = 6;
      int nsub = max(1, min(C::NT / ncol, nplanes / RING_MIN_PLANES));
      const int sub = (nplanes + nsub - 1) / nsub;
      nsub = (nplanes + sub - 1) / sub;
      for (int it = tid; it < ncol * nsub; it += C::NT) {
        const int sidx = it / ncol;
        int ci, cj;
        ring_column(g, sb, c0 + (it - sidx * ncol), &ci, &cj);  // consecutive threads take consecutive columns
        const int za = ka + sidx * sub;
        dot += shell_march(g, w, ztab, p, q, ci, cj, za, min(za + sub, kb));
      }
    }
  };
  // In a launch of a single wave every CTA would reach a trailing ring phase at the same time, with nothing left to
  // hide its latency behind: with ring_stagger the CTAs of odd parity march first and stream afterwards, so that at any
  // time about half of the resident CTAs keep HBM busy.
  const bool ring_first = RING && sb.ring_stagger && ((blockIdx.x + blockIdx.y + blockIdx.z) & 1);
  if (RING && ring_first) ring_phase();

  // an empty range (ka >= kb) stages and computes nothing, but still takes part in the grid reduction
  const int kstart = ka - HALO, kend = ka < kb ? kb + HALO : ka - HALO;
  int kp = kstart;
  int issued = kstart;  // next plane to stage
#pragma unroll
  for (int s = 0; s < ST - 3; ++s) {
    if (issued < kend) { issue_plane(S + s * C::PLANE); ++issued; }
    cp_async_commit();
  }
  int slot = 0;
  const int my_off = (YB * ty) * C::SX + 2 * tx;
  double* qdst = q + (long long)(ka - g.k0 + HALO) * g.plane + (long long)g.pitch * j0 + i0;  // output plane ka

  auto advance_pipeline = [&]() {
    cp_async_wait<ST - 4>();
    __syncthreads();
    const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;
    if (issued < kend) { issue_plane(S + fs * C::PLANE); ++issued; }
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

  // z coefficients of incoming plane k (ZGEN): rows k-2..k+2 of the per-plane table; a row outside the grid gives 0
  auto zstep = [&](const int k) {
    auto row = [&](const int ko, const int col) {
      return (ko >= 0 && ko < g.nz) ? ztab[(long long)ko * SHELL_ZSTRIDE + col] : 0.0;
    };
    // L(ko, s) = azz T_z(ko, s) + dz B_z(ko, s), tap s = k - ko at column 2 + s (T) / 7 + s (B); C_z taps at 10,11,12
    auto L = [&](const int ko) { return fma(U.azz, row(ko, 2 + (k - ko)), U.dz * row(ko, 7 + (k - ko))); };
    ZStep Z;
    Z.mz = row(k, 13);
    Z.l0 = L(k);
    Z.c0 = row(k, 11);
    Z.lm1 = L(k - 1);
    Z.lp1 = L(k + 1);
    Z.lm2 = L(k - 2);
    Z.cm2 = row(k - 2, 12);   // C_z(k-2, +2)
    Z.lp2 = L(k + 2);
    Z.cp2 = row(k + 2, 10);   // C_z(k+2, -2)
    return Z;
  };
  if (ZGEN) {
    for (int t = tid; t < kend - kstart && t < BOX_ZGEN_MAX_STEPS; t += C::NT) zsm[t] = zstep(kstart + t);
    // visible after the barrier of the first advance_pipeline()
  }
#define LSB_BOX_STEP(PH)                                                                                      \
  {                                                                                                           \
    if (kp >= kend) break;                                                                                    \
    advance_pipeline();                                                                                       \
    if (ZGEN) {                                                                                               \
      if (kp >= 0 && kp < g.nz) {                                                                             \
        const ZStep Z = zsm[kp - kstart];                                                                     \
        zgen_plane_fast<YB, C::SX>(S + slot * C::PLANE + my_off, U, Z, A[(PH + 0) % 5], A[(PH + 1) % 5],      \
                                   A[(PH + 2) % 5], A[(PH + 3) % 5], A[(PH + 4) % 5]);                        \
      }                                                                                                       \
    } else {                                                                                                  \
      uni_plane_fast<YB, C::SX>(S + slot * C::PLANE + my_off, U, A[(PH + 0) % 5], A[(PH + 1) % 5],            \
                                A[(PH + 2) % 5], A[(PH + 3) % 5], A[(PH + 4) % 5]);                           \
    }                                                                                                         \
    store_out(A[(PH + 0) % 5]);                                                                               \
    _Pragma("unroll") for (int r = 0; r < YB; ++r) A[(PH + 0) % 5][r][0] = A[(PH + 0) % 5][r][1] = 0.0;       \
    slot = (slot + 1 == ST) ? 0 : slot + 1;                                                                   \
    ++kp;                                                                                                     \
  }
  if (ka < kb) {
    for (;;) {
      LSB_BOX_STEP(0)
      LSB_BOX_STEP(1)
      LSB_BOX_STEP(2)
      LSB_BOX_STEP(3)
      LSB_BOX_STEP(4)
    }
  }
#undef LSB_BOX_STEP
  cp_async_wait<0>();

  if (RING && !ring_first) ring_phase();

  const double bsum = block_sum<C::NT>(dot, red);
  const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  double total;
  if (grid_sum_last_block<C::NT>(bsum, partials, counter, nblocks, bid, red, &total)) sc->pq_part[ZGEN ? 3 : 0] = total;
}



}  // namespace lsb

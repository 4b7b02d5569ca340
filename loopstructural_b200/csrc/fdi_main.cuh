// K1, large grids: building blocks of the FDI regularisation normal operator q = (R^T R + ridge^2 I) p as k_fdi_fused
// (csrc/fdi_fused.cuh) applies it -- the register-tiled plane step of the interior, where R^T R is the constant 25-tap
// composite kernel of LI/_finite_difference_interpolator.py:884-1050, and the per-plane z coefficient table of the
// planes next to a z face.  (Round 1 launched three kernels on three streams from this file: k_fdi_box for the interior,
// its z-general twin for the z faces and the gather kernel k_fdi_shell for the x/y ring; the single launch replaced them
// once it had won on every grid, profiles/r2h_matvec_plans.jsonl.)  k_fdi_normal (fdi_stencil.cuh) stays the
// single-kernel path for grids too small to have a box.
//
// Why a register tile: ncu of k_fdi_normal at 256^3 (profiles/r01b) shows 103-168 warp instructions and 40 fp64
// operations per 32 nodes, two to five resident warps per scheduler, "wait" and "short scoreboard" as the top stalls:
// bound by instruction latency, not by HBM.  In the plane step a thread owns a 2 (x) by YB (y) patch of the tile:
//   * every staged row is read once with three 16-byte shared loads and reused by the YB outputs that touch it
//     (3.5 LDS.128 per node pair instead of 11);
//   * the row sums p(i-2)+p(i+2) are formed once per staged row and serve the in-row tap, the xz tap and the xy
//     corner taps of the rows two above and below;
//   * 20 fp64 operations per node, every coefficient a launch constant, 2 YB independent chains per thread;
//   * the five output-plane accumulators rotate statically (five copies of the plane body);
//   * the ring keeps planes kp-1 and kp-2 resident so that p(k) for the fused p.q reduction is re-read from shared
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

// ---------------------------------------------------------------------------------------------
// The true boundary operators along one axis (no ghost cells): T = D^T diag(m) D / 6, C = G^T diag(m) G,
// B = sum_{r in {1,n-2}} d_r d_r^T / 6, m(r) = [1 <= r <= n-2] -- the row masks of
// StructuredGrid.neighbour_global_indexes / FDI.assemble_borders (LC/supports/_3d_structured_grid.py:164-293,
// LI/_finite_difference_interpolator.py:1052-1144) after row normalisation (LI/_discrete_interpolator.py:361-366).
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

}  // namespace lsb

// K1: matrix-free normal operator of the FDI regularisation, q = (R^T R + ridge^2 I) p.
//
// What it replaces: the twelve stencil families FiniteDifferenceInterpolator assembles as explicit
// rows (assemble_inner / assemble_borders / _assemble_operator,
// LI/_finite_difference_interpolator.py:643-692,1052-1164) and the product A^T A scipy then forms
// (LI/_solver_strategy.py:64).  The reference's own matrix-free statement of the same operator is
// _build_fused_cg_regularisation_operator (LI/_finite_difference_interpolator.py:884-1050): a 5x5x5
// composite kernel in the deep interior plus an exact two-pass shell.
//
// Formulation used here (exact at every node, no shell special case).  After row normalisation
// (LI/_discrete_interpolator.py:361-366) the families are
//   daa : rows  d = [1,-2,1]/sqrt(6) along axis a, one row per node interior on ALL axes
//   dab : rows  -1/2 G_a (x) G_b, G = [-1,0,1], same row set
//   da_lower/upper: rows d along a on the planes index 1 and n_a-2, all positions on the other axes
// With m_a(i) = [1 <= i <= n_a-2] the row mask factorises, so
//   R^T R = sum_a  ew_aa^2 m_b m_c T_a  +  sum_{a<b} ew_ab^2/4 m_c C_a (x) C_b  +  sum_a ew_da^2 B_a
//   T_a = D^T diag(m_a) D / 6 (1-D, 5 taps),  C_a = G^T diag(m_a) G (1-D, taps -2,0,2),
//   B_a = sum_{r in {1,n_a-2}} d_r d_r^T.
// The 1-D boundary rows of C_a are reproduced exactly by reading ghost values
//   p(-1):=p(1), p(-2):=p(0), p(n):=p(n-2), p(n+1):=p(n-1)
// (they make G p vanish on the invalid rows r=-1,0,n-1,n); with the same ghosts T_a differs from
// the constant 5-tap stencil only at the two nodes next to each end, by
//   q(0) -= (p0-p1),  q(1) += (p0-p1)/3          (and mirrored at the upper end).
// So the kernel is the constant-coefficient composite stencil on ghost-extended data, three 0/1
// masks, and a few predicated corrections executed by boundary lanes only.
//
// Data movement: z-marching.  A CTA owns an xy tile and a chunk of planes; every incoming plane
// (tile + halo 2) is staged once in shared memory, each thread scatters its contribution into
// five register accumulators (output planes k'-2..k'+2), so p is read once and q written once:
// 16 B/dof algorithmic traffic.
#pragma once
#include "common.cuh"

namespace lsb {

// ghost index map of the header comment; -1 = no such node (tile overhang)
__device__ __forceinline__ int ghost_map(int i, int n) {
  if (i >= 0 && i < n) return i;
  if (i == -1) return 1;
  if (i == -2) return 0;
  if (i == n) return n - 2;
  if (i == n + 1) return n - 1;
  return -1;
}

template <int XB, int TXT, int TYT, int ST = 4>
struct StencilCfg {
  static constexpr int NT = TXT * TYT;
  static constexpr int TX = TXT * XB;       // tile width in nodes
  static constexpr int TY = TYT;
  static constexpr int SX = TX + 2 * HALO;  // staged row length (even)
  static constexpr int SY = TY + 2 * HALO;
  static constexpr int PLANE = SX * SY;
  static constexpr int STAGES = ST;         // ring of staged planes (ST-1 in flight)
  static constexpr int LD8 = (PLANE + NT - 1) / NT;        // 8-byte copies per thread and plane
  static constexpr int LD16 = (PLANE / 2 + NT - 1) / NT;   // 16-byte copies per thread and plane
};

// cp.async (LDGSTS): global -> shared without a register round trip, zero-filled when !valid
#ifdef LSB_EMU
inline void cp_async_8(double* smem_dst, const double* gsrc, bool valid) { smem_dst[0] = valid ? gsrc[0] : 0.0; }
inline void cp_async_16(double* smem_dst, const double* gsrc, bool valid) {
  smem_dst[0] = valid ? gsrc[0] : 0.0;
  smem_dst[1] = valid ? gsrc[1] : 0.0;
}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async_8(double* smem_dst, const double* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_16(double* smem_dst, const double* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

// the five staged rows a thread needs for its XB outputs: rc/rm2/rp2 hold columns i0-2..i0+XB+1 of
// rows j, j-2, j+2; rm1/rp1 hold columns i0..i0+XB-1 of rows j-1, j+1
template <int XB, int SX>
__device__ __forceinline__ void load_rows(const double* __restrict__ Sc, double (&rc)[XB + 4], double (&rm2)[XB + 4],
                                          double (&rp2)[XB + 4], double (&rm1)[XB], double (&rp1)[XB]) {
  if constexpr ((XB & 1) == 0) {
    // 16-byte shared loads: lanes are 16 B apart, conflict free
#pragma unroll
    for (int e = 0; e < XB + 4; e += 2) {
      const double2 a = *reinterpret_cast<const double2*>(Sc + e);
      const double2 b = *reinterpret_cast<const double2*>(Sc + e - 2 * SX);
      const double2 c = *reinterpret_cast<const double2*>(Sc + e + 2 * SX);
      rc[e] = a.x, rc[e + 1] = a.y;
      rm2[e] = b.x, rm2[e + 1] = b.y;
      rp2[e] = c.x, rp2[e + 1] = c.y;
    }
#pragma unroll
    for (int e = 0; e < XB; e += 2) {
      const double2 a = *reinterpret_cast<const double2*>(Sc + e + 2 - SX);
      const double2 b = *reinterpret_cast<const double2*>(Sc + e + 2 + SX);
      rm1[e] = a.x, rm1[e + 1] = a.y;
      rp1[e] = b.x, rp1[e + 1] = b.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < XB + 4; ++e) {
      rc[e] = Sc[e];
      rm2[e] = Sc[e - 2 * SX];
      rp2[e] = Sc[e + 2 * SX];
    }
#pragma unroll
    for (int e = 0; e < XB; ++e) {
      rm1[e] = Sc[e + 2 - SX];
      rp1[e] = Sc[e + 2 + SX];
    }
  }
}

// ---- 1-D coefficient functions (thread-constant, evaluated once in the prologue) -------------
// coefficient of phat(i+s), s=-2..2, in the ghost-extended, end-corrected operator T_a (see header)
__device__ __forceinline__ double coef_T(int i, int n, int s) {
  double c = (s == 0) ? 1.0 : ((s == 1 || s == -1) ? (-2.0 / 3.0) : (1.0 / 6.0));
  if (i == 0) { if (s == 0) c -= 1.0; if (s == 1) c += 1.0; }
  if (i == 1) { if (s == -1) c += 1.0 / 3.0; if (s == 0) c -= 1.0 / 3.0; }
  if (i == n - 1) { if (s == 0) c -= 1.0; if (s == -1) c += 1.0; }
  if (i == n - 2) { if (s == 1) c += 1.0 / 3.0; if (s == 0) c -= 1.0 / 3.0; }
  return c;
}
// coefficient of p(i+s) in B_a = sum_{r in {1,n-2}} d_r d_r^T (d = [1,-2,1]/sqrt 6)
__device__ __forceinline__ double coef_B(int i, int n, int s) {
  if (n < 3) return 0.0;
  double c = 0.0;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const int r = which ? n - 2 : 1;
    const int o = r - i;   // row centre relative to this node
    const int t = s - o;   // tap relative to the row centre
    if (o >= -1 && o <= 1 && t >= -1 && t <= 1) c += (o == 0 ? -2.0 : 1.0) * (t == 0 ? -2.0 : 1.0);
  }
  return c * (1.0 / 6.0);
}

// Per-thread constants of the fast path.
template <int XB>
struct StencilCoef {
  double cx[XB][5];  // same-plane x taps -2..2; [2] is the merged centre coefficient
  double cyT[4];     // y taps -2,-1,+1,+2 of (ayy T_y - 2 byz [|t|=2]); multiplied by mx at use
  double mx[XB];     // 0/1 mask of the output column
  double zt0, zt1;   // azz my * {1/6, -2/3}; multiplied by mx at use
  double ua;         // bxz * my
};

// One incoming plane, regular case (3 <= kp <= nz-2): all masks and end corrections are folded
// into thread-constant coefficients.  am2..ap2 accumulate the output planes kp-2..kp+2.
template <int XB>
__device__ __forceinline__ void plane_fast(const StencilCoef<XB>& C, const StencilWeights& w, const GridDesc& g,
                                           const bool ybnd, const int j, const double (&rc)[XB + 4],
                                           const double (&rm2)[XB + 4], const double (&rp2)[XB + 4],
                                           const double (&rm1)[XB], const double (&rp1)[XB], double (&am2)[XB],
                                           double (&am1)[XB], double (&a0)[XB], double (&ap1)[XB],
                                           double (&ap2)[XB]) {
#pragma unroll
  for (int e = 0; e < XB; ++e) {
    const double pc = rc[e + 2];
    double S = C.cx[e][0] * rc[e];
    S = fma(C.cx[e][1], rc[e + 1], S);
    S = fma(C.cx[e][2], pc, S);
    S = fma(C.cx[e][3], rc[e + 3], S);
    S = fma(C.cx[e][4], rc[e + 4], S);
    const double sy2 = rm2[e + 2] + rp2[e + 2];
    const double sx2 = rc[e] + rc[e + 4];
    const double cs = (rm2[e] + rm2[e + 4]) + (rp2[e] + rp2[e + 4]);
    S = fma(w.bxy, fma(-2.0, sy2, cs), S);
    // everything masked by mx: y line, yz mixed term, z line
    double Y = C.cyT[0] * rm2[e + 2];
    Y = fma(C.cyT[1], rm1[e], Y);
    Y = fma(C.cyT[2], rp1[e], Y);
    Y = fma(C.cyT[3], rp2[e + 2], Y);
    S = fma(C.mx[e], Y, S);
    if (ybnd) {  // warp uniform: dy border rows touch the three rows next to either y end
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const double v = t == 0 ? rm2[e + 2] : t == 1 ? rm1[e] : t == 2 ? pc : t == 3 ? rp1[e] : rp2[e + 2];
        S = fma(w.dy * coef_B(j, g.ny, t - 2), v, S);
      }
    }
    const double pm = C.mx[e] * pc;
    const double uym = C.mx[e] * fma(2.0, pc, -sy2);
    const double u = fma(C.ua, fma(2.0, pc, -sx2), w.byz * uym);
    am2[e] = fma(C.zt0, pm, am2[e]) - u;
    am1[e] = fma(C.zt1, pm, am1[e]);
    a0[e] += S;
    ap1[e] = fma(C.zt1, pm, ap1[e]);
    ap2[e] = fma(C.zt0, pm, ap2[e]) - u;
  }
}

// One incoming plane, general case (ghost planes, mz = 0 planes, z-end corrections): every mask
// explicit.  Runs on at most 5 planes at either end of the z range.
template <int XB>
__device__ __forceinline__ void plane_general(const GridDesc& g, const StencilWeights& w, const int kp, const int i0,
                                              const int j, const double (&rc)[XB + 4], const double (&rm2)[XB + 4],
                                              const double (&rp2)[XB + 4], const double (&rm1)[XB],
                                              const double (&rp1)[XB], const double (&prev1)[XB],
                                              const double (&prev2)[XB], double (&am2)[XB], double (&am1)[XB],
                                              double (&a0)[XB], double (&ap1)[XB], double (&ap2)[XB]) {
  constexpr double t0 = 1.0 / 6.0, t1 = -2.0 / 3.0;
  const double mz = (kp > 0 && kp < g.nz - 1) ? 1.0 : 0.0;
  const double my = (j > 0 && j < g.ny - 1) ? 1.0 : 0.0;
#pragma unroll
  for (int e = 0; e < XB; ++e) {
    const int i = i0 + e;
    const double mx = (i > 0 && i < g.nx - 1) ? 1.0 : 0.0;
    const double pc = rc[e + 2];
    double X = 0.0, Y = 0.0, bord = 0.0;
#pragma unroll
    for (int s = -2; s <= 2; ++s) {
      const double vx = rc[e + 2 + s];
      const double vy = s == -2 ? rm2[e + 2] : s == -1 ? rm1[e] : s == 0 ? pc : s == 1 ? rp1[e] : rp2[e + 2];
      X = fma(coef_T(i, g.nx, s), vx, X);
      Y = fma(coef_T(j, g.ny, s), vy, Y);
      bord = fma(w.dx * coef_B(i, g.nx, s), vx, bord);
      bord = fma(w.dy * coef_B(j, g.ny, s), vy, bord);
    }
    const double sx2 = rc[e] + rc[e + 4], sy2 = rm2[e + 2] + rp2[e + 2];
    const double cs = (rm2[e] + rm2[e + 4]) + (rp2[e] + rp2[e + 4]);
    const double XY = 4.0 * pc - 2.0 * (sx2 + sy2) + cs;
    const double same = mz * (w.axx * my * X + w.ayy * mx * Y + w.bxy * XY) + w.ridge2 * pc + bord;
    const double u = w.bxz * my * (2.0 * pc - sx2) + w.byz * mx * (2.0 * pc - sy2);
    const double azm = w.azz * mx * my;
    const double zc = azm * pc;
    am2[e] += t0 * zc - u;
    am1[e] += t1 * zc;
    a0[e] += same + zc + 2.0 * u;
    ap1[e] += t1 * zc;
    ap2[e] += t0 * zc - u;
    // z ends: corrections of T_z and the dz border rows
    if (kp == 1) {
      const double dh = prev1[e] - pc;  // p(0)-p(1)
      am1[e] -= azm * dh;
      a0[e] += azm * dh * (1.0 / 3.0);
    }
    if (kp == g.nz - 1) {
      const double dh = pc - prev1[e];  // p(nz-1)-p(nz-2)
      a0[e] -= azm * dh;
      am1[e] += azm * dh * (1.0 / 3.0);
    }
    if (g.nz >= 3) {
      const double ez = (prev2[e] - 2.0 * prev1[e] + pc) * (w.dz * (1.0 / 6.0));
      if (kp == 2) { am2[e] += ez; am1[e] -= 2.0 * ez; a0[e] += ez; }
      if (kp == g.nz - 1) { am2[e] += ez; am1[e] -= 2.0 * ez; a0[e] += ez; }
    }
  }
}

template <int XB>
__device__ __forceinline__ double store_plane(const GridDesc& g, double* __restrict__ q, const int ko, const int j,
                                              const int i0, const double (&out)[XB], const double (&pk)[XB]) {
  double dot = 0.0;
  double* dst = q + (long long)(ko - g.k0 + HALO) * g.plane + g.pitch * j + i0;
  if (XB == 2 && (g.pitch & 1) == 0 && i0 + 1 < g.nx) {
    *reinterpret_cast<double2*>(dst) = make_double2(out[0], out[XB - 1]);
    dot = fma(pk[0], out[0], pk[XB - 1] * out[XB - 1]);
  } else {
#pragma unroll
    for (int e = 0; e < XB; ++e)
      if (i0 + e < g.nx) {
        dst[e] = out[e];
        dot = fma(pk[e], out[e], dot);
      }
  }
  return dot;
}

// VEC16: planes are staged with 16-byte cp.async (needs an even row pitch: every aligned pair of the
// ghost-extended row then maps to an aligned contiguous pair of real nodes); otherwise 8-byte copies.
template <int XB, int TXT, int TYT, int MINB, int ST, bool VEC16>
__global__ void __launch_bounds__(TXT* TYT, MINB)
    k_fdi_normal(const GridDesc g, const StencilWeights w, const double* __restrict__ p, double* __restrict__ q,
                 const int cz, double* __restrict__ partials, unsigned int* __restrict__ counter,
                 CgScalars* __restrict__ sc, const int cg_mode) {
  using C = StencilCfg<XB, TXT, TYT, ST>;
  LSB_DYN_SMEM(S);  // STAGES planes
  __shared__ double red[C::NT / 32];

  if (cg_mode && sc->done) return;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + TXT * ty;
  const int x0 = blockIdx.x * C::TX, y0 = blockIdx.y * C::TY;
  const int ka = g.k0 + blockIdx.z * cz;
  const int kb = min(ka + cz, g.k1);

  // per-thread copy descriptors: staged slot -> in-plane global offset of its ghost-mapped source
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
      const int gi = ghost_map(x0 - HALO + sx, g.nx), gj = ghost_map(y0 - HALO + sy, g.ny);
      if (gi >= 0 && gj >= 0) off[t] = gi + g.pitch * gj;
    }
  }
  auto issue_plane = [&](const int kp, double* dstS) {
    const int ks = ghost_map(kp, g.nz);  // always valid: kp within 2 of [0,nz)
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

  const int j = y0 + ty;
  const int i0 = x0 + XB * tx;
  const bool ybnd = (j <= 2 || j >= g.ny - 3);
  const bool row_ok = j < g.ny;

  // canonical accumulators: A[d] collects output plane kp-2+d of the next incoming plane kp;
  // P1/P2 are this thread's p at planes kp-1 / kp-2
  double A[5][XB], P1[XB], P2[XB];
#pragma unroll
  for (int e = 0; e < XB; ++e) {
    P1[e] = P2[e] = 0.0;
#pragma unroll
    for (int d = 0; d < 5; ++d) A[d][e] = 0.0;
  }
  double dot = 0.0;

  const int kend = kb + HALO;
  int kp = ka - HALO;

  // pipeline prologue: ST-1 planes in flight
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) {
    if (kp + s < kend) issue_plane(kp + s, S + s * C::PLANE);
    cp_async_commit();
  }
  int slot = 0;  // ring slot of plane kp
  const int my_off = (ty + HALO) * C::SX + XB * tx;

  // wait for plane kp, then refill the slot consumed by the previous step with plane kp+ST-1
  auto advance_pipeline = [&]() {
    cp_async_wait<ST - 2>();
    __syncthreads();
    const int fs = slot == 0 ? ST - 1 : slot - 1;
    if (kp + ST - 1 < kend) issue_plane(kp + ST - 1, S + fs * C::PLANE);
    cp_async_commit();
  };
  auto next_slot = [&]() { slot = (slot + 1 == ST) ? 0 : slot + 1; };

  // general planes with shifting accumulators (z ends only)
  auto slow_planes = [&](const int kstop) {
    for (; kp < kstop; ++kp) {
      advance_pipeline();
      double rc[XB + 4], rm2[XB + 4], rp2[XB + 4], rm1[XB], rp1[XB];
      load_rows<XB, C::SX>(S + slot * C::PLANE + my_off, rc, rm2, rp2, rm1, rp1);
      plane_general<XB>(g, w, kp, i0, j, rc, rm2, rp2, rm1, rp1, P1, P2, A[0], A[1], A[2], A[3], A[4]);
      if (kp - HALO >= ka && row_ok) dot += store_plane<XB>(g, q, kp - HALO, j, i0, A[0], P2);
#pragma unroll
      for (int e = 0; e < XB; ++e) {
        A[0][e] = A[1][e]; A[1][e] = A[2][e]; A[2][e] = A[3][e]; A[3][e] = A[4][e]; A[4][e] = 0.0;
        P2[e] = P1[e];
        P1[e] = rc[e + 2];
      }
      next_slot();
    }
  };

  slow_planes(min(3, kend));

  // regular planes: five statically rotated copies of the plane body, no accumulator moves
  const int kfast = min(g.nz - 1, kend);  // exclusive
  if (kp < kfast) {
    StencilCoef<XB> K;
    {
      const double my = (j > 0 && j < g.ny - 1) ? 1.0 : 0.0;
      K.ua = w.bxz * my;
      K.zt0 = w.azz * my * (1.0 / 6.0);
      K.zt1 = w.azz * my * (-2.0 / 3.0);
      K.cyT[0] = w.ayy * coef_T(j, g.ny, -2) - 2.0 * w.byz;
      K.cyT[1] = w.ayy * coef_T(j, g.ny, -1);
      K.cyT[2] = w.ayy * coef_T(j, g.ny, 1);
      K.cyT[3] = w.ayy * coef_T(j, g.ny, 2) - 2.0 * w.byz;
#pragma unroll
      for (int e = 0; e < XB; ++e) {
        const int i = i0 + e;
        const double mx = (i > 0 && i < g.nx - 1) ? 1.0 : 0.0;
        K.mx[e] = mx;
#pragma unroll
        for (int s = 0; s < 5; ++s) K.cx[e][s] = my * w.axx * coef_T(i, g.nx, s - 2) + w.dx * coef_B(i, g.nx, s - 2);
        K.cx[e][0] -= 2.0 * (w.bxy + K.ua);
        K.cx[e][4] -= 2.0 * (w.bxy + K.ua);
        // centre: y-line centre, mixed-term centres, z-line centre, ridge
        K.cx[e][2] += mx * w.ayy * coef_T(j, g.ny, 0) + 4.0 * (w.bxy + K.ua + w.byz * mx) + w.azz * mx * my + w.ridge2;
      }
    }
    double H[5][XB];
#pragma unroll
    for (int e = 0; e < XB; ++e) {
      H[4][e] = P1[e];
      H[3][e] = P2[e];
      H[0][e] = H[1][e] = H[2][e] = 0.0;
    }
    int phase = 0;
#define LSB_FAST_STEP(PH)                                                                                     \
  {                                                                                                           \
    if (kp >= kfast) { phase = PH; break; }                                                                   \
    advance_pipeline();                                                                                       \
    double rc[XB + 4], rm2[XB + 4], rp2[XB + 4], rm1[XB], rp1[XB];                                            \
    load_rows<XB, C::SX>(S + slot * C::PLANE + my_off, rc, rm2, rp2, rm1, rp1);                               \
    plane_fast<XB>(K, w, g, ybnd, j, rc, rm2, rp2, rm1, rp1, A[(PH + 0) % 5], A[(PH + 1) % 5], A[(PH + 2) % 5], \
                   A[(PH + 3) % 5], A[(PH + 4) % 5]);                                                         \
    if (kp - HALO >= ka && row_ok) dot += store_plane<XB>(g, q, kp - HALO, j, i0, A[(PH + 0) % 5], H[(PH + 3) % 5]); \
    _Pragma("unroll") for (int e = 0; e < XB; ++e) {                                                          \
      A[(PH + 0) % 5][e] = 0.0;                                                                               \
      H[PH][e] = rc[e + 2];                                                                                   \
    }                                                                                                         \
    next_slot();                                                                                              \
    ++kp;                                                                                                     \
  }
    for (;;) {
      LSB_FAST_STEP(0)
      LSB_FAST_STEP(1)
      LSB_FAST_STEP(2)
      LSB_FAST_STEP(3)
      LSB_FAST_STEP(4)
    }
#undef LSB_FAST_STEP
    // back to canonical order for the trailing general planes
    double T[5][XB];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
      for (int e = 0; e < XB; ++e) T[d][e] = A[d][e];
#pragma unroll
    for (int ph = 0; ph < 5; ++ph)
      if (phase == ph) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
          for (int e = 0; e < XB; ++e) A[d][e] = T[(ph + d) % 5][e];
#pragma unroll
        for (int e = 0; e < XB; ++e) {
          P1[e] = H[(ph + 4) % 5][e];
          P2[e] = H[(ph + 3) % 5][e];
        }
      }
  }

  slow_planes(kend);
  cp_async_wait<0>();

  // p . q over the nodes this CTA wrote, reduced in a fixed order
  const double bsum = block_sum<C::NT>(dot, red);
  const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  double total;
  if (grid_sum_last_block<C::NT>(bsum, partials, counter, nblocks, bid, red, &total)) {
    sc->pq_part[0] = total;
  }
}

// diag(R^T R + ridge^2 I), analytic (used for the column scaling,
// LI/_discrete_interpolator.py:815-833, and checked against the explicit Gram diagonal).
__device__ __forceinline__ double tii(int i, int n) {  // 6 * T[i,i]
  auto v = [&](int r) { return (r >= 1 && r <= n - 2) ? 1.0 : 0.0; };
  return v(i - 1) + 4.0 * v(i) + v(i + 1);
}
__device__ __forceinline__ double cii(int i, int n) {  // C[i,i]
  auto v = [&](int r) { return (r >= 1 && r <= n - 2) ? 1.0 : 0.0; };
  return v(i - 1) + v(i + 1);
}
__device__ __forceinline__ double bii(int i, int n) {  // 6 * B[i,i]
  if (n < 3) return 0.0;
  double s = 0.0;
  for (int which = 0; which < 2; ++which) {
    const int r = which ? n - 2 : 1;
    const int o = i - r;
    if (o == 0) s += 4.0;
    else if (o == 1 || o == -1) s += 1.0;
  }
  return s;
}

__global__ void k_fdi_diag(const GridDesc g, const StencilWeights w, double* __restrict__ diag) {
  const long long n = (long long)g.nx * g.ny * g.nzl;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.nx);
    const long long r = t / g.nx;
    const int j = (int)(r % g.ny);
    const int k = (int)(r / g.ny) + g.k0;
    const double mx = (i > 0 && i < g.nx - 1), my = (j > 0 && j < g.ny - 1), mz = (k > 0 && k < g.nz - 1);
    double d = w.ridge2;
    d += (w.axx * my * mz * tii(i, g.nx) + w.ayy * mx * mz * tii(j, g.ny) + w.azz * mx * my * tii(k, g.nz)) * (1.0 / 6.0);
    d += (w.dx * bii(i, g.nx) + w.dy * bii(j, g.ny) + w.dz * bii(k, g.nz)) * (1.0 / 6.0);
    d += w.bxy * mz * cii(i, g.nx) * cii(j, g.ny) + w.bxz * my * cii(i, g.nx) * cii(k, g.nz) +
         w.byz * mx * cii(j, g.ny) * cii(k, g.nz);
    diag[(long long)(k - g.k0 + HALO) * g.plane + (long long)g.pitch * j + i] += d;
  }
}

}  // namespace lsb

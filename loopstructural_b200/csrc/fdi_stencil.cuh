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

template <int XB, int TXT, int TYT>
struct StencilCfg {
  static constexpr int NT = TXT * TYT;
  static constexpr int TX = TXT * XB;       // tile width in nodes
  static constexpr int TY = TYT;
  static constexpr int SX = TX + 2 * HALO;  // staged row length
  static constexpr int SY = TY + 2 * HALO;
  static constexpr int PLANE = SX * SY;
  static constexpr int LD = (PLANE + NT - 1) / NT;  // staged elements per thread
};

// border-family contribution at 1-D position i for the rows r in {1, n-2}; v[o+2] = p(i+o), o=-2..2
__device__ __forceinline__ double border_1d(int i, int n, const double v[5]) {
  double s = 0.0;
  if (n < 3) return s;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const int r = which ? n - 2 : 1;
    const int o = r - i;  // row centre relative to this node
    if (o >= -1 && o <= 1) {
      const double e = v[o + 1] - 2.0 * v[o + 2] + v[o + 3];
      s += (o == 0 ? -2.0 : 1.0) * e;
    }
  }
  return s * (1.0 / 6.0);
}

template <int XB, int TXT, int TYT>
__global__ void __launch_bounds__(TXT* TYT)
    k_fdi_normal(const GridDesc g, const StencilWeights w, const double* __restrict__ p, double* __restrict__ q,
                 const int cz, double* __restrict__ partials, unsigned int* __restrict__ counter,
                 CgScalars* __restrict__ sc, const int cg_mode) {
  using C = StencilCfg<XB, TXT, TYT>;
  __shared__ __align__(16) double S[2][C::PLANE];
  __shared__ double red[C::NT / 32];

  if (cg_mode && sc->done) return;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + TXT * ty;
  const int x0 = blockIdx.x * C::TX, y0 = blockIdx.y * C::TY;
  const int ka = g.k0 + blockIdx.z * cz;
  const int kb = min(ka + cz, g.k1);

  // staged element -> in-plane global offset (thread constant)
  long long off[C::LD];
#pragma unroll
  for (int t = 0; t < C::LD; ++t) {
    const int idx = tid + t * C::NT;
    off[t] = -1;
    if (idx < C::PLANE) {
      const int sy = idx / C::SX, sx = idx - sy * C::SX;
      const int gi = ghost_map(x0 - HALO + sx, g.nx), gj = ghost_map(y0 - HALO + sy, g.ny);
      if (gi >= 0 && gj >= 0) off[t] = gi + (long long)g.pitch * gj;
    }
  }

  const int j = y0 + ty;
  const int i0 = x0 + XB * tx;
  const double my = (j > 0 && j < g.ny - 1) ? 1.0 : 0.0;
  double mx[XB];
#pragma unroll
  for (int e = 0; e < XB; ++e) mx[e] = (i0 + e > 0 && i0 + e < g.nx - 1) ? 1.0 : 0.0;
  const bool ybnd = (j <= 2 || j >= g.ny - 3);
  const bool xbnd = (i0 <= 2 || i0 + XB - 1 >= g.nx - 3);

  constexpr double t0 = 1.0 / 6.0, t1 = -2.0 / 3.0;
  double acc[XB][5];
  double prev1[XB], prev2[XB];
#pragma unroll
  for (int e = 0; e < XB; ++e) {
    prev1[e] = prev2[e] = 0.0;
#pragma unroll
    for (int d = 0; d < 5; ++d) acc[e][d] = 0.0;
  }
  double dot = 0.0;

  auto plane_ptr = [&](int kp) -> const double* {
    const int ks = ghost_map(kp, g.nz);  // always valid: kp within 2 of [0,nz)
    return p + (long long)(ks - g.k0 + HALO) * g.plane;
  };

  // prologue: stage the first incoming plane
  {
    const double* src = plane_ptr(ka - HALO);
#pragma unroll
    for (int t = 0; t < C::LD; ++t) {
      const int idx = tid + t * C::NT;
      if (idx < C::PLANE) S[0][idx] = off[t] >= 0 ? __ldg(src + off[t]) : 0.0;
    }
  }
  __syncthreads();

  int cur = 0;
  for (int kp = ka - HALO; kp < kb + HALO; ++kp) {
    // prefetch the next plane into registers while this one is consumed
    double nxt[C::LD];
    const bool more = (kp + 1 < kb + HALO);
    if (more) {
      const double* src = plane_ptr(kp + 1);
#pragma unroll
      for (int t = 0; t < C::LD; ++t) nxt[t] = off[t] >= 0 ? __ldg(src + off[t]) : 0.0;
    }

    const double mz = (kp > 0 && kp < g.nz - 1) ? 1.0 : 0.0;
    const double* Sc = &S[cur][(ty + HALO) * C::SX + XB * tx];
    double rc[XB + 4], rm2[XB + 4], rp2[XB + 4], rm1[XB], rp1[XB];
    if constexpr ((XB & 1) == 0) {
      // 16-byte shared loads: lanes are 16 B apart, conflict free
#pragma unroll
      for (int e = 0; e < XB + 4; e += 2) {
        const double2 a = *reinterpret_cast<const double2*>(Sc + e);
        const double2 b = *reinterpret_cast<const double2*>(Sc + e - 2 * C::SX);
        const double2 c = *reinterpret_cast<const double2*>(Sc + e + 2 * C::SX);
        rc[e] = a.x, rc[e + 1] = a.y;
        rm2[e] = b.x, rm2[e + 1] = b.y;
        rp2[e] = c.x, rp2[e + 1] = c.y;
      }
#pragma unroll
      for (int e = 0; e < XB; e += 2) {
        const double2 a = *reinterpret_cast<const double2*>(Sc + e + 2 - C::SX);
        const double2 b = *reinterpret_cast<const double2*>(Sc + e + 2 + C::SX);
        rm1[e] = a.x, rm1[e + 1] = a.y;
        rp1[e] = b.x, rp1[e + 1] = b.y;
      }
    } else {
#pragma unroll
      for (int e = 0; e < XB + 4; ++e) {
        rc[e] = Sc[e];
        rm2[e] = Sc[e - 2 * C::SX];
        rp2[e] = Sc[e + 2 * C::SX];
      }
#pragma unroll
      for (int e = 0; e < XB; ++e) {
        rm1[e] = Sc[e + 2 - C::SX];
        rp1[e] = Sc[e + 2 + C::SX];
      }
    }

#pragma unroll
    for (int e = 0; e < XB; ++e) {
      const int i = i0 + e;
      const double pc = rc[e + 2];
      const double sx1 = rc[e + 1] + rc[e + 3], sx2 = rc[e] + rc[e + 4];
      const double sy1 = rm1[e] + rp1[e], sy2 = rm2[e + 2] + rp2[e + 2];
      const double cs = (rm2[e] + rm2[e + 4]) + (rp2[e] + rp2[e + 4]);
      double X = pc + t1 * sx1 + t0 * sx2;
      double Y = pc + t1 * sy1 + t0 * sy2;
      double bord = 0.0;
      if (xbnd) {
        if (i == 0) X -= (pc - rc[e + 3]);
        if (i == 1) X += (rc[e + 1] - pc) * (1.0 / 3.0);
        if (i == g.nx - 1) X -= (pc - rc[e + 1]);
        if (i == g.nx - 2) X += (rc[e + 3] - pc) * (1.0 / 3.0);
        const double v[5] = {rc[e], rc[e + 1], pc, rc[e + 3], rc[e + 4]};
        bord += w.dx * border_1d(i, g.nx, v);
      }
      if (ybnd) {
        if (j == 0) Y -= (pc - rp1[e]);
        if (j == 1) Y += (rm1[e] - pc) * (1.0 / 3.0);
        if (j == g.ny - 1) Y -= (pc - rm1[e]);
        if (j == g.ny - 2) Y += (rp1[e] - pc) * (1.0 / 3.0);
        const double v[5] = {rm2[e + 2], rm1[e], pc, rp1[e], rp2[e + 2]};
        bord += w.dy * border_1d(j, g.ny, v);
      }
      const double XY = 4.0 * pc - 2.0 * (sx2 + sy2) + cs;
      const double same = mz * (w.axx * my * X + w.ayy * mx[e] * Y + w.bxy * XY) + w.ridge2 * pc + bord;
      const double u = w.bxz * my * (2.0 * pc - sx2) + w.byz * mx[e] * (2.0 * pc - sy2);
      const double azm = w.azz * mx[e] * my;
      const double zc = azm * pc;
      acc[e][0] += t0 * zc - u;
      acc[e][1] += t1 * zc;
      acc[e][2] += same + zc + 2.0 * u;
      acc[e][3] += t1 * zc;
      acc[e][4] += t0 * zc - u;
      // z ends: corrections of T_z and the dz border rows (uniform branches)
      if (kp == 1) {
        const double dh = prev1[e] - pc;  // p(0)-p(1)
        acc[e][1] -= azm * dh;
        acc[e][2] += azm * dh * (1.0 / 3.0);
      }
      if (kp == g.nz - 1) {
        const double dh = pc - prev1[e];  // p(nz-1)-p(nz-2)
        acc[e][2] -= azm * dh;
        acc[e][1] += azm * dh * (1.0 / 3.0);
      }
      if (g.nz >= 3) {
        const double ez = (prev2[e] - 2.0 * prev1[e] + pc) * (w.dz * (1.0 / 6.0));
        if (kp == 2) {
          acc[e][0] += ez;
          acc[e][1] -= 2.0 * ez;
          acc[e][2] += ez;
        }
        if (kp == g.nz - 1) {
          acc[e][0] += ez;
          acc[e][1] -= 2.0 * ez;
          acc[e][2] += ez;
        }
      }
    }

    // output plane kp-2 is complete
    const int ko = kp - HALO;
    if (ko >= ka && ko < kb && j < g.ny) {
      double* dst = q + (long long)(ko - g.k0 + HALO) * g.plane + (long long)g.pitch * j + i0;
      if (XB == 2 && (g.pitch & 1) == 0 && i0 + 1 < g.nx) {
        *reinterpret_cast<double2*>(dst) = make_double2(acc[0][0], acc[XB - 1][0]);
        dot += prev2[0] * acc[0][0] + prev2[XB - 1] * acc[XB - 1][0];
      } else {
#pragma unroll
        for (int e = 0; e < XB; ++e)
          if (i0 + e < g.nx) {
            dst[e] = acc[e][0];
            dot += prev2[e] * acc[e][0];
          }
      }
    }
#pragma unroll
    for (int e = 0; e < XB; ++e) {
      acc[e][0] = acc[e][1];
      acc[e][1] = acc[e][2];
      acc[e][2] = acc[e][3];
      acc[e][3] = acc[e][4];
      acc[e][4] = 0.0;
      prev2[e] = prev1[e];
      prev1[e] = rc[e + 2];
    }

    if (more) {
#pragma unroll
      for (int t = 0; t < C::LD; ++t) {
        const int idx = tid + t * C::NT;
        if (idx < C::PLANE) S[cur ^ 1][idx] = nxt[t];
      }
    }
    __syncthreads();
    cur ^= 1;
  }

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

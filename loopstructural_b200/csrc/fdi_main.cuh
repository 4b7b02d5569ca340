// K1, large grids: the FDI regularisation normal operator q = (R^T R + ridge^2 I) p split by region.
//   k_fdi_box    the interior box 4 <= i <= nx-5, 3 <= j <= ny-4, 3 <= k <= nz-4 (92 % of a 256^3 grid, 98 % of a
//                1024^3 one), where R^T R is the constant 25-tap composite kernel of
//                LI/_finite_difference_interpolator.py:884-1050;
//   k_fdi_shell  every other node, one thread per node in gather form with the true boundary operators
//                (row masks of LC/supports/_3d_structured_grid.py:164-293, border families of
//                LI/_finite_difference_interpolator.py:1052-1144), on a side stream beside the box kernel.
// k_fdi_normal (fdi_stencil.cuh) stays the single-kernel path for grids too small to have a box.
//
// Why: ncu of k_fdi_normal at 256^3 (profiles/r01b) shows 103-168 warp instructions and 40 fp64 operations per
// 32 nodes, two to five resident warps per scheduler, "wait" and "short scoreboard" as the top stalls: bound by
// instruction latency, not by HBM.  In k_fdi_box a thread owns a 2 (x) by YB (y) patch of the tile:
//   * every staged row is read once with three 16-byte shared loads and reused by the YB outputs that touch it
//     (3.5 LDS.128 per node pair instead of 11);
//   * the row sums p(i-2)+p(i+2) are formed once per staged row and serve the in-row tap, the xz tap and the xy
//     corner taps of the rows two above and below;
//   * 20 fp64 operations per node, every coefficient a launch constant, 2 YB independent chains per thread;
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

// ---------------------------------------------------------------------------------------------
// k_fdi_box: the interior box 4 <= i <= nx-5, 3 <= j <= ny-4, 3 <= k <= nz-4 only.  Every input plane is a
// real plane, every coefficient a launch constant: no ghost mapping, no masks, no general path.  Everything
// outside the box is k_fdi_shell's.  Output planes [zlo, zhi) of this handle, z-chunks of cz planes.
// (A persistent-CTA variant with an equal split of the (tile, plane) list was measured and dropped: it loses the
// L2 sharing of the tile halos between CTAs that march through the same planes together -- 573 us instead
// of 375 us at 512^3, profiles/r01h.)
template <int YB, int TYT, int ST, int MINB, bool VEC16>
__global__ void __launch_bounds__(32 * TYT, MINB)
    k_fdi_box(const GridDesc g, const UniConsts U, const double* __restrict__ p, double* __restrict__ q, const int cz,
              const int zlo, const int zhi, double* __restrict__ partials, unsigned int* __restrict__ counter,
              CgScalars* __restrict__ sc, const int cg_mode) {
  using C = MainCfg<YB, TYT, ST>;
  static_assert(ST >= 5, "ring needs planes kp-2..kp plus at least two in flight");
  extern __shared__ __align__(16) double S[];
  __shared__ double red[C::NT / 32];

  if (cg_mode && sc->done) return;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + 32 * ty;
  const int x0 = C::X_FIRST + blockIdx.x * C::TX, y0 = C::Y_FIRST + blockIdx.y * C::TY;
  const int ka = zlo + blockIdx.z * cz;
  const int kb = min(ka + cz, zhi);
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
  auto issue_plane = [&](double* dstS) {
#pragma unroll
    for (int t = 0; t < LD; ++t) {
      const int idx = tid + t * C::NT;
      if (idx < NSLOT) {
        const bool ok = off[t] >= 0;
        if (VEC16) cp_async_16(dstS + 2 * idx, src + (ok ? off[t] : 0), ok);
        else cp_async_8(dstS + idx, src + (ok ? off[t] : 0), ok);
      }
    }
    src += g.plane;
  };

  double A[5][YB][2];
#pragma unroll
  for (int d = 0; d < 5; ++d)
#pragma unroll
    for (int r = 0; r < YB; ++r) A[d][r][0] = A[d][r][1] = 0.0;
  double dot = 0.0;

  const int kstart = ka - HALO, kend = kb + HALO;
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

#define LSB_BOX_STEP(PH)                                                                                      \
  {                                                                                                           \
    if (kp >= kend) break;                                                                                    \
    advance_pipeline();                                                                                       \
    uni_plane_fast<YB, C::SX>(S + slot * C::PLANE + my_off, U, A[(PH + 0) % 5], A[(PH + 1) % 5],              \
                              A[(PH + 2) % 5], A[(PH + 3) % 5], A[(PH + 4) % 5]);                             \
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

  const double bsum = block_sum<C::NT>(dot, red);
  const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  double total;
  if (grid_sum_last_block<C::NT>(bsum, partials, counter, nblocks, bid, red, &total)) sc->pq_part[0] = total;
}

}  // namespace lsb

namespace lsb {

// ---------------------------------------------------------------------------------------------
// k_fdi_shell: every owned node OUTSIDE the interior box, one thread per node, gather form.
//   q = ridge^2 p + sum_a (a_aa m_b m_c T_a + d_a B_a) p + sum_{a<b} b_ab m_c (C_a (x) C_b) p
// with the true 1-D boundary operators (no ghost cells): T = D^T diag(m) D / 6, C = G^T diag(m) G,
// B = sum_{r in {1,n-2}} d_r d_r^T / 6, m(r) = [1 <= r <= n-2] -- the row masks of
// StructuredGrid.neighbour_global_indexes / FDI.assemble_borders (LC/supports/_3d_structured_grid.py:164-293,
// LI/_finite_difference_interpolator.py:1052-1144) after row normalisation (LI/_discrete_interpolator.py:361-366).
// 25 loads per node, all hitting L1/L2 beside the box kernel that streams the same planes.
struct ShellBox {
  int ilo, ihi, jlo, jhi, klo, khi;  // the box that is NOT computed here (klo/khi clipped to the owned planes)
  int n_full_planes;                 // owned planes outside [klo,khi): every node
  int full_blocks;                   // CTAs per such plane: ceil(nx*ny / NT)
  int yrow_blocks;                   // per plane inside [klo,khi): CTAs of the rows outside [jlo,jhi), every column
  int xcol_blocks;                   //   and of the columns outside [ilo,ihi) of the rows [jlo,jhi)
};

__device__ __forceinline__ double vrow(int r, int n) { return (r >= 1 && r <= n - 2) ? 1.0 : 0.0; }
// 1-D data of one axis of a shell node: clamped tap offsets (a clamped tap always has coefficient 0), the
// row-validity flags m(c-1), m(c), m(c+1) and whether border rows touch the node.  GEN = false: the axis is
// known to be regular at this node (3 <= c <= n-4), everything is a constant.
template <bool GEN>
struct Axis1D {
  int o[5];
  double va, vb, vc;
  bool border;
  __device__ __forceinline__ Axis1D(int c, int n, int stride) {
    if (GEN) {
      va = vrow(c - 1, n); vb = vrow(c, n); vc = vrow(c + 1, n);
      border = c <= 2 || c >= n - 3;
#pragma unroll
      for (int s = -2; s <= 2; ++s) o[s + 2] = (min(max(c + s, 0), n - 1) - c) * stride;
    } else {
      va = vb = vc = 1.0;
      border = false;
#pragma unroll
      for (int s = -2; s <= 2; ++s) o[s + 2] = s * stride;
    }
  }
};

template <bool GX, bool GY, bool GZ>
__device__ __forceinline__ double shell_node(const GridDesc& g, const StencilWeights& w, const double* __restrict__ p,
                                             double* __restrict__ q, const int cx, const int cy, const int cz) {
  const Axis1D<GX> X(cx, g.nx, 1);
  const Axis1D<GY> Y(cy, g.ny, g.pitch);
  const Axis1D<GZ> Z(cz, g.nz, (int)g.plane);
  const long long base = (long long)(cz - g.k0 + HALO) * g.plane + cy * g.pitch + cx;
  const double* pb = p + base;
  const double p0 = pb[0];
  // the 24 neighbours first: independent loads, one round trip
  const double x0 = pb[X.o[0]], x1 = pb[X.o[1]], x3 = pb[X.o[3]], x4 = pb[X.o[4]];
  const double y0 = pb[Y.o[0]], y1 = pb[Y.o[1]], y3 = pb[Y.o[3]], y4 = pb[Y.o[4]];
  const double z0 = pb[Z.o[0]], z1 = pb[Z.o[1]], z3 = pb[Z.o[3]], z4 = pb[Z.o[4]];
  const double xy00 = pb[X.o[0] + Y.o[0]], xy04 = pb[X.o[0] + Y.o[4]], xy40 = pb[X.o[4] + Y.o[0]], xy44 = pb[X.o[4] + Y.o[4]];
  const double xz00 = pb[X.o[0] + Z.o[0]], xz04 = pb[X.o[0] + Z.o[4]], xz40 = pb[X.o[4] + Z.o[0]], xz44 = pb[X.o[4] + Z.o[4]];
  const double yz00 = pb[Y.o[0] + Z.o[0]], yz04 = pb[Y.o[0] + Z.o[4]], yz40 = pb[Y.o[4] + Z.o[0]], yz44 = pb[Y.o[4] + Z.o[4]];
  double acc = w.ridge2 * p0;
  // lines: a_aa m_b m_c T_a + d_a B_a,  T = D^T diag(m) D / 6
  auto line = [&](const auto& A, const double a6, const double d, const int c, const int n, const double t0,
                  const double t1, const double t3, const double t4) {
    const double L0 = a6 * A.va, L1 = -2.0 * a6 * (A.va + A.vb), L2 = a6 * (A.va + 4.0 * A.vb + A.vc),
                 L3 = -2.0 * a6 * (A.vb + A.vc), L4 = a6 * A.vc;
    double v = L2 * p0;
    v = fma(L0, t0, v); v = fma(L1, t1, v); v = fma(L3, t3, v); v = fma(L4, t4, v);
    if (A.border && n >= 3) {
      // border rows d_1 = [1,-2,1] on nodes 0..2 and d_{n-2} on nodes n-3..n-1: d(c) (d . p) / 6
      if (c <= 2) {
        const double e0 = c == 0 ? p0 : (c == 1 ? t1 : t0), e1 = c == 0 ? t3 : (c == 1 ? p0 : t1),
                     e2 = c == 0 ? t4 : (c == 1 ? t3 : p0);
        v = fma(d * (c == 1 ? -2.0 / 6.0 : 1.0 / 6.0), e0 - 2.0 * e1 + e2, v);
      }
      const int cu = n - 1 - c;
      if (cu <= 2) {
        const double e0 = cu == 0 ? p0 : (cu == 1 ? t3 : t4), e1 = cu == 0 ? t1 : (cu == 1 ? p0 : t3),
                     e2 = cu == 0 ? t0 : (cu == 1 ? t1 : p0);
        v = fma(d * (cu == 1 ? -2.0 / 6.0 : 1.0 / 6.0), e0 - 2.0 * e1 + e2, v);
      }
    }
    return v;
  };
  acc += line(X, w.axx * Y.vb * Z.vb * (1.0 / 6.0), w.dx, cx, g.nx, x0, x1, x3, x4);
  acc += line(Y, w.ayy * X.vb * Z.vb * (1.0 / 6.0), w.dy, cy, g.ny, y0, y1, y3, y4);
  acc += line(Z, w.azz * X.vb * Y.vb * (1.0 / 6.0), w.dz, cz, g.nz, z0, z1, z3, z4);
  // mixed terms: b_ab m_c (C_a (x) C_b) p,  C = [-va, va+vc, -vc] at offsets -2, 0, +2
  auto mixed = [&](const auto& A, const auto& B, const double bw, const double a0, const double a4, const double b0,
                   const double b4, const double m00, const double m04, const double m40, const double m44) {
    const double cb = B.va + B.vc;
    const double rowm = fma(-B.vc, m04, fma(-B.va, m00, cb * a0));
    const double row0 = fma(-B.vc, b4, fma(-B.va, b0, cb * p0));
    const double rowp = fma(-B.vc, m44, fma(-B.va, m40, cb * a4));
    return bw * fma(-A.vc, rowp, fma(-A.va, rowm, (A.va + A.vc) * row0));
  };
  acc += mixed(X, Y, w.bxy * Z.vb, x0, x4, y0, y4, xy00, xy04, xy40, xy44);
  acc += mixed(X, Z, w.bxz * Y.vb, x0, x4, z0, z4, xz00, xz04, xz40, xz44);
  acc += mixed(Y, Z, w.byz * X.vb, y0, y4, z0, z4, yz00, yz04, yz40, yz44);
  q[base] = acc;
  return p0 * acc;
}

// One CTA = NT consecutive shell nodes of one plane: a full plane next to a z face (general on every axis),
// the y rows of an inner plane (z regular), or the x columns of its remaining rows (y and z regular).
template <int NT>
__global__ void __launch_bounds__(NT, 3)
    k_fdi_shell(const GridDesc g, const StencilWeights w, const ShellBox sb, const double* __restrict__ p,
                double* __restrict__ q, double* __restrict__ partials, unsigned int* __restrict__ counter,
                CgScalars* __restrict__ sc, const int cg_mode) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  double dot = 0.0;
  const int nfull = sb.n_full_planes * sb.full_blocks;
  const int b = blockIdx.x;
  if (b < nfull) {
    const int zi = b / sb.full_blocks;
    const int r = (b - zi * sb.full_blocks) * NT + threadIdx.x;
    const int cz = zi < sb.klo - g.k0 ? g.k0 + zi : sb.khi + (zi - (sb.klo - g.k0));
    const int cy = r / g.nx, cx = r - cy * g.nx;
    if (cy < g.ny) dot = shell_node<true, true, true>(g, w, p, q, cx, cy, cz);
  } else {
    const int per = sb.yrow_blocks + sb.xcol_blocks;
    const int b2 = b - nfull;
    const int kz = b2 / per, bb = b2 - kz * per;
    const int cz = sb.klo + kz;
    if (bb < sb.yrow_blocks) {
      const int r = bb * NT + threadIdx.x;
      const int jj = r / g.nx, cx = r - jj * g.nx;
      const int cy = jj < sb.jlo ? jj : sb.jhi + (jj - sb.jlo);
      if (cy < g.ny) dot = shell_node<true, true, false>(g, w, p, q, cx, cy, cz);
    } else {
      const int nx_out = g.nx - (sb.ihi - sb.ilo);
      const int r = (bb - sb.yrow_blocks) * NT + threadIdx.x;
      const int jr = r / nx_out, ii = r - jr * nx_out;
      const int cy = sb.jlo + jr, cx = ii < sb.ilo ? ii : sb.ihi + (ii - sb.ilo);
      if (cy < sb.jhi) dot = shell_node<true, false, false>(g, w, p, q, cx, cy, cz);
    }
  }
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[2] = total;
}

}  // namespace lsb

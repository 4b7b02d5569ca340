// K1, large grids, ONE launch: q = (R^T R + ridge^2 I) p over every owned node of a handle, fused p.q.
//
// Round 1 ran three launches on three streams (k_fdi_box uniform, k_fdi_box<ZGEN> for the z faces, k_fdi_shell for the
// x/y ring).  The ring kernel -- one thread per column gathering from global memory -- is latency bound (39 us alone
// at 256^3 for 5 % of the nodes, profiles/r02c) and cannot co-reside with the register-hungry box CTAs, so ~28 us of it
// trailed every matvec; the z-face launch queued behind the box for the same reason.  Here every node is computed by a
// CTA of the same launch, all staged through shared memory, in three roles:
//
//   BOX       tile of 64 x TY nodes over the interior columns 4 <= i < nx-4 and rows 3 <= j < ny-3, a chunk of planes.
//             Its output planes 3 <= k < nz-3 are the register-tiled constant-coefficient march of k_fdi_box
//             (uni_plane_fast: 20 fp64 operations per node, every coefficient a launch constant, nothing else in the
//             loop).  A chunk that owns planes next to a z face (k < 3 or k >= nz-3) first computes those <= 3 planes in
//             GATHER form from the <= 5 planes it stages at once (z_face_block): general z coefficients per output
//             plane, scalar shared-memory reads, few registers -- it runs once per CTA, not per plane step.
//   X STRIP   the 4 columns next to an x face, NT rows, a chunk of planes: lanes along y, a thread owns the 4 nodes of its
//             row.  Y STRIP: the 3 rows next to a y face over the box columns, NT columns: lanes along x, a thread owns
//             the 3 nodes of its column.  One routine serves both (strip_cta<AXIS>): along the strip's NORMAL axis the
//             local index c = 0 is the face (the far side is staged mirrored), so the end corrections of T, the border
//             rows and the ghost cells of C are compile-time constants per c; along its TANGENT axis the five taps come
//             from a per-thread table computed once (constant except next to the other faces, i.e. in the corners);
//             z coefficients per plane from a table in shared memory.
//
// A strip node costs several box nodes, but strips are 5 % of a 256^3 grid (2.7 % of 512^3); their CTAs are ordinary
// members of the launch, dispatched first, chunked short so that none of them outlasts a box CTA (plan_fused).
// (Measured dead ends, profiles/r2b-r2d: y edges and z faces inside the box march -- inlined they push the hot loop into
// register spills, out of line the warps that own edge rows serialise 6 calls per plane step and hold their CTA back.)
//
// Operator (fdi_main.cuh / fdi_stencil.cuh): q = ridge^2 p + sum_a (a_aa m_b m_c T_a + d_a B_a) p
//   + sum_{a<b} b_ab m_c (C_a (x) C_b) p, masks m(r) = [1 <= r <= n-2]; reference: assemble_inner / assemble_borders
//   (LI/_finite_difference_interpolator.py:1052-1164), row normalisation (LI/_discrete_interpolator.py:361-366).
#pragma once
#include "fdi_main.cuh"

namespace lsb {

// ---- TMA staging of the box tiles (cp.async.bulk.tensor, sm_90+; here sm_100a) ------------------------------------------
// A node vector is described to the TMA unit once per handle as a 3-D tensor (x, y, stored plane) of doubles
// (cuTensorMapEncodeTiled, csrc/api.cu); a plane of a tile with its halo is then ONE instruction of ONE thread:
// box (SX, SY, 1) at (x0-2, y0-2, plane) -> shared memory, bytes counted on an mbarrier, columns / rows beyond the grid
// zero-filled by the hardware.  Against the cp.async (LDGSTS) staging this removes six copy instructions, their address
// arithmetic and the six-entry offset table from every thread of every plane step.
struct alignas(64) TmapArg {
  unsigned long long opaque[16];  // a CUtensorMap (128 bytes), passed as a __grid_constant__ kernel parameter
};
#ifndef LSB_EMU
__device__ __forceinline__ unsigned smem_u32(const void* ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const TmapArg* tm, const int x, const int y, const int z,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
#else  // the CPU emulation instantiates the cp.async path only; these are never called
inline void mbar_init(unsigned long long*, int) {}
inline void mbar_fence_init() {}
inline void mbar_expect_tx(unsigned long long*, unsigned) {}
inline void mbar_wait(unsigned long long*, unsigned) {}
inline void tma_load_3d(void*, const TmapArg*, int, int, int, unsigned long long*) {}
#endif

struct FusedPlan {
  int tiles_x, tiles_y;   // box tiles over columns [4, nx-4) x rows [3, ny-3)
  int cz, chunks;         // z chunks of the box CTAs over the owned planes: the first holds lo_len planes, the others cz
  int lo_len;             //   (a chunk that owns a z face is shorter: it also runs z_face_block)
  int n_box;              // tiles_x * tiles_y * chunks
  int xs_y;               // x-strip CTAs per side and z chunk: ceil(ny / NT)
  int ys_x;               // y-strip CTAs per side and z chunk: ceil((nx - 8) / NT)
  int scz, schunks;       // z chunks of the strip CTAs
  int n_xs, n_ys;         // 2 * xs_y * schunks, 2 * ys_x * schunks
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
// index of a non-uniform plane in the table of a strip CTA: planes 0..4 -> 0..4, nz-5..nz-1 -> 5..9 (entries are
// computed per plane, so the two ranges may overlap on grids of fewer than 10 planes)
__device__ __forceinline__ int z_slot(const int kp, const int nz) { return kp < 5 ? kp : 5 + (kp - (nz - 5)); }

// end corrections of the ghost-cell 1-D operators at index i of n (fdi_stencil.cuh header): T differs from the constant
// 5-tap stencil by  q(0) -= p0 - p1,  q(1) += (p0 - p1)/3  (mirrored at the upper end); border rows B = d d^T / 6 on the
// rows 1 and n-2.  As coefficients of the taps p(i-2) .. p(i+2) on ghost data (tc: constant T plus corrections,
// bc: border rows), computed once per thread.
__device__ __forceinline__ void line_coefficients(const int i, const int n, double (&tc)[5], double (&bc)[5]) {
  tc[0] = tc[4] = 1.0 / 6.0;
  tc[1] = tc[3] = -2.0 / 3.0;
  tc[2] = 1.0;
#pragma unroll
  for (int s = 0; s < 5; ++s) bc[s] = 0.0;
  if (i == 0) { tc[2] -= 1.0; tc[3] += 1.0; }                // -(p0 - p(+1))
  if (i == 1) { tc[1] += 1.0 / 3.0; tc[2] -= 1.0 / 3.0; }    // (p(-1) - p0) / 3
  if (i == n - 1) { tc[2] -= 1.0; tc[1] += 1.0; }
  if (i == n - 2) { tc[3] += 1.0 / 3.0; tc[2] -= 1.0 / 3.0; }
  // lower border row, centred on index 1, touches 0,1,2; upper, centred on n-2, touches n-3,n-2,n-1
  constexpr double s6 = 1.0 / 6.0;
  if (i == 0) { bc[2] += s6; bc[3] -= 2.0 * s6; bc[4] += s6; }
  if (i == 1) { bc[1] -= 2.0 * s6; bc[2] += 4.0 * s6; bc[3] -= 2.0 * s6; }
  if (i == 2) { bc[0] += s6; bc[1] -= 2.0 * s6; bc[2] += s6; }
  if (i == n - 1) { bc[2] += s6; bc[1] -= 2.0 * s6; bc[0] += s6; }
  if (i == n - 2) { bc[3] -= 2.0 * s6; bc[2] += 4.0 * s6; bc[1] -= 2.0 * s6; }
  if (i == n - 3) { bc[4] += s6; bc[3] -= 2.0 * s6; bc[2] += s6; }
}

// ---------------------------------------------------------------------------------------------------------------
// STRIP roles.  AXIS = 0: x strip (normal axis x, W = 4 columns, tangent axis y: a thread owns a row); AXIS = 1: y strip
// (normal axis y, W = 3 rows, tangent axis x: a thread owns a column).  side 1 is staged mirrored along the normal axis
// (local index c <-> global n-1-c).  t0 = first tangent index of this CTA, [t_lo, t_hi) = the tangent range the strip
// owns (x strips: every row; y strips: the box columns).  Staged plane: (NT + 4) tangent positions x (W + 2) normal
// positions; element (tangent tt, normal c) at tt * ST_T + c * ST_C.
template <int AXIS>
struct StripGeom {
  static constexpr int W = AXIS == 0 ? 4 : 3;
  static constexpr int NC = W + 2;
};

// z coefficients of a plane step as a strip uses them: t = azz T_z (masked part), b = dz B_z (unmasked), c = C_z
struct ZStrip {
  double mz, t0, tm1, tp1, tm2, tp2, b0, bm1, bp1, bm2, bp2, c0, cm2, cp2;
};
__device__ __forceinline__ ZStrip z_strip(const ZFull& Z) {
  ZStrip R;
  R.mz = Z.mz;
  R.t0 = Z.t0; R.tm1 = Z.tm1; R.tp1 = Z.tp1; R.tm2 = Z.tm2; R.tp2 = Z.tp2;
  R.b0 = Z.l0 - Z.t0; R.bm1 = Z.lm1 - Z.tm1; R.bp1 = Z.lp1 - Z.tp1; R.bm2 = Z.lm2 - Z.tm2; R.bp2 = Z.lp2 - Z.tp2;
  R.c0 = Z.c0; R.cm2 = Z.cm2; R.cp2 = Z.cp2;
  return R;
}

template <int NT, int ST, int AXIS>
__device__ __forceinline__ double strip_cta(const GridDesc& g, const StencilWeights& w, const ZStrip* __restrict__ zsm,
                                            const double* __restrict__ p, double* __restrict__ q, const int side,
                                            const int t0, const int t_lo, const int t_hi, const int ka, const int kb,
                                            double* __restrict__ S, const int tid) {
  constexpr int W = StripGeom<AXIS>::W, NC = StripGeom<AXIS>::NC;
  constexpr int TT = NT + 2 * HALO;
  constexpr int PLANE = TT * NC;
  constexpr int ST_T = AXIS == 0 ? NC : 1;   // x strip: rows of NC columns; y strip: NC rows of TT columns
  constexpr int ST_C = AXIS == 0 ? 1 : TT;
  constexpr int LD = (PLANE + NT - 1) / NT;
  const int n_t = AXIS == 0 ? g.ny : g.nx;   // length of the tangent axis
  const int n_n = AXIS == 0 ? g.nx : g.ny;
  // weights seen from the strip: normal / tangent second derivatives, border rows, mixed terms with z
  const double a_n = AXIS == 0 ? w.axx : w.ayy, a_t = AXIS == 0 ? w.ayy : w.axx;
  const double d_n = AXIS == 0 ? w.dx : w.dy, d_t = AXIS == 0 ? w.dy : w.dx;
  const double b_nz = AXIS == 0 ? w.bxz : w.byz, b_tz = AXIS == 0 ? w.byz : w.bxz;

  int off[LD];
#pragma unroll
  for (int t = 0; t < LD; ++t) {
    const int idx = tid + t * NT;
    off[t] = -1;
    if (idx < PLANE) {
      // the staged layout follows global memory: x strip (tt, c) with c fastest; y strip (c, tt) with tt fastest
      const int tt = AXIS == 0 ? idx / NC : idx % TT;
      const int c = AXIS == 0 ? idx % NC : idx / TT;
      const int gt = ghost_map(t0 - HALO + tt, n_t);  // tangent index: mirrored outside the grid, -1 beyond
      const int gn = side ? n_n - 1 - c : c;
      if (gt >= 0) off[t] = AXIS == 0 ? gn + g.pitch * gt : gt + g.pitch * gn;
    }
  }
  const int jt = t0 + tid;                       // this thread's tangent index
  const bool mine = jt >= t_lo && jt < t_hi;
  const double mt = (jt >= 1 && jt <= n_t - 2) ? 1.0 : 0.0;
  double tc[5], bc[5];
  line_coefficients(jt, n_t, tc, bc);
  // the tangent line with both families folded in: mz (a_t T_t) is applied per step, d_t B_t is not masked
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    tc[s] *= a_t;
    bc[s] *= d_t;
  }
  const double an_mt = a_n * mt, bnz_mt = b_nz * mt;
  // in-plane offsets of this thread's W outputs
  int qo[W];
#pragma unroll
  for (int c = 0; c < W; ++c) qo[c] = AXIS == 0 ? g.pitch * jt + (side ? g.nx - 1 - c : c) : g.pitch * (side ? g.ny - 1 - c : c) + jt;
  double A[5][W];
#pragma unroll
  for (int dd = 0; dd < 5; ++dd)
#pragma unroll
    for (int c = 0; c < W; ++c) A[dd][c] = 0.0;
  double dot = 0.0;

  const int kstart = ka - HALO, kend = ka < kb ? kb + HALO : ka - HALO;
  const double* src = p + (long long)(kstart - g.k0 + HALO) * g.plane;
  int issued = kstart;
  auto issue_plane = [&](double* dstS) {
    if (issued >= 0 && issued < g.nz) {
#pragma unroll
      for (int t = 0; t < LD; ++t) {
        const int idx = tid + t * NT;
        if (idx < PLANE) {
          const bool ok = off[t] >= 0;
          cp_async_8(dstS + idx, src + (ok ? off[t] : 0), ok);  // idx IS the staged offset (see the layout above)
        }
      }
    }
    src += g.plane;
    ++issued;
  };
#pragma unroll
  for (int s = 0; s < ST - 3; ++s) {
    if (issued < kend) issue_plane(S + s * PLANE);
    cp_async_commit();
  }
  int slot = 0;
  double* qpl = q + (long long)(ka - g.k0 + HALO) * g.plane;  // output plane ka
  for (int kp = kstart; kp < kend; ++kp) {
    cp_async_wait<ST - 4>();
    __syncthreads();
    {
      const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;
      if (issued < kend) issue_plane(S + fs * PLANE);
      cp_async_commit();
    }
    if (kp >= 0 && kp < g.nz) {
      const ZStrip Z = zsm[z_step_uniform(kp, g.nz) ? 10 : z_slot(kp, g.nz)];
      const double* V = S + slot * PLANE + tid * ST_T;  // tangent offset -2 of this thread's window, normal index 0
      double v[5][NC];
#pragma unroll
      for (int tt = 0; tt < 5; ++tt) {
        if (AXIS == 0) {  // a row of 6 doubles, 48 bytes apart: three 16-byte loads
#pragma unroll
          for (int c = 0; c < NC; c += 2) {
            const double2 t2 = lds2(V + tt * ST_T + c);
            v[tt][c] = t2.x;
            v[tt][c + 1] = t2.y;
          }
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) v[tt][c] = V[tt * ST_T + c * ST_C];
        }
      }
#pragma unroll
      for (int c = 0; c < W; ++c) {
        // ghost cells of the normal axis: p(-1) := p(1), p(-2) := p(0)
        auto ng = [&](const int tt, const int cc) { return cc >= 0 ? v[tt][cc] : (cc == -1 ? v[tt][1] : v[tt][0]); };
        const double p0 = v[2][c];
        const double nm1 = ng(2, c - 1), nm2 = ng(2, c - 2), np1 = v[2][c + 1], np2 = v[2][c + 2];
        double Nl = fma(1.0 / 6.0, nm2 + np2, fma(-2.0 / 3.0, nm1 + np1, p0));  // T along the normal axis ...
        if (c == 0) Nl -= (v[2][0] - v[2][1]);                                   // ... with its end corrections
        if (c == 1) Nl += (v[2][0] - v[2][1]) * (1.0 / 3.0);
        const double gn0 = 2.0 * v[0][c] - ng(0, c - 2) - v[0][c + 2];           // C_normal p at tangent offsets -2, 0, +2
        const double gn2 = 2.0 * p0 - nm2 - np2;
        const double gn4 = 2.0 * v[4][c] - ng(4, c - 2) - v[4][c + 2];
        const double NTm = 2.0 * gn2 - gn0 - gn4;                                // (C_normal (x) C_tangent) p
        double inplane = fma(an_mt, Nl, w.bxy * NTm);                            // the part m_z switches off
        double rest = w.ridge2 * p0;                                             // the part it does not
        if (c <= 2) {                                                            // the border row next to the face
          const double e2 = v[2][0] - 2.0 * v[2][1] + v[2][2];
          rest = fma(d_n * (c == 1 ? -2.0 / 6.0 : 1.0 / 6.0), e2, rest);
        }
        double G = bnz_mt * gn2;
        double pm = 0.0;
        if (c > 0) {  // m_normal = 1: the tangent line, the tangent-z mixed term and the z line exist
          double Tl = tc[0] * v[0][c];
          Tl = fma(tc[1], v[1][c], Tl); Tl = fma(tc[2], p0, Tl); Tl = fma(tc[3], v[3][c], Tl); Tl = fma(tc[4], v[4][c], Tl);
          inplane += Tl;
          G = fma(b_tz, 2.0 * p0 - (v[0][c] + v[4][c]), G);
          pm = mt * p0;
        }
        double tb = bc[0] * v[0][c];                                             // d_t B_tangent p (zero table away from the corners)
        tb = fma(bc[1], v[1][c], tb); tb = fma(bc[2], p0, tb); tb = fma(bc[3], v[3][c], tb); tb = fma(bc[4], v[4][c], tb);
        rest += tb;
        A[2][c] += fma(Z.mz, inplane, rest) + fma(Z.t0, pm, Z.b0 * p0) + Z.c0 * G;
        A[1][c] += fma(Z.tm1, pm, Z.bm1 * p0);
        A[3][c] += fma(Z.tp1, pm, Z.bp1 * p0);
        A[0][c] += fma(Z.tm2, pm, Z.bm2 * p0) + Z.cm2 * G;
        A[4][c] += fma(Z.tp2, pm, Z.bp2 * p0) + Z.cp2 * G;
      }
    }
    // output plane kp-2 is complete
    if (kp - HALO >= ka) {
      if (mine) {
        const int s2 = slot >= 2 ? slot - 2 : slot + ST - 2;
        const double* P = S + s2 * PLANE + (tid + HALO) * ST_T;
#pragma unroll
        for (int c = 0; c < W; ++c) dot = fma(P[c * ST_C], A[0][c], dot);
        if (AXIS == 0 && (g.pitch & 1) == 0 && (g.nx & 1) == 0) {  // the four nodes of a row: one aligned 32-byte sector
          double2* dst = reinterpret_cast<double2*>(qpl + qo[side ? 3 : 0]);
          dst[0] = side ? make_double2(A[0][3], A[0][2]) : make_double2(A[0][0], A[0][1]);
          dst[1] = side ? make_double2(A[0][1], A[0][0]) : make_double2(A[0][2], A[0][3]);
        } else {
#pragma unroll
          for (int c = 0; c < W; ++c) qpl[qo[c]] = A[0][c];  // y strips: coalesced along x
        }
      }
      qpl += g.plane;
    }
#pragma unroll
    for (int c = 0; c < W; ++c) {
      A[0][c] = A[1][c]; A[1][c] = A[2][c]; A[2][c] = A[3][c]; A[3][c] = A[4][c]; A[4][c] = 0.0;
    }
    slot = (slot + 1 == ST) ? 0 : slot + 1;
  }
  cp_async_wait<0>();
  return dot;
}

// ---------------------------------------------------------------------------------------------------------------
// Output planes [f0, f1) next to a z face (at most 3) for the thread's 2 x YB patch of an interior tile, in gather form
// from the planes [s0, ..) staged side by side in S (at most ST - 1 of them; plane k at slot k - s0).  In-plane
// coefficients are the interior constants; the z coefficients of an output plane ko are row ko of the per-plane table
// (shell_fill_z_table: T_z taps at 0..4, B_z at 5..9, C_z at 10..12, m_z at 13).  Runs once per CTA.
template <int YB, int SX, int PLANE>
__device__ __forceinline__ double z_face_block(const GridDesc& g, const UniConsts& K, const double* __restrict__ ztab,
                                               const double* __restrict__ Sthr, double* __restrict__ qthr, const int f0,
                                               const int f1, const int s0, const int nrows, const bool c0ok, const bool c1ok) {
  // Sthr: staged element (row j0, column i0) of slot 0; qthr: q element (plane f0, row j0, column i0)
  double dot = 0.0;
  const double axx = K.cx1 * (-1.5), ayy = K.cy1 * (-1.5);  // cx1 = -2 axx / 3
  for (int ko = f0; ko < f1; ++ko) {
    const double* zr = ztab + (long long)ko * SHELL_ZSTRIDE;
    const double mz = zr[13];
    for (int r = 0; r < nrows; ++r) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (!(e == 0 ? c0ok : c1ok)) continue;
        auto at = [&](const int k, const int dj, const int di) { return Sthr[(k - s0) * PLANE + (r + dj) * SX + e + di]; };
        const double p0 = at(ko, 0, 0);
        const double xs1 = at(ko, 0, -1) + at(ko, 0, 1), xs2 = at(ko, 0, -2) + at(ko, 0, 2);
        const double ys1 = at(ko, -1, 0) + at(ko, 1, 0), ys2 = at(ko, -2, 0) + at(ko, 2, 0);
        const double cs = (at(ko, -2, -2) + at(ko, -2, 2)) + (at(ko, 2, -2) + at(ko, 2, 2));
        const double X = fma(1.0 / 6.0, xs2, fma(-2.0 / 3.0, xs1, p0));
        const double Y = fma(1.0 / 6.0, ys2, fma(-2.0 / 3.0, ys1, p0));
        const double XY = 4.0 * p0 - 2.0 * (xs2 + ys2) + cs;
        double acc = mz * (axx * X + ayy * Y + K.bxy * XY) + K.ridge2 * p0;
#pragma unroll
        for (int s = -2; s <= 2; ++s) {
          const int k = ko + s;
          if (k < 0 || k >= g.nz) continue;
          const double pk = s == 0 ? p0 : at(k, 0, 0);
          acc = fma(fma(K.azz, zr[2 + s], K.dz * zr[7 + s]), pk, acc);
          if (s == -2 || s == 0 || s == 2) {
            const double xk = s == 0 ? xs2 : at(k, 0, -2) + at(k, 0, 2);
            const double yk = s == 0 ? ys2 : at(k, -2, 0) + at(k, 2, 0);
            const double G = K.bxz * (2.0 * pk - xk) + K.byz * (2.0 * pk - yk);
            acc = fma(zr[11 + s / 2], G, acc);
          }
        }
        qthr[(long long)(ko - f0) * g.plane + (long long)r * g.pitch + e] = acc;
        dot = fma(p0, acc, dot);
      }
    }
  }
  return dot;
}

// ---------------------------------------------------------------------------------------------------------------
// The work of CTA `bid` of the plan: its part of q = N p, and (return value) this thread's share of p.q.  A device
// function so that the launch-per-matvec kernel below and the persistent CG kernel of small systems (csrc/cg_small.cuh,
// which walks the plan's CTAs in a loop) run the same code.  `phases`: TMA only, the parities of the ring's mbarriers,
// which live as long as the real CTA (initialised when `first`).
template <int TYT, int ST, bool VEC16, bool TMA>
__device__ __forceinline__ double fused_cta(const GridDesc& g, const UniConsts& U, const StencilWeights& w, const FusedPlan& pl,
                                            const double* __restrict__ p, double* __restrict__ q, const double* __restrict__ ztab,
                                            const TmapArg* tmap_ptr, const int bid, const bool first, unsigned& phases) {
  constexpr int YB = 4;
  using C = MainCfg<YB, TYT, ST>;
  static_assert(ST >= 6, "ring: planes kp-2..kp plus three in flight; the z-face block stages five planes at once");
  static_assert((C::NT + 2 * HALO) * StripGeom<0>::NC <= C::PLANE, "a strip plane must fit a box plane slot");
  LSB_DYN_SMEM(S);
  __shared__ ZStrip zsm[11];  // strips: z coefficients of the planes next to a z face; [10]: the interior constants
  __shared__ __align__(8) unsigned long long full[TMA ? ST : 1];  // TMA: one mbarrier per ring slot

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = tx + 32 * ty;
  double dot = 0.0;

  if constexpr (TMA) {
    if (first) {
      if (tid == 0) {
#pragma unroll
        for (int s2 = 0; s2 < ST; ++s2) mbar_init(&full[s2], 1);
        mbar_fence_init();
      }
      phases = 0;
      __syncthreads();
    }
  }
  if (bid >= pl.n_box) {
    // ---- strip roles (the last block indices: they fill the CTA slots the box tiles leave free)
    if (tid < 10) {
      const int k = tid < 5 ? tid : g.nz - 5 + (tid - 5);
      ZFull Zf{};
      if (k >= 0 && k < g.nz) Zf = z_full(g, ztab, U.azz, U.dz, k);
      zsm[tid] = z_strip(Zf);
    }
    if (tid == 10) zsm[10] = z_strip(z_uniform(U.azz));
    __syncthreads();
    const bool ys = bid >= pl.n_box + pl.n_xs;
    const int b = bid - pl.n_box - (ys ? pl.n_xs : 0);
    const int per_side = ys ? pl.ys_x : pl.xs_y;
    const int ch = b / (2 * per_side), rem = b - ch * 2 * per_side;
    const int side = rem / per_side, st = rem - side * per_side;
    const int ka = g.k0 + ch * pl.scz, kb = min(ka + pl.scz, g.k1);
    // a strip plane is small (NT + 4 rows of 6 doubles): the same shared memory holds a ring of STS of them, STS - 3 in
    // flight.  A strip row is 48 bytes of a 2 KB grid row, its copies are bound by DRAM latency, not bandwidth; with the 3
    // planes in flight of the box ring a strip step took 1.4 us (profiles/r2j_cta_trace_256.txt)
    constexpr int STS = (ST * C::PLANE) / ((C::NT + 2 * HALO) * StripGeom<0>::NC) < 12 ? (ST * C::PLANE) / ((C::NT + 2 * HALO) * StripGeom<0>::NC) : 12;
    static_assert(STS >= ST, "the strip ring is at least as deep as the box ring");
    if (ys) dot = strip_cta<C::NT, STS, 1>(g, w, zsm, p, q, side, C::X_FIRST + st * C::NT, C::X_FIRST, g.nx - C::X_FIRST, ka, kb, S, tid);
    else dot = strip_cta<C::NT, STS, 0>(g, w, zsm, p, q, side, st * C::NT, 0, g.ny, ka, kb, S, tid);
  } else {
    // ---- box role: interior tile, the march of k_fdi_box over the planes 3 <= k < nz-3 of its chunk
    const int b = bid;
    const int tiles = pl.tiles_x * pl.tiles_y;
    const int ch = b / tiles, trem = b - ch * tiles;
    const int by = trem / pl.tiles_x, bx = trem - by * pl.tiles_x;
    const int x0 = C::X_FIRST + bx * C::TX, y0 = C::Y_FIRST + by * C::TY;
    const int ka = ch == 0 ? g.k0 : g.k0 + pl.lo_len + (ch - 1) * pl.cz;
    const int kb = ch == 0 ? min(g.k0 + pl.lo_len, g.k1) : min(ka + pl.cz, g.k1);
    const int i0 = x0 + 2 * tx;
    const int j0 = y0 + YB * ty;
    const int nrows = max(0, min(YB, (g.ny - C::Y_FIRST) - j0));
    const int xlim = g.nx - C::X_FIRST;

    // Staging of one plane.  cp.async: slot idx = tid + t NT of the (SX x SY) tile (pairs of doubles when VEC16) <- the
    // in-plane offset off[t] of its source, computed once (every staged node of an interior tile is a real node or
    // overhang).  TMA: one thread arms the slot's mbarrier with the byte count and issues the tile copy.
    constexpr int LD = TMA ? 1 : (VEC16 ? C::LD16 : C::LD8);
    constexpr int NSLOT = VEC16 ? C::PLANE / 2 : C::PLANE;
    constexpr unsigned PLANE_BYTES = C::PLANE * sizeof(double);
    int off[LD];
    if constexpr (TMA) {
      off[0] = 0;
    } else {
#pragma unroll
      for (int t = 0; t < LD; ++t) {
        const int idx = tid + t * C::NT;
        off[t] = -1;
        if (idx < NSLOT) {
          const int e0 = VEC16 ? 2 * idx : idx;
          const int sy = e0 / C::SX, sx = e0 - sy * C::SX;
          const int gi = x0 - HALO + sx, gj = y0 - HALO + sy;  // >= 2, >= 1
          if (gi < g.nx && gj < g.ny) off[t] = gi + g.pitch * gj;
        }
      }
    }
    // plane k (global index, must exist) -> ring slot s
    auto stage = [&](const int s, const int k) {
      if constexpr (TMA) {
        if (tid == 0) {
          mbar_expect_tx(&full[s], PLANE_BYTES);
          tma_load_3d(S + s * C::PLANE, tmap_ptr, x0 - HALO, y0 - HALO, k - g.k0 + HALO, &full[s]);
        }
      } else {
        const double* src = p + (long long)(k - g.k0 + HALO) * g.plane;
        double* dstS = S + s * C::PLANE;
#pragma unroll
        for (int t = 0; t < LD; ++t) {
          const int idx = tid + t * C::NT;
          if (idx < NSLOT) {
            const bool ok = off[t] >= 0;
            if (VEC16) cp_async_16(dstS + 2 * idx, src + (ok ? off[t] : 0), ok);
            else cp_async_8(dstS + idx, src + (ok ? off[t] : 0), ok);
          }
        }
      }
    };
    auto wait_slot = [&](const int s) {  // TMA: until the tile of slot s has landed
      if constexpr (TMA) {
        mbar_wait(&full[s], (phases >> s) & 1u);
        phases ^= 1u << s;
      }
    };
    const int my_off = (YB * ty) * C::SX + 2 * tx;

    // ---- planes next to a z face owned by this chunk: gather form, once
    const int ua = max(ka, 3), ub = min(kb, g.nz - 3);  // the uniform range [ua, ub)
#pragma unroll 1
    for (int face = 0; face < 2; ++face) {
      const int f0 = face == 0 ? ka : max(ka, max(g.nz - 3, 3));
      const int f1 = face == 0 ? min(kb, min(3, g.nz)) : kb;
      if (f0 >= f1) continue;
      const int s0 = max(f0 - HALO, 0), s1 = min(f1 + HALO, g.nz);  // at most 5 planes
      for (int k = s0; k < s1; ++k) stage(k - s0, k);
      if constexpr (TMA) {
        for (int k = s0; k < s1; ++k) wait_slot(k - s0);
      } else {
        cp_async_commit();
        cp_async_wait<0>();
      }
      __syncthreads();
      dot += z_face_block<YB, C::SX, C::PLANE>(g, U, ztab, S + my_off + 2 * C::SX + 2,
                                               q + (long long)(f0 - g.k0 + HALO) * g.plane + (long long)g.pitch * j0 + i0, f0, f1,
                                               s0, nrows, i0 < xlim, i0 + 1 < xlim);
      __syncthreads();  // before the slots are staged again
    }

    // ---- the uniform march
    double A[5][YB][2];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
      for (int r = 0; r < YB; ++r) A[d][r][0] = A[d][r][1] = 0.0;
    const int kstart = ua - HALO, kend = ua < ub ? ub + HALO : ua - HALO;
    int kp = kstart;
    int issued = kstart;
    const double* src = p + (long long)(kstart - g.k0 + HALO) * g.plane;  // cp.async: next plane to stage
    // cp.async variant of `stage` with a running source pointer (no 64-bit multiply inside the march)
    auto stage_next = [&](const int s) {
      if constexpr (TMA) {
        stage(s, issued);
      } else {
        double* dstS = S + s * C::PLANE;
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
      }
      ++issued;
    };
#pragma unroll
    for (int s = 0; s < ST - 3; ++s) {
      if (issued < kend) stage_next(s);
      if constexpr (!TMA) cp_async_commit();
    }
    int slot = 0;
    double* qdst = q + (long long)(ua - g.k0 + HALO) * g.plane + (long long)g.pitch * j0 + i0;  // output plane ua

    auto advance_pipeline = [&]() {
      if constexpr (TMA) wait_slot(slot);
      else cp_async_wait<ST - 4>();
      __syncthreads();  // every thread is done with plane kp-3: its slot is staged again
      const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;
      if (issued < kend) stage_next(fs);
      if constexpr (!TMA) cp_async_commit();
    };
    auto store_out = [&](const double (&out)[YB][2]) {
      if (kp - HALO < ua) return;
      const int s2 = slot >= 2 ? slot - 2 : slot + ST - 2;
      const double* Sp = S + s2 * C::PLANE + my_off + 2 * C::SX + 2;
#pragma unroll
      for (int r = 0; r < YB; ++r) {
        if (r < nrows) {
          const double2 pv = lds2(Sp + r * C::SX);
          if (VEC16 && i0 + 1 < xlim) {  // (an odd last column of the box range -- odd nx -- is stored alone below)
            *reinterpret_cast<double2*>(qdst + (long long)r * g.pitch) = make_double2(out[r][0], out[r][1]);
            dot = fma(pv.x, out[r][0], dot);
            dot = fma(pv.y, out[r][1], dot);
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
    uni_plane_fast<YB, C::SX>(S + slot * C::PLANE + my_off, U, A[(PH + 0) % 5], A[(PH + 1) % 5], A[(PH + 2) % 5],   \
                              A[(PH + 3) % 5], A[(PH + 4) % 5]);                                                    \
    store_out(A[(PH + 0) % 5]);                                                                                     \
    _Pragma("unroll") for (int r = 0; r < YB; ++r) A[(PH + 0) % 5][r][0] = A[(PH + 0) % 5][r][1] = 0.0;             \
    slot = (slot + 1 == ST) ? 0 : slot + 1;                                                                         \
    ++kp;                                                                                                           \
  }
    if (ua < ub) {
      for (;;) {
        LSB_FUSED_STEP(0)
        LSB_FUSED_STEP(1)
        LSB_FUSED_STEP(2)
        LSB_FUSED_STEP(3)
        LSB_FUSED_STEP(4)
      }
    }
#undef LSB_FUSED_STEP
    if constexpr (!TMA) cp_async_wait<0>();
  }
  return dot;
}

// TMA = true: the box tiles are staged by the TMA unit (needs VEC16: an even row pitch, and `tmap` describing p);
// the strips keep cp.async (48-byte rows).
template <int TYT, int ST, int MINB, bool VEC16, bool TMA = false>
__global__ void __launch_bounds__(32 * TYT, MINB)
    k_fdi_fused(const GridDesc g, const UniConsts U, const StencilWeights w, const FusedPlan pl, const double* __restrict__ p,
                double* __restrict__ q, const double* __restrict__ ztab, double* __restrict__ partials,
                unsigned int* __restrict__ counter, CgScalars* __restrict__ sc, const int cg_mode,
                const LSB_GRID_CONSTANT TmapArg tmap) {
  constexpr int NT = 32 * TYT;
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const int tid = threadIdx.x + 32 * threadIdx.y;
  const unsigned long long t_begin = tid == 0 ? trace_begin() : 0ull;
  unsigned phases = 0;
  const double dot = fused_cta<TYT, ST, VEC16, TMA>(g, U, w, pl, p, q, ztab, &tmap, (int)blockIdx.x, true, phases);
  const double bsum = block_sum<NT>(dot, red);
  if (tid == 0) trace_end(t_begin, blockIdx.x);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[0] = total;
}

}  // namespace lsb

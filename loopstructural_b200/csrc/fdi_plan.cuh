// Launch planning of the tiled FDI operator (host code, no device calls): which planes, tiles and columns each of the
// three launches of csrc/fdi_main.cuh covers.  Shared by csrc/api.cu (the launches) and tests/hostsim (the CPU
// emulation of the same kernels), so that the plan itself is under test without a device.
#pragma once
#include <algorithm>
#include <functional>
#include <queue>
#include <vector>

#include "fdi_main.cuh"
#include "fdi_fused.cuh"

namespace lsb {

// Stencil launch plans.  IDs 0,1,2,4: k_fdi_box (interior box, uniform coefficients) beside k_fdi_shell (everything
// else); -1: k_fdi_normal everywhere (always used on grids too small to have a box).
// X(ID, YB rows per thread, TYT warps, ST ring planes, MINB CTAs per SM)
// (measured and dropped, profiles/r01g: <2,16,8,1> <2,8,6,3> <3,8,8,1> <3,4,8,2>)
#define LSB_BOX_CFGS(X) \
  X(0, 4, 4, 8, 2)      \
  X(1, 4, 8, 8, 1)      \
  X(2, 2, 8, 8, 2)      \
  X(4, 4, 4, 6, 3)

// z-chunk length that fills `slots` CTA slots in whole waves with the least halo re-staging
inline int pick_cz(int nplanes, long long tiles, long long slots, long long max_partials) {
  if (nplanes < 1) return 1;
  int best_cz = nplanes;
  double best = -1.0;
  const int max_nch = std::max(1, nplanes / 4);
  for (int nch = 1; nch <= max_nch && nch <= 4096; ++nch) {
    const int cz = (nplanes + nch - 1) / nch;
    const int real = (nplanes + cz - 1) / cz;
    const long long ctas = tiles * real;
    if (ctas > max_partials) break;
    const long long waves = (ctas + slots - 1) / slots;
    const double eff = (double)ctas / (double)(waves * slots) * (double)cz / (double)(cz + 2 * HALO);
    if (eff > best + 1e-9) {
      best = eff;
      best_cz = cz;
    }
  }
  return best_cz;
}


// Tiling of the box launches of a handle: tiles of 64 x (YB * TYT) nodes over the interior columns / rows, z chunks of
// `cz` planes over ALL owned planes (a chunk that touches a z face runs the z-general variant, the others the uniform
// one).  Returns false when the grid is too small to have a box (k_fdi_normal is used everywhere then).
struct BoxTiles {
  int cfg, ty, minb;
  int tiles_x, tiles_y, cz, chunks;
};
inline bool plan_box_tiles(const GridDesc& g, int cfg, const int cz_override, const int sm_count, const long long max_partials,
                           BoxTiles* out) {
  if (g.ny < 16 || g.nx < 16 || g.nz < 6) return false;
  int ty = 0, minb = 1;
  switch (cfg) {
#define X(ID, YB, TYT, ST, MINB) case ID: ty = YB * TYT; minb = MINB; break;
    LSB_BOX_CFGS(X)
#undef X
    default: cfg = 4; ty = 16; minb = 3;
  }
  BoxTiles b;
  b.cfg = cfg; b.ty = ty; b.minb = minb;
  b.tiles_x = (g.nx - 8 + 63) / 64;
  b.tiles_y = (g.ny - 6 + ty - 1) / ty;
  b.cz = cz_override ? cz_override : pick_cz(g.nzl, (long long)b.tiles_x * b.tiles_y, (long long)sm_count * minb, max_partials);
  b.cz = std::min(std::max(b.cz, 3), BOX_ZGEN_MAX_STEPS - 2 * HALO);  // a face chunk covers planes 0..2 / nz-3..nz-1
  b.chunks = g.nzl > 0 ? (g.nzl + b.cz - 1) / b.cz : 0;
  if ((long long)b.tiles_x * b.tiles_y * b.chunks > max_partials) return false;
  *out = b;
  return true;
}

// Output planes of the three box launches of a handle owning [k0,k1) of nz planes, z chunk cz:
//   out[0..1] = A, out[2..3] = B: ranges of the z-general launch (k_fdi_box<ZGEN>), next to the z faces the handle owns;
//               short (zface planes: that variant is ~1.3x slower per plane), A non-empty whenever B is;
//   out[4..5] = U: range of the uniform launch, every plane of it satisfies 3 <= k <= nz-4.
// A, U, B are disjoint and cover [k0,k1).  Exported as lsb200_plan_z_ranges for the CPU tests.
inline void plan_z_ranges(int nz, int k0, int k1, int cz, int zface, int out[6]) {
  const bool lo_face = k0 < 3, hi_face = k1 > nz - 3;
  const int zc = std::max(3, std::min(cz, zface));
  const int zl = lo_face ? std::min(k1, k0 + zc) : k0;
  const int zh = hi_face ? std::max(zl, k1 - zc) : k1;
  int za0 = 0, za1 = 0, zb0 = 0, zb1 = 0;
  if (lo_face) { za0 = k0; za1 = zl; }
  if (hi_face && zh < k1) { zb0 = zh; zb1 = k1; }
  if (za1 <= za0 && zb1 > zb0) {  // only the upper face: it becomes range A, the launch has one z block
    za0 = zb0; za1 = zb1; zb0 = zb1 = 0;
  }
  out[0] = za0; out[1] = za1; out[2] = zb0; out[3] = zb1;
  out[4] = zl; out[5] = std::max(zl, zh);
}

inline UniConsts make_uni_consts(const StencilWeights& w) {
  UniConsts U;
  U.cc = w.axx + w.ayy + w.azz + 4.0 * (w.bxy + w.bxz + w.byz) + w.ridge2;
  U.cx1 = w.axx * (-2.0 / 3.0);
  U.cx2 = w.axx * (1.0 / 6.0) - 2.0 * (w.bxy + w.bxz);
  U.cy1 = w.ayy * (-2.0 / 3.0);
  U.cy2 = w.ayy * (1.0 / 6.0) - 2.0 * (w.bxy + w.byz);
  U.bxy = w.bxy;
  U.kz = w.azz * (1.0 / 6.0) - 2.0 * (w.bxz + w.byz);
  U.bxz = w.bxz; U.byz = w.byz;
  U.zt1 = w.azz * (-2.0 / 3.0);
  U.cxy = w.axx + w.ayy + 4.0 * w.bxy;
  U.cx2xy = w.axx * (1.0 / 6.0) - 2.0 * w.bxy;
  U.cy2xy = w.ayy * (1.0 / 6.0) - 2.0 * w.bxy;
  U.g0 = 2.0 * (w.bxz + w.byz);
  U.azz = w.azz; U.dz = w.dz; U.ridge2 = w.ridge2;
  return U;
}

// The ring of columns k_fdi_shell computes for a handle (everything outside 4 <= i < nx-4, 3 <= j < ny-3, over every
// owned plane; the z faces belong to k_fdi_box<ZGEN>) and its z chunking; returns the number of CTAs of SNT threads.
inline long long make_shell_box(const GridDesc& g, const int shell_cz, const int SNT, ShellBox* out) {
  ShellBox sb;
  sb.ilo = 4; sb.ihi = g.nx - 4;
  sb.jlo = 3; sb.jhi = g.ny - 3;
  sb.klo = g.k0; sb.khi = g.k1;
  const int nzm = sb.khi - sb.klo;
  sb.ring_cols = (g.ny - (sb.jhi - sb.jlo)) * g.nx + (sb.jhi - sb.jlo) * (g.nx - (sb.ihi - sb.ilo));
  sb.ring_blocks = (sb.ring_cols + SNT - 1) / SNT;
  // z chunks of the ring: about 50 k marching threads in flight (measured: 16 planes at 256^3, 64 at 512^3, profiles/r01r)
  int chunks = shell_cz > 0 ? (nzm + shell_cz - 1) / shell_cz : (int)(50000 / std::max(sb.ring_cols, 1));
  chunks = std::max(1, std::min(chunks, std::max(1, nzm / 8)));
  sb.ring_cz = nzm > 0 ? (nzm + chunks - 1) / chunks : 1;
  sb.ring_chunks = nzm > 0 ? (nzm + sb.ring_cz - 1) / sb.ring_cz : 0;
  sb.capA0 = g.k0; sb.capA1 = sb.klo;
  sb.capB0 = sb.khi; sb.capB1 = g.k1;
  sb.cap_blocks = (sb.capA1 > sb.capA0 || sb.capB1 > sb.capB0) ? (g.nx * g.ny + SNT - 1) / SNT : 0;
  sb.ring_stagger = 0;
  *out = sb;
  return (long long)sb.ring_blocks * sb.ring_chunks + 2LL * sb.cap_blocks;
}

// ---- the single-launch plan (k_fdi_fused, csrc/fdi_fused.cuh) ------------------------------------------------------
// X(TYT warps per CTA, ST ring planes, MINB CTAs per SM): tiles of 64 x 4 TYT nodes
#define LSB_FUSED_CFGS(X) \
  X(4, 6, 3)              \
  X(3, 6, 4)

struct FusedChoice {
  int tyt, minb;
  FusedPlan pl;
  long long ctas;
  double cost;  // estimated critical path in box plane steps, the quantity the planner minimises
};

// Cost model (units: one plane step of a box CTA with the SM fully populated; measured on a B200, profiles/r2e):
//   box CTA      planes + 4 halo steps, + FACE_COST when the chunk owns planes next to a z face (z_face_block);
//   x / y strip  (planes + 4) * XS_REL / YS_REL (general coefficients: 344 / 476 instructions per warp step against 242 of
//                a TMA-staged box step, profiles/r2f).
// The launch is simulated as the hardware runs it: CTAs in block-index order onto the first free of (SMs x CTAs per SM)
// slots; the plan with the shortest makespan wins.
constexpr double FUSED_FACE_COST = 12.0, FUSED_XS_REL = 1.45, FUSED_YS_REL = 2.0;
constexpr double FUSED_CTA_COST = 4.0;  // prologue of a CTA (offset table, first planes in flight) in plane steps

inline FusedChoice fused_candidate(const GridDesc& g, const int tyt, const int minb, const int nch, const int scz_in,
                                   const int sm_count) {
  FusedChoice c;
  c.tyt = tyt;
  c.minb = minb;
  FusedPlan& pl = c.pl;
  const int ty = 4 * tyt, nt = 32 * tyt;
  pl.tiles_x = (g.nx - 8 + 63) / 64;
  pl.tiles_y = (g.ny - 6 + ty - 1) / ty;
  // chunks of cz planes, the first / last shorter by FACE_COST planes when they own a z face
  const bool lo_face = g.k0 < 3, hi_face = g.k1 > g.nz - 3;
  const int d_lo = (lo_face && nch > 1) ? (int)FUSED_FACE_COST : 0, d_hi = (hi_face && nch > 1) ? (int)FUSED_FACE_COST : 0;
  pl.cz = std::max(1, (g.nzl + d_lo + d_hi + nch - 1) / nch);
  pl.lo_len = std::max(1, std::min(pl.cz - d_lo, g.nzl));
  pl.chunks = 1 + (g.nzl > pl.lo_len ? (g.nzl - pl.lo_len + pl.cz - 1) / pl.cz : 0);
  pl.n_box = pl.tiles_x * pl.tiles_y * pl.chunks;
  pl.xs_y = (g.ny + nt - 1) / nt;
  pl.ys_x = (g.nx - 8 + nt - 1) / nt;
  const int scz = std::max(2, std::min(scz_in, g.nzl));
  pl.schunks = (g.nzl + scz - 1) / scz;
  pl.scz = (g.nzl + pl.schunks - 1) / pl.schunks;
  pl.schunks = (g.nzl + pl.scz - 1) / pl.scz;
  pl.n_xs = 2 * pl.xs_y * pl.schunks;
  pl.n_ys = 2 * pl.ys_x * pl.schunks;
  c.ctas = (long long)pl.n_box + pl.n_xs + pl.n_ys;
  // costs, in block-index order (the order in which the hardware hands CTAs to free slots): box tiles chunk by chunk,
  // then x strips, then y strips
  const int last_len = g.nzl - pl.lo_len - std::max(0, pl.chunks - 2) * pl.cz;  // planes of the last chunk (chunks >= 2)
  const long long slots = (long long)sm_count * minb;
  std::priority_queue<double, std::vector<double>, std::greater<double>> free_at;  // list scheduling on `slots` slots
  double makespan = 0.0;
  auto run = [&](const double cost, const long long count) {
    for (long long i = 0; i < count; ++i) {
      double start = 0.0;
      if ((long long)free_at.size() >= slots) {
        start = free_at.top();
        free_at.pop();
      }
      free_at.push(start + cost);
      makespan = std::max(makespan, start + cost);
    }
  };
  for (int ch = 0; ch < pl.chunks; ++ch) {
    const int len = ch == 0 ? pl.lo_len : (ch == pl.chunks - 1 ? last_len : pl.cz);
    const bool lo = lo_face && ch == 0, hi = hi_face && ch == pl.chunks - 1;
    run(FUSED_CTA_COST + len + 2 * HALO + (lo ? FUSED_FACE_COST : 0.0) + (hi ? FUSED_FACE_COST : 0.0), (long long)pl.tiles_x * pl.tiles_y);
  }
  run(FUSED_CTA_COST + (pl.scz + 2 * HALO) * FUSED_XS_REL, pl.n_xs);
  run(FUSED_CTA_COST + (pl.scz + 2 * HALO) * FUSED_YS_REL, pl.n_ys);
  // a CTA of a 96-thread configuration holds a quarter of an SM, one of 128 threads a third: scale to SM time
  c.cost = makespan * (32.0 * tyt * minb) / 384.0;
  return c;
}

// The plan of a handle, or false when the grid is too small for it (k_fdi_normal is used everywhere then).
// cfg_override: TYT (3 or 4), 0 = choose; cz_override / scz_override: chunk lengths, 0 = choose.
inline bool plan_fused(const GridDesc& g, const int sm_count, const long long max_partials, const int cfg_override,
                       const int cz_override, const int scz_override, FusedChoice* out) {
  if (g.nx < 16 || g.ny < 16 || g.nz < 6 || g.nzl < 1) return false;
  bool have = false;
  FusedChoice best{};
  // strip chunkings: every count of chunks up to 48 (chunk lengths nzl, nzl/2, ...), at least 6 planes each
  std::vector<int> scz_list;
  for (int n = 1; n <= 48; ++n) {
    const int scz = (g.nzl + n - 1) / n;
    if (scz < 6 && n > 1) break;
    if (scz_list.empty() || scz_list.back() != scz) scz_list.push_back(scz);
  }
#define X(TYT, ST, MINB)                                                                                \
  if (cfg_override == 0 || cfg_override == TYT) {                                                       \
    const int max_nch = cz_override > 0 ? 1 : std::max(1, std::min(g.nzl / 8, 24));                                   \
    for (int nch = 1; nch <= max_nch && nch <= 4096; ++nch) {                                           \
      const int n = cz_override > 0 ? std::max(1, (g.nzl + cz_override - 1) / cz_override) : nch;       \
      bool too_many = false;                                                                            \
      for (const int scz : scz_list) {                                                                  \
        const FusedChoice c = fused_candidate(g, TYT, MINB, n, scz_override > 0 ? scz_override : scz, sm_count); \
        if (c.ctas > max_partials) { too_many = true; continue; }                                       \
        if (!have || c.cost < best.cost - 1e-9) { best = c; have = true; }                              \
        if (scz_override > 0) break;                                                                    \
      }                                                                                                 \
      if (too_many && nch > 1) break;                                                                   \
    }                                                                                                   \
  }
  LSB_FUSED_CFGS(X)
#undef X
  if (!have) return false;
  *out = best;
  return true;
}

}  // namespace lsb

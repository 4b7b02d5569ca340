// Launch planning of the FDI operator (host code, no device calls): which tiles, z chunks and strips the CTAs of
// k_fdi_fused (csrc/fdi_fused.cuh) take.  Shared by csrc/api.cu (the launch) and tests/hostsim (the CPU emulation of
// the same kernel), so that the plan itself is under test without a device.
#pragma once
#include <algorithm>
#include <functional>
#include <queue>
#include <vector>

#include "fdi_main.cuh"
#include "fdi_fused.cuh"

namespace lsb {

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

// ---- the single-launch plan (k_fdi_fused, csrc/fdi_fused.cuh) ------------------------------------------------------
// X(TYT warps per CTA, ST ring planes, MINB CTAs per SM): tiles of 64 x 4 TYT nodes.  (64 x 8 tiles with 5 CTAs per SM,
// X(2, 6, 5), were measured: no faster at 256^3 (69.2 us), slower at 512^3 and on slabs, and the 20 % extra staging
// traffic costs the vector kernels of the iteration their L2, profiles/r2ae_tiles_64x8.jsonl)
#define LSB_FUSED_CFGS(X) \
  X(4, 6, 3)              \
  X(3, 6, 4)

struct FusedChoice {
  int tyt, minb;
  FusedPlan pl;
  long long ctas;
  double cost;  // estimated critical path in box plane steps, the quantity the planner minimises
};

// Cost model (units: one plane step of a box CTA with the SM fully populated; measured on a B200 with the per-CTA
// timeline of a launch, lsb200_debug_cta_trace, profiles/r2j_cta_trace_256.txt -- 320 box CTAs of 44-56 planes, then 80
// strip CTAs of 26 planes):
//   box CTA      planes + 4 halo steps, + FACE_COST when the chunk owns planes next to a z face (z_face_block);
//   x / y strip  (planes + 4) * XS_REL / YS_REL: a strip step is 1.4 us against the 1.0 us of a box step although it moves
//                a quarter of the nodes (general coefficients, 8-byte copies: 344 / 476 instructions per warp step against
//                242 of a TMA-staged box step).  A "recalibration" to 0.35 / 0.45 from a misread timeline made every
//                256^3 matvec 117 us instead of 71 (profiles/r2o_cost_model.jsonl): strips are short and many on purpose.
// The launch is simulated as the hardware runs it: CTAs in block-index order onto the first free of (SMs x CTAs per SM)
// slots; the plan with the shortest makespan wins.  LSB200_FUSED_MODEL=face,xs,ys overrides the three constants.
static double FUSED_FACE_COST = 12.0, FUSED_XS_REL = 1.45, FUSED_YS_REL = 2.0;
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

// TEST INFRASTRUCTURE ONLY: runs the FDI regularisation operator -- the single launch k_fdi_fused
// (loopstructural_b200/csrc/fdi_fused.cuh, unmodified, with the launch plan of csrc/fdi_plan.cuh that csrc/api.cu's
// launch_stencil uses) and the small-grid kernel k_fdi_normal -- on the CPU through tests/hostsim/cuda_emu.h, for ONE
// handle: a full grid or a z slab [k0,k1) whose halo planes are filled from the global vector as the NCCL halo exchange
// would.  tests/test_stencil_emu.py compares it with the oracle.
//
//   g++ -O1 -std=c++20 -pthread -mfma -DLSB_EMU -I/usr/local/cuda/include -fPIC -fvisibility=hidden -Wl,-Bsymbolic -shared \
//       -o tests/hostsim/libstencil_emu.so tests/hostsim/stencil_emu.cpp
#include "cuda_emu.h"

#include <string.h>

#include "../../loopstructural_b200/csrc/fdi_plan.cuh"

using namespace lsb;

namespace {
constexpr int MAX_PARTIALS = 1 << 18;  // as in csrc/api.cu

StencilWeights weights_from(const double* ew, double ridge2) {  // make_weights of csrc/api.cu
  StencilWeights w;
  w.axx = ew[0] * ew[0]; w.ayy = ew[1] * ew[1]; w.azz = ew[2] * ew[2];
  w.bxy = 0.25 * ew[3] * ew[3]; w.bxz = 0.25 * ew[4] * ew[4]; w.byz = 0.25 * ew[5] * ew[5];
  w.dx = ew[6] * ew[6]; w.dy = ew[7] * ew[7]; w.dz = ew[8] * ew[8];
  w.ridge2 = ridge2;
  return w;
}
}  // namespace

#define HOSTSIM_API __attribute__((visibility("default")))
extern "C" {

// dims = {nx, ny, nz, k0, k1}; ew = the nine effective row weights (LSB200_W_* order: mixed terms as given to
// lsb200_set_regularisation, i.e. already divided by 4 ... see weights_from); p_global: nx*ny*nz; q_owned: nx*ny*(k1-k0).
// k_fdi_normal everywhere (grids too small for box tiles, LSB200_STENCIL_MAIN=-1).  info[6] = cz, info[7] = CTAs launched.
// pq_out = p.q over the owned planes as the kernel sums it.
HOSTSIM_API int emu_fdi_normal(const int* dims, const double* ew, double ridge2, int cz_override, const double* p_global,
                               double* q_owned, double* pq_out, int* info) {
  GridDesc g{};
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
  g.k0 = dims[3]; g.k1 = dims[4]; g.nzl = g.k1 - g.k0;
  g.pitch = g.nx;
  g.plane = (long long)g.pitch * g.ny;
  g.nloc = g.plane * g.nzl;
  g.nall = g.plane * (g.nzl + 2 * HALO);
  for (int a = 0; a < 3; ++a) { g.origin[a] = 0.0; g.step[a] = 1.0; }
  const StencilWeights w = weights_from(ew, ridge2);
  // node vectors as ensure_vectors lays them out: zeroed guard plane, nzl + 4 planes, guard plane
  const size_t guard = ((size_t)g.plane + 31) & ~(size_t)31;
  std::vector<double> pst((size_t)g.nall + 2 * guard, 0.0), qst((size_t)g.nall + 2 * guard, nan(""));
  double* p = pst.data() + guard;
  double* q = qst.data() + guard;
  for (int k = g.k0 - HALO; k < g.k1 + HALO; ++k)  // owned planes + what the halo exchange delivers
    if (k >= 0 && k < g.nz)
      memcpy(p + (size_t)(k - g.k0 + HALO) * g.plane, p_global + (size_t)k * g.plane, sizeof(double) * g.plane);
  std::vector<double> ztab((size_t)g.nz * SHELL_ZSTRIDE);
  shell_fill_z_table(g.nz, ztab.data());
  std::vector<double> partials((size_t)6 * MAX_PARTIALS, 0.0);
  unsigned int counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CgScalars sc{};
  const bool vec16 = (g.pitch % 2) == 0;
  int ctas = 0;
  for (int i = 0; i < 9; ++i) info[i] = 0;
  {
    // k_fdi_normal everywhere: plan_stencil<2,32,8>(h, 2) of csrc/api.cu with one z chunk per 16 planes
    using C = StencilCfg<2, 32, 8, 4>;
    const int tiles_x = (g.nx + C::TX - 1) / C::TX, tiles_y = (g.ny + C::TY - 1) / C::TY;
    const int cz = cz_override > 0 ? cz_override : std::max(4, std::min(g.nzl, 16));
    const dim3 grid(tiles_x, tiles_y, (g.nzl + cz - 1) / cz);
    const size_t smem = sizeof(double) * C::PLANE * 4;
    if (vec16) lsb_emu::launch(grid, dim3(32, 8), smem, [&] { k_fdi_normal<2, 32, 8, 2, 4, true>(g, w, p, q, cz, partials.data(), counters + 0, &sc, 0); });
    else lsb_emu::launch(grid, dim3(32, 8), smem, [&] { k_fdi_normal<2, 32, 8, 2, 4, false>(g, w, p, q, cz, partials.data(), counters + 0, &sc, 0); });
    ctas += grid.x * grid.y * grid.z;
    info[6] = cz;
    *pq_out = sc.pq_part[0];
  }
  info[7] = ctas;
  for (int k = 0; k < g.nzl; ++k)
    memcpy(q_owned + (size_t)k * g.plane, q + (size_t)(k + HALO) * g.plane, sizeof(double) * g.plane);
  // nothing may be written outside the owned planes (halo and guard planes keep their sentinel)
  for (size_t i = 0; i < qst.size(); ++i) {
    const bool owned = i >= guard + (size_t)HALO * g.plane && i < guard + (size_t)(HALO + g.nzl) * g.plane;
    if (!owned && qst[i] == qst[i]) return -2;
  }
  return 0;
}

// The single-launch plan (k_fdi_fused + plan_fused, csrc/fdi_fused.cuh / fdi_plan.cuh), same conventions as
// emu_fdi_matvec.  cfg: TYT (3 or 4), 0 = the planner's choice; cz / scz: forced chunk lengths of the box / strip CTAs.
// info[0] = TYT, [1] = cz, [2] = chunks, [3] = box CTAs, [4] = scz, [5] = strip CTAs, [6] = tiles_x, [7] = tiles_y.
HOSTSIM_API int emu_fdi_fused(const int* dims, const double* ew, double ridge2, int cfg, int cz_override, int scz_override, int sm_count,
                              const double* p_global, double* q_owned, double* pq_out, int* info) {
  GridDesc g{};
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
  g.k0 = dims[3]; g.k1 = dims[4]; g.nzl = g.k1 - g.k0;
  g.pitch = g.nx;
  g.plane = (long long)g.pitch * g.ny;
  g.nloc = g.plane * g.nzl;
  g.nall = g.plane * (g.nzl + 2 * HALO);
  for (int a = 0; a < 3; ++a) { g.origin[a] = 0.0; g.step[a] = 1.0; }
  const StencilWeights w = weights_from(ew, ridge2);
  const size_t guard = ((size_t)g.plane + 31) & ~(size_t)31;
  std::vector<double> pst((size_t)g.nall + 2 * guard, 0.0), qst((size_t)g.nall + 2 * guard, nan(""));
  double* p = pst.data() + guard;
  double* q = qst.data() + guard;
  for (int k = g.k0 - HALO; k < g.k1 + HALO; ++k)
    if (k >= 0 && k < g.nz)
      memcpy(p + (size_t)(k - g.k0 + HALO) * g.plane, p_global + (size_t)k * g.plane, sizeof(double) * g.plane);
    else  // planes outside the grid are never read: poison them
      for (long long t = 0; t < g.plane; ++t) p[(size_t)(k - g.k0 + HALO) * g.plane + t] = nan("");
  std::vector<double> ztab((size_t)g.nz * SHELL_ZSTRIDE);
  shell_fill_z_table(g.nz, ztab.data());
  std::vector<double> partials((size_t)MAX_PARTIALS, 0.0);
  unsigned int counter = 0;
  CgScalars sc{};
  FusedChoice fc;
  if (!plan_fused(g, sm_count, MAX_PARTIALS, cfg, cz_override, scz_override, &fc)) return -3;
  const UniConsts U = make_uni_consts(w);
  const bool vec16 = (g.pitch % 2) == 0;
  const unsigned ctas = (unsigned)(fc.pl.n_box + fc.pl.n_xs + fc.pl.n_ys);
  switch (fc.tyt) {
#define X(TYT, ST, MINB)                                                                                             \
  case TYT: {                                                                                                        \
    const size_t smem = sizeof(double) * MainCfg<4, TYT, ST>::PLANE * ST;                                            \
    if (vec16) lsb_emu::launch(dim3(ctas), dim3(32, TYT), smem, [&] { k_fdi_fused<TYT, ST, MINB, true>(g, U, w, fc.pl, p, q, ztab.data(), partials.data(), &counter, &sc, 0, TmapArg{}); }); \
    else lsb_emu::launch(dim3(ctas), dim3(32, TYT), smem, [&] { k_fdi_fused<TYT, ST, MINB, false>(g, U, w, fc.pl, p, q, ztab.data(), partials.data(), &counter, &sc, 0, TmapArg{}); }); \
  } break;
    LSB_FUSED_CFGS(X)
#undef X
    default: return -4;
  }
  info[0] = fc.tyt; info[1] = fc.pl.cz; info[2] = fc.pl.chunks; info[3] = fc.pl.n_box; info[4] = fc.pl.scz; info[5] = fc.pl.n_xs + fc.pl.n_ys;
  info[6] = fc.pl.tiles_x; info[7] = fc.pl.tiles_y; info[8] = 2;
  *pq_out = sc.pq_part[0];
  for (int k = 0; k < g.nzl; ++k)
    memcpy(q_owned + (size_t)k * g.plane, q + (size_t)(k + HALO) * g.plane, sizeof(double) * g.plane);
  for (size_t i = 0; i < qst.size(); ++i) {
    const bool owned = i >= guard + (size_t)HALO * g.plane && i < guard + (size_t)(HALO + g.nzl) * g.plane;
    if (!owned && qst[i] == qst[i]) return -2;
  }
  return 0;
}

HOSTSIM_API void emu_set_fused_model(double face, double xs, double ys) { FUSED_FACE_COST = face; FUSED_XS_REL = xs; FUSED_YS_REL = ys; }

// the plan alone (no launch): info = {TYT, cz, chunks, lo_len, box CTAs, scz, x-strip CTAs, y-strip CTAs, MINB}, *cost = model cost
HOSTSIM_API int emu_plan_fused(const int* dims, int cfg, int cz_override, int scz_override, int sm_count, int* info, double* cost) {
  GridDesc g{};
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
  g.k0 = dims[3]; g.k1 = dims[4]; g.nzl = g.k1 - g.k0;
  FusedChoice fc;
  if (!plan_fused(g, sm_count, MAX_PARTIALS, cfg, cz_override, scz_override, &fc)) return -3;
  info[0] = fc.tyt; info[1] = fc.pl.cz; info[2] = fc.pl.chunks; info[3] = fc.pl.lo_len; info[4] = fc.pl.n_box;
  info[5] = fc.pl.scz; info[6] = fc.pl.n_xs; info[7] = fc.pl.n_ys; info[8] = fc.minb;
  *cost = fc.cost;
  return 0;
}

}  // extern "C"

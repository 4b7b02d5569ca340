// TEST INFRASTRUCTURE ONLY: the P1 / TetMesh kernels (loopstructural_b200/csrc/p1.cuh, unmodified) on the CPU through
// tests/hostsim/cuda_emu.h, for ONE handle -- the full mesh or a z slab [k0,k1) with its halo planes filled as the NCCL
// exchange would.  tests/test_p1_emu.py compares with the oracle (explicit face-jump rows, point rows).
//
//   g++ -O1 -std=c++20 -pthread -mfma -w -DLSB_EMU -I/usr/local/cuda/include -fPIC -fvisibility=hidden -Wl,-Bsymbolic \
//       -shared -o tests/hostsim/libp1_emu.so tests/hostsim/p1_emu.cpp
#include "cuda_emu.h"

#include <string.h>

#include <mutex>

#include "../../include/lsb200.h"

// atomics of the row kernels: the emulated threads of a CTA run concurrently
namespace { std::mutex g_atomic; }
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) {
  std::lock_guard<std::mutex> l(g_atomic);
  const unsigned long long old = *a;
  *a = old + v;
  return old;
}
inline double atomicAdd(double* a, double v) {
  std::lock_guard<std::mutex> l(g_atomic);
  const double old = *a;
  *a = old + v;
  return old;
}
inline unsigned long long atomicExch(unsigned long long* a, unsigned long long v) {
  std::lock_guard<std::mutex> l(g_atomic);
  const unsigned long long old = *a;
  *a = v;
  return old;
}
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }

#include "../../loopstructural_b200/csrc/p1.cuh"
#include "../../loopstructural_b200/csrc/p1_march.cuh"

using namespace lsb;

namespace {
constexpr int MAX_PARTIALS = 1 << 18;
GridDesc make_grid(const int* dims, const double* origin, const double* step, const bool even_pitch = false) {
  GridDesc g{};
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
  g.k0 = dims[3]; g.k1 = dims[4]; g.nzl = g.k1 - g.k0;
  g.pitch = even_pitch ? (g.nx + 1) & ~1 : g.nx;  // the library pads rows to an even length (api.cu: lsb200_create)
  g.plane = (long long)g.pitch * g.ny;
  g.nloc = g.plane * g.nzl;
  g.nall = g.plane * (g.nzl + 2 * HALO);
  for (int a = 0; a < 3; ++a) { g.origin[a] = origin[a]; g.step[a] = step[a]; }
  return g;
}
}  // namespace

#define HOSTSIM_API __attribute__((visibility("default")))
extern "C" {

// q = (w2 R^T R + ridge2 I) p on the owned planes and diag of the same; dims = {nx, ny, nz, k0, k1}
// tiles != 0: the k_p1_tiles plan of launch_regularisation with tile configuration tiles - 1 (LSB_P1_TILE_CFGS); returns 1
// then if the tile kernel was used, 0 if the plan fell back to the generic kernel (no interior).  tiles = 10 + ID of
// LSB_P1_MARCH_CFGS: k_p1_march, cz = z planes per chunk (0: the plan's own choice).  p_global, q_owned, diag_owned are dense
// (nx * ny per plane); the emulated device arrays have the library's even row pitch.
HOSTSIM_API int emu_p1_normal(const int* dims, const double* step, const long long* m, const double* table, double w2,
                              double ridge2, int tiles, int cz, const double* p_global, double* q_owned, double* diag_owned, double* pq_out) {
  const double origin[3] = {0.0, 0.0, 0.0};
  const GridDesc g = make_grid(dims, origin, step, true);
  std::vector<double> coef, dg;
  std::vector<int> off;
  size_t bad = 0;
  if (p1_compact_tables(g, m, table, coef, off, dg, &bad)) return -1;
  const P1Tables tb{{(int)m[0], (int)m[1], (int)m[2]}, coef.data(), off.data(), dg.data()};
  const size_t guard = ((size_t)g.plane + 31) & ~(size_t)31;
  std::vector<double> pst((size_t)g.nall + 2 * guard, 0.0), qst((size_t)g.nall + 2 * guard, nan("")), dst((size_t)g.nall + 2 * guard, 0.0);
  double* p = pst.data() + guard;
  double* q = qst.data() + guard;
  double* d = dst.data() + guard;
  for (int k = g.k0 - HALO; k < g.k1 + HALO; ++k)
    if (k >= 0 && k < g.nz)
      for (int j = 0; j < g.ny; ++j)
        memcpy(p + (size_t)(k - g.k0 + HALO) * g.plane + (size_t)j * g.pitch, p_global + ((size_t)k * g.ny + j) * g.nx, sizeof(double) * g.nx);
  std::vector<double> partials((size_t)MAX_PARTIALS, 0.0);
  unsigned int counter = 0;
  CgScalars sc{};
  constexpr int NT = 256;
  const long long n = (long long)g.nx * g.ny * g.nzl;
  const int nb = (int)std::max<long long>(1, std::min<long long>((n + NT - 1) / NT, 5));
  int used_tiles = 0;
  P1TilePlan pl;
  P1ClassTaps cls;
  const int mi[3] = {(int)m[0], (int)m[1], (int)m[2]};
  switch (tiles - 1) {
#define X(ID, TY, TZ, NW)                                                                                                \
  case ID:                                                                                                               \
    if (p1_plan_tiles(g, mi, table, TY, TZ, 32 * NW, MAX_PARTIALS, &pl, &cls)) {                                         \
      used_tiles = 1;                                                                                                    \
      lsb_emu::launch(dim3(pl.n_bnd_cta + pl.tiles_x * pl.tiles_y * pl.tiles_z), dim3(32 * NW),                          \
                      sizeof(double) * P1TileGeom<TY, TZ>::ELEMS, [&] {                                                   \
                        k_p1_tiles<TY, TZ, NW>(g, tb, cls, pl, w2, ridge2, p, q, partials.data(), &counter, &sc, 0);     \
                      });                                                                                                \
    }                                                                                                                    \
    break;
    LSB_P1_TILE_CFGS(X)
#undef X
#define M(ID, NW, NB, MINB)                                                                                              \
  case 9 + ID: {                                                                                                         \
    using MC = P1MarchCfg<NW, NB>;                                                                                       \
    P1MarchPlan mp;                                                                                                      \
    P1MarchTaps mt;                                                                                                      \
    if (p1_plan_march<NW, NB>(g, mi, table, 3, 2, MAX_PARTIALS, cz, &mp, &mt)) {                                         \
      used_tiles = 1;                                                                                                    \
      lsb_emu::launch(dim3(mp.n_cta + mp.b.n_bnd_cta), dim3(MC::NT), sizeof(double) * MC::PLANE * MC::ST,                \
                      [&] { k_p1_march<NW, NB, MINB>(g, tb, mt, mp, w2, ridge2, p, q, partials.data(), &counter, &sc, 0); }); \
    }                                                                                                                    \
  } break;
    LSB_P1_MARCH_CFGS(M)
    case 19: {  // tiles = 20: the split plan -- k_p1_march<4, 0, 3> marching only, then k_p1_boundary
      using MC = P1MarchCfg<4, 0>;
      P1MarchPlan mp;
      P1MarchTaps mt;
      if (p1_plan_march<4, 0>(g, mi, table, 3, 2, MAX_PARTIALS, cz, &mp, &mt, true)) {
        used_tiles = 1;
        CgScalars sc2{};
        unsigned int counter2 = 0;
        std::vector<double> partials2((size_t)MAX_PARTIALS, 0.0);
        if (mp.n_cta > 0)
          lsb_emu::launch(dim3(mp.n_cta), dim3(MC::NT), sizeof(double) * MC::PLANE * MC::ST,
                          [&] { k_p1_march<4, 0, 3>(g, tb, mt, mp, w2, ridge2, p, q, partials.data(), &counter, &sc, 0); });
        lsb_emu::launch(dim3(3), dim3(64), 0, [&] { k_p1_boundary<64>(g, tb, mp.b, w2, ridge2, p, q, partials2.data(), &counter2, &sc2, 0); });
        sc.pq_part[0] += sc2.pq_part[2];
      }
    } break;
#undef M
  }
  if (!used_tiles)
    lsb_emu::launch(dim3(nb), dim3(NT), 0, [&] { k_p1_normal<NT>(g, tb, w2, ridge2, p, q, partials.data(), &counter, &sc, 0); });
  lsb_emu::launch(dim3(nb), dim3(NT), 0, [&] { k_p1_diag(g, tb, w2, ridge2, d); });
  for (int k = 0; k < g.nzl; ++k)
    for (int j = 0; j < g.ny; ++j) {
      memcpy(q_owned + ((size_t)k * g.ny + j) * g.nx, q + (size_t)(k + HALO) * g.plane + (size_t)j * g.pitch, sizeof(double) * g.nx);
      memcpy(diag_owned + ((size_t)k * g.ny + j) * g.nx, d + (size_t)(k + HALO) * g.plane + (size_t)j * g.pitch, sizeof(double) * g.nx);
    }
  *pq_out = sc.pq_part[0];
  for (size_t i = 0; i < qst.size(); ++i) {
    const bool owned = i >= guard + (size_t)HALO * g.plane && i < guard + (size_t)(HALO + g.nzl) * g.plane;
    const bool padding = owned && (int)((i - guard) % (size_t)g.pitch) >= g.nx;
    if ((!owned || padding) && (qst[i] == qst[i] || dst[i] != 0.0)) return -2;
  }
  return used_tiles;
}

// k_p1_point_rows on a slab: rows as (n_out, 4) values, GLOBAL node ids (-1 = empty row), b, w^2; returns rows kept
HOSTSIM_API long long emu_p1_point_rows(const int* dims, const double* origin, const double* step, int kind, long long n,
                                        const double* data, int ncols, double weight, double* vals, long long* cols, double* b,
                                        double* w2) {
  const GridDesc g = make_grid(dims, origin, step);
  const int mult = kind == LSB200_ROWS_NORMAL ? 3 : 1;
  const long long nr = n * mult;
  std::vector<double> v((size_t)nr * 4), bb((size_t)nr), ww((size_t)nr);
  std::vector<int> cc((size_t)nr * 4);
  RowBlock rb{};
  rb.n = nr; rb.k = 4; rb.vals = v.data(); rb.cols = cc.data(); rb.b = bb.data(); rb.w2 = ww.data();
  unsigned long long status[2] = {0, 0};
  lsb_emu::launch(dim3(2), dim3(128), 0, [&] { k_p1_point_rows(g, kind, n, data, ncols, weight, rb, 0, status); });
  for (long long r = 0; r < nr; ++r) {
    for (int k = 0; k < 4; ++k) {
      vals[r * 4 + k] = v[(size_t)k * nr + r];
      const int lc = cc[(size_t)k * nr + r];
      cols[r * 4 + k] = lc < 0 ? -1 : (long long)lc + (long long)(g.k0 - HALO) * g.plane;  // storage index -> global node id
    }
    b[r] = bb[r];
    w2[r] = ww[r];
  }
  return (long long)status[0];
}

// lsb200_add_face_vector_rows on a slab: k_p1_face_vector_vals then k_rows_ingest; rows as (n, 8) values, GLOBAL node ids
// (-1 = row not kept by this slab), b, w^2; returns rows kept (or -1: a node id outside the mesh / outside the halo planes)
HOSTSIM_API long long emu_p1_face_vector_rows(const int* dims, const double* origin, const double* step, long long n,
                                              const long long* nodes_a, const long long* nodes_b, const double* vec, double weight,
                                              double* vals, long long* cols, double* b, double* w2) {
  const GridDesc g = make_grid(dims, origin, step);
  std::vector<double> tv((size_t)n * 8), tb((size_t)n), tw((size_t)n), v((size_t)n * 8), bb((size_t)n), ww((size_t)n);
  std::vector<long long> tc((size_t)n * 8);
  std::vector<int> cc((size_t)n * 8);
  lsb_emu::launch(dim3(2), dim3(64), 0, [&] { k_p1_face_vector_vals(g, n, nodes_a, nodes_b, vec, weight, tv.data(), tc.data(), tb.data(), tw.data()); });
  RowBlock rb{};
  rb.n = n; rb.k = 8; rb.vals = v.data(); rb.cols = cc.data(); rb.b = bb.data(); rb.w2 = ww.data();
  unsigned long long status[4] = {0, 0, 0, 0};
  lsb_emu::launch(dim3(2), dim3(64), 0, [&] { k_rows_ingest(g, n, 8, tv.data(), tc.data(), tb.data(), tw.data(), 1, rb, 0, status); });
  if (status[1] || status[2]) return -1;
  for (long long r = 0; r < n; ++r) {
    for (int k = 0; k < 8; ++k) {
      vals[r * 8 + k] = v[(size_t)k * n + r];
      const int lc = cc[(size_t)k * n + r];
      cols[r * 8 + k] = lc < 0 ? -1 : (long long)lc + (long long)(g.k0 - HALO) * g.plane;
    }
    b[r] = bb[r];
    w2[r] = ww[r];
  }
  return (long long)status[0];
}

// lsb200_add_directional_rows on a slab (a GRID handle: the row kernels are shared): rows as (6 * ncent, 4) values, GLOBAL
// node ids (-1 = empty slot), w^2 (0 = dead row); vec_global is the (nx*ny*nz, 3) field; returns rows kept
HOSTSIM_API long long emu_fdi_directional_rows(const int* dims, const double* origin, const double* step, const double* vec_global,
                                               double weight, double* vals, long long* cols, double* w2) {
  const GridDesc g = make_grid(dims, origin, step, true);
  const int ka = std::max(1, g.k0 - 1), kb = std::min(g.nz - 1, g.k1 + 1);
  if (g.nx < 3 || g.ny < 3 || kb <= ka) return 0;
  const long long ncent = (long long)(g.nx - 2) * (g.ny - 2) * (kb - ka), nr = 6 * ncent;
  std::vector<double> v((size_t)nr * 4), bb((size_t)nr), ww((size_t)nr);
  std::vector<int> cc((size_t)nr * 4);
  RowBlock rb{};
  rb.n = nr; rb.k = 4; rb.vals = v.data(); rb.cols = cc.data(); rb.b = bb.data(); rb.w2 = ww.data();
  unsigned long long status[4] = {0, 0, 0, 0};
  lsb_emu::launch(dim3(3), dim3(64), 0, [&] {
    k_fdi_directional_rows(g, ka, kb, vec_global + 3 * (long long)g.nx * g.ny * ka, weight, rb, status);
  });
  for (long long r = 0; r < nr; ++r) {
    for (int k = 0; k < 4; ++k) {
      vals[r * 4 + k] = v[(size_t)k * nr + r];
      const int lc = cc[(size_t)k * nr + r];
      if (lc < 0) { cols[r * 4 + k] = -1; continue; }
      // storage index (even row pitch) -> global node id
      const long long kl = lc / g.plane, rem = lc - kl * g.plane;
      const long long j = rem / g.pitch, i = rem - j * g.pitch;
      cols[r * 4 + k] = ((kl + g.k0 - HALO) * g.ny + j) * g.nx + i;
    }
    w2[r] = ww[r];
  }
  return (long long)status[0];
}

}  // extern "C"

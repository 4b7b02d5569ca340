// K5, third form: the P1 face-jump operator q = (w^2 R^T R + ridge^2 I) p by z-marching register tiles.
//
// k_p1_tiles (p1.cuh) computes a node from 15 / 33 shared-memory loads, one per tap: 41 us for 1.64 M nodes, bound by
// instruction issue and load latency (profiles/r2_final_ncu_full_k_p1_tiles_details.csv).  Here, as in the FDI box march
// (fdi_main.cuh), a thread owns a 2 x 2 patch of columns -- one column of each (x, y) parity class -- and marches along z
// in SCATTER form: an incoming plane is read once, as the 6 x 6 window around the patch (18 LDS.128 for 4 nodes), and
// every tap it feeds is applied to the accumulators of the five output planes k-2 .. k+2.  Which taps exist is topology
// (p1_march_taps.inc, generated from p1_tables.py and checked against the support's table when the plan is made), so the
// multiply-adds are straight-line code on registers with constant-bank coefficients: 96 DFMA + 18 LDS.128 per thread and
// plane against 4 x (24 LDS.64 + 24 DFMA + address arithmetic).  Planes are staged by 16-byte cp.async into a ring of six
// (three in flight); the ring keeps the plane of the output being finished, so p for the fused p.q comes from shared memory.
// Interior nodes only (2 <= i <= n-3 on every axis); the boundary shell (10 % of config 3) is the generic form
// (p1_node_generic: three dependent gathers per tap, about a microsecond per node and thread).  Done by the marching
// threads before their first plane it added 6 us to every CTA of an 11 us march (profiles/r2ai_p1_march_trace.txt), so a
// CTA has NB more warps that do nothing else: their gathers run under the march's arithmetic, and the marching warps
// meet at a named barrier of their own.
#pragma once
#include "p1.cuh"
#include "p1_march_taps.inc"

namespace lsb {

struct P1MarchTaps {
  double coef[8][LSB_P1M_MAXTAPS];  // class = pi + 2 pj + 4 pk, entries in the order of LSB_P1M_TAPS
};
struct P1MarchPlan {
  P1TilePlan b;            // kfirst / klast and the enumeration of the boundary nodes (p1_boundary_node)
  int tiles_x, tiles_y;    // tiles of 64 x TY interior nodes
  int cz, chunks;          // z chunks of the owned interior planes [kfirst, klast)
  int n_cta;               // tiles_x * tiles_y * chunks
  int split;               // the boundary nodes are another launch's (k_p1_boundary): the CTAs of this plan only march
};

// X(ID, marching warps, boundary warps, CTAs per SM the kernel is compiled for); LSB200_P1_MARCH = ID (0: k_p1_tiles instead).
// Measured on config 3 (118^3 nodes, profiles/r2ak_p1_march.txt), operator / CG iteration of the persistent loop in us:
// k_p1_tiles 41.1 / 58.3; 1: 27.5 / 54.3; 3: 25.2 / 48.0.  (Patches of 2 x 1 columns, twice the marching warps with half
// the accumulators each, were no faster: 28.4 / 49.2 -- the march is not short of warps.)
#define LSB_P1_MARCH_CFGS(X) \
  X(1, 4, 0, 3)              \
  X(2, 4, 2, 2)              \
  X(3, 4, 4, 2)              \
  X(4, 4, 2, 3)
#define LSB_P1_MARCH_DEFAULT 3
constexpr int P1M_BND_PER_THREAD = 6;  // boundary nodes per thread the plan accepts before it launches more CTAs than slots
template <int NW, int NB = 0>
struct P1MarchCfg {
  static constexpr int NM = 32 * NW;               // marching threads
  static constexpr int NBT = NB ? 32 * NB : NM;    // threads of a tile CTA that take boundary nodes
  static constexpr int NT = 32 * (NW + NB);
  static constexpr int TX = 64, TY = 2 * NW, SX = TX + 2 * HALO, SY = TY + 2 * HALO, PLANE = SX * SY, ST = 6;
  static constexpr int LD16 = (PLANE / 2 + NM - 1) / NM;
};

// barrier of the first `count` threads of a CTA (whole warps), id 1: the marching warps of a warp-specialised CTA
__device__ __forceinline__ void p1m_named_sync(const int count) {
#ifdef LSB_EMU
  lsb_emu::named_sync(1, count);
#else
  asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory");
#endif
}

// host: plan and coefficients, or false (no interior, an entry of the table outside the baked pattern, too many CTAs)
template <int NW, int NB>
inline bool p1_plan_march(const GridDesc& g, const int m[3], const double* table, const int sm_count, const int ctas_per_sm,
                          const long long max_blocks, const int cz_override, P1MarchPlan* out, P1MarchTaps* T,
                          const bool split = false) {
  using C = P1MarchCfg<NW, NB>;
  const int nc[3] = {g.nx - 1, g.ny - 1, g.nz - 1};
  for (int a = 0; a < 3; ++a)
    if (nc[a] == m[a] || nc[a] < 6) return false;
  if (g.pitch % 2 != 0) return false;
  P1MarchPlan pl{};
  P1ClassTaps unused;
  if (!p1_plan_tiles(g, m, table, 8, 4, C::NT, max_blocks, &pl.b, &unused)) return false;
  pl.tiles_x = (g.nx - 4 + C::TX - 1) / C::TX;
  pl.tiles_y = (g.ny - 4 + C::TY - 1) / C::TY;
  const int planes = pl.b.klast - pl.b.kfirst;
  // chunks: as many as fit the CTA slots at once (a second wave would double the launch), at least 4 planes per chunk
  // (each chunk re-stages 4 halo planes; small meshes leave most slots empty and are bound by the length of a march)
  const long long tiles = (long long)pl.tiles_x * pl.tiles_y, slots = (long long)sm_count * ctas_per_sm;
  const long long nb = pl.b.nbz + pl.b.nby + pl.b.nbx;
  int chunks = (int)std::max<long long>(1, slots / tiles);
  chunks = std::max(1, std::min(chunks, std::max(1, planes / 4)));
  if (cz_override > 0) chunks = std::max(1, (planes + cz_override - 1) / cz_override);
  for (;; --chunks) {
    pl.cz = planes > 0 ? (planes + chunks - 1) / chunks : 1;
    pl.chunks = planes > 0 ? (planes + pl.cz - 1) / pl.cz : 0;
    pl.n_cta = (int)(tiles * pl.chunks);
    // the boundary shell: spread over the boundary threads of the tile CTAs and boundary-only CTAs in the slots the
    // tiles leave free, as few nodes per thread as that allows (a thread walks its nodes one after the other, each a
    // chain of three dependent loads), at most P1M_BND_PER_THREAD before more CTAs than slots are launched
    const long long free_slots = std::max<long long>(0, slots - pl.n_cta);
    const long long threads = (long long)pl.n_cta * C::NBT + free_slots * C::NT;
    long long per = threads > 0 ? (nb + threads - 1) / threads : P1M_BND_PER_THREAD;
    per = std::max<long long>(1, std::min<long long>(per, P1M_BND_PER_THREAD));
    const long long carried = (long long)pl.n_cta * C::NBT * per, per_cta = (long long)C::NT * per;
    pl.b.n_bnd_cta = (carried >= nb || split) ? 0 : (int)((nb - carried + per_cta - 1) / per_cta);
    // they too belong to the first wave: fewer, longer chunks until they fit (or nothing is left to give)
    if (cz_override > 0 || chunks <= 1 || (long long)pl.n_cta + pl.b.n_bnd_cta <= slots) break;
  }
  pl.split = split ? 1 : 0;
  if ((long long)pl.n_cta + pl.b.n_bnd_cta > max_blocks) return false;
  memset(T, 0, sizeof(*T));
  for (int cls = 0; cls < 8; ++cls) {
    const int pi = cls & 1, pj = (cls >> 1) & 1, pk = cls >> 2;
    const size_t row = ((size_t)(2 + pk) * (m[1] + 1) + (2 + pj)) * (m[0] + 1) + 2 + pi;
    bool used[125] = {false};
    for (int e = 0; e < LSB_P1M_NTAPS[cls]; ++e) {
      const int t = (LSB_P1M_TAPS[cls][e][2] + 2) * 25 + (LSB_P1M_TAPS[cls][e][1] + 2) * 5 + (LSB_P1M_TAPS[cls][e][0] + 2);
      const double c = table[row * 125 + t];
      if (!(c == c)) return false;
      T->coef[cls][e] = c;
      used[t] = true;
    }
    for (int t = 0; t < 125; ++t)
      if (!used[t] && table[row * 125 + t] != 0.0) return false;  // a coupling the baked pattern does not know
  }
  *out = pl;
  return true;
}

// The work of CTA `vb` of `nvb`: its tile chunk and its share of the boundary nodes; returns this thread's share of p.q.
template <int NW, int NB>
__device__ __forceinline__ double p1_march_cta(const GridDesc& g, const P1Tables& tb, const P1MarchTaps& T, const P1MarchPlan& pl,
                                               const double w2, const double ridge2, const double* __restrict__ p,
                                               double* __restrict__ q, const int vb, const int nvb) {
  using C = P1MarchCfg<NW, NB>;
  constexpr int NT = C::NM, SX = C::SX, ST = C::ST, LD = C::LD16, NSLOT = C::PLANE / 2;  // NT: the marching threads
  LSB_DYN_SMEM(S);
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  double dot = 0.0;
  const bool has_tile = vb < pl.n_cta;
  const bool marcher = has_tile && tid < NT;
  int x0 = 0, y0 = 0, ka = 0, kb = 0;
  int off[LD];
  int issued = 0, kstart = 0, kend = 0;
  const double* src = p;
  auto stage_next = [&](const int s) {
    double* dstS = S + s * C::PLANE;
#pragma unroll
    for (int t = 0; t < LD; ++t) {
      const int idx = tid + t * NT;
      if (idx < NSLOT) {
        const bool ok = off[t] >= 0;
        cp_async_16(dstS + 2 * idx, src + (ok ? off[t] : 0), ok);
      }
    }
    src += g.plane;
    ++issued;
  };
  if (marcher) {
    const int tiles = pl.tiles_x * pl.tiles_y;
    const int ch = vb / tiles, trem = vb - ch * tiles;
    const int by = trem / pl.tiles_x, bx = trem - by * pl.tiles_x;
    x0 = 2 + C::TX * bx;
    y0 = 2 + C::TY * by;
    ka = pl.b.kfirst + ch * pl.cz;
    kb = min(ka + pl.cz, pl.b.klast);
    // staged pair idx <- in-plane offset of its source (pairs start on even columns: x0 - 2 is a multiple of 64)
#pragma unroll
    for (int t = 0; t < LD; ++t) {
      const int idx = tid + t * NT;
      off[t] = -1;
      if (idx < NSLOT) {
        const int e0 = 2 * idx;
        const int sy = e0 / SX, sx = e0 - sy * SX;
        const int gi = x0 - HALO + sx, gj = y0 - HALO + sy;
        if (gi < g.nx && gj < g.ny) off[t] = gi + g.pitch * gj;  // gi even, nx <= pitch: the partner is a node or row padding (zero)
      }
    }
    kstart = ka - HALO;
    kend = kb + HALO;  // input planes [kstart, kend): all inside the grid (2 <= ka, kb <= nz - 2)
    src = p + (long long)(kstart - g.k0 + HALO) * g.plane;
    issued = kstart;
#pragma unroll
    for (int s = 0; s < ST - 3; ++s) {
      if (issued < kend) stage_next(s);
      cp_async_commit();
    }
  }
  if (!pl.split && (!has_tile || NB == 0 || tid >= NT)) {
    // ---- a share of the boundary nodes: the boundary warps of a tile CTA (NB = 0: its marching threads, while their
    // first planes are on their way), every thread of a boundary-only CTA
    const long long nb = pl.b.nbz + pl.b.nby + pl.b.nbx;
    const long long all = (long long)pl.n_cta * C::NBT + (long long)(nvb - pl.n_cta) * C::NT;
    const long long me = has_tile ? (long long)vb * C::NBT + (NB ? tid - NT : tid)
                                  : (long long)pl.n_cta * C::NBT + (long long)(vb - pl.n_cta) * C::NT + tid;
    for (long long t = me; t < nb; t += all) {
      int i, j, kl;
      p1_boundary_node(g, pl.b, t, &i, &j, &kl);
      dot += p1_node_generic(g, tb, w2, ridge2, p, q, i, j, kl);
    }
  }
  if (marcher) {
    const int i0 = x0 + 2 * lane, j0 = y0 + 2 * wy;
    const bool xok = i0 + 1 <= g.nx - 3, xok0 = i0 <= g.nx - 3;  // columns i0, i0 + 1 interior?
    const bool yok0 = j0 <= g.ny - 3, yok1 = j0 + 1 <= g.ny - 3;
    const int my_off = (2 * wy) * SX + 2 * lane;  // staged element (row j0 - 2, column i0 - 2)
    double A[5][2][2];
#pragma unroll
    for (int d = 0; d < 5; ++d) A[d][0][0] = A[d][0][1] = A[d][1][0] = A[d][1][1] = 0.0;
    int kp = kstart, slot = 0;
    double* qdst = q + (long long)(ka - g.k0 + HALO) * g.plane + (long long)g.pitch * j0 + i0;
#define LSB_P1M_STEP(PH)                                                                                               \
  {                                                                                                                    \
    if (kp >= kend) break;                                                                                             \
    cp_async_wait<ST - 4>();                                                                                           \
    if (NB == 0) __syncthreads(); /* every marcher is done with the plane of step - 3: its slot is staged again */     \
    else p1m_named_sync(NT);                                                                                           \
    {                                                                                                                  \
      const int fs = slot >= 3 ? slot - 3 : slot + ST - 3;                                                             \
      if (issued < kend) stage_next(fs);                                                                               \
      cp_async_commit();                                                                                               \
    }                                                                                                                  \
    {                                                                                                                  \
      const double* Sc = S + slot * C::PLANE + my_off;                                                                 \
      double w[6][6];                                                                                                  \
      _Pragma("unroll") for (int yy = 0; yy < 6; ++yy) {                                                               \
        _Pragma("unroll") for (int xx = 0; xx < 6; xx += 2) {                                                          \
          const double2 t2 = *reinterpret_cast<const double2*>(Sc + yy * SX + xx);                                     \
          w[yy][xx] = t2.x;                                                                                            \
          w[yy][xx + 1] = t2.y;                                                                                        \
        }                                                                                                              \
      }                                                                                                                \
      if (kp & 1) {                                                                                                    \
        LSB_P1M_PLANE_1(LSB_P1M_ACC_##PH, LSB_P1M_W, LSB_P1M_C)                                                        \
      } else {                                                                                                         \
        LSB_P1M_PLANE_0(LSB_P1M_ACC_##PH, LSB_P1M_W, LSB_P1M_C)                                                        \
      }                                                                                                                \
    }                                                                                                                  \
    if (kp - HALO >= ka) { /* output plane kp - 2 is complete */                                                       \
      const int s2 = slot >= 2 ? slot - 2 : slot + ST - 2;                                                             \
      const double* Sp = S + s2 * C::PLANE + my_off + 2 * SX + 2;                                                      \
      _Pragma("unroll") for (int ry = 0; ry < 2; ++ry) {                                                               \
        if (ry == 0 ? yok0 : yok1) {                                                                                   \
          const double2 pv = *reinterpret_cast<const double2*>(Sp + ry * SX);                                          \
          const double v0 = fma(w2, A[(PH) % 5][ry][0], ridge2 * pv.x), v1 = fma(w2, A[(PH) % 5][ry][1], ridge2 * pv.y); \
          if (xok) {                                                                                                   \
            *reinterpret_cast<double2*>(qdst + (long long)ry * g.pitch) = make_double2(v0, v1);                        \
            dot = fma(pv.x, v0, dot);                                                                                  \
            dot = fma(pv.y, v1, dot);                                                                                  \
          } else if (xok0) {                                                                                           \
            qdst[(long long)ry * g.pitch] = v0;                                                                        \
            dot = fma(pv.x, v0, dot);                                                                                  \
          }                                                                                                            \
        }                                                                                                              \
      }                                                                                                                \
      qdst += g.plane;                                                                                                 \
    }                                                                                                                  \
    A[(PH) % 5][0][0] = A[(PH) % 5][0][1] = A[(PH) % 5][1][0] = A[(PH) % 5][1][1] = 0.0;                               \
    slot = (slot + 1 == ST) ? 0 : slot + 1;                                                                            \
    ++kp;                                                                                                              \
  }
#define LSB_P1M_W(yy, xx) w[yy][xx]
#define LSB_P1M_C(cls, e) T.coef[cls][e]
#define LSB_P1M_ACC_0(dz, ry, rx) A[(0 + 2 - (dz)) % 5][ry][rx]
#define LSB_P1M_ACC_1(dz, ry, rx) A[(1 + 2 - (dz)) % 5][ry][rx]
#define LSB_P1M_ACC_2(dz, ry, rx) A[(2 + 2 - (dz)) % 5][ry][rx]
#define LSB_P1M_ACC_3(dz, ry, rx) A[(3 + 2 - (dz)) % 5][ry][rx]
#define LSB_P1M_ACC_4(dz, ry, rx) A[(4 + 2 - (dz)) % 5][ry][rx]
    if (ka < kb) {
      for (;;) {
        LSB_P1M_STEP(0)
        LSB_P1M_STEP(1)
        LSB_P1M_STEP(2)
        LSB_P1M_STEP(3)
        LSB_P1M_STEP(4)
      }
    }
#undef LSB_P1M_STEP
#undef LSB_P1M_W
#undef LSB_P1M_C
#undef LSB_P1M_ACC_0
#undef LSB_P1M_ACC_1
#undef LSB_P1M_ACC_2
#undef LSB_P1M_ACC_3
#undef LSB_P1M_ACC_4
    cp_async_wait<0>();
  }
  return dot;
}

template <int NW, int NB, int MINB>
__global__ void __launch_bounds__(32 * (NW + NB), MINB)
    k_p1_march(const GridDesc g, const P1Tables tb, const LSB_GRID_CONSTANT P1MarchTaps T, const P1MarchPlan pl, const double w2,
               const double ridge2, const double* __restrict__ p, double* __restrict__ q, double* __restrict__ partials,
               unsigned int* __restrict__ counter, CgScalars* __restrict__ sc, const int cg_mode) {
  constexpr int NT = 32 * (NW + NB);
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const int tid = threadIdx.x;
  const unsigned long long t_begin = tid == 0 ? trace_begin() : 0ull;
  const double dot = p1_march_cta<NW, NB>(g, tb, T, pl, w2, ridge2, p, q, (int)blockIdx.x, (int)gridDim.x);
  const double bsum = block_sum<NT>(dot, red);
  if (tid == 0) trace_end(t_begin, blockIdx.x);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[0] = total;
}

// The boundary shell of a split plan: one thread per node at full occupancy (no shared memory, ~40 registers), launched on
// a side stream beside k_p1_march, whose CTAs leave a third of an SM's registers free; its share of p.q goes to pq_part[2].
template <int NT>
__global__ void __launch_bounds__(NT)
    k_p1_boundary(const GridDesc g, const P1Tables tb, const P1TilePlan b, const double w2, const double ridge2,
                  const double* __restrict__ p, double* __restrict__ q, double* __restrict__ partials,
                  unsigned int* __restrict__ counter, CgScalars* __restrict__ sc, const int cg_mode) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const long long nb = b.nbz + b.nby + b.nbx;
  double dot = 0.0;
  for (long long t = blockIdx.x * (long long)NT + threadIdx.x; t < nb; t += (long long)gridDim.x * NT) {
    int i, j, kl;
    p1_boundary_node(g, b, t, &i, &j, &kl);
    dot += p1_node_generic(g, tb, w2, ridge2, p, q, i, j, kl);
  }
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[2] = total;
}

}  // namespace lsb

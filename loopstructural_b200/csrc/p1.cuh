// P1 / TetMesh twin of the path (K4, K5, K8): P1Interpolator on the structured tetrahedral mesh.
//
// Mesh (LC/supports/_3d_structured_tetra.py): every cell of the lattice is cut into 5 tets, with the mask
// depending on the cell parity (i+j+k)%2 (:28-34); tet id = local + 5 (cx + ncx cy + ncx ncy cz) (:401-419).
//
// K5, regularisation (P1Interpolator.minimise_edge_jumps, LI/_p1interpolator.py:115-166): one row
// n . (grad phi_T1 - grad phi_T2) = 0 per interior shared face, row-normalised over its 8 listed entries,
// weight cgw.  The reference stores ~10 rows per cell (15.9 M rows at 8 M tets).  Here the normal operator
// R^T R is applied matrix-free as a variable-coefficient node stencil: on this lattice the row of R^T R of a
// node depends only on (a) the node parity and (b) how many cells (0, 1, >= 2) lie on either side of it on
// each axis, and it reaches +-2 nodes.  The host assembles R^T R explicitly ONCE on a tiny mesh (<= 6 cells per
// axis, same step vector, same cell-count parity) that contains every such class (p1_tables.py) and uploads
// its rows as a table; k_p1_normal maps each node to its representative (p1_rep) and gathers its <= 36 taps.
// 16 B/node of algorithmic traffic (read p, write q) instead of 96 B/row + 16 B/node.
#pragma once
#include <vector>

#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "fdi_stencil.cuh"
#include "rows.cuh"

namespace lsb {

// The dense table rows (125 taps, at most 33 non-zero) are compacted on the host into fixed-width lists of
// (coefficient, storage offset); `diag` keeps the centre tap for the column scaling.
constexpr int P1_TAPS = 36;
struct P1Tables {
  int m[3];             // cells per axis of the small mesh the table was assembled on
  const double* coef;   // [rows][P1_TAPS], zero padded
  const int* off;       // [rows][P1_TAPS], element offset of the tap in the node vector (0 for padding)
  const double* diag;   // [rows] centre coefficient
};

// host: compact the dense table (rows of 125 taps, tap (dz+2)*25 + (dy+2)*5 + (dx+2)) into fixed-width
// (coefficient, storage offset) lists for a handle with the layout g.  Returns 0, 1 = NaN in the table, 2 = a row has
// more than P1_TAPS non-zero taps (*bad_row).  Shared by lsb200_set_p1_tables and tests/hostsim.
inline int p1_compact_tables(const GridDesc& g, const long long m[3], const double* table, std::vector<double>& coef,
                             std::vector<int>& off, std::vector<double>& diag, size_t* bad_row) {
  const size_t rows = (size_t)(m[0] + 1) * (m[1] + 1) * (m[2] + 1);
  coef.assign(rows * P1_TAPS, 0.0);
  diag.assign(rows, 0.0);
  off.assign(rows * P1_TAPS, 0);
  for (size_t r = 0; r < rows; ++r) {
    int cnt = 0;
    for (int t = 0; t < 125; ++t) {
      const double c = table[r * 125 + t];
      if (!(c == c)) return 1;
      if (t == 62) diag[r] = c;
      if (c == 0.0) continue;
      if (cnt == P1_TAPS) { *bad_row = r; return 2; }
      const int dz = t / 25 - 2, dy = (t / 5) % 5 - 2, dx = t % 5 - 2;
      coef[r * P1_TAPS + cnt] = c;
      off[r * P1_TAPS + cnt] = (int)(dz * g.plane + dy * g.pitch + dx);
      ++cnt;
    }
  }
  return 0;
}

// representative node of the small mesh for node i of an axis with nc cells (m cells in the small mesh):
// same count (0, 1, >= 2) of cells on either side, same parity
__device__ __forceinline__ int p1_rep(int i, int nc, int m) {
  if (nc == m) return i;
  if (i < 2) return i;
  if (i > nc - 2) return m - (nc - i);
  return 2 + ((i - 2) & 1);
}
__device__ __forceinline__ int p1_row(const GridDesc& g, const P1Tables& tb, int i, int j, int k) {
  const int rx = p1_rep(i, g.nx - 1, tb.m[0]), ry = p1_rep(j, g.ny - 1, tb.m[1]), rz = p1_rep(k, g.nz - 1, tb.m[2]);
  return (rz * (tb.m[1] + 1) + ry) * (tb.m[0] + 1) + rx;
}

// a node whose class is one of the eight interior ones on every axis (p1_rep: 2 + parity): 2 <= i <= nc - 2, on a
// mesh with more cells than the small mesh of the table (nc > m)
__device__ __forceinline__ bool p1_interior_axis(int i, int nc, int m) { return nc != m && i >= 2 && i <= nc - 2; }

// The generic form of a node: (coefficient, offset) lists and p gathered from global memory.  One thread per node in
// k_p1_normal (meshes too small to have an interior: an axis of <= 6 cells) and in the boundary role of k_p1_tiles.
__device__ __forceinline__ double p1_node_generic(const GridDesc& g, const P1Tables& tb, const double w2, const double ridge2,
                                                  const double* __restrict__ p, double* __restrict__ q, const int i, const int j,
                                                  const int kl) {
  const int row = p1_row(g, tb, i, j, g.k0 + kl);  // the class of a node depends on its GLOBAL plane
  const double* __restrict__ C = tb.coef + (long long)row * P1_TAPS;
  const int* __restrict__ O = tb.off + (long long)row * P1_TAPS;
  const double* pb = p + ((long long)(kl + HALO) * g.plane + (long long)j * g.pitch + i);
  // a non-zero coefficient implies that the neighbour exists (same cell pattern as the representative);
  // padding entries have coefficient 0 and offset 0
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int e = 0; e < P1_TAPS; e += 4) {
    a0 = fma(__ldg(C + e), pb[__ldg(O + e)], a0);
    a1 = fma(__ldg(C + e + 1), pb[__ldg(O + e + 1)], a1);
    a2 = fma(__ldg(C + e + 2), pb[__ldg(O + e + 2)], a2);
    a3 = fma(__ldg(C + e + 3), pb[__ldg(O + e + 3)], a3);
  }
  const double p0 = pb[0];
  const double v = fma(w2, (a0 + a1) + (a2 + a3), ridge2 * p0);
  q[pb - p] = v;
  return p0 * v;
}

template <int NT>
__global__ void __launch_bounds__(NT)
    k_p1_normal(const GridDesc g, const P1Tables tb, const double w2, const double ridge2, const double* __restrict__ p,
                double* __restrict__ q, double* __restrict__ partials, unsigned int* __restrict__ counter,
                CgScalars* __restrict__ sc, const int cg_mode) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const long long n = (long long)g.nx * g.ny * g.nzl;  // the owned planes [k0,k1); the taps reach into the halo planes
  double dot = 0.0;
  for (long long t = blockIdx.x * (long long)NT + threadIdx.x; t < n; t += (long long)gridDim.x * NT) {
    const int i = (int)(t % g.nx);
    const long long r = t / g.nx;
    dot += p1_node_generic(g, tb, w2, ridge2, p, q, i, (int)(r % g.ny), (int)(r / g.ny));
  }
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[0] = total;
}

// ---- k_p1_tiles: the same operator from shared-memory tiles, one launch -----------------------------------------------
// k_p1_normal spends 108 loads per node (36 coefficients, 36 offsets, 36 values of p), two thirds of them table reads
// that depend only on the node's CLASS; it is gather-latency bound (66 us for 1.64 M nodes, 0.06 of the HBM roofline,
// profiles/r01p).  On the interior (2 <= i <= n-3 on every axis) there are eight classes, the parities of (i, j, k):
//   * a CTA stages a tile of p (64 x TY x TZ nodes + 2 halo nodes per side) in shared memory, DE-INTERLEAVED along x
//     (even columns first, then odd), so that 32 nodes of equal x parity are 32 consecutive doubles;
//   * a warp computes a row of 32 same-class nodes at a time.  The class is a compile-time constant of the unrolled
//     code, so coefficients and shared-memory offsets are constant-bank operands (P1ClassTaps, a kernel parameter):
//     a tap is one LDS + one DFMA, no table load and no per-thread register array;
//   * nodes whose parity sum is even belong to corner tets only: 15 taps against 33 (checked on the host when the plan
//     is made, p1_plan_tiles), their loops are 16 long.
// The boundary nodes (a shell two nodes thick, 10 % of config 3) are enumerated compactly and computed in the generic form,
// an equal share by every CTA between issuing the copies of its tile and waiting for them: the gathers of the one hide
// the latency of the other.  (CTAs of their own for the boundary held every CTA slot for the first 9 us of a 38 us launch:
// they need no shared memory but are launched with the tile's, profiles/r2j.)
constexpr int P1_TAPS_EVEN = 16;
struct P1ClassTaps {
  double coef[8][P1_TAPS];  // class = pi + 2 pj + 4 pk; zero padded
  int soff[8][P1_TAPS];     // BYTE offset of the tap in the de-interleaved tile, from (row base + lane)
  int centre[2];            // byte offset of the node itself, by pi
};
struct P1TilePlan {
  int tiles_x, tiles_y, tiles_z;
  int z_base;          // global plane of the tile lattice's first layer (even)
  int kfirst, klast;   // owned interior planes [kfirst, klast)
  int n_bnd_cta;       // CTAs beyond the tiles that only take a share of the boundary (slabs without interior planes)
  long long nbz, nby, nbx;  // boundary nodes: whole planes / the 4 boundary rows of interior planes / the 4 boundary columns
};
// X(ID, TY, TZ, warps per CTA): tiles of 64 x TY x TZ nodes
#define LSB_P1_TILE_CFGS(X) \
  X(0, 8, 8, 8)             \
  X(1, 8, 4, 8)             \
  X(2, 4, 4, 4)
template <int TY, int TZ>
struct P1TileGeom {
  static constexpr int TX = 64, HX = TX / 2 + HALO, SY = TY + 2 * HALO, SZ = TZ + 2 * HALO;
  static constexpr int ELEMS = 2 * HX * SY * SZ;
};

// host: tile plan and class tables of a handle for tiles of 64 x ty x tz nodes, or false when the mesh has no interior
// (an axis that is its own small mesh) or a class needs more taps than its loop holds -- k_p1_normal does everything then
inline bool p1_plan_tiles(const GridDesc& g, const int m[3], const double* table, const int ty, const int tz, const int nt,
                          const long long max_blocks, P1TilePlan* out, P1ClassTaps* T) {
  const int nc[3] = {g.nx - 1, g.ny - 1, g.nz - 1};
  for (int a = 0; a < 3; ++a)
    if (nc[a] == m[a] || nc[a] < 6) return false;
  P1TilePlan pl;
  pl.kfirst = std::min(std::max(g.k0, 2), g.k1);
  pl.klast = std::max(pl.kfirst, std::min(g.k1, g.nz - 2));
  pl.z_base = pl.kfirst & ~1;
  pl.tiles_x = (g.nx - 4 + 63) / 64;
  pl.tiles_y = (g.ny - 4 + ty - 1) / ty;
  pl.tiles_z = pl.klast > pl.kfirst ? (pl.klast - pl.z_base + tz - 1) / tz : 0;
  const long long nzi = pl.klast - pl.kfirst;
  pl.nbz = (long long)(g.nzl - nzi) * g.nx * g.ny;
  pl.nby = nzi * 4 * g.nx;
  pl.nbx = nzi * 4 * (g.ny - 4);
  const long long nb = pl.nbz + pl.nby + pl.nbx;
  const long long n_tiles = (long long)pl.tiles_x * pl.tiles_y * pl.tiles_z;
  pl.n_bnd_cta = n_tiles * 2 * nt >= nb ? 0 : (int)((nb - n_tiles * 2 * nt + 2 * nt - 1) / (2 * nt));  // <= two nodes per thread
  if ((long long)pl.tiles_x * pl.tiles_y * pl.tiles_z + pl.n_bnd_cta > max_blocks) return false;
  const int HX = 32 + HALO, SY = ty + 2 * HALO;
  memset(T, 0, sizeof(*T));
  for (int cls = 0; cls < 8; ++cls) {
    const int pi = cls & 1, pj = (cls >> 1) & 1, pk = cls >> 2;
    const size_t row = ((size_t)(2 + pk) * (m[1] + 1) + (2 + pj)) * (m[0] + 1) + 2 + pi;
    const int budget = ((pi + pj + pk) & 1) ? P1_TAPS : P1_TAPS_EVEN;
    int cnt = 0;
    for (int t = 0; t < 125; ++t) {
      const double c = table[row * 125 + t];
      if (!(c == c)) return false;
      if (c == 0.0) continue;
      if (cnt == budget) return false;
      const int dz = t / 25 - 2, dy = (t / 5) % 5 - 2, dx = t % 5 - 2;
      T->coef[cls][cnt] = c;
      T->soff[cls][cnt] = 8 * ((((dz + 2) * SY + (dy + 2)) * 2 + ((pi + dx) & 1)) * HX + ((pi + 2 + dx) >> 1));
      ++cnt;
    }
  }
  for (int pi = 0; pi < 2; ++pi) T->centre[pi] = 8 * (((2 * SY + 2) * 2 + pi) * HX + 1);
  *out = pl;
  return true;
}

// boundary node t of the plan's enumeration -> (i, j, local plane)
__device__ __forceinline__ void p1_boundary_node(const GridDesc& g, const P1TilePlan& pl, long long t, int* i, int* j, int* kl) {
  if (t < pl.nbz) {
    const long long per = (long long)g.nx * g.ny;
    const int lp = (int)(t / per);
    const int r = (int)(t - lp * per);
    const int nlo = pl.kfirst - g.k0;  // boundary planes below the interior ones
    *kl = lp < nlo ? lp : (pl.klast - g.k0) + (lp - nlo);
    *j = r / g.nx;
    *i = r - *j * g.nx;
    return;
  }
  t -= pl.nbz;
  if (t < pl.nby) {
    const int per = 4 * g.nx;
    const int kz = (int)(t / per);
    const int r = (int)(t - (long long)kz * per);
    const int jr = r / g.nx;
    *kl = pl.kfirst - g.k0 + kz;
    *j = jr < 2 ? jr : g.ny - 4 + jr;
    *i = r - jr * g.nx;
    return;
  }
  t -= pl.nby;
  const int per = 4 * (g.ny - 4);
  const int kz = (int)(t / per);
  const int r = (int)(t - (long long)kz * per);
  const int ic = r & 3;
  *kl = pl.kfirst - g.k0 + kz;
  *j = 2 + (r >> 2);
  *i = ic < 2 ? ic : g.nx - 4 + ic;
}

// one row of 32 same-class nodes: NTAP taps, every coefficient and offset a constant-bank operand
__device__ __forceinline__ double p1_at(const double* __restrict__ Sb, const int byte_off) {
  return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(Sb) + byte_off);
}
template <int CLS, int NTAP>
__device__ __forceinline__ double p1_row_taps(const P1ClassTaps& T, const double* __restrict__ Sb) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int e = 0; e < NTAP; e += 4) {
    a0 = fma(T.coef[CLS][e], p1_at(Sb, T.soff[CLS][e]), a0);
    a1 = fma(T.coef[CLS][e + 1], p1_at(Sb, T.soff[CLS][e + 1]), a1);
    a2 = fma(T.coef[CLS][e + 2], p1_at(Sb, T.soff[CLS][e + 2]), a2);
    a3 = fma(T.coef[CLS][e + 3], p1_at(Sb, T.soff[CLS][e + 3]), a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// The work of CTA `vb` of `nvb` (tiles first, then the boundary-only CTAs): its tile and its share of the boundary nodes;
// returns this thread's share of p.q.  Shared by k_p1_tiles and the persistent CG kernel of small systems (cg_small.cuh).
template <int TY, int TZ, int NW>
__device__ __forceinline__ double p1_tiles_cta(const GridDesc& g, const P1Tables& tb, const P1ClassTaps& T, const P1TilePlan& pl,
                                               const double w2, const double ridge2, const double* __restrict__ p,
                                               double* __restrict__ q, const int vb, const int nvb) {
  using G = P1TileGeom<TY, TZ>;
  constexpr int NT = 32 * NW, HX = G::HX, SY = G::SY;
  static_assert(TY % 2 == 0 && TZ % 2 == 0, "tiles start on even rows / planes: the class of a row is a compile-time constant");
  LSB_DYN_SMEM(S);
  const int tid = threadIdx.x;
  double dot = 0.0;
  const int n_tiles = pl.tiles_x * pl.tiles_y * pl.tiles_z;
  const bool has_tile = vb < n_tiles;
  int x0 = 0, y0 = 0, z0 = 0;
  if (has_tile) {
    const int b = vb;
    const int txy = pl.tiles_x * pl.tiles_y;
    const int bz = b / txy, brem = b - bz * txy;
    const int by = brem / pl.tiles_x, bx = brem - by * pl.tiles_x;
    x0 = 2 + 64 * bx; y0 = 2 + TY * by; z0 = pl.z_base + TZ * bz;  // first output node of the tile
    // stage the nodes (x0-2 .., y0-2 .., z0-2 ..): consecutive threads read consecutive x.  (A warp per staged row and a
    // thread per pair of x-neighbours were both tried to save address arithmetic: 46 and 44 us against 41, profiles/r2ac,
    // r2_final_p1_tiles.jsonl.)
    for (int idx = tid; idx < G::ELEMS; idx += NT) {
      const int sx = idx % (2 * HX), r = idx / (2 * HX);
      const int sy = r % SY, sz = r / SY;
      const int gi = x0 - HALO + sx, gj = y0 - HALO + sy, gk = z0 - HALO + sz;
      const bool ok = gi < g.nx && gj < g.ny && gk >= 0 && gk < g.nz && gk >= g.k0 - HALO && gk < g.k1 + HALO;
      const double* src = p + ((long long)(gk - g.k0 + HALO) * g.plane + (long long)gj * g.pitch + gi);
      cp_async_8(S + ((sz * SY + sy) * 2 + (sx & 1)) * HX + (sx >> 1), ok ? src : p, ok);
    }
    cp_async_commit();
  }
  {
    // ---- this CTA's share of the boundary nodes, while the tile is on its way
    const long long nb = pl.nbz + pl.nby + pl.nbx;
    for (long long t = (long long)vb * NT + tid; t < nb; t += (long long)nvb * NT) {
      int i, j, kl;
      p1_boundary_node(g, pl, t, &i, &j, &kl);
      dot += p1_node_generic(g, tb, w2, ridge2, p, q, i, j, kl);
    }
  }
  if (has_tile) {
    cp_async_wait<0>();
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int ROWS = (TY / 2) * (TZ / 2);  // rows of one class in a tile
    for (int r = warp; r < ROWS; r += NW) {
      const int ya = 2 * (r % (TY / 2)), za = 2 * (r / (TY / 2));
#define LSB_P1_CLASS(CLS)                                                                                           \
  {                                                                                                                 \
    constexpr int pi = (CLS) & 1, pj = ((CLS) >> 1) & 1, pk = (CLS) >> 2;                                           \
    const int y = ya + pj, z = za + pk;                                                                             \
    const int i = x0 + 2 * lane + pi, j = y0 + y, k = z0 + z;                                                       \
    const double* Sb = S + (z * SY + y) * 2 * HX + lane;                                                            \
    const double acc = p1_row_taps<CLS, ((pi + pj + pk) & 1) ? P1_TAPS : P1_TAPS_EVEN>(T, Sb);                      \
    const double p0 = p1_at(Sb, T.centre[pi]);                                                                      \
    const double v = fma(w2, acc, ridge2 * p0);                                                                     \
    if (i <= g.nx - 3 && j <= g.ny - 3 && k >= pl.kfirst && k < pl.klast) {                                         \
      q[(long long)(k - g.k0 + HALO) * g.plane + (long long)j * g.pitch + i] = v;                                   \
      dot = fma(p0, v, dot);                                                                                        \
    }                                                                                                               \
  }
      LSB_P1_CLASS(0) LSB_P1_CLASS(1) LSB_P1_CLASS(2) LSB_P1_CLASS(3)
      LSB_P1_CLASS(4) LSB_P1_CLASS(5) LSB_P1_CLASS(6) LSB_P1_CLASS(7)
#undef LSB_P1_CLASS
    }
  }
  return dot;
}

template <int TY, int TZ, int NW>
__global__ void __launch_bounds__(32 * NW)
    k_p1_tiles(const GridDesc g, const P1Tables tb, const LSB_GRID_CONSTANT P1ClassTaps T, const P1TilePlan pl, const double w2,
               const double ridge2, const double* __restrict__ p, double* __restrict__ q, double* __restrict__ partials,
               unsigned int* __restrict__ counter, CgScalars* __restrict__ sc, const int cg_mode) {
  constexpr int NT = 32 * NW;
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const int tid = threadIdx.x;
  const unsigned long long t_begin = tid == 0 ? trace_begin() : 0ull;
  const double dot = p1_tiles_cta<TY, TZ, NW>(g, tb, T, pl, w2, ridge2, p, q, (int)blockIdx.x, (int)gridDim.x);
  const double bsum = block_sum<NT>(dot, red);
  if (tid == 0) trace_end(t_begin, blockIdx.x);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) sc->pq_part[0] = total;
}

// diag(R^T R) w^2 + ridge^2 (for the column scaling, LI/_discrete_interpolator.py:815-833)
__global__ void k_p1_diag(const GridDesc g, const P1Tables tb, const double w2, const double ridge2,
                          double* __restrict__ diag) {
  const long long n = (long long)g.nx * g.ny * g.nzl;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t % g.nx);
    const long long r = t / g.nx;
    const int j = (int)(r % g.ny), kl = (int)(r / g.ny);
    diag[(long long)(kl + HALO) * g.plane + (long long)j * g.pitch + i] += fma(w2, tb.diag[p1_row(g, tb, i, j, g.k0 + kl)], ridge2);
  }
}

// ---- point location ------------------------------------------------------------------------------------
__constant__ int c_tet_even[5][4] = {{7, 1, 2, 4}, {6, 2, 4, 7}, {5, 1, 4, 7}, {0, 1, 2, 4}, {3, 1, 2, 7}};
__constant__ int c_tet_odd[5][4] = {{0, 6, 5, 3}, {7, 3, 5, 6}, {4, 0, 5, 6}, {2, 0, 3, 6}, {1, 0, 3, 5}};
// corner s of tet t of a cell (LC/supports/_3d_structured_tetra.py:28-34); the host copy serves tests/hostsim only
LSB_HD int tet_corner(const bool even, const int t, const int s) {
#ifdef __CUDA_ARCH__
  return even ? c_tet_even[t][s] : c_tet_odd[t][s];
#else
  constexpr int ev[5][4] = {{7, 1, 2, 4}, {6, 2, 4, 7}, {5, 1, 4, 7}, {0, 1, 2, 4}, {3, 1, 2, 7}};
  constexpr int od[5][4] = {{0, 6, 5, 3}, {7, 3, 5, 6}, {4, 0, 5, 6}, {2, 0, 3, 6}, {1, 0, 3, 5}};
  return even ? ev[t][s] : od[t][s];
#endif
}

struct TetLoc {
  bool inside;
  int cz;          // z index of the cell (global)
  int node[4];     // storage indices of the tet's nodes (relative to the handle's planes)
  double c[4];     // barycentric coordinates (shape functions)
  double v[4][3];  // vertex positions
};

// unfused arithmetic in numpy's order, so that points on tet faces take the same side as in the reference
LSB_HD void cross_rn(const double a[3], const double b[3], double o[3]) {
  o[0] = rn_sub(rn_mul(a[1], b[2]), rn_mul(a[2], b[1]));
  o[1] = rn_sub(rn_mul(a[2], b[0]), rn_mul(a[0], b[2]));
  o[2] = rn_sub(rn_mul(a[0], b[1]), rn_mul(a[1], b[0]));
}
LSB_HD double dot_rn(const double a[3], const double b[3]) {
  return rn_add(rn_add(rn_mul(a[0], b[0]), rn_mul(a[1], b[1])), rn_mul(a[2], b[2]));
}

// TetMesh.get_element_for_location (LC/supports/_3d_structured_tetra.py:274-365): cell by floor + clip, the five
// tets of the cell, barycentrics by scalar triple products, the LAST tet with all c >= 0 wins; strict inside
// (:264-272) on top of the tolerant one of position_to_cell_index (LC/supports/_3d_base_structured.py:255-293).
LSB_HD TetLoc locate_tet(const GridDesc& g, const double x[3]) {
  TetLoc L;
  const int nc[3] = {g.nx - 1, g.ny - 1, g.nz - 1};
  bool ins_tol = true, ins_strict = true;
  int cell[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double maxi = rn_add(g.origin[a], rn_mul((double)nc[a], g.step[a]));
    const double ext = fmax(fabs(maxi - g.origin[a]), 1.0);
    const double tol = rn_mul(rn_mul(2.220446049250313e-16, 32.0), ext);
    ins_tol = ins_tol && (x[a] >= g.origin[a] - tol) && (x[a] <= maxi + tol);
    // TetMesh.inside: origin + step_vector * nsteps_cells
    ins_strict = ins_strict && (x[a] > g.origin[a]) && (x[a] < rn_add(g.origin[a], rn_mul(g.step[a], (double)nc[a])));
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int c = 0;
    if (ins_tol) {
      const double f = floor((x[a] - g.origin[a]) / g.step[a]);
      c = (int)fmin(fmax(f, 0.0), (double)(nc[a] - 1));
    }
    cell[a] = c;
  }
  const bool even = ((cell[0] + cell[1] + cell[2]) & 1) == 0;
  int best = -1;
  L.inside = false;
#pragma unroll 1
  for (int t = 0; t < 5; ++t) {
    double v[4][3];
    int corner[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      corner[s] = tet_corner(even, t, s);
      const int off[3] = {corner[s] & 1, (corner[s] >> 1) & 1, (corner[s] >> 2) & 1};
#pragma unroll
      for (int a = 0; a < 3; ++a) v[s][a] = rn_add(g.origin[a], rn_mul(g.step[a], (double)(cell[a] + off[a])));
    }
    double vap[3], vbp[3], vab[3], vac[3], vad[3], vbc[3], vbd[3], cr[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      vap[a] = x[a] - v[0][a]; vbp[a] = x[a] - v[1][a];
      vab[a] = v[1][a] - v[0][a]; vac[a] = v[2][a] - v[0][a]; vad[a] = v[3][a] - v[0][a];
      vbc[a] = v[2][a] - v[1][a]; vbd[a] = v[3][a] - v[1][a];
    }
    cross_rn(vbd, vbc, cr); const double va = dot_rn(vbp, cr) / 6.0;
    cross_rn(vac, vad, cr); const double vb = dot_rn(vap, cr) / 6.0;
    const double vol = dot_rn(vab, cr) / 6.0;
    cross_rn(vad, vab, cr); const double vc = dot_rn(vap, cr) / 6.0;
    cross_rn(vab, vac, cr); const double vd = dot_rn(vap, cr) / 6.0;
    const double c0 = va / vol, c1 = vb / vol, c2 = vc / vol, c3 = vd / vol;
    if (c0 >= 0.0 && c1 >= 0.0 && c2 >= 0.0 && c3 >= 0.0) {  // later tets overwrite earlier ones
      best = t;
      L.c[0] = c0; L.c[1] = c1; L.c[2] = c2; L.c[3] = c3;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int off[3] = {corner[s] & 1, (corner[s] >> 1) & 1, (corner[s] >> 2) & 1};
        L.node[s] = (int)((long long)(cell[2] + off[2] - g.k0 + HALO) * g.plane + (long long)(cell[1] + off[1]) * g.pitch + cell[0] + off[0]);
#pragma unroll
        for (int a = 0; a < 3; ++a) L.v[s][a] = v[s][a];
      }
    }
  }
  L.inside = ins_tol && ins_strict && best >= 0;
  L.cz = cell[2];
  return L;
}

// shape-function gradients G = E^-1 [[-1,1,0,0],[-1,0,1,0],[-1,0,0,1]], E rows = v1-v0, v2-v0, v3-v0
// (TetMesh.get_element_gradients, LC/supports/_3d_structured_tetra.py:421-486); returns |det E| / 6 (element_size)
LSB_HD double tet_gradients(const double v[4][3], double G[3][4]) {
  double E[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int a = 0; a < 3; ++a) E[r][a] = v[r + 1][a] - v[0][a];
  const double det = E[0][0] * (E[1][1] * E[2][2] - E[1][2] * E[2][1]) - E[0][1] * (E[1][0] * E[2][2] - E[1][2] * E[2][0]) +
                     E[0][2] * (E[1][0] * E[2][1] - E[1][1] * E[2][0]);
  const double id = 1.0 / det;
  double inv[3][3];  // inverse of E
  inv[0][0] = (E[1][1] * E[2][2] - E[1][2] * E[2][1]) * id;
  inv[0][1] = (E[0][2] * E[2][1] - E[0][1] * E[2][2]) * id;
  inv[0][2] = (E[0][1] * E[1][2] - E[0][2] * E[1][1]) * id;
  inv[1][0] = (E[1][2] * E[2][0] - E[1][0] * E[2][2]) * id;
  inv[1][1] = (E[0][0] * E[2][2] - E[0][2] * E[2][0]) * id;
  inv[1][2] = (E[0][2] * E[1][0] - E[0][0] * E[1][2]) * id;
  inv[2][0] = (E[1][0] * E[2][1] - E[1][1] * E[2][0]) * id;
  inv[2][1] = (E[0][1] * E[2][0] - E[0][0] * E[2][1]) * id;
  inv[2][2] = (E[0][0] * E[1][1] - E[0][1] * E[1][0]) * id;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    G[d][0] = -(inv[d][0] + inv[d][1] + inv[d][2]);
    G[d][1] = inv[d][0];
    G[d][2] = inv[d][1];
    G[d][3] = inv[d][2];
  }
  return fabs(det) / 6.0;
}

// K4: P1 rows from constraint points (LI/_p1interpolator.py:73-113,248-296).  value: the 4 barycentrics, b = value,
// w = cpw (the per-point weight column is not used); normal: three rows grad phi[d,:] with b = n_d, w = npw;
// tangent / orthogonal: v_hat . grad phi = b, w = tpw * tet volume; element vector (fold rows): v . grad phi = b, w as given.
// Rows are L2-normalised; mult rows per point at
// row0 + 3*point + d for normals (the reference flattens the (n,3,4) block point-major), else row0 + point.
__global__ void k_p1_point_rows(const GridDesc g, const int kind, const long long n, const double* __restrict__ data,
                                const int ncols, const double weight, RowBlock rb, const long long row0,
                                unsigned long long* __restrict__ status) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double* d = data + r * ncols;
    const double x[3] = {d[0], d[1], d[2]};
    const TetLoc L = locate_tet(g, x);
    bool ok = L.inside;
    const int used = kind == LSB200_ROWS_VALUE ? 4 : (kind == LSB200_ROWS_TANGENT || kind == LSB200_ROWS_ELEMENT_VECTOR ? 7 : 6);
    for (int c = 0; c < used; ++c) ok = ok && (d[c] == d[c]);
    // slab: keep the row when its cell touches an owned plane (as k_fdi_point_rows does)
    ok = ok && (L.cz + 1 >= g.k0) && (L.cz < g.k1);
    const int mult = kind == LSB200_ROWS_NORMAL ? 3 : 1;
    double A[3][4], bb[3] = {0.0, 0.0, 0.0};
    double ww = weight;
    if (ok) {
      if (kind == LSB200_ROWS_VALUE) {
#pragma unroll
        for (int c = 0; c < 4; ++c) A[0][c] = L.c[c];
        bb[0] = d[3];
      } else {
        double G[3][4];
        const double vol = tet_gradients(L.v, G);
        if (kind == LSB200_ROWS_NORMAL) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c < 4; ++c) A[a][c] = G[a][c];
            bb[a] = d[3 + a];
          }
        } else if (kind == LSB200_ROWS_ELEMENT_VECTOR) {
          // fold rows (LI/_discrete_fold_interpolator.py:116-190): the volume factor of A and B cancels in the row
          // normalisation, the weight is the caller's
#pragma unroll
          for (int c = 0; c < 4; ++c) A[0][c] = d[3] * G[0][c] + d[4] * G[1][c] + d[5] * G[2][c];
          bb[0] = d[6];
        } else {
          const double len = sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
          double vv[3] = {d[3], d[4], d[5]};
          if (len > 0.0) { vv[0] /= len; vv[1] /= len; vv[2] /= len; }
#pragma unroll
          for (int c = 0; c < 4; ++c) A[0][c] = vv[0] * G[0][c] + vv[1] * G[1][c] + vv[2] * G[2][c];
          bb[0] = d[6];
          ww = weight * vol;
        }
      }
      atomicAdd(status, (unsigned long long)mult);
    }
    for (int m = 0; m < mult; ++m) {
      const long long ro = row0 + (long long)mult * r + m;
      double nrm2 = 0.0;
      if (ok)
#pragma unroll
        for (int c = 0; c < 4; ++c) nrm2 += A[m][c] * A[m][c];
      const double len = sqrt(nrm2);
      const bool good = ok && (len == len);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        rb.vals[c * rb.n + ro] = good ? (len > 0.0 ? A[m][c] / len : A[m][c]) : 0.0;
        rb.cols[c * rb.n + ro] = good ? L.node[c] : -1;
      }
      rb.b[ro] = good ? (len > 0.0 ? bb[m] / len : bb[m]) : 0.0;
      rb.w2[ro] = good ? ww * ww : 0.0;
    }
  }
}

// Direction-weighted face-jump rows, P1Interpolator.minimise_edge_jumps with a vector per shared face
// (LI/_p1interpolator.py:115-166; the 'fold regularisation n' families of DiscreteFoldInterpolator, and every
// DirectionalRegularisation term of a P1 interpolator): v . (grad phi_a - grad phi_b) = 0 over the 4 + 4 listed nodes of the
// two tetrahedra (global node ids; the three shared nodes appear twice, as in the reference before its coo -> csc merge).
// Writes the (n, 8) values and column ids, b = 0 and w = weight: what k_rows_ingest then normalises and stores.
__global__ void k_p1_face_vector_vals(const GridDesc g, const long long n, const long long* __restrict__ nodes_a,
                                      const long long* __restrict__ nodes_b, const double* __restrict__ vec,
                                      const double weight, double* __restrict__ vals, long long* __restrict__ cols,
                                      double* __restrict__ b, double* __restrict__ w) {
  const long long nxny = (long long)g.nx * g.ny, ntot = nxny * g.nz;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double v[3] = {vec[3 * r], vec[3 * r + 1], vec[3 * r + 2]};
    b[r] = 0.0;
    w[r] = weight;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const long long* nd = (side ? nodes_b : nodes_a) + 4 * r;
      double X[4][3], G[3][4];
      bool ok = true;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const long long id = nd[c];
        ok = ok && id >= 0 && id < ntot;
        const long long kz = id / nxny, rem = id - kz * nxny;
        const long long jy = rem / g.nx, ix = rem - jy * g.nx;
        X[c][0] = g.origin[0] + g.step[0] * (double)ix;
        X[c][1] = g.origin[1] + g.step[1] * (double)jy;
        X[c][2] = g.origin[2] + g.step[2] * (double)kz;
        cols[8 * r + 4 * side + c] = id;  // an id outside the mesh is reported by k_rows_ingest
      }
      if (ok) tet_gradients(X, G);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double a = ok ? v[0] * G[0][c] + v[1] * G[1][c] + v[2] * G[2][c] : 0.0;
        vals[8 * r + 4 * side + c] = side ? -a : a;
      }
    }
  }
}

// K8: TetMesh.evaluate_value (LC/supports/_3d_structured_tetra.py:208-229) behind _evaluate_value_local
// (LI/_discrete_interpolator.py:1342-1362): NaN outside, 0.0 for rows holding a NaN coordinate.
__global__ void k_eval_value_tet(const GridDesc g, const double* __restrict__ c, const long long n,
                                 const double* __restrict__ xyz, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    double v;
    if (x[0] != x[0] || x[1] != x[1] || x[2] != x[2]) {
      v = 0.0;
    } else {
      const TetLoc L = locate_tet(g, x);
      if (!L.inside) v = nan("");
      else v = L.c[0] * c[L.node[0]] + L.c[1] * c[L.node[1]] + L.c[2] * c[L.node[2]] + L.c[3] * c[L.node[3]];
    }
    out[r] = v;
  }
}

// TetMesh.evaluate_gradient (:231-262): NaN outside
__global__ void k_eval_gradient_tet(const GridDesc g, const double* __restrict__ c, const long long n,
                                    const double* __restrict__ xyz, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    const TetLoc L = locate_tet(g, x);
    double gr[3] = {nan(""), nan(""), nan("")};
    if (L.inside) {
      double G[3][4];
      tet_gradients(L.v, G);
#pragma unroll
      for (int a = 0; a < 3; ++a)
        gr[a] = G[a][0] * c[L.node[0]] + G[a][1] * c[L.node[1]] + G[a][2] * c[L.node[2]] + G[a][3] * c[L.node[3]];
    }
    out[3 * r] = gr[0];
    out[3 * r + 1] = gr[1];
    out[3 * r + 2] = gr[2];
  }
}

}  // namespace lsb

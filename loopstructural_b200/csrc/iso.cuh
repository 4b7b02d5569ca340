// Isosurface of a scalar volume on a regular lattice (SURVEY.md section 8 f, rank 4): what BaseFeature.surfaces /
// LoopIsosurfacer.fit (LS/modelling/features/_base_geological_feature.py:317-357, LS/utils/_surface.py:66-186) get from
// skimage.measure.marching_cubes(volume, level, spacing=..., mask=...) -- vertices, triangles, vertex normals.
//
// scikit-image is a third-party dependency of the reference that is absent from this image (its Lewiner tables cannot be
// consulted here), so this is NOT a restatement of its triangulation: it is marching TETRAHEDRA on the same lattice.
// Every cell is cut into the six Kuhn tetrahedra around its main diagonal (one per order of the three axes: corner 0 ->
// + e_a -> + e_b -> corner 7), which matches across cell faces, so the surface is watertight wherever the volume is
// defined; a vertex sits on a lattice edge (axis edge, face diagonal or body diagonal) whose end values straddle the
// level, by linear interpolation, exactly one vertex per such edge (shared by every triangle around it).  The surface is
// the level set of the same piecewise-linear interpolant of the same samples; it has more triangles than Lewiner's.
// PARITY UNPINNED against scikit-image; pinned against oracle/iso_oracle.py (the same construction in numpy) and
// analytic properties (tests/test_iso_*.py).
//
// Volume: C order, v[(i0 * n1 + i1) * n2 + i2]; NaN samples (the reference's mask = ~isnan) disable every tetrahedron that
// touches them.  Vertex id = rank of its edge among the crossed edges in edge-index order (edge = 7 * node + type); triangles
// are emitted cell by cell, tetrahedron by tetrahedron: the output is a deterministic function of the volume.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace lsb {

struct IsoGrid {
  long long n0, n1, n2;  // samples per axis
  double spacing[3];
  double level;
};

// the seven edge directions that start at a node: bit a of (type + 1) = step along axis a (axis 0 = bit 0)
LSB_HD void iso_edge_dir(const int type, int d[3]) {
  const int m = type + 1;
  d[0] = m & 1;
  d[1] = (m >> 1) & 1;
  d[2] = (m >> 2) & 1;
}
LSB_HD long long iso_node(const IsoGrid& g, const long long i0, const long long i1, const long long i2) {
  return (i0 * g.n1 + i1) * g.n2 + i2;
}
// a sample counts as inside when it is >= level
LSB_HD bool iso_in(const double v, const double level) { return v >= level; }

// the six tetrahedra of a cell: corners as offsets from the cell's lowest node, along the axis order perm
LSB_HD void iso_tet_corners(const int t, int c[4][3]) {
  constexpr int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  int cur[3] = {0, 0, 0};
  for (int a = 0; a < 3; ++a) c[0][a] = 0;
  for (int s = 0; s < 3; ++s) {
    cur[perm[t][s]] = 1;
    for (int a = 0; a < 3; ++a) c[s + 1][a] = cur[a];
  }
}

// edge index of the lattice edge between two corners of a cell (offsets oa < ob componentwise)
LSB_HD long long iso_edge_of(const IsoGrid& g, const long long i0, const long long i1, const long long i2, const int oa[3],
                             const int ob[3]) {
  const int type = ((ob[0] - oa[0]) | ((ob[1] - oa[1]) << 1) | ((ob[2] - oa[2]) << 2)) - 1;
  return 7 * iso_node(g, i0 + oa[0], i1 + oa[1], i2 + oa[2]) + type;
}

#ifdef __CUDACC__
// flag[e] = 1 when edge e exists, both end samples are finite and they straddle the level
__global__ void k_iso_edge_flags(const IsoGrid g, const double* __restrict__ v, int* __restrict__ flag) {
  const long long n = g.n0 * g.n1 * g.n2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < 7 * n; e += (long long)gridDim.x * blockDim.x) {
    const long long node = e / 7;
    const int type = (int)(e - 7 * node);
    int d[3];
    iso_edge_dir(type, d);
    const long long i2 = node % g.n2, i1 = (node / g.n2) % g.n1, i0 = node / (g.n2 * g.n1);
    int f = 0;
    if (i0 + d[0] < g.n0 && i1 + d[1] < g.n1 && i2 + d[2] < g.n2) {
      const double a = v[node], b = v[iso_node(g, i0 + d[0], i1 + d[1], i2 + d[2])];
      f = (a == a && b == b && iso_in(a, g.level) != iso_in(b, g.level)) ? 1 : 0;
    }
    flag[e] = f;
  }
}

// vertex and normal of every crossed edge; id = exclusive scan of the flags
__global__ void k_iso_vertices(const IsoGrid g, const double* __restrict__ v, const int* __restrict__ flag,
                               const int* __restrict__ id, double* __restrict__ verts, double* __restrict__ normals) {
  const long long n = g.n0 * g.n1 * g.n2;
  auto sample = [&](long long i0, long long i1, long long i2) { return v[iso_node(g, i0, i1, i2)]; };
  // gradient of the volume at a node by central differences (one-sided at the border or next to a NaN), per unit length
  auto grad = [&](const long long i[3], double gr[3]) {
    const long long nn[3] = {g.n0, g.n1, g.n2};
    for (int a = 0; a < 3; ++a) {
      long long lo[3] = {i[0], i[1], i[2]}, hi[3] = {i[0], i[1], i[2]};
      if (i[a] > 0) lo[a] -= 1;
      if (i[a] + 1 < nn[a]) hi[a] += 1;
      double vl = sample(lo[0], lo[1], lo[2]), vh = sample(hi[0], hi[1], hi[2]);
      const double vc = sample(i[0], i[1], i[2]);
      if (!(vl == vl)) { vl = vc; lo[a] = i[a]; }
      if (!(vh == vh)) { vh = vc; hi[a] = i[a]; }
      const double dist = (double)(hi[a] - lo[a]) * g.spacing[a];
      gr[a] = dist > 0.0 ? (vh - vl) / dist : 0.0;
    }
  };
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < 7 * n; e += (long long)gridDim.x * blockDim.x) {
    if (!flag[e]) continue;
    const long long node = e / 7;
    const int type = (int)(e - 7 * node);
    int d[3];
    iso_edge_dir(type, d);
    const long long ia[3] = {node / (g.n2 * g.n1), (node / g.n2) % g.n1, node % g.n2};
    const long long ib[3] = {ia[0] + d[0], ia[1] + d[1], ia[2] + d[2]};
    const double a = v[node], b = sample(ib[0], ib[1], ib[2]);
    const double t = (g.level - a) / (b - a);
    double ga[3], gb[3], nr[3];
    grad(ia, ga);
    grad(ib, gb);
    double len2 = 0.0;
    for (int k = 0; k < 3; ++k) {
      verts[3 * (long long)id[e] + k] = ((double)ia[k] + t * (double)d[k]) * g.spacing[k];
      nr[k] = -(ga[k] + t * (gb[k] - ga[k]));  // towards decreasing values (skimage's gradient_direction='descent')
      len2 += nr[k] * nr[k];
    }
    const double len = sqrt(len2);
    for (int k = 0; k < 3; ++k) normals[3 * (long long)id[e] + k] = len > 0.0 ? nr[k] / len : 0.0;
  }
}

// triangles of tetrahedron t of a cell: count (0, 1, 2), and -- when tri != nullptr -- the vertex ids, oriented so that
// the geometric normal points from the inside corners (>= level) to the outside ones
__device__ __forceinline__ int iso_tet_triangles(const IsoGrid& g, const double* __restrict__ v, const int* __restrict__ id,
                                                 const long long i0, const long long i1, const long long i2, const int t,
                                                 int (*tri)[3]) {
  int c[4][3];
  iso_tet_corners(t, c);
  double val[4];
  int in_mask = 0;
  for (int s = 0; s < 4; ++s) {
    val[s] = v[iso_node(g, i0 + c[s][0], i1 + c[s][1], i2 + c[s][2])];
    if (!(val[s] == val[s])) return 0;
    if (iso_in(val[s], g.level)) in_mask |= 1 << s;
  }
  if (in_mask == 0 || in_mask == 15) return 0;
  const int n_in = __popc(in_mask);
  // corner lists: `lone` side first
  int in_c[4], out_c[4], ni = 0, no = 0;
  for (int s = 0; s < 4; ++s) {
    if (in_mask & (1 << s)) in_c[ni++] = s;
    else out_c[no++] = s;
  }
  if (!tri) return n_in == 2 ? 2 : 1;
  auto edge_vertex = [&](const int a, const int b) {  // corners a < b in march order: offsets are componentwise ordered
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return id[iso_edge_of(g, i0, i1, i2, c[lo], c[hi])];
  };
  // direction inside -> outside (difference of the corner centroids), for the orientation
  double dir[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < 3; ++k) {
    double ci = 0.0, co = 0.0;
    for (int s = 0; s < ni; ++s) ci += c[in_c[s]][k];
    for (int s = 0; s < no; ++s) co += c[out_c[s]][k];
    dir[k] = (co / no - ci / ni) * g.spacing[k];
  }
  // positions of the cut points (in cell-local units) for the orientation test
  auto cut = [&](const int a, const int b, double p[3]) {
    const double tt = (g.level - val[a]) / (val[b] - val[a]);
    for (int k = 0; k < 3; ++k) p[k] = (c[a][k] + tt * (c[b][k] - c[a][k])) * g.spacing[k];
  };
  auto orient = [&](int (&q)[3], const int (&ea)[3], const int (&eb)[3]) {
    double p0[3], p1[3], p2[3];
    cut(ea[0], eb[0], p0);
    cut(ea[1], eb[1], p1);
    cut(ea[2], eb[2], p2);
    const double u[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, w[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    const double nx = u[1] * w[2] - u[2] * w[1], ny = u[2] * w[0] - u[0] * w[2], nz = u[0] * w[1] - u[1] * w[0];
    if (nx * dir[0] + ny * dir[1] + nz * dir[2] < 0.0) {
      const int tmp = q[1];
      q[1] = q[2];
      q[2] = tmp;
    }
  };
  if (n_in == 1 || n_in == 3) {
    const int lone = n_in == 1 ? in_c[0] : out_c[0];
    const int* others = n_in == 1 ? out_c : in_c;
    const int ea[3] = {lone, lone, lone}, eb[3] = {others[0], others[1], others[2]};
    tri[0][0] = edge_vertex(lone, others[0]);
    tri[0][1] = edge_vertex(lone, others[1]);
    tri[0][2] = edge_vertex(lone, others[2]);
    orient(tri[0], ea, eb);
    return 1;
  }
  // two in, two out: the quad (i0 o0) (i0 o1) (i1 o1) (i1 o0), cut along (i0 o0)-(i1 o1)
  {
    const int a0 = in_c[0], a1 = in_c[1], b0 = out_c[0], b1 = out_c[1];
    const int ea0[3] = {a0, a0, a1}, eb0[3] = {b0, b1, b1};
    tri[0][0] = edge_vertex(a0, b0);
    tri[0][1] = edge_vertex(a0, b1);
    tri[0][2] = edge_vertex(a1, b1);
    orient(tri[0], ea0, eb0);
    const int ea1[3] = {a0, a1, a1}, eb1[3] = {b0, b1, b0};
    tri[1][0] = edge_vertex(a0, b0);
    tri[1][1] = edge_vertex(a1, b1);
    tri[1][2] = edge_vertex(a1, b0);
    orient(tri[1], ea1, eb1);
  }
  return 2;
}

// count[cell] = triangles of the cell (cells in C order over (n0-1, n1-1, n2-1))
__global__ void k_iso_cell_counts(const IsoGrid g, const double* __restrict__ v, int* __restrict__ count) {
  const long long c1 = g.n1 - 1, c2 = g.n2 - 1, nc = (g.n0 - 1) * c1 * c2;
  for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < nc; cell += (long long)gridDim.x * blockDim.x) {
    const long long i2 = cell % c2, i1 = (cell / c2) % c1, i0 = cell / (c2 * c1);
    int cnt = 0;
    for (int t = 0; t < 6; ++t) cnt += iso_tet_triangles(g, v, nullptr, i0, i1, i2, t, nullptr);
    count[cell] = cnt;
  }
}

__global__ void k_iso_triangles(const IsoGrid g, const double* __restrict__ v, const int* __restrict__ id,
                                const int* __restrict__ first, int* __restrict__ tris) {
  const long long c1 = g.n1 - 1, c2 = g.n2 - 1, nc = (g.n0 - 1) * c1 * c2;
  for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < nc; cell += (long long)gridDim.x * blockDim.x) {
    const long long i2 = cell % c2, i1 = (cell / c2) % c1, i0 = cell / (c2 * c1);
    long long at = first[cell];
    for (int t = 0; t < 6; ++t) {
      int tri[2][3];
      const int k = iso_tet_triangles(g, v, id, i0, i1, i2, t, tri);
      for (int s = 0; s < k; ++s) {
        tris[3 * at] = tri[s][0];
        tris[3 * at + 1] = tri[s][1];
        tris[3 * at + 2] = tri[s][2];
        ++at;
      }
    }
  }
}
#endif

}  // namespace lsb

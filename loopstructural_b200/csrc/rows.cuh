// K2/K3: scattered constraint rows (value, gradient, normal, tangent, interface, fold, orthogonality ...).
//
// K2 applies q += A_d^T W A_d p for ELL blocks of row-normalised rows -- the part of
// A^T A (LI/_solver_strategy.py:64) that comes from DiscreteInterpolator.add_constraints_to_least_squares
// (LI/_discrete_interpolator.py:316-389) families other than the regularisation.  Rows enter the
// least-squares system multiplied by w (build_matrix, :684-755), so the normal operator carries w^2.
// K3 generates the FDI point rows on the device (trilinear weights / trilinear gradient matrix,
// LC/supports/_3d_structured_grid.py:90-146,393-454) instead of numpy temporaries.
#pragma once
#include "common.cuh"

namespace lsb {

// ELL block, column-major: entry j of row r at [j*n + r].
struct RowBlock {
  long long n;
  int k;
  double* vals;
  int* cols;   // LOCAL storage index (halo planes included); -1 = padding
  double* w2;  // w^2
  double* b;
  int ordered;  // host bookkeeping: compacted and sorted by lsb200_finalize_system
};

__device__ __forceinline__ bool owned(const GridDesc& g, int c) {
  return c >= HALO * g.plane && c < (long long)(g.nzl + HALO) * g.plane;
}

// non-returning fp64 add (RED.E.ADD.F64): the issuing thread does not wait for the L2 round trip, unlike atomicAdd,
// whose return value the compiler keeps alive through ATOMG
__device__ __forceinline__ void red_add(double* addr, double v) {
#ifdef LSB_EMU
  *addr += v;
#else
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
#endif
}

#ifndef LSB_EMU  // warp-level kernels: not part of the CPU emulation of the stencil / P1 operators
// q += A_d^T W A_d p over one ELL block whose rows lsb200_finalize_system has compacted (no dead rows) and sorted by
// their first column, so that (a) a warp's gathers and adds fall into neighbouring sectors and (b) the rows one
// constraint point produces (2 gradient rows, 3 normal rows, several points of one cell) sit in adjacent lanes.  One
// thread per row: K independent gathers, t = w^2 (a . p); then lanes whose K columns all equal those of the lane below
// form a segment, the per-column products a_j t are summed over the segment with shuffles, and the segment's first lane
// issues ONE non-returning add per column -- the warp-aggregated scatter.  p . q part: sum_r t_r (a_r . p) over owned
// columns, fixed-order block / grid reduction.
// rows [vb NT, ...) in steps of nvb NT, one row per thread (tid in [0, NT)); returns this thread's share of p . q
template <int NT, int K>
__device__ __forceinline__ double rows_normal_cta(const GridDesc& g, const RowBlock& rb, const double* __restrict__ p,
                                                  double* __restrict__ q, const int tid, const int vb, const int nvb) {
  double dot = 0.0;
  const unsigned lane = (unsigned)tid & 31u;
  const long long stride = (long long)nvb * NT;
  for (long long base = vb * (long long)NT; base < rb.n; base += stride) {  // block-uniform trip count
    const long long r = base + tid;
    const bool active = r < rb.n;
    int c[K];
    double a[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      c[j] = active ? rb.cols[j * rb.n + r] : -1;
      a[j] = active ? rb.vals[j * rb.n + r] : 0.0;
    }
    double t = 0.0, town = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (c[j] >= 0) {
        const double v = a[j] * p[c[j]];
        t += v;
        if (owned(g, c[j])) town += v;
      }
    }
    t *= active ? rb.w2[r] : 0.0;
    dot = fma(town, t, dot);
    // segments of lanes with identical column lists
    bool same = active && lane > 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int below = __shfl_up_sync(0xffffffffu, c[j], 1);
      same = same && (below == c[j]);
    }
    const unsigned heads = __ballot_sync(0xffffffffu, !same);
    const unsigned above = lane == 31u ? 0u : (heads >> (lane + 1u));
    const int len = same ? 0 : (above ? __ffs(above) : (int)(32u - lane));  // head lanes: lanes in my segment
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double v = a[j] * t;
      for (int s = 1; s < maxlen; ++s) {
        const double x = __shfl_down_sync(0xffffffffu, a[j] * t, s);
        if (s < len) v += x;
      }
      if (!same && c[j] >= 0 && v != 0.0 && owned(g, c[j])) red_add(q + c[j], v);
    }
  }
  return dot;
}

template <int NT, int K>
__global__ void __launch_bounds__(NT)
    k_rows_normal(const GridDesc g, const RowBlock rb, const double* __restrict__ p, double* __restrict__ q,
                  double* __restrict__ partials, unsigned int* __restrict__ counter, CgScalars* __restrict__ sc,
                  const int cg_mode, const int first) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const double dot = rows_normal_cta<NT, K>(g, rb, p, q, (int)threadIdx.x, (int)blockIdx.x, (int)gridDim.x);
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) {
    sc->pq_part[1] = first ? total : sc->pq_part[1] + total;
  }
}

// generic width (blocks wider than 16): one thread per row, one non-returning add per entry
template <int NT>
__device__ __forceinline__ double rows_normal_wide_cta(const GridDesc& g, const RowBlock& rb, const double* __restrict__ p,
                                                       double* __restrict__ q, const int tid, const int vb, const int nvb) {
  double dot = 0.0;
  for (long long r = vb * (long long)NT + tid; r < rb.n; r += (long long)nvb * NT) {
    double t = 0.0, town = 0.0;
    for (int j = 0; j < rb.k; ++j) {
      const int c = rb.cols[j * rb.n + r];
      if (c >= 0) {
        const double v = rb.vals[j * rb.n + r] * p[c];
        t += v;
        if (owned(g, c)) town += v;
      }
    }
    t *= rb.w2[r];
    dot = fma(town, t, dot);
    if (t != 0.0) {
      for (int j = 0; j < rb.k; ++j) {
        const int c = rb.cols[j * rb.n + r];
        if (c >= 0 && owned(g, c)) red_add(q + c, rb.vals[j * rb.n + r] * t);
      }
    }
  }
  return dot;
}

template <int NT>
__global__ void __launch_bounds__(NT)
    k_rows_normal_wide(const GridDesc g, const RowBlock rb, const double* __restrict__ p, double* __restrict__ q,
                       double* __restrict__ partials, unsigned int* __restrict__ counter, CgScalars* __restrict__ sc,
                       const int cg_mode, const int first) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  const double dot = rows_normal_wide_cta<NT>(g, rb, p, q, (int)threadIdx.x, (int)blockIdx.x, (int)gridDim.x);
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) {
    sc->pq_part[1] = first ? total : sc->pq_part[1] + total;
  }
}

// ---- compaction and ordering of a block at finalize -------------------------------------------------------------
// key of a row: its first column (the lowest corner of its cell for point rows; the smallest column of an ingested row);
// dead rows -- weight 0 (points outside the support or, on a z slab, in cells that touch no owned plane) or no column --
// get the largest key and are cut off after the sort.  status[0] counts the live rows.
__global__ void k_rows_sort_keys(const RowBlock rb, unsigned int* __restrict__ keys, int* __restrict__ index,
                                 unsigned long long* __restrict__ status) {
  unsigned long long live = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rb.n; r += (long long)gridDim.x * blockDim.x) {
    const int c0 = rb.cols[r];
    const bool ok = c0 >= 0 && rb.w2[r] != 0.0;
    keys[r] = ok ? (unsigned int)c0 : 0xffffffffu;
    index[r] = (int)r;
    live += ok ? 1ull : 0ull;
  }
  live = __reduce_add_sync(0xffffffffu, (unsigned)live);
  if ((threadIdx.x & 31) == 0 && live) atomicAdd(status, live);
}

// dst row r = src row order[r], r < dst.n
__global__ void k_rows_permute(const RowBlock src, const int* __restrict__ order, RowBlock dst) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < dst.n; r += (long long)gridDim.x * blockDim.x) {
    const long long o = order[r];
    for (int j = 0; j < src.k; ++j) {
      dst.vals[j * dst.n + r] = src.vals[j * src.n + o];
      dst.cols[j * dst.n + r] = src.cols[j * src.n + o];
    }
    dst.w2[r] = src.w2[o];
    dst.b[r] = src.b[o];
  }
}

#endif  // LSB_EMU

// rhs += A_d^T W b
__global__ void k_rows_rhs(const GridDesc g, const RowBlock rb, double* __restrict__ rhs) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rb.n; r += (long long)gridDim.x * blockDim.x) {
    const double t = rb.w2[r] * rb.b[r];
    if (t == 0.0) continue;
    for (int j = 0; j < rb.k; ++j) {
      const int c = rb.cols[j * rb.n + r];
      if (c >= 0 && owned(g, c)) atomicAdd(rhs + c, rb.vals[j * rb.n + r] * t);
    }
  }
}

// diag += diag(A_d^T W A_d).  Listed duplicate columns of one row are merged first, as the
// reference's coo->csc conversion does (LI/_discrete_interpolator.py:384-386).
__global__ void k_rows_diag(const GridDesc g, const RowBlock rb, double* __restrict__ diag) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rb.n; r += (long long)gridDim.x * blockDim.x) {
    const double w2 = rb.w2[r];
    if (w2 == 0.0) continue;
    for (int j = 0; j < rb.k; ++j) {
      const int c = rb.cols[j * rb.n + r];
      if (c < 0 || !owned(g, c)) continue;
      bool dup = false;
      for (int jj = 0; jj < j; ++jj) dup |= (rb.cols[jj * rb.n + r] == c);
      if (dup) continue;
      double s = 0.0;
      for (int jj = j; jj < rb.k; ++jj)
        if (rb.cols[jj * rb.n + r] == c) s += rb.vals[jj * rb.n + r];
      atomicAdd(diag + c, w2 * s * s);
    }
  }
}

// Ingest host-format rows (row-major vals/global cols) into an ELL block: optional row
// normalisation (LI/_discrete_interpolator.py:361-366), global -> local column ids, w -> w^2.
// status[0] counts rows kept, status[1] flags a kept row with a column outside owned+halo planes, status[2] a column id
// outside [0, nx*ny*nz) (status[3]: one such id).
__global__ void k_rows_ingest(const GridDesc g, const long long n, const int k, const double* __restrict__ vals,
                              const long long* __restrict__ cols, const double* __restrict__ b,
                              const double* __restrict__ w, const int normalise, RowBlock rb,
                              const long long row0, unsigned long long* __restrict__ status) {
  const long long nxny = (long long)g.nx * g.ny;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    double nrm2 = 0.0;
    bool touches = false, bad = false, nan = false;
    for (int j = 0; j < k; ++j) {
      const double a = vals[r * k + j];
      nrm2 += a * a;
      nan |= (a != a);
      const long long gc = cols[r * k + j];
      if (gc < 0 || gc >= nxny * g.nz) {
        if (atomicExch(status + 2, 1ull) == 0ull) status[3] = (unsigned long long)gc;
        bad = true;
      }
      const long long kz = gc < 0 ? -1 - HALO + g.k0 : gc / nxny;
      touches |= (kz >= g.k0 && kz < g.k1);
      bad |= (kz < g.k0 - HALO || kz >= g.k1 + HALO);
    }
    double bb = b[r];
    nan |= (bb != bb) || (w[r] != w[r]);
    const double len = sqrt(nrm2);
    const double inv = (normalise && len > 0.0) ? 1.0 / len : 1.0;
    const bool keep = touches && !nan;
    if (keep && bad) atomicExch(status + 1, 1ull);
    const bool ok = keep && !bad;
    if (ok) atomicAdd(status, 1ull);
    const long long ro = row0 + r;
    for (int j = 0; j < k; ++j) {
      int lc = -1;
      double a = 0.0;
      if (ok) {
        const long long gi = cols[r * k + j];
        const long long kz = gi / nxny, rem = gi - kz * nxny;
        const long long jy = rem / g.nx, ix = rem - jy * g.nx;
        lc = (int)((kz - g.k0 + HALO) * g.plane + jy * g.pitch + ix);
        // match numpy: a / len (division, not multiplication by the reciprocal)
        a = (normalise && len > 0.0) ? vals[r * k + j] / len : vals[r * k + j];
      }
      rb.vals[j * rb.n + ro] = a;
      rb.cols[j * rb.n + ro] = lc;
    }
    (void)inv;
    rb.b[ro] = ok ? ((normalise && len > 0.0) ? bb / len : bb) : 0.0;
    rb.w2[ro] = ok ? w[r] * w[r] : 0.0;
  }
}

// ---- K3: FDI rows from constraint points ---------------------------------------------------

struct CellLoc {
  bool inside;
  int ci[3];
  double lc[3];  // local coordinates clipped to [0,1]
};

// position_to_cell_index / inside / position_to_local_coordinates
// (LC/supports/_3d_base_structured.py:251-293; LC/supports/_3d_structured_grid.py:126-146)
LSB_HD CellLoc locate(const GridDesc& g, const double x[3]) {
  CellLoc L;
  L.inside = true;
  const int nc[3] = {g.nx - 1, g.ny - 1, g.nz - 1};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double maxi = rn_add(g.origin[a], rn_mul((double)nc[a], g.step[a]));
    const double ext = fmax(fabs(maxi - g.origin[a]), 1.0);
    const double tol = rn_mul(rn_mul(2.220446049250313e-16, 32.0), ext);
    L.inside = L.inside && (x[a] >= g.origin[a] - tol) && (x[a] <= maxi + tol);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int c = 0;
    double l = 0.0;
    if (L.inside) {
      const double f = floor((x[a] - g.origin[a]) / g.step[a]);
      c = (int)fmin(fmax(f, 0.0), (double)(nc[a] - 1));
      const double co = rn_add(g.origin[a], rn_mul(g.step[a], (double)c));
      l = fmin(fmax((x[a] - co) / g.step[a], 0.0), 1.0);
    }
    L.ci[a] = c;
    L.lc[a] = l;
  }
  return L;
}

// trilinear shape functions, corner c has offsets (c&1,(c>>1)&1,(c>>2)&1)
// (LC/supports/_3d_structured_grid.py:90-124)
LSB_HD void trilinear_w(const double l[3], double wgt[8]) {
  const double x = l[0], y = l[1], z = l[2];
  wgt[0] = (1 - x) * (1 - y) * (1 - z);
  wgt[1] = x * (1 - y) * (1 - z);
  wgt[2] = (1 - x) * y * (1 - z);
  wgt[3] = x * y * (1 - z);
  wgt[4] = (1 - x) * (1 - y) * z;
  wgt[5] = x * (1 - y) * z;
  wgt[6] = (1 - x) * y * z;
  wgt[7] = x * y * z;
}

// derivative of the shape functions / step (LC/supports/_3d_structured_grid.py:393-454)
LSB_HD void trilinear_grad(const GridDesc& g, const double l[3], double T[3][8]) {
  const double x = l[0], y = l[1], z = l[2];
  T[0][0] = (1 - z) * (y - 1); T[0][1] = (1 - y) * (1 - z); T[0][2] = -y * (1 - z); T[0][3] = y * (1 - z);
  T[0][4] = -(1 - y) * z;      T[0][5] = (1 - y) * z;       T[0][6] = -y * z;       T[0][7] = y * z;
  T[1][0] = (x - 1) * (1 - z); T[1][1] = -x * (1 - z);      T[1][2] = (1 - x) * (1 - z); T[1][3] = x * (1 - z);
  T[1][4] = -(1 - x) * z;      T[1][5] = -x * z;            T[1][6] = (1 - x) * z;  T[1][7] = x * z;
  T[2][0] = -(1 - x) * (1 - y); T[2][1] = -x * (1 - y);     T[2][2] = -(1 - x) * y; T[2][3] = -x * y;
  T[2][4] = (1 - x) * (1 - y); T[2][5] = x * (1 - y);       T[2][6] = (1 - x) * y;  T[2][7] = x * y;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) T[a][c] /= g.step[a];
}

LSB_HD long long corner_local(const GridDesc& g, const int ci[3], int c) {
  const int i = ci[0] + (c & 1), j = ci[1] + ((c >> 1) & 1), k = ci[2] + ((c >> 2) & 1);
  return (long long)(k - g.k0 + HALO) * g.plane + (long long)j * g.pitch + i;
}

// get_vectors (LC/math/_maths.py:276-298): strike and dip vectors of a unit normal
__device__ __forceinline__ void strike_dip(const double nrm[3], double strike[3], double dip[3]) {
  const double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
  const double n[3] = {nrm[0] / len, nrm[1] / len, nrm[2] / len};
  const double phi = atan2(n[1], n[0]);
  const double deg = -phi * (180.0 / 3.14159265358979323846);
  const double s = deg * (3.14159265358979323846 / 180.0);
  double sv[3] = {sin(-s), -cos(-s), 0.0};
  const double sl = sqrt(sv[0] * sv[0] + sv[1] * sv[1]);
  sv[0] /= sl; sv[1] /= sl;
  double dv[3] = {sv[1] * n[2] - sv[2] * n[1], sv[2] * n[0] - sv[0] * n[2], sv[0] * n[1] - sv[1] * n[0]};
  const double dl = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    strike[a] = sv[a] * len;
    dip[a] = dv[a] / dl * len;
  }
}

// Directional regularisation of a grid, FiniteDifferenceInterpolator.minimise_directional_gradient_change
// (LI/_finite_difference_interpolator.py:1207-1340): at every node interior on all axes the six second-derivative masks
// (LI/_operator.py:29-35) as rows of their own, weighted by the components of the direction v at the node --
// vx^2 fxx + vy^2 fyy + vz^2 fzz + 2 vx vy fxy + 2 vx vz fxz + 2 vy vz fyz = 0 -- each row L2-normalised like every row
// (dxx: (1,-2,1)/sqrt 6; the mixed masks: four entries of +-1/2), b = 0, w = weight * component; rows whose weight is 0 or
// NaN are dead (the reference drops them, :1325-1329) and are cut off when the block is compacted at finalize.
// Centres: planes [ka, kb) = the interior planes whose stencil touches an owned plane; vec holds the (node, 3) vectors
// of the planes from ka on; row of (op, centre c) = op * ncent + c.
__global__ void k_fdi_directional_rows(const GridDesc g, const int ka, const int kb, const double* __restrict__ vec,
                                       const double weight, RowBlock rb, unsigned long long* __restrict__ status) {
  const int mx = g.nx - 2, my = g.ny - 2;
  const long long ncent = (long long)mx * my * (kb - ka);
  const double c3 = 1.0 / sqrt(6.0), c3m = -2.0 / sqrt(6.0);
  unsigned long long kept = 0;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncent; c += (long long)gridDim.x * blockDim.x) {
    const int i = 1 + (int)(c % mx);
    const long long rr = c / mx;
    const int j = 1 + (int)(rr % my), k = ka + (int)(rr / my);
    const long long vn = ((long long)(k - ka) * g.ny + j) * g.nx + i;
    const double vx = vec[3 * vn], vy = vec[3 * vn + 1], vz = vec[3 * vn + 2];
    const int centre = (int)((long long)(k - g.k0 + HALO) * g.plane + (long long)j * g.pitch + i);
    const int sx = 1, sy = g.pitch, sz = (int)g.plane;
    const double comp[6] = {vx * vx, vy * vy, vz * vz, 2.0 * vx * vy, 2.0 * vx * vz, 2.0 * vy * vz};
    const int sa[6] = {sx, sy, sz, sx, sx, sy}, sb[6] = {0, 0, 0, sy, sz, sz};
#pragma unroll
    for (int op = 0; op < 6; ++op) {
      const long long ro = (long long)op * ncent + c;
      const double rw = weight * comp[op];
      const bool live = fabs(rw) > 0.0;  // false for NaN
      if (op < 3) {
        rb.cols[0 * rb.n + ro] = centre - sa[op]; rb.vals[0 * rb.n + ro] = c3;
        rb.cols[1 * rb.n + ro] = centre;          rb.vals[1 * rb.n + ro] = c3m;
        rb.cols[2 * rb.n + ro] = centre + sa[op]; rb.vals[2 * rb.n + ro] = c3;
        rb.cols[3 * rb.n + ro] = -1;              rb.vals[3 * rb.n + ro] = 0.0;
      } else {
        // mask value at (a, b) offsets: -(-1,-1), +(+1,-1), +(-1,+1), -(+1,+1) (Operator.Dxy_mask and its axis swaps)
        rb.cols[0 * rb.n + ro] = centre - sa[op] - sb[op]; rb.vals[0 * rb.n + ro] = -0.5;
        rb.cols[1 * rb.n + ro] = centre + sa[op] - sb[op]; rb.vals[1 * rb.n + ro] = 0.5;
        rb.cols[2 * rb.n + ro] = centre - sa[op] + sb[op]; rb.vals[2 * rb.n + ro] = 0.5;
        rb.cols[3 * rb.n + ro] = centre + sa[op] + sb[op]; rb.vals[3 * rb.n + ro] = -0.5;
      }
      for (int e = 4; e < rb.k; ++e) {
        rb.cols[e * rb.n + ro] = -1;
        rb.vals[e * rb.n + ro] = 0.0;
      }
      rb.b[ro] = 0.0;
      rb.w2[ro] = live ? rw * rw : 0.0;
      kept += live ? 1ull : 0ull;
    }
  }
  if (kept) atomicAdd(status, kept);
}

// One thread per constraint point; writes `mult` rows per point into rb at row0 + m*n + point.
// Points outside the support (or on a slab: in cells that touch no owned plane) become empty rows.
__global__ void k_fdi_point_rows(const GridDesc g, const int kind, const long long n, const double* __restrict__ data,
                                 const int ncols, const double weight, RowBlock rb, const long long row0,
                                 unsigned long long* __restrict__ status) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double* d = data + r * ncols;
    const double x[3] = {d[0], d[1], d[2]};
    CellLoc L = locate(g, x);
    bool ok = L.inside;
    for (int c = 0; c < ncols; ++c) ok = ok && (d[c] == d[c]);
    // slab: keep the row when its cell touches an owned plane
    ok = ok && (L.ci[2] + 1 >= g.k0) && (L.ci[2] < g.k1);
    const int mult = kind == LSB200_ROWS_GRADIENT ? 2 : (kind == LSB200_ROWS_NORMAL ? 3 : 1);
    double A[3][8];
    double bb[3] = {0.0, 0.0, 0.0};
    double ww = weight;
    if (ok) {
      if (kind == LSB200_ROWS_VALUE) {
        trilinear_w(L.lc, A[0]);
        bb[0] = d[3];
        ww = weight * d[4];
      } else {
        double T[3][8];
        trilinear_grad(g, L.lc, T);
        if (kind == LSB200_ROWS_NORMAL) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c < 8; ++c) A[a][c] = T[a][c];
            bb[a] = d[3 + a];
          }
          ww = weight / 3.0 * d[6];
        } else {
          const double len = sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
          double v[3] = {d[3], d[4], d[5]};
          if (len > 0.0 || kind == LSB200_ROWS_GRADIENT) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              v[a] /= len;
#pragma unroll
              for (int c = 0; c < 8; ++c) T[a][c] /= len;
            }
          }
          if (kind == LSB200_ROWS_GRADIENT) {
            double sv[3], dv[3];
            strike_dip(v, sv, dv);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              A[0][c] = sv[0] * T[0][c] + sv[1] * T[1][c] + sv[2] * T[2][c];
              A[1][c] = dv[0] * T[0][c] + dv[1] * T[1][c] + dv[2] * T[2][c];
            }
          } else {  // tangent / orthogonal
#pragma unroll
            for (int c = 0; c < 8; ++c) A[0][c] = v[0] * T[0][c] + v[1] * T[1][c] + v[2] * T[2][c];
            bb[0] = d[6];
          }
        }
      }
      atomicAdd(status, (unsigned long long)mult);
    }
    for (int m = 0; m < mult; ++m) {
      const long long ro = row0 + (long long)m * n + r;
      double nrm2 = 0.0;
      if (ok)
#pragma unroll
        for (int c = 0; c < 8; ++c) nrm2 += A[m][c] * A[m][c];
      const double len = sqrt(nrm2);
      ok = ok && (len == len);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        rb.vals[c * rb.n + ro] = ok ? (len > 0.0 ? A[m][c] / len : A[m][c]) : 0.0;
        rb.cols[c * rb.n + ro] = ok ? (int)corner_local(g, L.ci, c) : -1;
      }
      rb.b[ro] = ok ? (len > 0.0 ? bb[m] / len : bb[m]) : 0.0;
      rb.w2[ro] = ok ? ww * ww : 0.0;
    }
  }
}

// Interface rows N(x_i) - N(x_j) = 0 (FDI.add_interface_constraints,
// LI/_finite_difference_interpolator.py:297-340): 16 listed entries per pair, normalised over the listed
// entries (duplicate columns merge only later in the reference, LI/_discrete_interpolator.py:361-366,384-386).
// The pairing itself (all / star / chain per interface id) is host logic, as in the reference (:271-295).
__global__ void k_fdi_pair_rows(const GridDesc g, const long long n, const double* __restrict__ xi,
                                const double* __restrict__ xj, const double weight, RowBlock rb,
                                unsigned long long* __restrict__ status) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double a[3] = {xi[3 * r], xi[3 * r + 1], xi[3 * r + 2]};
    const double b[3] = {xj[3 * r], xj[3 * r + 1], xj[3 * r + 2]};
    const CellLoc La = locate(g, a), Lb = locate(g, b);
    bool ok = La.inside && Lb.inside;
    // a slab can only apply a row whose two cells both lie within its owned + halo planes
    const bool touches = (La.ci[2] + 1 >= g.k0 && La.ci[2] < g.k1) || (Lb.ci[2] + 1 >= g.k0 && Lb.ci[2] < g.k1);
    const bool reach = La.ci[2] >= g.k0 - HALO && La.ci[2] + 1 < g.k1 + HALO && Lb.ci[2] >= g.k0 - HALO &&
                       Lb.ci[2] + 1 < g.k1 + HALO;
    if (ok && touches && !reach) atomicExch(status + 1, 1ull);
    ok = ok && touches && reach;
    double wa[8], wb[8];
    double nrm2 = 0.0;
    if (ok) {
      trilinear_w(La.lc, wa);
      trilinear_w(Lb.lc, wb);
#pragma unroll
      for (int c = 0; c < 8; ++c) nrm2 += wa[c] * wa[c] + wb[c] * wb[c];
      atomicAdd(status, 1ull);
    }
    const double len = sqrt(nrm2);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      rb.vals[c * rb.n + r] = ok ? (len > 0.0 ? wa[c] / len : wa[c]) : 0.0;
      rb.cols[c * rb.n + r] = ok ? (int)corner_local(g, La.ci, c) : -1;
      rb.vals[(c + 8) * rb.n + r] = ok ? (len > 0.0 ? -wb[c] / len : -wb[c]) : 0.0;
      rb.cols[(c + 8) * rb.n + r] = ok ? (int)corner_local(g, Lb.ci, c) : -1;
    }
    rb.b[r] = 0.0;
    rb.w2[r] = ok ? weight * weight : 0.0;
  }
}

// ---- deterministic application of the data rows (lsb200_set_options(..., deterministic = 1)) -----------------
// The atomics of k_rows_normal / k_rows_rhs / k_rows_diag add in arrival order, so two solves of the same system
// differ in the last bits (and, on ill-conditioned systems, by an iteration).  RowTranspose is the node-major view
// of one ELL block, built once per finalize by a stable radix sort of (node, row) keys: entries [off[u], off[u+1])
// of `entry` belong to node `node[u]`, ordered by row.  Then  t = W A p  is one pass over the rows and
// q += A^T t  one pass over the touched nodes, every sum in a fixed order and no atomics.
struct RowTranspose {
  int n_nodes;         // touched nodes
  int* node;           // [n_nodes] storage index
  int* off;            // [n_nodes + 1] segment offsets into entry
  int* entry;          // [valid entries] ELL entry index e = j * n + r
  double* t;           // [n rows] scratch: w^2 (a . p)
};

__global__ void k_rows_keys(const RowBlock rb, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const long long E = rb.n * rb.k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const int c = rb.cols[e];
    const unsigned long long r = (unsigned long long)(e % rb.n);
    keys[e] = c < 0 ? ~0ull : (((unsigned long long)(unsigned int)c << 32) | r);  // padding sorts to the end
    vals[e] = (int)e;
  }
}
__global__ void k_rows_node_of_key(const long long n, const unsigned long long* __restrict__ keys, int* __restrict__ nodes) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    nodes[e] = keys[e] == ~0ull ? -1 : (int)(keys[e] >> 32);
}

// t_r = w_r^2 (a_r . p); pq part = sum_r t_r (a_r . p) restricted to owned columns, as k_rows_normal counts it
template <int NT>
__global__ void __launch_bounds__(NT)
    k_rows_t(const GridDesc g, const RowBlock rb, const RowTranspose tr, const double* __restrict__ p,
             double* __restrict__ partials, unsigned int* __restrict__ counter, CgScalars* __restrict__ sc,
             const int cg_mode, const int first) {
  __shared__ double red[NT / 32];
  if (cg_mode && sc->done) return;
  double dot = 0.0;
  for (long long r = blockIdx.x * (long long)NT + threadIdx.x; r < rb.n; r += (long long)gridDim.x * NT) {
    double t = 0.0, town = 0.0;
    for (int j = 0; j < rb.k; ++j) {
      const int c = rb.cols[j * rb.n + r];
      if (c >= 0) {
        const double v = rb.vals[j * rb.n + r] * p[c];
        t += v;
        if (owned(g, c)) town += v;
      }
    }
    t *= rb.w2[r];
    tr.t[r] = t;
    dot += town * t;
  }
  const double bsum = block_sum<NT>(dot, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total))
    sc->pq_part[1] = first ? total : sc->pq_part[1] + total;
}

// mode 0: q[node] += sum_e a_e t_row(e);  mode 1: rhs[node] += sum_e a_e w_r^2 b_r;
// mode 2: diag[node] += sum_rows w_r^2 (sum of the row's entries on this node)^2  (entries of a node are ordered by row)
__global__ void k_rows_gather(const GridDesc g, const RowBlock rb, const RowTranspose tr, double* __restrict__ out,
                              const int mode, const CgScalars* __restrict__ sc, const int cg_mode) {
  if (cg_mode && sc->done) return;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < tr.n_nodes; u += gridDim.x * blockDim.x) {
    const int node = tr.node[u];
    if (!owned(g, node)) continue;
    double acc = 0.0;
    if (mode == 2) {
      long long prev = -1;
      double s = 0.0, w2 = 0.0;
      for (int t = tr.off[u]; t < tr.off[u + 1]; ++t) {
        const int e = tr.entry[t];
        const long long r = e % rb.n;
        if (r != prev) {
          acc += w2 * s * s;
          s = 0.0;
          w2 = rb.w2[r];
          prev = r;
        }
        s += rb.vals[e];
      }
      acc += w2 * s * s;
    } else {
      for (int t = tr.off[u]; t < tr.off[u + 1]; ++t) {
        const int e = tr.entry[t];
        const long long r = e % rb.n;
        acc += rb.vals[e] * (mode == 0 ? tr.t[r] : rb.w2[r] * rb.b[r]);
      }
    }
    out[node] += acc;
  }
}

}  // namespace lsb

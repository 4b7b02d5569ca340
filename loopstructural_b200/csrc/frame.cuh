// K9: fault-frame evaluation -- the largest consumer of the solved fields (SURVEY.md section 8 f, rank 1).
// A fault frame is three scalar fields solved one after the other (three handles, three c vectors).  The reference
// moves a point cloud through a fault by gathering all three values and the gradient of the middle field at every
// point, `steps` times, each gather a separate numpy pass over the cloud:
//   FaultSegment.apply_to_points        LS/modelling/features/fault/_fault_segment.py:315-423
//   FaultSegment.evaluate_displacement  LS/modelling/features/fault/_fault_segment.py:283-313
//   profile functions                   LS/modelling/features/fault/_fault_function.py:12-16,40-140,198-226,308-369
//   world <-> local transforms          LC/geometry/_bounding_box.py:118-176,702-761 behind
//                                       GeologicalInterpolator.evaluate_value / evaluate_gradient
//                                       (LI/_geological_interpolator.py:723-741)
// Every step of that loop is independent per point, so here ONE thread carries one point through all steps:
// xyz is read once (24 B), the moved point written once (24 B), the 3 x 8 corner values per step come from L2 / HBM
// as random 8-byte gathers exactly as in k_eval_value_grid.  Each field is a grid (FDI, the reference's default for
// faults) or a tet mesh (PLI / fold interpolators); fault-frame features must carry no faults / regions of their own
// -- the host mirror checks and otherwise leaves the call to the reference.
//
// The per-point code is LSB_HD: tests/hostsim runs the same functions on the CPU against the oracle.
#pragma once
#include "common.cuh"
#include "p1.cuh"
#include "rows.cuh"

namespace lsb {

// one solved field of the frame
struct FrameField {
  int kind;  // LSB200_SUPPORT_GRID: trilinear cells; LSB200_SUPPORT_TETMESH: P1 on the 5-tet decomposition
  int pad;
  GridDesc g;
  const double* c;  // node vector of the handle (halo planes included)
  double w2l[12];   // world -> local, rows of the 3 x 4 affine: l_a = w2l[4a] x + w2l[4a+1] y + w2l[4a+2] z + w2l[4a+3]
  double l2w[9];    // local -> world rotation for gradients (row-major 3 x 3)
};
struct FrameFields {
  FrameField f[3];
};

// one 1-D profile: kinds of LS/modelling/features/fault/_fault_function.py
struct Profile1D {
  int kind;  // LSB200_PROFILE_ZEROS / ONES / CUBIC / SMOOTH_PEAK
  int pad;
  double w[4];    // CubicFunction.w: w0 v^3 + w1 v^2 + w2 v + w3
  double lim[2];  // CubicFunction.lim: clamped to the value at the limit outside
  double min_v, max_v;
};
struct FaultProfile {
  int has_function;  // 0: FaultSegment.faultfunction is None, d = [gx > 0]
  int composite[3];  // per axis: Composite(pos, neg) (0 at v == 0); else `pos` for every v
  double scale;      // FaultDisplacement.scale
  Profile1D pos[3], neg[3];
};

LSB_HD double cubic_at(const double w[4], const double v) {
  // w0*v**3 + w1*v**2 + w2*v + w3, left to right, every product and sum rounded once
  const double v2 = rn_mul(v, v);
  return rn_add(rn_add(rn_add(rn_mul(w[0], rn_mul(v2, v)), rn_mul(w[1], v2)), rn_mul(w[2], v)), w[3]);
}

LSB_HD double profile1d(const Profile1D& P, const double v) {
  switch (P.kind) {
    case 0: return 0.0;
    case 1: return 1.0;
    case 3: {  // smooth_peak: x^4 - 2 x^2 + 1 on [-1, 1], 0 outside
      if (!(v >= -1.0 && v <= 1.0)) return 0.0;
      const double v2 = rn_mul(v, v);
      return rn_add(rn_add(rn_mul(v2, v2), -rn_mul(2.0, v2)), 1.0);
    }
    default: {  // CubicFunction.__call__
      double e = cubic_at(P.w, v);
      if (v > P.lim[1]) e = cubic_at(P.w, P.lim[1]);
      if (v < P.lim[0]) e = cubic_at(P.w, P.lim[0]);
      if (e > P.max_v) e = P.max_v;
      if (e < P.min_v) e = P.min_v;
      return e;
    }
  }
}

LSB_HD double profile_axis(const FaultProfile& F, const int a, const double v) {
  if (!F.composite[a]) return profile1d(F.pos[a], v);
  if (v > 0.0) return profile1d(F.pos[a], v);
  if (v < 0.0) return profile1d(F.neg[a], v);
  return 0.0;
}

// d of one step before `* displacement` (_fault_segment.py:389-403); *ok = none of the three values is NaN
LSB_HD double fault_step_d(const FaultProfile& F, const double gx, const double gy, const double gz, const double offset,
                           bool* ok) {
  *ok = (gx == gx) && (gy == gy) && (gz == gz);
  if (!*ok) return 0.0;
  if (!F.has_function) return gx > 0.0 ? 1.0 : 0.0;
  return rn_mul(rn_mul(rn_mul(F.scale, profile_axis(F, 0, rn_add(gx, offset))), profile_axis(F, 1, gy)), profile_axis(F, 2, gz));
}

// value (and, when grad != nullptr, world-frame gradient) of one field at a WORLD point: project, locate, gather.
// Value: 0.0 when a projected coordinate is NaN, NaN outside the support; gradient: 0 outside a grid, NaN outside a
// tet mesh (StructuredGrid.evaluate_gradient vs TetMesh.evaluate_gradient, LC/supports/_3d_structured_grid.py:356-357,
// LC/supports/_3d_structured_tetra.py:247-248).
// TET = false: every field of the call is a grid (the tet branch is not compiled: 124 instead of 252 registers).
template <bool TET = true>
LSB_HD double frame_field_at(const FrameField& F, const double xw[3], double* grad) {
  double l[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
    l[a] = rn_add(rn_add(rn_add(rn_mul(F.w2l[4 * a], xw[0]), rn_mul(F.w2l[4 * a + 1], xw[1])), rn_mul(F.w2l[4 * a + 2], xw[2])),
                  F.w2l[4 * a + 3]);
  const bool nanrow = (l[0] != l[0]) || (l[1] != l[1]) || (l[2] != l[2]);
  double v = nanrow ? 0.0 : nan("");
  double gl[3] = {0.0, 0.0, 0.0};
  if (TET && F.kind == LSB200_SUPPORT_TETMESH) {
    const TetLoc T = locate_tet(F.g, l);
    double cv[4] = {0.0, 0.0, 0.0, 0.0};
    if (T.inside) {
#pragma unroll
      for (int k = 0; k < 4; ++k) cv[k] = ld_node(F.c + T.node[k]);
      v = T.c[0] * cv[0] + T.c[1] * cv[1] + T.c[2] * cv[2] + T.c[3] * cv[3];
    }
    if (grad) {
      if (T.inside) {
        double G[3][4];
        tet_gradients(T.v, G);
#pragma unroll
        for (int a = 0; a < 3; ++a) gl[a] = G[a][0] * cv[0] + G[a][1] * cv[1] + G[a][2] * cv[2] + G[a][3] * cv[3];
      } else {
        gl[0] = gl[1] = gl[2] = nan("");
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
        grad[a] = rn_add(rn_add(rn_mul(F.l2w[3 * a], gl[0]), rn_mul(F.l2w[3 * a + 1], gl[1])), rn_mul(F.l2w[3 * a + 2], gl[2]));
    }
    return v;
  }
  const CellLoc L = locate(F.g, l);
  if (L.inside) {
    double cv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cv[k] = ld_node(F.c + corner_local(F.g, L.ci, k));
    double wgt[8];
    trilinear_w(L.lc, wgt);
    v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += wgt[k] * cv[k];
    if (grad) {
      double T[3][8];
      trilinear_grad(F.g, L.lc, T);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        gl[0] += T[0][k] * cv[k];
        gl[1] += T[1][k] * cv[k];
        gl[2] += T[2][k] * cv[k];
      }
    }
  }
  if (grad) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
      grad[a] = rn_add(rn_add(rn_mul(F.l2w[3 * a], gl[0]), rn_mul(F.l2w[3 * a + 1], gl[1])), rn_mul(F.l2w[3 * a + 2], gl[2]));
  }
  return v;
}

// FaultSegment.apply_to_points for one point: x is moved in place
template <bool TET = true>
LSB_HD void fault_displace_point(const FrameFields& FF, const FaultProfile& P, const double displacement, const int steps,
                                 double x[3]) {
  double g[3];
  double gx = frame_field_at<TET>(FF.f[0], x, nullptr);
  double gy = frame_field_at<TET>(FF.f[1], x, g);
  double gz = frame_field_at<TET>(FF.f[2], x, nullptr);
  bool ok;
  const double d0 = fault_step_d(P, gx, gy, gz, 0.0, &ok);
  if (!(fabs(d0) > 0.0)) return;  // mask = |d| > 0: only these points enter the loop
  const double inv_steps = 1.0 / (double)steps;
  for (int i = 0; i < steps; ++i) {
    if (i > 0) {
      gx = frame_field_at<TET>(FF.f[0], x, nullptr);
      gy = frame_field_at<TET>(FF.f[1], x, g);
      gz = frame_field_at<TET>(FF.f[2], x, nullptr);
    }
    const double d = rn_mul(fault_step_d(P, gx, gy, gz, 0.0, &ok), displacement);
    double gmag = 0.0;
    if (ok) gmag = sqrt(rn_add(rn_add(rn_mul(g[0], g[0]), rn_mul(g[1], g[1])), rn_mul(g[2], g[2])));
    if (gmag > 0.0) {
      g[0] /= gmag;
      g[1] /= gmag;
      g[2] /= gmag;
    }
    const double s = rn_mul(inv_steps, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a] = rn_add(x[a], rn_mul(g[a], s));
  }
}

// FaultSegment.evaluate_displacement for one point
template <bool TET = true>
LSB_HD double fault_displacement_point(const FrameFields& FF, const FaultProfile& P, const double displacement,
                                       const double offset, const double x[3]) {
  const double gx = frame_field_at<TET>(FF.f[0], x, nullptr);
  const double gy = frame_field_at<TET>(FF.f[1], x, nullptr);
  const double gz = frame_field_at<TET>(FF.f[2], x, nullptr);
  bool ok;
  return rn_mul(fault_step_d(P, gx, gy, gz, offset, &ok), displacement);
}

// ---- frame-orthogonality rows generated from a solved field (SURVEY.md section 8 f, rank 2) -------------------
// GeologicalFeatureBuilder.install_gradient_constraint (LS/modelling/features/builders/_geological_feature_builder.py:
// 293-318): the gradient of ANOTHER feature at the barycentre of every `estep`-th cell of this support, normalised,
// becomes one row  v . grad(phi) = b  of FDI.add_gradient_orthogonal_constraints
// (LI/_finite_difference_interpolator.py:478-539), row-normalised by add_constraints_to_least_squares
// (LI/_discrete_interpolator.py:361-366).  O(N / estep) rows: the reference evaluates the other feature on the host
// over ALL N barycentres, einsums and builds a COO matrix; here one thread makes one row from the device-resident
// field.  The barycentres are this support's LOCAL coordinates and the reference passes them to the other feature's
// evaluate_gradient, which projects them as world coordinates (rows come out empty where that lands outside the other
// support): kept as is, pinned by tests/golden/orth_world.npz.
// Row r <- cell e = r * estep.  Returns false when the row is empty on this handle (slab: cell touches no owned plane).
template <bool TET = true>
LSB_HD bool field_orthogonal_row(const GridDesc& g, const FrameField& other, const long long e, const double bval,
                                 double A[8], int ci[3], double* b_out) {
  const int ncx = g.nx - 1, ncy = g.ny - 1;
  const long long cxy = (long long)ncx * ncy;
  const int cell[3] = {(int)(e % ncx), (int)((e / ncx) % ncy), (int)(e / cxy)};
  double x[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)  // StructuredGrid.cell_centres: origin + step * cell + step * 0.5
    x[a] = rn_add(rn_add(g.origin[a], rn_mul(g.step[a], (double)cell[a])), rn_mul(g.step[a], 0.5));
  double v[3];
  frame_field_at<TET>(other, x, v);
  const double n1 = sqrt(rn_add(rn_add(rn_mul(v[0], v[0]), rn_mul(v[1], v[1])), rn_mul(v[2], v[2])));
  if (n1 > 0.0) { v[0] /= n1; v[1] /= n1; v[2] /= n1; }
  const CellLoc L = locate(g, x);
  ci[0] = L.ci[0]; ci[1] = L.ci[1]; ci[2] = L.ci[2];
  const bool ok = L.inside && (L.ci[2] + 1 >= g.k0) && (L.ci[2] < g.k1);
  double T[3][8];
  trilinear_grad(g, L.lc, T);
  // add_gradient_orthogonal_constraints normalises the vector (and T) once more
  const double n2 = sqrt(rn_add(rn_add(rn_mul(v[0], v[0]), rn_mul(v[1], v[1])), rn_mul(v[2], v[2])));
  if (n2 > 0.0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      v[a] /= n2;
#pragma unroll
      for (int c = 0; c < 8; ++c) T[a][c] /= n2;
    }
  }
  double nrm2 = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    A[c] = v[0] * T[0][c] + v[1] * T[1][c] + v[2] * T[2][c];
    nrm2 += A[c] * A[c];
  }
  const double len = sqrt(nrm2);
  if (len > 0.0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) A[c] /= len;
  }
  *b_out = len > 0.0 ? bval / len : bval;
  return ok;
}

// body of k_fdi_field_orthogonal_rows for row r (shared with tests/hostsim)
template <bool TET = true>
LSB_HD bool field_orthogonal_row_store(const GridDesc& g, const FrameField& other, const long long estep, const double weight,
                                       const double bval, const RowBlock& rb, const long long row0, const long long r) {
  double A[8], b;
  int ci[3];
  const bool ok = field_orthogonal_row<TET>(g, other, r * estep, bval, A, ci, &b);
  const long long ro = row0 + r;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    rb.vals[c * rb.n + ro] = ok ? A[c] : 0.0;
    rb.cols[c * rb.n + ro] = ok ? (int)corner_local(g, ci, c) : -1;
  }
  rb.b[ro] = ok ? b : 0.0;
  rb.w2[ro] = ok ? weight * weight : 0.0;
  return ok;
}

// ---- evaluation on a regular lattice without an xyz input (SURVEY.md section 8 f, rank 4) ----------------------
// BaseFeature.scalar_field / gradient_norm_scalar_field / GeologicalModel.evaluate_model
// (LS/modelling/features/_base_geological_feature.py:359-449) evaluate a feature on BoundingBox.regular_grid
// (LC/geometry/_bounding_box.py:532-573): np.linspace per axis -> meshgrid -> an (N,3) array that is projected and
// gathered.  The lattice is 9 numbers; generating the coordinates in the kernel leaves 8 B (value) or 24 B (gradient)
// of HBM traffic per point instead of 24 B in + 8 B out on top of building the array on the host.
struct Lattice {
  double start[3], stop[3], delta[3];  // np.linspace(start, stop, n): delta = (stop - start) / (n - 1)
  long long n[3];
  int order_f;  // 1: flatten(order='F'), first axis fastest; 0: order='C', last axis fastest
};

// point `idx` of the flattened lattice: np.linspace's values, x_i = i * delta + start, the last one exactly `stop`
LSB_HD void lattice_point(const Lattice& lat, const long long idx, double x[3]) {
  long long i[3];
  if (lat.order_f) {
    i[0] = idx % lat.n[0];
    i[1] = (idx / lat.n[0]) % lat.n[1];
    i[2] = idx / (lat.n[0] * lat.n[1]);
  } else {
    i[2] = idx % lat.n[2];
    i[1] = (idx / lat.n[2]) % lat.n[1];
    i[0] = idx / (lat.n[2] * lat.n[1]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (lat.n[a] == 1) x[a] = lat.start[a];
    else if (i[a] == lat.n[a] - 1) x[a] = lat.stop[a];
    else x[a] = rn_add(rn_mul((double)i[a], lat.delta[a]), lat.start[a]);
  }
}

#ifdef __CUDACC__
// points [first, first + m) of the lattice: value (m) or world gradient (m,3) of one field
template <bool TET>
__global__ void __launch_bounds__(256)
    k_eval_lattice(const FrameField F, const Lattice lat, const long long first, const long long m, const int gradient,
                   double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x) {
    double x[3], g[3];
    lattice_point(lat, first + r, x);
    const double v = frame_field_at<TET>(F, x, gradient ? g : nullptr);
    if (gradient) {
      out[3 * r] = g[0];
      out[3 * r + 1] = g[1];
      out[3 * r + 2] = g[2];
    } else {
      out[r] = v;
    }
  }
}

// The P1 / TetMesh twin: install_gradient_constraint hands TetMesh.barycentre[::estep] (mean of the four vertices,
// LC/supports/_3d_structured_tetra.py:106-118) and the other feature's normalised gradient there to
// P1Interpolator.add_gradient_orthogonal_constraints (LI/_p1interpolator.py:248-296), which is the tangent family of
// k_p1_point_rows (v_hat . grad phi = b, weight w * tet volume).  This kernel makes that family's input on the device --
// row r: x,y,z of the barycentre of tet r * estep, the normalised gradient of `other` there, b -- and k_p1_point_rows
// turns it into rows; nothing of size N crosses the bus.
template <bool TET>
LSB_HD void p1_barycentre_vector(const GridDesc& g, const FrameField& other, const long long e, const double bval, double* d) {
  const int ncx = g.nx - 1, ncy = g.ny - 1;
  const long long cell_id = e / 5;
  const int local = (int)(e - 5 * cell_id);
  const int cell[3] = {(int)(cell_id % ncx), (int)((cell_id / ncx) % ncy), (int)(cell_id / ((long long)ncx * ncy))};
  const bool even = ((cell[0] + cell[1] + cell[2]) & 1) == 0;
  double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int corner = tet_corner(even, local, s);
    const int off[3] = {corner & 1, (corner >> 1) & 1, (corner >> 2) & 1};
#pragma unroll
    for (int a = 0; a < 3; ++a) x[a] = rn_add(x[a], rn_add(g.origin[a], rn_mul(g.step[a], (double)(cell[a] + off[a]))));
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) x[a] = x[a] / 4.0;
  double v[3];
  frame_field_at<TET>(other, x, v);
  const double n1 = sqrt(rn_add(rn_add(rn_mul(v[0], v[0]), rn_mul(v[1], v[1])), rn_mul(v[2], v[2])));
  if (n1 > 0.0) { v[0] /= n1; v[1] /= n1; v[2] /= n1; }
  d[0] = x[0]; d[1] = x[1]; d[2] = x[2];
  d[3] = v[0]; d[4] = v[1]; d[5] = v[2];
  d[6] = bval;
}

template <bool TET>
__global__ void __launch_bounds__(128)
    k_p1_barycentre_vectors(const GridDesc g, const FrameField other, const long long estep, const long long n_rows,
                            const double bval, double* __restrict__ data, unsigned long long* __restrict__ status) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    double* d = data + 7 * r;
    p1_barycentre_vector<TET>(g, other, r * estep, bval, d);
    if (d[3] != d[3] || d[4] != d[4] || d[5] != d[5]) atomicAdd(status + 2, 1ull);  // outside a tet-mesh source: see the caller
  }
}

template <bool TET>
__global__ void __launch_bounds__(128)
    k_fdi_field_orthogonal_rows(const GridDesc g, const FrameField other, const long long estep, const long long n_rows,
                                const double weight, const double bval, const RowBlock rb, const long long row0,
                                unsigned long long* __restrict__ status) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    if (field_orthogonal_row_store<TET>(g, other, estep, weight, bval, rb, row0, r)) atomicAdd(status, 1ull);
    const double a0 = rb.vals[row0 + r];
    if (a0 != a0) atomicAdd(status + 2, 1ull);  // NaN gradient of the other field (outside a tet-mesh source)
  }
}

// the three values (n,3) and, when grad_out != nullptr, the gradient (n,3) of field `gfield`: one pass over xyz
template <bool TET>
__global__ void __launch_bounds__(256)
    k_frame_eval(const FrameFields FF, const int gfield, const long long n, const double* __restrict__ xyz,
                 double* __restrict__ val_out, double* __restrict__ grad_out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int f = 0; f < 3; ++f) val_out[3 * r + f] = frame_field_at<TET>(FF.f[f], x, (grad_out && f == gfield) ? g : nullptr);
    if (grad_out) {
      grad_out[3 * r] = g[0];
      grad_out[3 * r + 1] = g[1];
      grad_out[3 * r + 2] = g[2];
    }
  }
}

// value (n) or world gradient (n,3) of ONE field at world points that are already on the device (the tail of a fault
// chain: GeologicalFeature.evaluate_value / evaluate_gradient after _apply_faults)
template <bool TET>
__global__ void __launch_bounds__(256)
    k_field_eval(const FrameField F, const int gradient, const long long n, const double* __restrict__ xyz,
                 double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    double g[3];
    const double v = frame_field_at<TET>(F, x, gradient ? g : nullptr);
    if (gradient) {
      out[3 * r] = g[0];
      out[3 * r + 1] = g[1];
      out[3 * r + 2] = g[2];
    } else {
      out[r] = v;
    }
  }
}

template <bool TET>
__global__ void __launch_bounds__(256)
    k_fault_displace(const FrameFields FF, const FaultProfile P, const double displacement, const int steps,
                     const long long n, const double* __restrict__ xyz, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    fault_displace_point<TET>(FF, P, displacement, steps, x);
    out[3 * r] = x[0];
    out[3 * r + 1] = x[1];
    out[3 * r + 2] = x[2];
  }
}

template <bool TET>
__global__ void __launch_bounds__(256)
    k_fault_displacement(const FrameFields FF, const FaultProfile P, const double displacement, const double offset,
                         const long long n, const double* __restrict__ xyz, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    out[r] = fault_displacement_point<TET>(FF, P, displacement, offset, x);
  }
}
#endif

}  // namespace lsb

// TEST INFRASTRUCTURE ONLY (not part of liblsb200.so, never loaded by the product): runs the per-point code of the
// fault-frame kernels (loopstructural_b200/csrc/frame.cuh -- the very LSB_HD functions k_frame_eval /
// k_fault_displace / k_fault_displacement call per thread) in a host loop, so that tests/ can compare it with the
// oracle and the golden fixtures in a container that has no GPU.  The kernels' launch code is not exercised here;
// tests/test_zz_fault_gpu.py does that on the device.
//
//   nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared -o tests/hostsim/libframe_hostsim.so tests/hostsim/frame_hostsim.cu
#include "../../include/lsb200.h"

#include <string.h>

#include <vector>

#include "../../loopstructural_b200/csrc/frame.cuh"

using namespace lsb;

namespace {
struct HostFrame {
  FrameFields FF;
  std::vector<double> store[3];  // node vectors in the handle's layout: two halo planes either side
};

int g_kinds[3] = {LSB200_SUPPORT_GRID, LSB200_SUPPORT_GRID, LSB200_SUPPORT_GRID};  // set by hostsim_set_kinds

void build(HostFrame& H, const int64_t* nsteps, const double* origin, const double* step, const double* const c[3],
           const double* transforms) {
  for (int f = 0; f < 3; ++f) {
    FrameField& F = H.FF.f[f];
    F.kind = g_kinds[f];
    F.pad = 0;
    GridDesc& g = F.g;
    g.nx = (int)nsteps[3 * f]; g.ny = (int)nsteps[3 * f + 1]; g.nz = (int)nsteps[3 * f + 2];
    g.k0 = 0; g.k1 = g.nz; g.nzl = g.nz;
    g.pitch = g.nx;
    g.plane = (long long)g.pitch * g.ny;
    g.nloc = g.plane * g.nzl;
    g.nall = g.plane * (g.nzl + 2 * HALO);
    for (int a = 0; a < 3; ++a) { g.origin[a] = origin[3 * f + a]; g.step[a] = step[3 * f + a]; }
    H.store[f].assign((size_t)g.nall, 0.0);
    memcpy(H.store[f].data() + (size_t)HALO * g.plane, c[f], sizeof(double) * (size_t)g.nloc);
    F.c = H.store[f].data();
    if (transforms) {
      for (int i = 0; i < 12; ++i) F.w2l[i] = transforms[21 * f + i];
      for (int i = 0; i < 9; ++i) F.l2w[i] = transforms[21 * f + 12 + i];
    } else {
      for (int i = 0; i < 12; ++i) F.w2l[i] = (i % 5 == 0) ? 1.0 : 0.0;
      for (int i = 0; i < 9; ++i) F.l2w[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
  }
}
}  // namespace

#define HOSTSIM_API __attribute__((visibility("default")))
extern "C" {

// support kind of the three fields of the following calls (LSB200_SUPPORT_GRID / LSB200_SUPPORT_TETMESH)
HOSTSIM_API void hostsim_set_kinds(int k0, int k1, int k2) { g_kinds[0] = k0; g_kinds[1] = k1; g_kinds[2] = k2; }

HOSTSIM_API int hostsim_frame_eval(const int64_t* nsteps, const double* origin, const double* step, const double* c0, const double* c1,
                       const double* c2, const double* transforms, int gradient_of, int64_t n, const double* xyz,
                       double* values, double* gradient) {
  HostFrame H;
  const double* c[3] = {c0, c1, c2};
  build(H, nsteps, origin, step, c, transforms);
  for (int64_t r = 0; r < n; ++r) {  // body of k_frame_eval
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    double g[3] = {0.0, 0.0, 0.0};
    for (int f = 0; f < 3; ++f) values[3 * r + f] = frame_field_at(H.FF.f[f], x, (gradient && f == gradient_of) ? g : nullptr);
    if (gradient) { gradient[3 * r] = g[0]; gradient[3 * r + 1] = g[1]; gradient[3 * r + 2] = g[2]; }
  }
  return 0;
}

HOSTSIM_API int hostsim_fault_displace(const int64_t* nsteps, const double* origin, const double* step, const double* c0,
                           const double* c1, const double* c2, const double* transforms,
                           const lsb200_fault_profile* profile, double displacement, int steps, int64_t n,
                           const double* xyz, double* moved) {
  HostFrame H;
  const double* c[3] = {c0, c1, c2};
  build(H, nsteps, origin, step, c, transforms);
  FaultProfile P;
  memcpy(&P, profile, sizeof(P));
  for (int64_t r = 0; r < n; ++r) {  // body of k_fault_displace
    double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    fault_displace_point(H.FF, P, displacement, steps, x);
    moved[3 * r] = x[0]; moved[3 * r + 1] = x[1]; moved[3 * r + 2] = x[2];
  }
  return 0;
}

HOSTSIM_API int hostsim_fault_displacement(const int64_t* nsteps, const double* origin, const double* step, const double* c0,
                               const double* c1, const double* c2, const double* transforms,
                               const lsb200_fault_profile* profile, double displacement, double offset, int64_t n,
                               const double* xyz, double* d) {
  HostFrame H;
  const double* c[3] = {c0, c1, c2};
  build(H, nsteps, origin, step, c, transforms);
  FaultProfile P;
  memcpy(&P, profile, sizeof(P));
  for (int64_t r = 0; r < n; ++r) {  // body of k_fault_displacement
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    d[r] = fault_displacement_point(H.FF, P, displacement, offset, x);
  }
  return 0;
}

// body of k_fdi_field_orthogonal_rows: rows of `target` (nsteps/origin/step at index 0 of the three-field arrays are the
// TARGET support, index 1 the OTHER field whose c is c1 and whose transform is transforms[21..42)); outputs are
// (n_rows, 8) row-major values / local column ids (-1 = empty), b and w^2.
HOSTSIM_API int hostsim_field_orthogonal_rows(const int64_t* nsteps, const double* origin, const double* step, const double* c_other,
                                  const double* transforms, int64_t estep, double weight, double bval, int64_t n_rows,
                                  double* vals, int64_t* cols, double* b, double* w2) {
  HostFrame H;
  std::vector<double> dummy((size_t)(nsteps[0] * nsteps[1] * nsteps[2]), 0.0), dummy2((size_t)(nsteps[6] * nsteps[7] * nsteps[8]), 0.0);
  const double* c[3] = {dummy.data(), c_other, dummy2.data()};
  build(H, nsteps, origin, step, c, transforms);
  const GridDesc& g = H.FF.f[0].g;
  std::vector<double> v((size_t)n_rows * 8), bb((size_t)n_rows), ww((size_t)n_rows);
  std::vector<int> cc((size_t)n_rows * 8);
  RowBlock rb{};
  rb.n = n_rows; rb.k = 8; rb.vals = v.data(); rb.cols = cc.data(); rb.b = bb.data(); rb.w2 = ww.data();
  int64_t kept = 0;
  for (int64_t r = 0; r < n_rows; ++r) kept += field_orthogonal_row_store(g, H.FF.f[1], estep, weight, bval, rb, 0, r) ? 1 : 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    for (int k = 0; k < 8; ++k) {
      vals[r * 8 + k] = v[(size_t)k * n_rows + r];
      const int lc = cc[(size_t)k * n_rows + r];
      cols[r * 8 + k] = lc < 0 ? -1 : (int64_t)lc - (int64_t)HALO * g.plane;  // storage index -> global node id (full handle, pitch = nx)
    }
    b[r] = bb[r];
    w2[r] = ww[r];
  }
  return (int)kept;
}

// body of k_eval_lattice for the field at index 0 of the three-field arrays
HOSTSIM_API int hostsim_eval_lattice(const int64_t* nsteps, const double* origin, const double* step, const double* c0,
                         const double* transforms, const double* start, const double* stop, const int64_t* n, int order_f,
                         int gradient, double* out) {
  HostFrame H;
  std::vector<double> d1((size_t)(nsteps[3] * nsteps[4] * nsteps[5]), 0.0), d2((size_t)(nsteps[6] * nsteps[7] * nsteps[8]), 0.0);
  const double* c[3] = {c0, d1.data(), d2.data()};
  build(H, nsteps, origin, step, c, transforms);
  Lattice lat;
  long long total = 1;
  for (int a = 0; a < 3; ++a) {  // make_lattice of csrc/api.cu
    lat.start[a] = start[a]; lat.stop[a] = stop[a]; lat.n[a] = n[a];
    lat.delta[a] = n[a] > 1 ? (stop[a] - start[a]) / (double)(n[a] - 1) : 0.0;
    total *= n[a];
  }
  lat.order_f = order_f ? 1 : 0;
  for (long long r = 0; r < total; ++r) {
    double x[3], g[3];
    lattice_point(lat, r, x);
    const double v = frame_field_at(H.FF.f[0], x, gradient ? g : nullptr);
    if (gradient) { out[3 * r] = g[0]; out[3 * r + 1] = g[1]; out[3 * r + 2] = g[2]; }
    else out[r] = v;
  }
  return 0;
}

}  // extern "C"

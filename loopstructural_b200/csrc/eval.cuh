// K7: batched evaluation of the solved field and its gradient on a StructuredGrid.
// One thread per query point; xyz (n,3) row-major in the support's local coordinates.
#pragma once
#include "common.cuh"
#include "rows.cuh"

namespace lsb {

// StructuredGrid.evaluate_value (LC/supports/_3d_structured_grid.py:295-322) behind
// DiscreteInterpolator._evaluate_value_local (LI/_discrete_interpolator.py:1342-1362):
// NaN outside the grid, 0.0 for rows holding a NaN coordinate.
__global__ void k_eval_value_grid(const GridDesc g, const double* __restrict__ c, const long long n,
                                  const double* __restrict__ xyz, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    double v;
    if (x[0] != x[0] || x[1] != x[1] || x[2] != x[2]) {
      v = 0.0;
    } else {
      const CellLoc L = locate(g, x);
      if (!L.inside) {
        v = nan("");
      } else {
        double wgt[8];
        trilinear_w(L.lc, wgt);
        v = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += wgt[k] * __ldg(c + corner_local(g, L.ci, k));
      }
    }
    out[r] = v;
  }
}

// StructuredGrid.evaluate_gradient (LC/supports/_3d_structured_grid.py:324-391): rows outside the
// grid (NaN coordinates included) are 0 because T is zero-initialised (:356-357).
__global__ void k_eval_gradient_grid(const GridDesc g, const double* __restrict__ c, const long long n,
                                     const double* __restrict__ xyz, double* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double x[3] = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
    double gr[3] = {0.0, 0.0, 0.0};
    const CellLoc L = locate(g, x);
    if (L.inside) {
      double T[3][8];
      trilinear_grad(g, L.lc, T);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double cv = __ldg(c + corner_local(g, L.ci, k));
        gr[0] += T[0][k] * cv;
        gr[1] += T[1][k] * cv;
        gr[2] += T[2][k] * cv;
      }
    }
    out[3 * r] = gr[0];
    out[3 * r + 1] = gr[1];
    out[3 * r + 2] = gr[2];
  }
}

}  // namespace lsb

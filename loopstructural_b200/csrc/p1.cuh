// P1 / TetMesh kernels (K4, K5, K8) -- placeholder while the FDI path is brought up.
#pragma once
#include "common.cuh"
#include "rows.cuh"

namespace lsb {

static inline int p1_launch_face_normal(const GridDesc&, double, double, const double*, double*, double*, unsigned int*,
                                        CgScalars*, int, int, cudaStream_t, long long*) {
  return 1;
}
static inline int p1_launch_face_diag(const GridDesc&, double, double, double*, int, cudaStream_t, long long*) { return 1; }
__global__ void k_p1_point_rows(const GridDesc, const int, const long long, const double*, const int, const double, RowBlock,
                                const long long, unsigned long long*) {}
__global__ void k_eval_value_tet(const GridDesc, const double*, const long long, const double*, double*) {}
__global__ void k_eval_gradient_tet(const GridDesc, const double*, const long long, const double*, double*) {}

}  // namespace lsb

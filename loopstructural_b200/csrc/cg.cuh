// K6: device-resident conjugate gradients.
//
// Replaces scipy.sparse.linalg.cg as called by solve_with_cg (LI/_solver_strategy.py:48-71) on the
// ridge-augmented, column-scaled system of preprocess_main_system (LI/_solver_pipeline.py:86-124).
// scipy's loop (M = None):   test ||r|| < max(atol, rtol ||b||) at the top of every iteration;
//   rho = r.r ; p = r + (rho/rho_prev) p ; q = A p ; alpha = rho / p.q ; x += alpha p ; r -= alpha q.
// The reference runs it on y = S^-1 c with A = S N S.  Substituting x = S y, r = S^-1 r_y, p = S p_y
// gives the same iterates in c-space without ever forming S N S:
//   rho = r . S^2 r (= ||r_y||^2, so the stopping test is unchanged) ; p = S^2 r + beta p ; q = N p ;
//   alpha = rho / p.q ; x += alpha p ; r -= alpha q.
// All scalars live in CgScalars on the device; kernels after convergence are no-ops (sc->done), so
// the host only polls.
#pragma once
#include "common.cuh"

namespace lsb {

// advance the scalar state after rho_new is known (one thread)
__device__ __forceinline__ void cg_scalar_step(CgScalars* sc, double rho_new) {
  sc->beta = rho_new / sc->rho;
  sc->rho = rho_new;
  sc->iter += 1;
  if (sc->iter >= sc->maxiter) {
    sc->done = 1;
    sc->info = (int)(sc->maxiter > 2147483647LL ? 2147483647LL : sc->maxiter);
  } else if (rho_new < sc->thr2) {
    sc->done = 1;
    sc->info = 0;
  }
}

// p = S^2 r + beta p   (s2 == nullptr: no scaling)
template <int NT>
__global__ void __launch_bounds__(NT)
    k_cg_update_p(const long long n, const double* __restrict__ s2, const double* __restrict__ r,
                  double* __restrict__ p, const CgScalars* __restrict__ sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  const long long n2 = n >> 1;
  const double2* r2 = reinterpret_cast<const double2*>(r);
  const double2* s22 = reinterpret_cast<const double2*>(s2);
  double2* p2 = reinterpret_cast<double2*>(p);
  for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n2; i += (long long)gridDim.x * NT) {
    double2 rv = r2[i], pv = p2[i];
    if (s2) {
      const double2 sv = s22[i];
      rv.x *= sv.x;
      rv.y *= sv.y;
    }
    pv.x = rv.x + beta * pv.x;
    pv.y = rv.y + beta * pv.y;
    p2[i] = pv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long i = n - 1;
    p[i] = (s2 ? s2[i] * r[i] : r[i]) + beta * p[i];
  }
}

// x += alpha p ; r -= alpha q ; rho_new = r . S^2 r
template <int NT>
__global__ void __launch_bounds__(NT)
    k_cg_update_xr(const long long n, const double* __restrict__ p, const double* __restrict__ q,
                   double* __restrict__ x, double* __restrict__ r, const double* __restrict__ s2,
                   double* __restrict__ partials, unsigned int* __restrict__ counter, CgScalars* __restrict__ sc) {
  __shared__ double red[NT / 32];
  if (sc->done) return;
  const double pq = sc->nranks == 1 ? (((sc->pq_part[0] + sc->pq_part[2]) + sc->pq_part[3]) + sc->pq_part[4]) + sc->pq_part[1] : sc->pq;
  const double alpha = sc->rho / pq;
  const long long n2 = n >> 1;
  const double2* p2 = reinterpret_cast<const double2*>(p);
  const double2* q2 = reinterpret_cast<const double2*>(q);
  const double2* s22 = reinterpret_cast<const double2*>(s2);
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n2; i += (long long)gridDim.x * NT) {
    const double2 pv = p2[i], qv = q2[i];
    double2 xv = x2[i], rv = r2[i];
    xv.x += alpha * pv.x;
    xv.y += alpha * pv.y;
    rv.x -= alpha * qv.x;
    rv.y -= alpha * qv.y;
    x2[i] = xv;
    r2[i] = rv;
    if (s2) {
      const double2 sv = s22[i];
      acc += rv.x * sv.x * rv.x + rv.y * sv.y * rv.y;
    } else {
      acc += rv.x * rv.x + rv.y * rv.y;
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long i = n - 1;
    x[i] += alpha * p[i];
    const double rv = r[i] - alpha * q[i];
    r[i] = rv;
    acc += (s2 ? rv * s2[i] * rv : rv * rv);
  }
  const double bsum = block_sum<NT>(acc, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) {
    sc->rho_new = total;
    if (sc->nranks == 1) cg_scalar_step(sc, total);
  }
}

// multi-rank: scalar steps run as tiny kernels after the all-reduce
__global__ void k_cg_pack_pq(CgScalars* sc) {
  if (sc->done) return;
  sc->pq = (((sc->pq_part[0] + sc->pq_part[2]) + sc->pq_part[3]) + sc->pq_part[4]) + sc->pq_part[1];
}
__global__ void k_cg_scalar_step(CgScalars* sc) {
  if (sc->done) return;
  cg_scalar_step(sc, sc->rho_new);
}

// r = b - q (q may be nullptr) ; *out = sum r . S^2 r.  Used for ||S b||^2 and the initial residual.
template <int NT>
__global__ void __launch_bounds__(NT)
    k_cg_residual(const long long n, const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r,
                  const double* __restrict__ s2, double* __restrict__ partials, unsigned int* __restrict__ counter,
                  double* __restrict__ out) {
  __shared__ double red[NT / 32];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const double rv = q ? b[i] - q[i] : b[i];
    if (r) r[i] = rv;
    acc += s2 ? rv * s2[i] * rv : rv * rv;
  }
  const double bsum = block_sum<NT>(acc, red);
  double total;
  if (grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total)) *out = total;
}

// s2 = 1/diag where diag > 0, else 1 (compute_column_scaling_matrix, LI/_discrete_interpolator.py:815-833:
// S = 1/||column||, 1 for empty columns; here s2 = S^2)
__global__ void k_scaling_from_diag(const long long n, const double* __restrict__ diag, double* __restrict__ s2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = diag[i];
    // S = 1/sqrt(d) as the reference computes it; S^2 formed from S so that S*S is what multiplies r
    const double s = d > 0.0 ? 1.0 / sqrt(d) : 1.0;
    s2[i] = s * s;
  }
}

}  // namespace lsb

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

// ---- multi-rank scalar reductions over peer-mapped memory ----------------------------------------------------------
// Round 1 all-reduced p.q and r.S^2 r with two 8-byte ncclAllReduce calls per iteration, serial on the critical path
// (~70 us at 2 ranks, more at 8: SCALE_r01).  Here every rank owns a small mailbox in device memory that its peers have
// mapped (cudaIpcOpenMemHandle, lsb200_comm_init): a rank STORES its partial into slot [parity][kind][rank] of every
// mailbox, fences, and stores the sequence number of the iteration into the matching flag; a reader spins on the flags
// of its OWN mailbox (local memory) and sums the partials in rank order, so that every rank forms bit-identical
// scalars and the result does not depend on arrival order.  Two parities: a rank can be at most one reduction ahead of
// the slowest reader of the previous one (it needs that reader's partial of the current reduction to get further).
constexpr int PEER_MAX_RANKS = 16;
struct PeerMailbox {
  double val[2][2][PEER_MAX_RANKS];                 // [parity][kind: 0 = p.q, 1 = rho][writer rank]
  unsigned long long flag[2][2][PEER_MAX_RANKS];    // sequence number of the iteration the value belongs to
  unsigned long long halo[2][2];                    // [parity][0: from the rank below, 1: from the rank above]: the halo planes of p of that iteration have landed
};
struct PeerTable {
  PeerMailbox* box[PEER_MAX_RANKS];  // box[r]: rank r's mailbox as mapped into this process (box[rank] = the local one)
  int rank, nranks;
};

__device__ __forceinline__ void peer_push(const PeerTable& pt, const int kind, const unsigned long long seq, const double v) {
  // one lane per peer
  const int r = threadIdx.x;
  if (r < pt.nranks) {
    PeerMailbox* mb = pt.box[r];
    const int par = (int)(seq & 1ull);
    mb->val[par][kind][pt.rank] = v;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&mb->flag[par][kind][pt.rank]) = seq;
  }
}

// sum of all ranks' partials of reduction (kind, seq), in rank order; ONE WARP calls it (lane r waits for rank r's flag, so
// the waits overlap; one system-scope fence for the warp), every lane gets the value.  (The first version polled the ranks
// one after the other from thread 0 of EVERY block of the x / r update: 1184 blocks x 8 ranks x a system-scope fence per
// iteration -- at 8 ranks a real solve ran at 502 us per iteration against 316 us of kernels, profiles/r2aa.)
__device__ __forceinline__ double peer_pull_warp(const PeerTable& pt, const int kind, const unsigned long long seq) {
  const PeerMailbox* mb = pt.box[pt.rank];
  const int par = (int)(seq & 1ull);
  const int lane = threadIdx.x & 31;
  if (lane < pt.nranks) {
    const volatile unsigned long long* f = &mb->flag[par][kind][lane];
    // bounded: a peer that died must not hang this GPU (the sum is then wrong and the solve fails its residual check)
    for (long long spin = 0; *f != seq && spin < (1ll << 31); ++spin) {
    }
  }
  __syncwarp();
  __threadfence_system();
  const double v = lane < pt.nranks ? *reinterpret_cast<const volatile double*>(&mb->val[par][kind][lane]) : 0.0;
  double s = 0.0;
  for (int r = 0; r < pt.nranks; ++r) s += __shfl_sync(0xffffffffu, v, r);  // rank order: the same bits on every rank
  return s;
}

// after the matvec: this rank's p.q (all parts) to every peer, then everybody's into sc->pq.  <<<1, 32>>>
__global__ void k_cg_sync_pq(CgScalars* sc, const PeerTable pt) {
  if (sc->done) return;
  const unsigned long long seq = sc->epoch + (unsigned long long)sc->iter + 1ull;
  const double v = (((sc->pq_part[0] + sc->pq_part[2]) + sc->pq_part[3]) + sc->pq_part[4]) + sc->pq_part[1];
  peer_push(pt, 0, seq, v);
  const double pq = peer_pull_warp(pt, 0, seq);
  if (threadIdx.x == 0) sc->pq = pq;
}
// after the x / r update has pushed its partial of rho: gather, advance the scalar state.  <<<1, 32>>>
__global__ void k_cg_pull_rho(CgScalars* sc, const PeerTable pt) {
  if (sc->done) return;
  const double rho_new = peer_pull_warp(pt, 1, sc->epoch + (unsigned long long)sc->iter + 1ull);
  if (threadIdx.x == 0) {
    sc->rho_new = rho_new;
    cg_scalar_step(sc, rho_new);
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

// ---- halo planes of p over peer memory ------------------------------------------------------------------------------
// Round 1 (and this round's first half) shipped the two edge planes of p with ncclSend / ncclRecv on a side stream "while
// the interior planes are updated" -- but the NCCL kernel needs CTA slots the interior update (a grid that fills every
// SM) does not give up, so the exchange ran AFTER it: +80 us per iteration at 2 ranks (profiles/r2y_bench_2gpu.json).
// Here the kernel that updates an edge of p stores every value twice, into its own vector and into the halo planes of the
// neighbour (that rank's p, mapped through cudaIpc), and the last of its CTAs raises a flag in the neighbour's mailbox;
// the neighbour waits for its two flags in a one-warp kernel in front of its matvec.  No launch of a communication library
// inside the iteration.  A rank may overwrite a neighbour's halo planes as soon as it has this iteration's scalars: every
// rank pushed its part of rho after its own matvec had read them.
template <int NT>
__global__ void __launch_bounds__(NT)
    k_cg_update_p_edge(const long long n, const double* __restrict__ s2, const double* __restrict__ r, double* __restrict__ p,
                       double* __restrict__ nb_halo, PeerMailbox* __restrict__ nb_box, const int from, unsigned int* __restrict__ counter,
                       const CgScalars* __restrict__ sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  const long long n2 = n >> 1;  // n = HALO * plane, plane even: the caller checks
  const double2* r2 = reinterpret_cast<const double2*>(r);
  const double2* s22 = reinterpret_cast<const double2*>(s2);
  double2* p2 = reinterpret_cast<double2*>(p);
  double2* h2 = reinterpret_cast<double2*>(nb_halo);
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
    h2[i] = pv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative: the CTA's stores, ordered before this thread by the barrier, are out
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    if (t == gridDim.x - 1) {  // every CTA's stores are out: tell the neighbour
      __threadfence_system();
      const unsigned long long seq = sc->epoch + (unsigned long long)sc->iter + 1ull;
      *reinterpret_cast<volatile unsigned long long*>(&nb_box->halo[(int)(seq & 1ull)][from]) = seq;
    }
  }
}

// in front of the matvec: until the neighbours' halo planes of this iteration have landed.  <<<1, 32>>>
__global__ void k_cg_halo_wait(const CgScalars* sc, const PeerMailbox* mine, const int has_lo, const int has_hi) {
  if (sc->done || threadIdx.x > 1) return;
  if ((threadIdx.x == 0 && !has_lo) || (threadIdx.x == 1 && !has_hi)) return;
  const unsigned long long seq = sc->epoch + (unsigned long long)sc->iter + 1ull;
  const volatile unsigned long long* f = &mine->halo[(int)(seq & 1ull)][threadIdx.x];
  for (long long spin = 0; *f != seq && spin < (1ll << 31); ++spin) {
  }
  __threadfence_system();
}

// x += alpha p ; r -= alpha q ; rho_new = r . S^2 r
template <int NT>
__global__ void __launch_bounds__(NT)
    k_cg_update_xr(const long long n, const double* __restrict__ p, const double* __restrict__ q,
                   double* __restrict__ x, double* __restrict__ r, const double* __restrict__ s2,
                   double* __restrict__ partials, unsigned int* __restrict__ counter, CgScalars* __restrict__ sc,
                   const PeerTable* __restrict__ peers) {
  __shared__ double red[NT / 32];
  if (sc->done) return;
  double pq;
  if (sc->nranks == 1) {
    pq = (((sc->pq_part[0] + sc->pq_part[2]) + sc->pq_part[3]) + sc->pq_part[4]) + sc->pq_part[1];
  } else {
    pq = sc->pq;  // k_cg_sync_pq (peer mailboxes) or ncclAllReduce put the sum of all ranks there
  }
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
  __shared__ double s_total;
  __shared__ int s_push;
  const bool last = grid_sum_last_block<NT>(bsum, partials, counter, gridDim.x, blockIdx.x, red, &total);
  if (threadIdx.x == 0) s_push = 0;
  __syncthreads();
  if (last) {
    sc->rho_new = total;
    if (sc->nranks == 1) cg_scalar_step(sc, total);
    else if (peers) { s_total = total; s_push = 1; }
  }
  if (sc->nranks > 1 && peers) {
    __syncthreads();
    if (s_push) peer_push(*peers, 1, sc->epoch + (unsigned long long)sc->iter + 1ull, s_total);  // lanes 0..nranks-1 of the last block
  }
}

// multi-rank, NCCL fallback (no peer mapping): scalar steps run as tiny kernels after the all-reduce
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

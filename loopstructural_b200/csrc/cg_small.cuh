// K6 for small systems: several CG iterations in ONE persistent launch.
//
// A CG iteration of a large grid is four bandwidth-bound launches; on a small one (the README's Claudius model: 176 k
// nodes; the fault and fold frames of a geological model: 1e4-1e5 elements each, several per model) every launch is
// shorter than its own latency: 40 us per iteration for 6.5 launches even inside a CUDA graph (profiles/r01q), of which
// the kernels' work is a few microseconds.  Here the CTAs of one cooperative launch stay resident and walk the phases of
// an iteration themselves, with a grid-wide barrier where a phase needs all of the previous one:
//     p = S^2 r + beta p        | barrier |  q = N_reg p  (the CTAs of the operator's own launch plan, walked in a loop)
//     | barrier |  q += A_d^T W A_d p  | barrier: p.q |  x += alpha p, r -= alpha q  | barrier: r.S^2 r | scalar step
// The two dot products ride on the barriers: every CTA leaves its partial in a slot before it arrives and sums all slots
// in slot order after it leaves, so every CTA holds the same bits and no CTA waits for a broadcast.  Same recurrence, same
// stopping rule (cg.cuh: cg_scalar_step) as the launch-per-phase loop; CTA 0 publishes the scalar block once per launch
// for the host's polling.  One rank only (a slab's halo exchange is a host-enqueued NCCL call).
#pragma once
#include "cg.cuh"
#include "fdi_fused.cuh"
#include "p1.cuh"
#include "p1_march.cuh"
#include "rows.cuh"

namespace lsb {

constexpr int SMALL_MAX_BLOCKS = 3;     // ELL width classes of a handle
constexpr int SMALL_MAX_CTAS = 1024;    // slots of the barrier reductions

struct SmallCgArgs {
  long long n;                 // owned nodes (contiguous, halo planes excluded)
  double* x;                   // owned parts of the node vectors
  double* r;
  double* p;
  double* q;
  const double* s2;            // column scaling squared, or null
  const double* p_all;         // the same p / q from their first halo plane: what the operator kernels take
  double* q_all;
  RowBlock blocks[SMALL_MAX_BLOCKS];
  int nblocks;
  int iters;                   // iterations of this launch
  unsigned int* bar;           // [0] arrivals of all barriers of this launch (zeroed by the host in front of it)
  double* slots;               // 2 * SMALL_MAX_CTAS partial sums
  CgScalars* sc;
};

// grid barrier on ONE monotonic counter (zero at launch): a CTA adds its arrival and polls the same word until the
// arrivals of barrier number `gen` are complete -- one atomic round trip and one poll, where a counter that is reset plus
// a generation flag costs the last arriver a second round trip before anybody can leave.  `gen`: barriers passed so far,
// kept in a register by every thread.  The poll is bounded: a CTA that never arrives traps the launch instead of hanging it.
__device__ __forceinline__ unsigned int small_poll(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void small_barrier(unsigned int* bar, const unsigned int nctas, unsigned int& gen, const int tid) {
  __syncthreads();
  if (tid == 0) {
    const unsigned int target = (gen + 1u) * nctas;
    __threadfence();
    atomicAdd(&bar[0], 1u);
    unsigned int spins = 0;
    while (small_poll(&bar[0]) < target) {
      if (++spins > (1u << 25)) __trap();
    }
    __threadfence();
  }
  ++gen;
  __syncthreads();
}

// every CTA: the sum of all CTAs' partials, in slot order (identical bits everywhere); warp 0 reads, everyone gets it
template <int NT>
__device__ __forceinline__ double small_sum(const double* slots, const int nctas, const int tid, double* red) {
  if (tid < 32) {
    double s = 0.0;
    for (int i = tid; i < nctas; i += 32) s += __ldcg(slots + i);
    s = warp_sum(s);
    if (tid == 0) red[0] = s;
  }
  __syncthreads();
  const double v = red[0];
  __syncthreads();
  return v;
}

// REG: the operator of the handle.  FDI: k_fdi_fused's plan (TMA staging, even pitch); P1: k_p1_tiles' plan.
struct SmallFdi {
  GridDesc g;
  UniConsts U;
  StencilWeights w;
  FusedPlan pl;
  const double* ztab;
};
struct SmallP1 {
  GridDesc g;
  P1Tables tb;
  P1TilePlan pl;
  double w2, ridge2;
};

struct SmallP1M {
  GridDesc g;
  P1Tables tb;
  P1MarchPlan pl;
  double w2, ridge2;
};

template <int NT, class RegFn>
__device__ __forceinline__ void small_cg_loop(const SmallCgArgs& a, const GridDesc& g, const int tid, double* red, RegFn reg) {
  const int nctas = (int)gridDim.x, cta = (int)blockIdx.x;
  CgScalars* sc = a.sc;
  // the scalar state, identical in every CTA
  double rho = sc->rho, beta = sc->beta;
  const double thr2 = sc->thr2;
  long long iter = sc->iter;
  const long long maxiter = sc->maxiter;
  int done = sc->done, info = sc->info;
  double rho_new = sc->rho_new, pq = sc->pq;
  unsigned int gen = 0;  // the host zeroes the counter in front of every launch
  // diagnostics (lsb200_debug_small_phases): CTA 0 adds up how long it spends in each phase and at each barrier
  unsigned long long* tr = (cta == 0 && tid == 0) ? g_cta_trace : nullptr;
  unsigned long long t_prev = 0;
  if (tr) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_prev));
  auto mark = [&](const int k) {
    if (tr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      tr[k] += t - t_prev;
      t_prev = t;
    }
  };
  // every CTA has read the state before CTA 0 may publish a new one: the first barrier below orders that
  const long long n2 = a.n >> 1;
  for (int it = 0; it < a.iters && !done; ++it) {
    // ---- p = S^2 r + beta p
    {
      const double2* r2 = reinterpret_cast<const double2*>(a.r);
      const double2* s22 = reinterpret_cast<const double2*>(a.s2);
      double2* p2 = reinterpret_cast<double2*>(a.p);
      for (long long i = cta * (long long)NT + tid; i < n2; i += (long long)nctas * NT) {
        double2 rv = r2[i], pv = p2[i];
        if (a.s2) {
          const double2 sv = s22[i];
          rv.x *= sv.x;
          rv.y *= sv.y;
        }
        pv.x = rv.x + beta * pv.x;
        pv.y = rv.y + beta * pv.y;
        p2[i] = pv;
      }
      if ((a.n & 1) && cta == 0 && tid == 0) {
        const long long i = a.n - 1;
        a.p[i] = (a.s2 ? a.s2[i] * a.r[i] : a.r[i]) + beta * a.p[i];
      }
    }
    mark(0);
    small_barrier(a.bar, nctas, gen, tid);
    mark(1);
    // ---- q = N_reg p, this CTA's share of the operator's launch plan; then the data rows on top of it
    double dot = reg();
    mark(2);
    small_barrier(a.bar, nctas, gen, tid);
    mark(3);
    for (int b = 0; b < a.nblocks; ++b) {
      const RowBlock& rb = a.blocks[b];
      switch (rb.k) {
        case 4: dot += rows_normal_cta<NT, 4>(g, rb, a.p_all, a.q_all, tid, cta, nctas); break;
        case 8: dot += rows_normal_cta<NT, 8>(g, rb, a.p_all, a.q_all, tid, cta, nctas); break;
        case 16: dot += rows_normal_cta<NT, 16>(g, rb, a.p_all, a.q_all, tid, cta, nctas); break;
        default: dot += rows_normal_wide_cta<NT>(g, rb, a.p_all, a.q_all, tid, cta, nctas);
      }
    }
    {
      const double bsum = block_sum<NT>(dot, red);
      if (tid == 0) a.slots[cta] = bsum;
    }
    mark(4);
    small_barrier(a.bar, nctas, gen, tid);
    pq = small_sum<NT>(a.slots, nctas, tid, red);
    mark(5);
    // ---- x += alpha p ; r -= alpha q ; rho_new = r . S^2 r
    const double alpha = rho / pq;
    double acc = 0.0;
    {
      const double2* p2 = reinterpret_cast<const double2*>(a.p);
      const double2* q2 = reinterpret_cast<const double2*>(a.q);
      const double2* s22 = reinterpret_cast<const double2*>(a.s2);
      double2* x2 = reinterpret_cast<double2*>(a.x);
      double2* r2 = reinterpret_cast<double2*>(a.r);
      for (long long i = cta * (long long)NT + tid; i < n2; i += (long long)nctas * NT) {
        const double2 pv = p2[i], qv = q2[i];
        double2 xv = x2[i], rv = r2[i];
        xv.x += alpha * pv.x;
        xv.y += alpha * pv.y;
        rv.x -= alpha * qv.x;
        rv.y -= alpha * qv.y;
        x2[i] = xv;
        r2[i] = rv;
        if (a.s2) {
          const double2 sv = s22[i];
          acc += rv.x * sv.x * rv.x + rv.y * sv.y * rv.y;
        } else {
          acc += rv.x * rv.x + rv.y * rv.y;
        }
      }
      if ((a.n & 1) && cta == 0 && tid == 0) {
        const long long i = a.n - 1;
        a.x[i] += alpha * a.p[i];
        const double rv = a.r[i] - alpha * a.q[i];
        a.r[i] = rv;
        acc += (a.s2 ? rv * a.s2[i] * rv : rv * rv);
      }
    }
    {
      const double bsum = block_sum<NT>(acc, red);
      if (tid == 0) a.slots[SMALL_MAX_CTAS + cta] = bsum;
    }
    mark(6);
    small_barrier(a.bar, nctas, gen, tid);
    rho_new = small_sum<NT>(a.slots + SMALL_MAX_CTAS, nctas, tid, red);
    mark(7);
    // ---- scalar step (cg_scalar_step, on the local copy)
    beta = rho_new / rho;
    rho = rho_new;
    iter += 1;
    if (iter >= maxiter) {
      done = 1;
      info = (int)(maxiter > 2147483647LL ? 2147483647LL : maxiter);
    } else if (rho_new < thr2) {
      done = 1;
      info = 0;
    }
  }
  if (cta == 0 && tid == 0) {
    sc->rho = rho;
    sc->beta = beta;
    sc->iter = iter;
    sc->rho_new = rho_new;
    sc->pq = pq;
    sc->info = info;
    __threadfence();
    sc->done = done;
  }
}

#ifndef LSB_EMU
template <int TYT, int ST, int MINB>
__global__ void __launch_bounds__(32 * TYT, MINB)
    k_cg_small_fdi(const SmallCgArgs a, const SmallFdi f, const LSB_GRID_CONSTANT TmapArg tmap) {
  constexpr int NT = 32 * TYT;
  __shared__ double red[NT / 32];
  const int tid = threadIdx.x + 32 * threadIdx.y;
  unsigned phases = 0;
  bool first = true;
  const int nvb = f.pl.n_box + f.pl.n_xs + f.pl.n_ys;
  small_cg_loop<NT>(a, f.g, tid, red, [&]() {
    double dot = 0.0;
    for (int vb = (int)blockIdx.x; vb < nvb; vb += (int)gridDim.x) {
      dot += fused_cta<TYT, ST, true, true>(f.g, f.U, f.w, f.pl, a.p_all, a.q_all, f.ztab, &tmap, vb, first, phases);
      first = false;
      __syncthreads();  // the ring is free for the next CTA of the plan
    }
    return dot;
  });
}

template <int TY, int TZ, int NW>
__global__ void __launch_bounds__(32 * NW)
    k_cg_small_p1(const SmallCgArgs a, const SmallP1 f, const LSB_GRID_CONSTANT P1ClassTaps T) {
  constexpr int NT = 32 * NW;
  __shared__ double red[NT / 32];
  const int tid = threadIdx.x;
  const int nvb = f.pl.n_bnd_cta + f.pl.tiles_x * f.pl.tiles_y * f.pl.tiles_z;
  small_cg_loop<NT>(a, f.g, tid, red, [&]() {
    double dot = 0.0;
    for (int vb = (int)blockIdx.x; vb < nvb; vb += (int)gridDim.x) {
      dot += p1_tiles_cta<TY, TZ, NW>(f.g, f.tb, T, f.pl, f.w2, f.ridge2, a.p_all, a.q_all, vb, nvb);
      __syncthreads();  // the tile's shared memory is free for the next one
    }
    return dot;
  });
}

template <int NW, int NB, int MINB>
__global__ void __launch_bounds__(32 * (NW + NB), MINB)
    k_cg_small_p1m(const SmallCgArgs a, const SmallP1M f, const LSB_GRID_CONSTANT P1MarchTaps T) {
  constexpr int NT = 32 * (NW + NB);
  __shared__ double red[NT / 32];
  const int tid = threadIdx.x;
  const int nvb = f.pl.n_cta + f.pl.b.n_bnd_cta;
  small_cg_loop<NT>(a, f.g, tid, red, [&]() {
    double dot = 0.0;
    for (int vb = (int)blockIdx.x; vb < nvb; vb += (int)gridDim.x) {
      dot += p1_march_cta<NW, NB>(f.g, f.tb, T, f.pl, f.w2, f.ridge2, a.p_all, a.q_all, vb, nvb);
      __syncthreads();  // the ring is free for the next chunk
    }
    return dot;
  });
}
#endif

}  // namespace lsb

// TEST INFRASTRUCTURE ONLY.  A minimal emulation of the CUDA execution model for host tests: every CUDA thread of
// a CTA is an OS thread, __syncthreads() is a std::barrier, CTAs of a grid run one after the other, cp.async copies
// complete at issue.  Enough to run the block-level kernels of loopstructural_b200/csrc (k_fdi_box, k_fdi_shell,
// k_fdi_normal) unmodified on the CPU -- compiled by g++ with LSB_EMU defined, see the LSB_EMU hooks in common.cuh
// and fdi_stencil.cuh -- so that launch plans and kernels can be checked against the oracle in a container without
// a GPU.  Nothing in liblsb200.so is built with LSB_EMU.
#pragma once
#include <cuda_runtime.h>  // dim3, uint3, double2, make_double2 (host definitions)

#include <algorithm>
#include <barrier>
#include <cmath>
#include <mutex>
#include <thread>
#include <vector>

namespace lsb_emu {
inline thread_local uint3 t_idx;
inline uint3 b_idx;
inline dim3 b_dim, g_dim;
inline std::barrier<>* cta_barrier = nullptr;
inline std::barrier<>* named_barrier[16] = {};  // bar.sync id, count: created by the first thread that arrives
inline std::mutex named_mutex;
inline void named_sync(int id, int count) {
  {
    std::lock_guard<std::mutex> l(named_mutex);
    if (!named_barrier[id]) named_barrier[id] = new std::barrier<>((std::ptrdiff_t)count);
  }
  named_barrier[id]->arrive_and_wait();
}
inline std::vector<double> dyn;
inline double* dyn_smem() { return dyn.data(); }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  g_dim = grid;
  b_dim = block;
  dyn.assign(smem_bytes / 8 + 2, 0.0);
  const unsigned nt = block.x * block.y * block.z;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        b_idx = uint3{bx, by, bz};
        std::barrier<> bar((std::ptrdiff_t)nt);
        cta_barrier = &bar;
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx)
              th.emplace_back([=] {
                t_idx = uint3{tx, ty, tz};
                body();
                cta_barrier->arrive_and_drop();  // a thread that leaves the kernel no longer takes part in barriers
              });
        for (auto& t : th) t.join();
        for (auto& nb : named_barrier) {
          delete nb;
          nb = nullptr;
        }
      }
}
}  // namespace lsb_emu

#define threadIdx (lsb_emu::t_idx)
#define blockIdx (lsb_emu::b_idx)
#define blockDim (lsb_emu::b_dim)
#define gridDim (lsb_emu::g_dim)
#undef __shared__
#define __shared__ static
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#undef __launch_bounds__
#define __launch_bounds__(...)

using std::max;
using std::min;
inline void __syncthreads() { lsb_emu::cta_barrier->arrive_and_wait(); }
inline void __threadfence() {}
template <class T>
inline T __ldg(const T* p) { return *p; }
inline double __ldcg(const double* p) { return *p; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
// called by thread 0 of a CTA only, CTAs run one after the other
inline unsigned int atomicInc(unsigned int* a, unsigned int v) {
  const unsigned int old = *a;
  *a = (old >= v) ? 0u : old + 1u;
  return old;
}

// Shared device/host definitions of the lsb200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace lsb {

// Regular lattice shared by StructuredGrid and TetMesh.  Node counts are GLOBAL; a handle
// owns the z-planes [k0,k1).  Node vectors are stored with two halo planes on either side of
// the owned planes: local plane l = k - k0 + 2, element (i,j,l) at i + pitch*j + plane*l.
// Flat order follows the reference, gi = i + nx*j + nx*ny*k
// (LC/supports/_3d_base_structured.py:346-350), so z-slabs are contiguous.
struct GridDesc {
  int nx, ny, nz;   // global node counts
  int k0, k1;       // owned planes
  int nzl;          // k1 - k0
  int pitch;        // row pitch in elements (>= nx)
  long long plane;  // pitch * ny
  long long nloc;   // plane * nzl        (owned storage, halo excluded)
  long long nall;   // plane * (nzl + 4)  (owned + halo)
  double origin[3];
  double step[3];
};

constexpr int HALO = 2;

// Per-point geometry (point location, shape functions) is written once for the device and the host: the host
// instantiation exists for tests/hostsim only (a CPU harness that runs the very per-point code of a gather kernel
// against the oracle when no GPU is at hand); nothing in liblsb200.so calls it.
#define LSB_HD __host__ __device__ __forceinline__
// single-rounding add / multiply (never contracted into an FMA), numpy's operation order
LSB_HD double rn_add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
LSB_HD double rn_sub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  volatile double r = a - b;
  return r;
#endif
}
LSB_HD double rn_mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
// read-only load of a node value
LSB_HD double ld_node(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// Effective stencil weights squared, pre-multiplied by the constant factors of the
// row-normalised masks (see fdi_stencil.cuh).
struct StencilWeights {
  double axx, ayy, azz;  // ew_xx^2, ew_yy^2, ew_zz^2      (rows [1,-2,1]/sqrt6)
  double bxy, bxz, byz;  // ew_xy^2/4, ew_xz^2/4, ew_yz^2/4 (rows +-1/2 at the 4 corners)
  double dx, dy, dz;     // border families ew_dx^2, ...     (rows [1,-2,1]/sqrt6 on planes 1 and n-2)
  double ridge2;         // ridge_factor^2 (0 when the ridge is off)
};

// CG state kept on the device so that the loop never waits for the host.
struct CgScalars {
  double rho;       // r . S^2 r of the current residual (== ||r||^2 in the column-scaled variables)
  double rho_new;   // reduction target of the x/r update
  double pq;        // p . N p (sum of the parts below, all-reduced across ranks)
  double pq_part[5];  // [0] stencil kernel, [1] data rows, [2]/[3]/[4] y-/x-/z-shell launches of the stencil (0 when unused)
  double beta;
  double thr2;      // threshold^2
  long long iter;
  long long maxiter;
  int done;
  int info;
  int nranks;
  int pad;
  unsigned long long epoch;  // solve counter << 32: makes the sequence numbers of the peer mailboxes unique across solves
};

// Dynamic shared memory of a kernel.  (LSB_EMU: tests/hostsim/cuda_emu.h runs the block-level kernels on the CPU, one
// OS thread per CUDA thread, to check launch plans and kernels without a device; it is never defined in the library.)
#ifdef LSB_EMU
#define LSB_DYN_SMEM(name) double* name = lsb_emu::dyn_smem()
#else
#define LSB_DYN_SMEM(name) extern __shared__ __align__(128) double name[]  // 128 B: destination of TMA tile loads
#endif

// a kernel parameter the device code reads in place (constant bank): TMA descriptors, coefficient tables
#ifdef LSB_EMU
#define LSB_GRID_CONSTANT
#else
#define LSB_GRID_CONSTANT __grid_constant__
#endif

#ifndef LSB_EMU
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, result valid in thread 0.  Fixed order => deterministic.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red /* >= NT/32 doubles */) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  v = warp_sum(v);
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (tid < 32) {
    s = (tid < NT / 32) ? red[tid] : 0.0;
    s = warp_sum(s);
  }
  __syncthreads();
  return s;
}
#else
// emulation: same contract (sum valid in thread 0), plain sum through a CTA-wide scratch array
template <int NT>
inline double block_sum(double v, double* /*red*/) {
  static double scratch[NT];
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  scratch[tid] = v;
  __syncthreads();
  double s = 0.0;
  if (tid == 0)
    for (int i = 0; i < NT; ++i) s += scratch[i];
  __syncthreads();
  return s;
}
#endif

// Diagnostics (lsb200_debug_cta_trace): when set, thread 0 of every CTA of an instrumented kernel records
// {start ns, end ns, SM id} at [3 * block index].  Null in normal operation: one uniform load per CTA.
#ifndef LSB_EMU
__device__ unsigned long long* g_cta_trace = nullptr;
__device__ __forceinline__ unsigned long long trace_begin() {
  unsigned long long t = 0;
  if (g_cta_trace) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_end(const unsigned long long t0, const unsigned int bid) {
  unsigned long long* tr = g_cta_trace;
  if (tr) {
    unsigned long long t1;
    unsigned int sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    tr[3 * (size_t)bid] = t0;
    tr[3 * (size_t)bid + 1] = t1;
    tr[3 * (size_t)bid + 2] = sm;
  }
}
#else
inline unsigned long long trace_begin() { return 0; }
inline void trace_end(unsigned long long, unsigned int) {}
#endif

// "last block done" reduction of per-block partials in a fixed order.  Returns true in
// thread 0 of the last-arriving block, with *total holding the sum.
template <int NT>
__device__ __forceinline__ bool grid_sum_last_block(double block_val, double* partials, unsigned int* counter,
                                                    unsigned int nblocks, unsigned int bid, double* red,
                                                    double* total) {
  __shared__ bool s_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  if (tid == 0) {
    partials[bid] = block_val;
    __threadfence();
    unsigned int t = atomicInc(counter, nblocks - 1);
    s_last = (t == nblocks - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double s = 0.0;
  for (unsigned int i = tid; i < nblocks; i += NT) s += __ldcg(partials + i);
  s = block_sum<NT>(s, red);
  if (tid == 0) *total = s;
  return tid == 0;
}

}  // namespace lsb

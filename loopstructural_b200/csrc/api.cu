// C ABI of the lsb200 engine (include/lsb200.h): handle, device memory, solve driver.
// Host code only orchestrates; all arithmetic on node vectors runs in the kernels of
// fdi_stencil.cuh / rows.cuh / cg.cuh / eval.cuh / p1.cuh.  No CPU fallback anywhere.
#include "../../include/lsb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cub/cub.cuh>

#include <algorithm>
#include <string>
#include <vector>

#include "cg.cuh"
#include "cg_small.cuh"
#include "common.cuh"
#include "eval.cuh"
#include "fdi_stencil.cuh"
#include "fdi_main.cuh"
#include "fdi_plan.cuh"
#include "frame.cuh"
#include "hostpipe.cuh"
#include "iso.cuh"
#include "p1.cuh"
#include "p1_march.cuh"
#include "rows.cuh"

using namespace lsb;

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(LSB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKL()                                                                                         \
  do {                                                                                                \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess)                                                                            \
      return fail(LSB200_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------ NCCL (loaded at run time)
// The library ships without a link-time NCCL dependency; libnccl.so.2 (the copy torch already
// loaded, when the host process uses torch.distributed) is bound on first use.
namespace nccl {
typedef struct ncclComm* comm_t;
typedef struct { char internal[128]; } unique_id;
enum { Sum = 0 };
enum { Float64 = 8, Int8 = 0 };
static void* lib = nullptr;
static int (*GetUniqueId)(unique_id*) = nullptr;
static int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
static int (*CommDestroy)(comm_t) = nullptr;
static int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
static int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
static int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
static int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
static int (*GroupStart)() = nullptr;
static int (*GroupEnd)() = nullptr;
static const char* (*GetErrorString)(int) = nullptr;
static int load() {
  if (lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(LSB200_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(name)                                                                   \
  *(void**)(&name) = dlsym(lib, "nccl" #name);                                      \
  if (!name) return fail(LSB200_ERR_COMM, "libnccl lacks symbol nccl%s", #name);
  SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(AllGather) SYM(Send) SYM(Recv) SYM(GroupStart)
      SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
  return 0;
}
}  // namespace nccl
#define NCK(call)                                                                                     \
  do {                                                                                                \
    int e_ = (call);                                                                                  \
    if (e_ != 0) return fail(LSB200_ERR_COMM, "%s failed: %s", #call, nccl::GetErrorString(e_));       \
  } while (0)

// ------------------------------------------------------------------ context
constexpr int MAX_PARTIALS = 1 << 18;
constexpr int VEC_NT = 256;

struct lsb200_ctx {
  int kind = LSB200_SUPPORT_GRID;
  int device = 0;
  GridDesc g{};
  double ew[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double p1_w = 0.0;
  double* p1_table = nullptr;  // compacted rows of R^T R on the small mesh, see p1.cuh: coefficients,
  int* p1_off = nullptr;       //   tap offsets
  double* p1_diag = nullptr;   //   and centre taps
  int p1_m[3] = {0, 0, 0};
  bool p1_tiles = false;       // k_p1_tiles plan (meshes with an interior), else k_p1_normal everywhere
  int p1_tile_cfg = 0;
  P1TilePlan p1_plan{};
  P1ClassTaps p1_cls{};
  int p1_march = 0;            // configuration of the k_p1_march plan (LSB_P1_MARCH_CFGS), 0: not used
  int p1_march_nw = 0;         // its marching and boundary warps per CTA
  int p1_march_nb = 0;
  int p1_march_minb = 1;       // CTAs of it that fit an SM
  P1MarchPlan p1_mplan{};
  P1MarchTaps p1_mtaps{};
  bool p1_split = false;       // launches outside the persistent loop: k_p1_march<4, 0, 3> marching only (p1_splan), the
  P1MarchPlan p1_splan{};      // boundary shell by k_p1_boundary on a side stream
  double ridge = 1e-8;
  int add_ridge = 1, apply_scaling = 1, deterministic = 0;
  cudaStream_t stream = nullptr;
  double *x = nullptr, *r = nullptr, *p = nullptr, *q = nullptr, *rhs = nullptr, *diag = nullptr, *s2 = nullptr;
  std::vector<RowBlock> pending;  // as added
  std::vector<RowBlock> blocks;   // merged per width at finalize
  std::vector<RowTranspose> transposes;  // node-major views of `blocks` (deterministic mode only)
  long long rows_kept = 0;
  CgScalars* sc = nullptr;
  CgScalars* sc_host = nullptr;  // pinned
  double* partials = nullptr;    // 6 * MAX_PARTIALS: stencil, rows, x/r update, stencil y-, x-, z-shell
  unsigned int* counters = nullptr;
  unsigned long long* status = nullptr;  // 4 words
  bool finalized = false;
  bool have_c = false;
  nccl::comm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // scalar reductions over peer-mapped mailboxes (cg.cuh); null = the two ncclAllReduce calls of round 1
  PeerMailbox* mailbox = nullptr;      // this rank's mailbox (cudaMalloc, exported by cudaIpcGetMemHandle)
  PeerTable peer_host{};               // peers' mailboxes as mapped here
  PeerTable* peers = nullptr;          // device copy of peer_host
  // halo planes of p over peer memory (cg.cuh): the neighbours' p vectors as mapped here ([0] below, [1] above) and the
  // position of the halo planes this rank fills in them; null = ncclSend / ncclRecv
  std::vector<int> slab_k;             // [k0, k1) of every rank
  double* nb_p_raw[2] = {nullptr, nullptr};
  double* nb_halo[2] = {nullptr, nullptr};
  bool peer_halo = false;
  unsigned long long solve_epoch = 0;
  int cz = 0;              // k_fdi_normal (grids too small for box tiles): z chunk and grid
  dim3 sgrid;
  bool fused = false;      // the single-launch plan (k_fdi_fused): whenever the grid is large enough
  // small systems: the CG loop as one persistent launch per batch of iterations (csrc/cg_small.cuh)
  bool small_ok = false;
  unsigned int* small_bar = nullptr;
  double* small_slots = nullptr;
  int small_grid = 0;
  // the CUDA graph of one batch of iterations, kept from solve to solve: lsb200_finalize_system drops it (row blocks,
  // weights and the launch plan are kernel arguments by value)
  cudaGraphExec_t cg_graph = nullptr;
  long long cg_graph_launches = 0;
  bool tuned = false;      // the plan has been chosen among the planner's candidates by timing them (tune_fused)
  bool plan_forced = false;  // an environment switch fixed the plan: no tuning
  FusedChoice fz{};
  // TMA descriptors of the two node vectors the operator is applied to (p inside CG, x for the initial residual and
  // lsb200_residual_norm), one per box configuration (their boxes differ); tma_ok: built for this handle
  bool tma_ok = false;
  bool want_tma = true;    // LSB200_TMA=0 (read when the handle is planned): cp.async staging everywhere
  const double* tma_vec[2] = {nullptr, nullptr};
  TmapArg tma_map[2];
  long long guard = 0;     // elements of the guard plane in front of every node vector (ensure_vectors)
  double* ztab = nullptr;  // per-plane z coefficients of the planes next to a z face (nz * SHELL_ZSTRIDE)
  // side[1]: the halo exchange of a slab handle runs beside the interior part of the p update (fork/join by events)
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  bool overlap_shells = true;  // LSB200_OVERLAP=0: the exchange on the main stream
  int sm_count = 148;
  long long launches = 0;
  double setup_seconds = 0.0;
  HostPipe pipe;  // staging of the host-array entry points (hostpipe.cuh)
};

constexpr long long PIPE_CHUNK = 1 << 21;  // points per staged chunk (48 MB in, 16-96 MB out)
static int pipe_result(int rc, cudaError_t e, const char* what) {
  if (e != cudaSuccess) return fail(LSB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return rc;
}

extern "C" {
static void free_transposes(lsb200_ctx* h);  // defined beside lsb200_finalize_system
}

static int vec_blocks(const lsb200_ctx* h, long long n, int per_thread = 4) {
  long long b = (n + (long long)VEC_NT * per_thread - 1) / ((long long)VEC_NT * per_thread);
  long long cap = (long long)h->sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static StencilWeights make_weights(const lsb200_ctx* h) {
  StencilWeights w;
  const double* e = h->ew;
  w.axx = e[LSB200_W_DXX] * e[LSB200_W_DXX];
  w.ayy = e[LSB200_W_DYY] * e[LSB200_W_DYY];
  w.azz = e[LSB200_W_DZZ] * e[LSB200_W_DZZ];
  w.bxy = 0.25 * e[LSB200_W_DXY] * e[LSB200_W_DXY];
  w.bxz = 0.25 * e[LSB200_W_DXZ] * e[LSB200_W_DXZ];
  w.byz = 0.25 * e[LSB200_W_DYZ] * e[LSB200_W_DYZ];
  w.dx = e[LSB200_W_DX] * e[LSB200_W_DX];
  w.dy = e[LSB200_W_DY] * e[LSB200_W_DY];
  w.dz = e[LSB200_W_DZ] * e[LSB200_W_DZ];
  w.ridge2 = h->add_ridge ? h->ridge * h->ridge : 0.0;
  return w;
}

static double* owned_ptr(const lsb200_ctx* h, double* v) { return v + (long long)HALO * h->g.plane; }

// choose the z-chunk length of the stencil kernel: fill the SMs in whole waves
template <int XB, int TXT, int TYT>
static void plan_stencil(lsb200_ctx* h, int ctas_per_sm) {
  using C = StencilCfg<XB, TXT, TYT>;
  const int tiles_x = (h->g.nx + C::TX - 1) / C::TX, tiles_y = (h->g.ny + C::TY - 1) / C::TY;
  const long long tiles = (long long)tiles_x * tiles_y;
  const long long slots = (long long)h->sm_count * ctas_per_sm;
  int best_nch = 1;
  double best = -1.0;
  const int max_nch = std::max(1, h->g.nzl / 4);
  for (int nch = 1; nch <= max_nch && nch <= 4096; ++nch) {
    const int cz = (h->g.nzl + nch - 1) / nch;
    const int real = (h->g.nzl + cz - 1) / cz;
    const long long ctas = tiles * real;
    if (ctas > MAX_PARTIALS) break;
    const long long waves = (ctas + slots - 1) / slots;
    const double eff = (double)ctas / (double)(waves * slots) * (double)cz / (double)(cz + 2 * HALO);
    if (eff > best + 1e-9) {
      best = eff;
      best_nch = real;
      h->cz = cz;
    }
  }
  (void)best_nch;
  const int real = (h->g.nzl + h->cz - 1) / h->cz;
  h->sgrid = dim3(tiles_x, tiles_y, real);
}

static void plan_main(lsb200_ctx* h) {
  // the single-launch plan of a grid large enough to have interior box tiles; LSB200_STENCIL_MAIN=-1 keeps k_fdi_normal
  // everywhere (comparison measurements)
  h->fused = false;
  const GridDesc& g = h->g;
  if (h->kind != LSB200_SUPPORT_GRID) return;
  if (getenv("LSB200_STENCIL_MAIN") && atoi(getenv("LSB200_STENCIL_MAIN")) < 0) return;
  auto env_int = [](const char* name) { const char* e = getenv(name); return (e && atoi(e) > 0) ? atoi(e) : 0; };
  if (const char* e = getenv("LSB200_FUSED_MODEL")) sscanf(e, "%lf,%lf,%lf", &FUSED_FACE_COST, &FUSED_XS_REL, &FUSED_YS_REL);
  h->plan_forced = getenv("LSB200_FUSED_MODEL") || env_int("LSB200_FUSED_CFG") || env_int("LSB200_FUSED_CZ") || env_int("LSB200_FUSED_SCZ") ||
                   (getenv("LSB200_AUTOTUNE") && atoi(getenv("LSB200_AUTOTUNE")) == 0);
  h->tuned = false;
  if (!plan_fused(g, h->sm_count, MAX_PARTIALS, env_int("LSB200_FUSED_CFG"), env_int("LSB200_FUSED_CZ"), env_int("LSB200_FUSED_SCZ"),
                  &h->fz))
    return;  // no box: k_fdi_normal everywhere
  if (!h->ztab) {
    std::vector<double> zt((size_t)g.nz * SHELL_ZSTRIDE);
    shell_fill_z_table(g.nz, zt.data());
    if (cudaMalloc(&h->ztab, sizeof(double) * zt.size()) != cudaSuccess) return;
    if (cudaMemcpy(h->ztab, zt.data(), sizeof(double) * zt.size(), cudaMemcpyHostToDevice) != cudaSuccess) return;
  }
  h->want_tma = !(getenv("LSB200_TMA") && atoi(getenv("LSB200_TMA")) == 0);
  h->fused = true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
static bool encode_plane_tiles(const lsb200_ctx* h, const double* vec, const int box_x, const int box_y, TmapArg* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return false;
    }
    encode = (encode_fn)fn;
  }
  const GridDesc& g = h->g;
  static_assert(sizeof(CUtensorMap) == sizeof(TmapArg), "TmapArg carries a CUtensorMap");
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)(g.nzl + 2 * HALO)};
  const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * sizeof(double), (cuuint64_t)g.plane * sizeof(double)};
  const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(vec), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;
  memcpy(out, &tm, sizeof(tm));
  return true;
}

static int launch_stencil(lsb200_ctx* h, const double* p, double* q, int cg_mode) {
  const StencilWeights w = make_weights(h);
  const GridDesc& g = h->g;
  double* part = h->partials;
  unsigned int* cnt = h->counters + 0;
  const bool vec16 = (g.pitch % 2) == 0;
#define LSB_LAUNCH_AT(STREAM, GRID, CZ, SLOT, XB, TXT, TYT, MINB, ST)                                            \
  {                                                                                                              \
    const size_t smem = sizeof(double) * StencilCfg<XB, TXT, TYT, ST>::PLANE * ST;                               \
    double* part_ = h->partials + (size_t)(SLOT) * MAX_PARTIALS;                                                 \
    unsigned int* cnt_ = h->counters + (SLOT);                                                                   \
    if (vec16 && g.nx % 2 == 0) { /* pairs must not straddle the last column: its partner would be row padding */ \
      auto kern = k_fdi_normal<XB, TXT, TYT, MINB, ST, true>;                                                    \
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      kern<<<GRID, dim3(TXT, TYT), smem, STREAM>>>(g, w, p, q, CZ, part_, cnt_, h->sc, cg_mode);                  \
    } else {                                                                                                     \
      auto kern = k_fdi_normal<XB, TXT, TYT, MINB, ST, false>;                                                   \
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      kern<<<GRID, dim3(TXT, TYT), smem, STREAM>>>(g, w, p, q, CZ, part_, cnt_, h->sc, cg_mode);                  \
    }                                                                                                            \
    CKL();                                                                                                       \
    h->launches++;                                                                                               \
  }
#define LSB_LAUNCH(XB, TXT, TYT, MINB, ST) LSB_LAUNCH_AT(h->stream, h->sgrid, h->cz, 0, XB, TXT, TYT, MINB, ST)
  if (h->fused) {
    // one launch: interior box tiles (with their z-face planes), x strips and y strips, sized by plan_fused
    const UniConsts U = make_uni_consts(w);
    const FusedPlan& pl = h->fz.pl;
    const unsigned int ctas = (unsigned int)(pl.n_box + pl.n_xs + pl.n_ys);
    // TMA staging when the vector has a descriptor (built lazily for p and x; needs an even row pitch); LSB200_TMA=0
    // keeps cp.async everywhere
    const bool want_tma = h->want_tma;
    int tmi = -1;
    if (want_tma && vec16) {
      for (int t = 0; t < 2; ++t)
        if (h->tma_vec[t] == p) tmi = t;
      if (tmi < 0 && (p == h->p || p == h->x)) {
        const int t = p == h->p ? 0 : 1;
        bool ok = false;
        switch (h->fz.tyt) {
#define X(TYT, ST, MINB) case TYT: ok = encode_plane_tiles(h, p, MainCfg<4, TYT, ST>::SX, MainCfg<4, TYT, ST>::SY, &h->tma_map[t]); break;
          LSB_FUSED_CFGS(X)
#undef X
        }
        if (ok) {
          h->tma_vec[t] = p;
          tmi = t;
        }
      }
    }
    const TmapArg no_map{};
    switch (h->fz.tyt) {
#define X(TYT, ST, MINB)                                                                                          \
  case TYT: {                                                                                                     \
    const size_t smem = sizeof(double) * MainCfg<4, TYT, ST>::PLANE * ST;                                         \
    if (vec16 && tmi >= 0) {                                                                                      \
      auto kern = k_fdi_fused<TYT, ST, MINB, true, true>;                                                         \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
      kern<<<ctas, dim3(32, TYT), smem, h->stream>>>(g, U, w, pl, p, q, h->ztab, part, cnt, h->sc, cg_mode, h->tma_map[tmi]); \
    } else if (vec16) {                                                                                           \
      auto kern = k_fdi_fused<TYT, ST, MINB, true, false>;                                                        \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
      kern<<<ctas, dim3(32, TYT), smem, h->stream>>>(g, U, w, pl, p, q, h->ztab, part, cnt, h->sc, cg_mode, no_map); \
    } else {                                                                                                      \
      auto kern = k_fdi_fused<TYT, ST, MINB, false, false>;                                                       \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
      kern<<<ctas, dim3(32, TYT), smem, h->stream>>>(g, U, w, pl, p, q, h->ztab, part, cnt, h->sc, cg_mode, no_map); \
    }                                                                                                             \
  } break;
      LSB_FUSED_CFGS(X)
#undef X
      default: return fail(LSB200_ERR_STATE, "no fused stencil configuration %d", h->fz.tyt);
    }
    CKL();
    h->launches++;
    return 0;
  }
  LSB_LAUNCH(2, 32, 8, 2, 4);
#undef LSB_LAUNCH
#undef LSB_LAUNCH_AT
  return 0;
}

static void plan(lsb200_ctx* h) {
  plan_stencil<2, 32, 8>(h, 2);  // k_fdi_normal: small grids
  const char* cz = getenv("LSB200_STENCIL_CZ");
  if (cz && atoi(cz) > 0) {
    h->cz = atoi(cz);
    h->sgrid.z = (h->g.nzl + h->cz - 1) / h->cz;
  }
  plan_main(h);
}

// the regularisation part of the normal operator (writes every owned entry of q, sets pq_part[0])
static int launch_regularisation(lsb200_ctx* h, const double* p, double* q, int cg_mode) {
  if (h->kind == LSB200_SUPPORT_GRID) return launch_stencil(h, p, q, cg_mode);
  if (!h->p1_table) return fail(LSB200_ERR_STATE, "tet-mesh handle without regularisation tables (lsb200_set_p1_tables)");
  const P1Tables tb{{h->p1_m[0], h->p1_m[1], h->p1_m[2]}, h->p1_table, h->p1_off, h->p1_diag};
  const long long n = (long long)h->g.nx * h->g.ny * h->g.nzl;
  const int nb = (int)std::max<long long>(1, std::min<long long>((n + VEC_NT - 1) / VEC_NT, (long long)h->sm_count * 16));
  const double w2 = h->p1_w * h->p1_w, ridge2 = h->add_ridge ? h->ridge * h->ridge : 0.0;
  if (h->p1_split) {
    // two launches side by side: the march of the interior, the boundary shell at full occupancy on a side stream
    const P1MarchPlan& pl = h->p1_splan;
    auto kern = k_p1_march<4, 0, 3>;
    const size_t smem = sizeof(double) * P1MarchCfg<4, 0>::PLANE * P1MarchCfg<4, 0>::ST;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->side[0], h->ev_fork, 0));
    kern<<<(unsigned int)pl.n_cta, 128, smem, h->stream>>>(h->g, tb, h->p1_mtaps, pl, w2, ridge2, p, q, h->partials, h->counters + 0, h->sc, cg_mode);
    CKL();
    const long long nbn = pl.b.nbz + pl.b.nby + pl.b.nbx;
    const int nbb = (int)std::max<long long>(1, std::min<long long>((nbn + 255) / 256, (long long)h->sm_count * 4));
    k_p1_boundary<256><<<nbb, 256, 0, h->side[0]>>>(h->g, tb, pl.b, w2, ridge2, p, q, h->partials + (size_t)3 * MAX_PARTIALS,
                                                    h->counters + 3, h->sc, cg_mode);
    CKL();
    CK(cudaEventRecord(h->ev_join[0], h->side[0]));
    CK(cudaStreamWaitEvent(h->stream, h->ev_join[0], 0));
    h->launches += 2;
    return 0;
  }
  if (h->p1_march) {
    // one launch: the interior by z-marching register tiles, every CTA with its share of the boundary nodes (csrc/p1_march.cuh)
    const P1MarchPlan& pl = h->p1_mplan;
    const unsigned int ctas = (unsigned int)(pl.n_cta + pl.b.n_bnd_cta);
    switch (h->p1_march) {
#define X(ID, NW, NB, MINB)                                                                                 \
  case ID: {                                                                                                \
    auto kern = k_p1_march<NW, NB, MINB>;                                                                   \
    const size_t smem = sizeof(double) * P1MarchCfg<NW, NB>::PLANE * P1MarchCfg<NW, NB>::ST;                \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                     \
    kern<<<ctas, 32 * (NW + NB), smem, h->stream>>>(h->g, tb, h->p1_mtaps, pl, w2, ridge2, p, q, h->partials, h->counters + 0, h->sc, \
                                             cg_mode);                                                      \
  } break;
      LSB_P1_MARCH_CFGS(X)
#undef X
    }
    CKL();
    h->launches++;
    return 0;
  }
  if (h->p1_tiles) {
    // one launch: shared-memory tiles of the interior, every CTA with its share of the boundary nodes (csrc/p1.cuh)
    const P1TilePlan& pl = h->p1_plan;
    const unsigned int ctas = (unsigned int)(pl.n_bnd_cta + pl.tiles_x * pl.tiles_y * pl.tiles_z);
    switch (h->p1_tile_cfg) {
#define X(ID, TY, TZ, NW)                                                                                   \
  case ID: {                                                                                                \
    auto kern = k_p1_tiles<TY, TZ, NW>;                                                                     \
    const size_t smem = sizeof(double) * P1TileGeom<TY, TZ>::ELEMS;                                         \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                     \
    kern<<<ctas, 32 * NW, smem, h->stream>>>(h->g, tb, h->p1_cls, pl, w2, ridge2, p, q, h->partials, h->counters + 0, h->sc, \
                                             cg_mode);                                                      \
  } break;
      LSB_P1_TILE_CFGS(X)
#undef X
    }
    CKL();
    h->launches++;
    return 0;
  }
  k_p1_normal<VEC_NT><<<nb, VEC_NT, 0, h->stream>>>(h->g, tb, w2, ridge2, p, q, h->partials, h->counters + 0, h->sc, cg_mode);
  CKL();
  h->launches++;
  return 0;
}

static int launch_rows(lsb200_ctx* h, const double* p, double* q, int cg_mode) {
  bool first = true;
  const bool det = h->deterministic && h->transposes.size() == h->blocks.size();
  for (size_t bi = 0; bi < h->blocks.size(); ++bi) {
    const RowBlock& rb = h->blocks[bi];
    if (rb.n == 0) continue;
    long long nb = (rb.n + VEC_NT - 1) / VEC_NT;
    if (nb > h->sm_count * 16) nb = h->sm_count * 16;
    if (det) {
      const RowTranspose& tr = h->transposes[bi];
      k_rows_t<VEC_NT><<<(int)nb, VEC_NT, 0, h->stream>>>(h->g, rb, tr, p, h->partials + MAX_PARTIALS, h->counters + 1, h->sc,
                                                          cg_mode, first ? 1 : 0);
      CKL();
      const int ng = std::max(1, std::min((tr.n_nodes + 255) / 256, h->sm_count * 16));
      k_rows_gather<<<ng, 256, 0, h->stream>>>(h->g, rb, tr, q, 0, h->sc, cg_mode);
      CKL();
      h->launches += 2;
    } else {
      double* part = h->partials + MAX_PARTIALS;
      unsigned int* cnt = h->counters + 1;
      const int f = first ? 1 : 0;
      constexpr int RNT = 128;  // one row per thread, 2 K + 10 live values: small CTAs spread a short launch over every SM
      long long rnb = (rb.n + RNT - 1) / RNT;
      if (rnb > h->sm_count * 16) rnb = h->sm_count * 16;
      switch (rb.k) {
        case 4: k_rows_normal<RNT, 4><<<(int)rnb, RNT, 0, h->stream>>>(h->g, rb, p, q, part, cnt, h->sc, cg_mode, f); break;
        case 8: k_rows_normal<RNT, 8><<<(int)rnb, RNT, 0, h->stream>>>(h->g, rb, p, q, part, cnt, h->sc, cg_mode, f); break;
        case 16: k_rows_normal<RNT, 16><<<(int)rnb, RNT, 0, h->stream>>>(h->g, rb, p, q, part, cnt, h->sc, cg_mode, f); break;
        default: k_rows_normal_wide<VEC_NT><<<(int)nb, VEC_NT, 0, h->stream>>>(h->g, rb, p, q, part, cnt, h->sc, cg_mode, f);
      }
      CKL();
      h->launches++;
    }
    first = false;
  }
  return 0;
}

static int halo_exchange(lsb200_ctx* h, double* v, cudaStream_t stream = nullptr) {
  if (h->nranks == 1) return 0;
  if (!stream) stream = h->stream;
  const size_t cnt = (size_t)HALO * h->g.plane;
  double* lo_halo = v;                                                  // planes k0-2,k0-1
  double* lo_own = v + (long long)HALO * h->g.plane;                    // planes k0,k0+1
  double* hi_own = v + (long long)h->g.nzl * h->g.plane;                // planes k1-2,k1-1
  double* hi_halo = v + (long long)(h->g.nzl + HALO) * h->g.plane;      // planes k1,k1+1
  NCK(nccl::GroupStart());
  if (h->rank > 0) {
    NCK(nccl::Send(lo_own, cnt, nccl::Float64, h->rank - 1, h->comm, stream));
    NCK(nccl::Recv(lo_halo, cnt, nccl::Float64, h->rank - 1, h->comm, stream));
  }
  if (h->rank < h->nranks - 1) {
    NCK(nccl::Send(hi_own, cnt, nccl::Float64, h->rank + 1, h->comm, stream));
    NCK(nccl::Recv(hi_halo, cnt, nccl::Float64, h->rank + 1, h->comm, stream));
  }
  NCK(nccl::GroupEnd());
  return 0;
}

static int allreduce_scalar(lsb200_ctx* h, double* dev) {
  if (h->nranks == 1) return 0;
  NCK(nccl::AllReduce(dev, dev, 1, nccl::Float64, nccl::Sum, h->comm, h->stream));
  return 0;
}

static void pool_free(lsb200_ctx* h, void* p) {
  if (p) cudaFreeAsync(p, h->stream);
}

static int free_block(lsb200_ctx* h, RowBlock& rb) {
  pool_free(h, rb.vals);
  pool_free(h, rb.cols);
  pool_free(h, rb.w2);
  pool_free(h, rb.b);
  rb = RowBlock{};
  return 0;
}

static int alloc_block(lsb200_ctx* h, RowBlock& rb, long long n, int k) {
  rb.n = n;
  rb.k = k;
  const long long m = std::max<long long>(n, 1);
  CK(cudaMallocAsync(&rb.vals, sizeof(double) * m * k, h->stream));
  CK(cudaMallocAsync(&rb.cols, sizeof(int) * m * k, h->stream));
  CK(cudaMallocAsync(&rb.w2, sizeof(double) * m, h->stream));
  CK(cudaMallocAsync(&rb.b, sizeof(double) * m, h->stream));
  return 0;
}

static int ensure_vectors(lsb200_ctx* h, const bool solution_only = false) {
  // One zeroed guard plane in front of and behind the nzl + 4 stored planes: the gather kernels read the in-plane taps
  // of the first / last stored plane without clamping -- a tap outside the grid has coefficient 0 but is still loaded,
  // up to 2 rows + 2 nodes away.  On a single-handle grid those loads fall into the halo planes; on a z slab the first /
  // last stored plane IS a halo plane, and they would leave the allocation.
  // solution_only: a handle that only evaluates a field set by lsb200_set_c(_dev) (the replicated solution of a slab
  // solve, a fault-frame coordinate) holds x alone -- one node vector instead of seven.
  const size_t guard = ((size_t)h->g.plane + 31) & ~(size_t)31;  // keeps the 256-byte alignment of the base
  h->guard = (long long)guard;
  const size_t bytes = sizeof(double) * ((size_t)h->g.nall + 2 * guard);
  double** v[] = {&h->x, &h->r, &h->p, &h->q, &h->rhs, &h->diag, &h->s2};
  for (double** pv : v) {
    if (*pv) continue;
    if (solution_only && pv != &h->x) continue;
    double* raw = nullptr;
    cudaError_t e = cudaMalloc(&raw, bytes);
    if (e != cudaSuccess) return fail(LSB200_ERR_ALLOC, "cudaMalloc of a node vector (%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    CK(cudaMemsetAsync(raw, 0, bytes, h->stream));
    *pv = raw + guard;
  }
  return 0;
}

// host (nx*ny*nzl, unpitched) <-> device (pitched, halo planes) copies of the owned planes
static int copy_in(lsb200_ctx* h, double* dev, const double* host) {
  CK(cudaMemcpy2DAsync(owned_ptr(h, dev), sizeof(double) * h->g.pitch, host, sizeof(double) * h->g.nx,
                       sizeof(double) * h->g.nx, (size_t)h->g.ny * h->g.nzl, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
static int copy_out(lsb200_ctx* h, double* host, const double* dev) {
  CK(cudaMemcpy2DAsync(host, sizeof(double) * h->g.nx, owned_ptr(h, const_cast<double*>(dev)), sizeof(double) * h->g.pitch,
                       sizeof(double) * h->g.nx, (size_t)h->g.ny * h->g.nzl, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// ------------------------------------------------------------------ API
extern "C" {

int lsb200_abi_version(void) { return LSB200_ABI_VERSION; }
const char* lsb200_last_error(void) { return g_err; }
int lsb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static int create_common(int kind, const int64_t nsteps[3], const double origin[3], const double step[3], int64_t k0,
                         int64_t k1, int device, lsb200_handle* out) {
  if (!out || !nsteps || !origin || !step) return fail(LSB200_ERR_ARG, "null argument");
  *out = nullptr;
  for (int a = 0; a < 3; ++a) {
    if (nsteps[a] < 2) return fail(LSB200_ERR_ARG, "nsteps[%d]=%lld: need at least 2 nodes per axis", a, (long long)nsteps[a]);
    if (!(step[a] > 0.0)) return fail(LSB200_ERR_ARG, "step[%d] must be positive", a);
  }
  if (k0 < 0 || k1 > nsteps[2] || k1 <= k0) return fail(LSB200_ERR_ARG, "bad slab [%lld,%lld) of %lld planes", (long long)k0, (long long)k1, (long long)nsteps[2]);
  int ndev = lsb200_device_count();
  if (ndev <= 0) return fail(LSB200_ERR_CUDA, "no CUDA device available (this engine has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(LSB200_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  const long long plane = (long long)((nsteps[0] + 1) & ~1LL) * nsteps[1];
  if (plane * (k1 - k0 + 2 * HALO) >= 2147483647LL)
    return fail(LSB200_ERR_ARG, "local slab has %lld node slots; the limit per handle is 2^31-1 (use more z-slabs)", plane * (k1 - k0 + 2 * HALO));
  CK(cudaSetDevice(device));
  lsb200_ctx* h = new lsb200_ctx();
  h->kind = kind;
  h->device = device;
  GridDesc& g = h->g;
  g.nx = (int)nsteps[0]; g.ny = (int)nsteps[1]; g.nz = (int)nsteps[2];
  g.k0 = (int)k0; g.k1 = (int)k1; g.nzl = (int)(k1 - k0);
  g.pitch = (g.nx + 1) & ~1;  // even: 16-byte copies and the TMA descriptors need rows that start on 16 bytes
  g.plane = (long long)g.pitch * g.ny;
  g.nloc = g.plane * g.nzl;
  g.nall = g.plane * (g.nzl + 2 * HALO);
  for (int a = 0; a < 3; ++a) { g.origin[a] = origin[a]; g.step[a] = step[a]; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    // transient buffers and row blocks come from the stream-ordered pool; keep what is freed (cudaMalloc / cudaFree
    // cost 20-400 ms per set-up when they map and unmap memory, profiles/r01z)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  for (int t = 0; t < 3; ++t) {
    CK(cudaStreamCreateWithFlags(&h->side[t], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_join[t], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  { const char* e = getenv("LSB200_OVERLAP"); h->overlap_shells = !(e && atoi(e) == 0); }
  CK(cudaMalloc(&h->sc, sizeof(CgScalars)));
  CK(cudaMemset(h->sc, 0, sizeof(CgScalars)));
  CK(cudaMallocHost(&h->sc_host, sizeof(CgScalars)));
  CK(cudaMalloc(&h->partials, sizeof(double) * 6 * MAX_PARTIALS));
  CK(cudaMalloc(&h->counters, sizeof(unsigned int) * 8));
  CK(cudaMemset(h->counters, 0, sizeof(unsigned int) * 8));
  CK(cudaMalloc(&h->status, sizeof(unsigned long long) * 4));
  plan(h);
  *out = h;
  return 0;
}

int lsb200_create_grid(const int64_t nsteps[3], const double origin[3], const double step[3], int64_t k0, int64_t k1,
                       int device, lsb200_handle* out) {
  return create_common(LSB200_SUPPORT_GRID, nsteps, origin, step, k0, k1, device, out);
}

int lsb200_create_tetmesh(const int64_t nsteps[3], const double origin[3], const double step[3], int device,
                          lsb200_handle* out) {
  if (!nsteps) return fail(LSB200_ERR_ARG, "null argument");
  return create_common(LSB200_SUPPORT_TETMESH, nsteps, origin, step, 0, nsteps[2], device, out);
}

int lsb200_create_tetmesh_slab(const int64_t nsteps[3], const double origin[3], const double step[3], int64_t k0,
                               int64_t k1, int device, lsb200_handle* out) {
  if (!nsteps) return fail(LSB200_ERR_ARG, "null argument");
  return create_common(LSB200_SUPPORT_TETMESH, nsteps, origin, step, k0, k1, device, out);
}

int lsb200_destroy(lsb200_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (RowBlock& rb : h->pending) free_block(h, rb);
  for (RowBlock& rb : h->blocks) free_block(h, rb);
  free_transposes(h);
  double* v[] = {h->x, h->r, h->p, h->q, h->rhs, h->diag, h->s2};
  for (double* pv : v)
    if (pv) cudaFree(pv - h->guard);  // allocated with a guard plane in front, see ensure_vectors
  cudaFree(h->sc);
  cudaFreeHost(h->sc_host);
  cudaFree(h->partials);
  cudaFree(h->counters);
  cudaFree(h->status);
  if (h->cg_graph) cudaGraphExecDestroy(h->cg_graph);
  cudaFree(h->small_bar);
  cudaFree(h->small_slots);
  cudaFree(h->p1_table);
  cudaFree(h->p1_off);
  cudaFree(h->p1_diag);
  cudaFree(h->ztab);
  for (int r = 0; r < h->peer_host.nranks; ++r)
    if (h->peers && r != h->rank && h->peer_host.box[r]) cudaIpcCloseMemHandle(h->peer_host.box[r]);
  for (int side = 0; side < 2; ++side)
    if (h->nb_p_raw[side]) cudaIpcCloseMemHandle(h->nb_p_raw[side]);
  cudaFree(h->peers);
  cudaFree(h->mailbox);
  if (h->comm) nccl::CommDestroy(h->comm);
  h->pipe.destroy();
  for (int t = 0; t < 3; ++t) {
    if (h->side[t]) cudaStreamDestroy(h->side[t]);
    if (h->ev_join[t]) cudaEventDestroy(h->ev_join[t]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int64_t lsb200_local_dof(lsb200_handle h) { return h ? (int64_t)h->g.nx * h->g.ny * h->g.nzl : 0; }
int lsb200_support_kind(lsb200_handle h) { return h ? h->kind : -1; }

int lsb200_set_regularisation(lsb200_handle h, const double w[9]) {
  if (!h || !w) return fail(LSB200_ERR_ARG, "null argument");
  if (h->kind != LSB200_SUPPORT_GRID) return fail(LSB200_ERR_STATE, "stencil regularisation needs a grid handle");
  for (int i = 0; i < 9; ++i) {
    if (!(w[i] == w[i])) return fail(LSB200_ERR_ARG, "regularisation weight %d is NaN", i);
    h->ew[i] = w[i];
  }
  h->finalized = false;
  return 0;
}

int lsb200_set_p1_regularisation(lsb200_handle h, double w) {
  if (!h) return fail(LSB200_ERR_ARG, "null argument");
  if (h->kind != LSB200_SUPPORT_TETMESH) return fail(LSB200_ERR_STATE, "face-jump regularisation needs a tet-mesh handle");
  h->p1_w = w;
  h->finalized = false;
  return 0;
}

int lsb200_set_p1_tables(lsb200_handle h, const int64_t m[3], const double* table) {
  if (!h || !m || !table) return fail(LSB200_ERR_ARG, "null argument");
  if (h->kind != LSB200_SUPPORT_TETMESH) return fail(LSB200_ERR_STATE, "regularisation tables need a tet-mesh handle");
  const int nc[3] = {h->g.nx - 1, h->g.ny - 1, h->g.nz - 1};
  for (int a = 0; a < 3; ++a) {
    const bool ok = (nc[a] <= 6) ? (m[a] == nc[a]) : ((m[a] == 5 || m[a] == 6) && ((m[a] - nc[a]) % 2 == 0));
    if (!ok) return fail(LSB200_ERR_ARG, "axis %d: a small mesh of %lld cells cannot represent %d cells", a, (long long)m[a], nc[a]);
  }
  CK(cudaSetDevice(h->device));
  std::vector<double> coef, diag;
  std::vector<int> off;
  size_t bad = 0;
  const long long mm[3] = {(long long)m[0], (long long)m[1], (long long)m[2]};
  const int crc = p1_compact_tables(h->g, mm, table, coef, off, diag, &bad);
  if (crc == 1) return fail(LSB200_ERR_ARG, "NaN in the regularisation table");
  if (crc == 2) return fail(LSB200_ERR_ARG, "table row %zu has more than %d non-zero taps", bad, P1_TAPS);
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(h->p1_table); cudaFree(h->p1_off); cudaFree(h->p1_diag);
  h->p1_table = nullptr; h->p1_off = nullptr; h->p1_diag = nullptr;
  CK(cudaMalloc(&h->p1_table, sizeof(double) * coef.size()));
  CK(cudaMalloc(&h->p1_off, sizeof(int) * off.size()));
  CK(cudaMalloc(&h->p1_diag, sizeof(double) * diag.size()));
  CK(cudaMemcpy(h->p1_table, coef.data(), sizeof(double) * coef.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->p1_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->p1_diag, diag.data(), sizeof(double) * diag.size(), cudaMemcpyHostToDevice));
  for (int a = 0; a < 3; ++a) h->p1_m[a] = (int)m[a];
  // tile plan of the interior (LSB200_P1_TILES=0: the generic kernel everywhere; LSB200_P1_TILE_CFG: tile shape)
  const bool want_tiles = !(getenv("LSB200_P1_TILES") && atoi(getenv("LSB200_P1_TILES")) == 0);
  const int tile_cfg = getenv("LSB200_P1_TILE_CFG") ? atoi(getenv("LSB200_P1_TILE_CFG")) : 1;
  h->p1_tiles = false;
  if (want_tiles) {
    switch (tile_cfg) {
#define X(ID, TY, TZ, NW)                                                                                          \
  case ID:                                                                                                         \
    h->p1_tiles = p1_plan_tiles(h->g, h->p1_m, table, TY, TZ, 32 * NW, MAX_PARTIALS, &h->p1_plan, &h->p1_cls);     \
    h->p1_tile_cfg = ID;                                                                                           \
    break;
      LSB_P1_TILE_CFGS(X)
#undef X
    }
  }
  // the z-marching form of the interior (LSB200_P1_MARCH = configuration of LSB_P1_MARCH_CFGS, 0: off; LSB200_P1_MARCH_CZ: planes per chunk)
  h->p1_march = 0;
  const int march = getenv("LSB200_P1_MARCH") ? atoi(getenv("LSB200_P1_MARCH")) : LSB_P1_MARCH_DEFAULT;
  const int march_cz = getenv("LSB200_P1_MARCH_CZ") ? atoi(getenv("LSB200_P1_MARCH_CZ")) : 0;
  if (want_tiles && h->p1_tiles) {
    switch (march) {
#define X(ID, NW, NB, MINB)                                                                                          \
  case ID: {                                                                                                         \
    int per_sm = 0;                                                                                                  \
    const size_t smem = sizeof(double) * P1MarchCfg<NW, NB>::PLANE * P1MarchCfg<NW, NB>::ST;                         \
    cudaFuncSetAttribute(k_p1_march<NW, NB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_p1_march<NW, NB, MINB>, 32 * (NW + NB), smem) != cudaSuccess || per_sm < 1) { \
      cudaGetLastError();                                                                                            \
      break;                                                                                                         \
    }                                                                                                                \
    {                                                                                                                \
      /* the persistent CG loop walks the same plan: one wave there too */                                          \
      int per_sm_loop = 0;                                                                                           \
      cudaFuncSetAttribute(k_cg_small_p1m<NW, NB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_loop, k_cg_small_p1m<NW, NB, MINB>, 32 * (NW + NB), smem) == cudaSuccess && \
          per_sm_loop >= 1)                                                                                          \
        per_sm = std::min(per_sm, per_sm_loop);                                                                      \
      else                                                                                                           \
        cudaGetLastError();                                                                                          \
    }                                                                                                                \
    if (p1_plan_march<NW, NB>(h->g, h->p1_m, table, h->sm_count, per_sm, MAX_PARTIALS, march_cz, &h->p1_mplan, &h->p1_mtaps)) { \
      h->p1_march = ID;                                                                                              \
      h->p1_march_nw = NW;                                                                                           \
      h->p1_march_nb = NB;                                                                                           \
      h->p1_march_minb = per_sm;                                                                                     \
    }                                                                                                                \
  } break;
      LSB_P1_MARCH_CFGS(X)
#undef X
    }
  }
  // outside the persistent loop: the march and the boundary shell as two launches side by side (LSB200_P1_SPLIT=0: one)
  h->p1_split = false;
  if (h->p1_march && !(getenv("LSB200_P1_SPLIT") && atoi(getenv("LSB200_P1_SPLIT")) == 0)) {
    int per_sm = 0;
    const size_t smem = sizeof(double) * P1MarchCfg<4, 0>::PLANE * P1MarchCfg<4, 0>::ST;
    cudaFuncSetAttribute(k_p1_march<4, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_p1_march<4, 0, 3>, 128, smem) == cudaSuccess && per_sm >= 1) {
      // two marching CTAs per SM: the third CTA's registers are the boundary kernel's
      P1MarchTaps same;
      h->p1_split = p1_plan_march<4, 0>(h->g, h->p1_m, table, h->sm_count, std::min(per_sm, 2), MAX_PARTIALS, march_cz, &h->p1_splan,
                                        &same, true) && h->p1_splan.n_cta > 0;  // (a slab without interior planes: one launch)
    } else {
      cudaGetLastError();
    }
  }
  h->finalized = false;
  return 0;
}

int lsb200_set_options(lsb200_handle h, double ridge_factor, int add_ridge, int apply_scaling, int deterministic) {
  if (!h) return fail(LSB200_ERR_ARG, "null argument");
  h->ridge = ridge_factor;
  h->add_ridge = add_ridge;
  h->apply_scaling = apply_scaling;
  h->deterministic = deterministic;
  h->finalized = false;
  return 0;
}

static int width_class(int k) {
  int c = 4;
  while (c < k) c <<= 1;
  return c;
}

int lsb200_add_rows(lsb200_handle h, int64_t n_rows, int k, const double* vals, const int64_t* cols, const double* b,
                    const double* w, int normalise) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (n_rows < 0 || k < 1 || k > 64) return fail(LSB200_ERR_ARG, "bad row block shape (%lld,%d)", (long long)n_rows, k);
  if (n_rows == 0) return 0;
  if (!vals || !cols || !b || !w) return fail(LSB200_ERR_ARG, "null row arrays");
  CK(cudaSetDevice(h->device));
  double *dv = nullptr, *db = nullptr, *dw = nullptr;
  long long* dc = nullptr;
  RowBlock rb{};
  auto release = [&]() { pool_free(h, dv); pool_free(h, dc); pool_free(h, db); pool_free(h, dw); };
  auto stage = [&]() -> int {
    CK(cudaMallocAsync(&dv, sizeof(double) * n_rows * k, h->stream));
    CK(cudaMallocAsync(&dc, sizeof(long long) * n_rows * k, h->stream));
    CK(cudaMallocAsync(&db, sizeof(double) * n_rows, h->stream));
    CK(cudaMallocAsync(&dw, sizeof(double) * n_rows, h->stream));
    CK(cudaMemcpyAsync(dv, vals, sizeof(double) * n_rows * k, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dc, cols, sizeof(long long) * n_rows * k, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(db, b, sizeof(double) * n_rows, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dw, w, sizeof(double) * n_rows, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
    int rc = alloc_block(h, rb, n_rows, width_class(k));
    if (rc) return rc;
    CK(cudaMemsetAsync(rb.cols, 0xff, sizeof(int) * rb.n * rb.k, h->stream));
    CK(cudaMemsetAsync(rb.vals, 0, sizeof(double) * rb.n * rb.k, h->stream));
    const int nb = (int)std::min<long long>((n_rows + 255) / 256, 4096);
    k_rows_ingest<<<nb, 256, 0, h->stream>>>(h->g, n_rows, k, dv, dc, db, dw, normalise, rb, 0, h->status);
    CKL();
    h->launches++;
    return 0;
  };
  unsigned long long st[4] = {0, 0, 0, 0};
  int rc = stage();
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(LSB200_ERR_CUDA, "row ingestion failed: %s", cudaGetErrorString(e));
  }
  release();
  if (rc) {
    free_block(h, rb);
    return rc;
  }
  if (st[2]) {
    free_block(h, rb);
    return fail(LSB200_ERR_ARG, "column id %lld outside [0,%lld)", (long long)st[3], (long long)h->g.nx * h->g.ny * h->g.nz);
  }
  if (st[1]) {
    free_block(h, rb);
    return fail(LSB200_ERR_ARG, "a row touching the owned planes [%d,%d) has a column outside the halo planes", h->g.k0, h->g.k1);
  }
  h->rows_kept += (long long)st[0];
  h->pending.push_back(rb);
  h->finalized = false;
  return 0;
}

int lsb200_add_point_rows(lsb200_handle h, int kind, int64_t n, const double* data, int ncols, double weight,
                          int64_t* n_rows_out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (n_rows_out) *n_rows_out = 0;
  if (n < 0) return fail(LSB200_ERR_ARG, "negative point count");
  const int need = kind == LSB200_ROWS_VALUE ? 5 : 7;
  if (kind < LSB200_ROWS_VALUE || kind > LSB200_ROWS_ELEMENT_VECTOR) return fail(LSB200_ERR_ARG, "unknown row kind %d", kind);
  if (kind == LSB200_ROWS_ELEMENT_VECTOR && h->kind != LSB200_SUPPORT_TETMESH)
    return fail(LSB200_ERR_STATE, "element-vector rows need a tet-mesh handle");
  if (ncols < need) return fail(LSB200_ERR_ARG, "row kind %d needs %d data columns, got %d", kind, need, ncols);
  if (n == 0) return 0;
  if (!data) return fail(LSB200_ERR_ARG, "null data");
  CK(cudaSetDevice(h->device));
  const int mult = kind == LSB200_ROWS_GRADIENT ? 2 : (kind == LSB200_ROWS_NORMAL ? 3 : (kind == LSB200_ROWS_VALUE ? 1 : 1));
  double* dd = nullptr;
  CK(cudaMallocAsync(&dd, sizeof(double) * n * ncols, h->stream));
  CK(cudaMemcpyAsync(dd, data, sizeof(double) * n * ncols, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
  RowBlock rb{};
  const int k = h->kind == LSB200_SUPPORT_GRID ? 8 : 4;
  int rc = alloc_block(h, rb, n * mult, k);
  if (rc) return rc;
  const int nb = (int)std::min<long long>((n + 127) / 128, 4096);
  if (h->kind == LSB200_SUPPORT_GRID) {
    k_fdi_point_rows<<<nb, 128, 0, h->stream>>>(h->g, kind, n, dd, ncols, weight, rb, 0, h->status);
  } else {
    if (kind == LSB200_ROWS_GRADIENT) {
      // P1Interpolator.add_gradient_constraints is a no-op (LI/_p1interpolator.py:70-71)
      pool_free(h, dd);
      free_block(h, rb);
      return 0;
    }
    k_p1_point_rows<<<nb, 128, 0, h->stream>>>(h->g, kind, n, dd, ncols, weight, rb, 0, h->status);
  }
  CKL();
  h->launches++;
  unsigned long long st[2];
  CK(cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  pool_free(h, dd);
  if (n_rows_out) *n_rows_out = (int64_t)st[0];
  h->rows_kept += (long long)st[0];
  h->pending.push_back(rb);
  h->finalized = false;
  return 0;
}

int lsb200_add_face_vector_rows(lsb200_handle h, int64_t n, const int64_t* nodes_a, const int64_t* nodes_b,
                                const double* vectors, double weight, int64_t* n_rows_out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (n_rows_out) *n_rows_out = 0;
  if (n < 0) return fail(LSB200_ERR_ARG, "negative face count");
  if (h->kind != LSB200_SUPPORT_TETMESH) return fail(LSB200_ERR_STATE, "face-vector rows need a tet-mesh handle");
  if (n == 0) return 0;
  if (!nodes_a || !nodes_b || !vectors) return fail(LSB200_ERR_ARG, "null face arrays");
  CK(cudaSetDevice(h->device));
  long long *da = nullptr, *dbn = nullptr, *dc = nullptr;
  double *dvec = nullptr, *dv = nullptr, *db = nullptr, *dw = nullptr;
  RowBlock rb{};
  auto release = [&]() { pool_free(h, da); pool_free(h, dbn); pool_free(h, dc); pool_free(h, dvec); pool_free(h, dv); pool_free(h, db); pool_free(h, dw); };
  auto stage = [&]() -> int {
    CK(cudaMallocAsync(&da, sizeof(long long) * 4 * n, h->stream));
    CK(cudaMallocAsync(&dbn, sizeof(long long) * 4 * n, h->stream));
    CK(cudaMallocAsync(&dvec, sizeof(double) * 3 * n, h->stream));
    CK(cudaMallocAsync(&dv, sizeof(double) * 8 * n, h->stream));
    CK(cudaMallocAsync(&dc, sizeof(long long) * 8 * n, h->stream));
    CK(cudaMallocAsync(&db, sizeof(double) * n, h->stream));
    CK(cudaMallocAsync(&dw, sizeof(double) * n, h->stream));
    CK(cudaMemcpyAsync(da, nodes_a, sizeof(long long) * 4 * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dbn, nodes_b, sizeof(long long) * 4 * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dvec, vectors, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
    const int nb = (int)std::min<long long>((n + 255) / 256, 4096);
    k_p1_face_vector_vals<<<nb, 256, 0, h->stream>>>(h->g, n, da, dbn, dvec, weight, dv, dc, db, dw);
    CKL();
    int rc = alloc_block(h, rb, n, width_class(8));
    if (rc) return rc;
    CK(cudaMemsetAsync(rb.cols, 0xff, sizeof(int) * rb.n * rb.k, h->stream));
    CK(cudaMemsetAsync(rb.vals, 0, sizeof(double) * rb.n * rb.k, h->stream));
    k_rows_ingest<<<nb, 256, 0, h->stream>>>(h->g, n, 8, dv, dc, db, dw, 1, rb, 0, h->status);
    CKL();
    h->launches += 2;
    return 0;
  };
  unsigned long long st[4] = {0, 0, 0, 0};
  int rc = stage();
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(LSB200_ERR_CUDA, "face-vector rows failed: %s", cudaGetErrorString(e));
  }
  release();
  if (rc) {
    free_block(h, rb);
    return rc;
  }
  if (st[2]) {
    free_block(h, rb);
    return fail(LSB200_ERR_ARG, "node id %lld outside [0,%lld)", (long long)st[3], (long long)h->g.nx * h->g.ny * h->g.nz);
  }
  if (st[1]) {
    free_block(h, rb);
    return fail(LSB200_ERR_ARG, "a face touching the owned planes [%d,%d) has a node outside the halo planes", h->g.k0, h->g.k1);
  }
  if (n_rows_out) *n_rows_out = (int64_t)st[0];
  h->rows_kept += (long long)st[0];
  h->pending.push_back(rb);
  h->finalized = false;
  return 0;
}

int lsb200_add_directional_rows(lsb200_handle h, const double* vectors, double weight, int64_t* n_rows_out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (n_rows_out) *n_rows_out = 0;
  if (h->kind != LSB200_SUPPORT_GRID) return fail(LSB200_ERR_STATE, "directional stencil rows need a grid handle");
  if (!vectors) return fail(LSB200_ERR_ARG, "null vectors");
  const GridDesc& g = h->g;
  const int ka = std::max(1, g.k0 - 1), kb = std::min(g.nz - 1, g.k1 + 1);
  if (g.nx < 3 || g.ny < 3 || kb <= ka) return 0;  // no node is interior on all axes
  CK(cudaSetDevice(h->device));
  const long long per_plane = (long long)g.nx * g.ny;
  const long long ncent = (long long)(g.nx - 2) * (g.ny - 2) * (kb - ka);
  double* dv = nullptr;
  const size_t bytes = sizeof(double) * 3 * (size_t)per_plane * (size_t)(kb - ka);
  CK(cudaMallocAsync(&dv, bytes, h->stream));
  CK(cudaMemcpyAsync(dv, vectors + 3 * per_plane * ka, bytes, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
  RowBlock rb{};
  int rc = alloc_block(h, rb, 6 * ncent, 4);
  if (rc) return rc;
  const int nb = (int)std::min<long long>((ncent + 255) / 256, 8192);
  k_fdi_directional_rows<<<nb, 256, 0, h->stream>>>(g, ka, kb, dv, weight, rb, h->status);
  CKL();
  h->launches++;
  unsigned long long st[2];
  CK(cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  pool_free(h, dv);
  if (n_rows_out) *n_rows_out = (int64_t)st[0];
  h->rows_kept += (long long)st[0];
  h->pending.push_back(rb);
  h->finalized = false;
  return 0;
}

int lsb200_add_pair_rows(lsb200_handle h, int64_t n, const double* xyz_i, const double* xyz_j, double weight,
                         int64_t* n_rows_out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (n_rows_out) *n_rows_out = 0;
  if (n < 0) return fail(LSB200_ERR_ARG, "negative pair count");
  if (h->kind != LSB200_SUPPORT_GRID) return fail(LSB200_ERR_STATE, "interface rows need a grid handle");
  if (n == 0) return 0;
  if (!xyz_i || !xyz_j) return fail(LSB200_ERR_ARG, "null coordinates");
  CK(cudaSetDevice(h->device));
  double *di = nullptr, *dj = nullptr;
  CK(cudaMallocAsync(&di, sizeof(double) * 3 * n, h->stream));
  CK(cudaMallocAsync(&dj, sizeof(double) * 3 * n, h->stream));
  CK(cudaMemcpyAsync(di, xyz_i, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(dj, xyz_j, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
  RowBlock rb{};
  int rc = alloc_block(h, rb, n, 16);
  if (rc) return rc;
  const int nb = (int)std::min<long long>((n + 127) / 128, 4096);
  k_fdi_pair_rows<<<nb, 128, 0, h->stream>>>(h->g, n, di, dj, weight, rb, h->status);
  CKL();
  h->launches++;
  unsigned long long st[2];
  CK(cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  pool_free(h, di);
  pool_free(h, dj);
  if (st[1]) {
    free_block(h, rb);
    return fail(LSB200_ERR_ARG, "an interface pair spans planes outside the owned + halo planes [%d,%d) of this slab", h->g.k0 - HALO, h->g.k1 + HALO);
  }
  if (n_rows_out) *n_rows_out = (int64_t)st[0];
  h->rows_kept += (long long)st[0];
  h->pending.push_back(rb);
  h->finalized = false;
  return 0;
}

static int make_frame_field(const lsb200_ctx* h, const double* t, const char* what, int f, FrameField* F);  // below, frame section

int lsb200_add_field_orthogonal_rows(lsb200_handle h, lsb200_handle other, const double* other_transform,
                                     int64_t element_step, double weight, double b, int64_t* n_rows_out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (n_rows_out) *n_rows_out = 0;
  if (element_step < 1) return fail(LSB200_ERR_ARG, "element_step must be >= 1");
  if (other == h) return fail(LSB200_ERR_ARG, "a field cannot be made orthogonal to itself");
  FrameField F;
  int rc = make_frame_field(other, other_transform, "other field", 0, &F);
  if (rc) return rc;
  if (other->device != h->device) return fail(LSB200_ERR_ARG, "the two handles live on different devices");
  CK(cudaSetDevice(h->device));
  if (h->kind == LSB200_SUPPORT_TETMESH) {
    // P1 twin: barycentres of the tets e = 0, step, 2 step, ... and the other field's normalised gradient there, made on
    // the device (k_p1_barycentre_vectors), then the tangent family of k_p1_point_rows (weight w * tet volume)
    const long long n_el = 5LL * (h->g.nx - 1) * (h->g.ny - 1) * (h->g.nz - 1);
    const long long n = (n_el + element_step - 1) / element_step;
    if (n == 0) return 0;
    double* dd = nullptr;
    CK(cudaMallocAsync(&dd, sizeof(double) * 7 * n, h->stream));
    CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
    RowBlock rb{};
    rc = alloc_block(h, rb, n, 4);
    if (rc) return rc;
    const int nb = (int)std::min<long long>((n + 127) / 128, 4096);
    if (F.kind == LSB200_SUPPORT_TETMESH) k_p1_barycentre_vectors<true><<<nb, 128, 0, h->stream>>>(h->g, F, (long long)element_step, n, b, dd, h->status);
    else k_p1_barycentre_vectors<false><<<nb, 128, 0, h->stream>>>(h->g, F, (long long)element_step, n, b, dd, h->status);
    CKL();
    k_p1_point_rows<<<nb, 128, 0, h->stream>>>(h->g, LSB200_ROWS_TANGENT, n, dd, 7, weight, rb, 0, h->status);
    CKL();
    h->launches += 2;
    unsigned long long st[3];
    CK(cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    pool_free(h, dd);
    if (st[2]) {  // a NaN among the vectors: add_constraints_to_least_squares drops the WHOLE family (LI/_discrete_interpolator.py:357-359)
      free_block(h, rb);
      return 0;
    }
    if (n_rows_out) *n_rows_out = (int64_t)st[0];
    h->rows_kept += (long long)st[0];
    h->pending.push_back(rb);
    h->finalized = false;
    return 0;
  }
  const long long n_el = (long long)(h->g.nx - 1) * (h->g.ny - 1) * (h->g.nz - 1);
  const long long n = (n_el + element_step - 1) / element_step;  // element_idx[::step]
  if (n == 0) return 0;
  CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
  RowBlock rb{};
  rc = alloc_block(h, rb, n, 8);
  if (rc) return rc;
  const int nb = (int)std::min<long long>((n + 127) / 128, 4096);
  if (F.kind == LSB200_SUPPORT_TETMESH)
    k_fdi_field_orthogonal_rows<true><<<nb, 128, 0, h->stream>>>(h->g, F, (long long)element_step, n, weight, b, rb, 0, h->status);
  else
    k_fdi_field_orthogonal_rows<false><<<nb, 128, 0, h->stream>>>(h->g, F, (long long)element_step, n, weight, b, rb, 0, h->status);
  CKL();
  h->launches++;
  unsigned long long st[3];
  CK(cudaMemcpyAsync(st, h->status, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (st[2]) {  // a NaN gradient (outside a tet-mesh source): the reference drops the whole family
    free_block(h, rb);
    return 0;
  }
  if (n_rows_out) *n_rows_out = (int64_t)st[0];
  h->rows_kept += (long long)st[0];
  h->pending.push_back(rb);
  h->finalized = false;
  return 0;
}

int lsb200_clear_rows(lsb200_handle h) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  for (RowBlock& rb : h->pending) free_block(h, rb);
  for (RowBlock& rb : h->blocks) free_block(h, rb);
  free_transposes(h);
  h->pending.clear();
  h->blocks.clear();
  h->rows_kept = 0;
  h->finalized = false;
  return 0;
}

int64_t lsb200_num_rows(lsb200_handle h) { return h ? h->rows_kept : 0; }

// merge pending blocks into one ELL block per width class
static int merge_blocks(lsb200_ctx* h) {
  if (h->pending.empty()) return 0;
  std::vector<RowBlock> all = h->blocks;
  all.insert(all.end(), h->pending.begin(), h->pending.end());
  h->blocks.clear();
  h->pending.clear();
  std::vector<int> widths;
  for (const RowBlock& rb : all)
    if (std::find(widths.begin(), widths.end(), rb.k) == widths.end()) widths.push_back(rb.k);
  std::sort(widths.begin(), widths.end());
  for (int k : widths) {
    long long n = 0;
    int cnt = 0;
    for (const RowBlock& rb : all)
      if (rb.k == k) { n += rb.n; cnt++; }
    if (cnt == 1) {
      for (RowBlock& rb : all)
        if (rb.k == k) { h->blocks.push_back(rb); rb = RowBlock{}; }
      continue;
    }
    RowBlock m{};
    int rc = alloc_block(h, m, n, k);
    if (rc) return rc;
    long long off = 0;
    for (RowBlock& rb : all) {
      if (rb.k != k) continue;
      for (int j = 0; j < k; ++j) {
        CK(cudaMemcpyAsync(m.vals + j * m.n + off, rb.vals + j * rb.n, sizeof(double) * rb.n, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(m.cols + j * m.n + off, rb.cols + j * rb.n, sizeof(int) * rb.n, cudaMemcpyDeviceToDevice, h->stream));
      }
      CK(cudaMemcpyAsync(m.w2 + off, rb.w2, sizeof(double) * rb.n, cudaMemcpyDeviceToDevice, h->stream));
      CK(cudaMemcpyAsync(m.b + off, rb.b, sizeof(double) * rb.n, cudaMemcpyDeviceToDevice, h->stream));
      off += rb.n;
    }
    CK(cudaStreamSynchronize(h->stream));
    for (RowBlock& rb : all)
      if (rb.k == k) free_block(h, rb);
    h->blocks.push_back(m);
  }
  return 0;
}

// Drop the dead rows of a merged block and sort the rest by first column (k_rows_sort_keys / k_rows_permute): the block
// k_rows_normal streams every iteration then holds live rows only -- on a z slab every rank is handed all the points
// and keeps the few whose cell touches its planes -- in an order that keeps a warp's gathers and adds close together
// and the rows of one constraint point in adjacent lanes.
static int compact_and_order(lsb200_ctx* h, RowBlock& rb) {
  if (rb.ordered || rb.n == 0) return 0;
  if (rb.n >= 2147483647LL) return fail(LSB200_ERR_ARG, "a row block of %lld rows is too large", rb.n);
  const int n = (int)rb.n;
  unsigned int *k_in = nullptr, *k_out = nullptr;
  int *i_in = nullptr, *i_out = nullptr;
  void* tmp = nullptr;
  CK(cudaMallocAsync(&k_in, sizeof(unsigned int) * n, h->stream));
  CK(cudaMallocAsync(&k_out, sizeof(unsigned int) * n, h->stream));
  CK(cudaMallocAsync(&i_in, sizeof(int) * n, h->stream));
  CK(cudaMallocAsync(&i_out, sizeof(int) * n, h->stream));
  CK(cudaMemsetAsync(h->status, 0, sizeof(unsigned long long) * 4, h->stream));
  const int nb = std::min((n + 255) / 256, 4096);
  k_rows_sort_keys<<<nb, 256, 0, h->stream>>>(rb, k_in, i_in, h->status);
  CKL();
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, i_in, i_out, n, 0, 32, h->stream));
  CK(cudaMallocAsync(&tmp, tmp_bytes, h->stream));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, i_in, i_out, n, 0, 32, h->stream));  // stable
  unsigned long long live = 0;
  CK(cudaMemcpyAsync(&live, h->status, sizeof(live), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  RowBlock out{};
  if (live > 0) {
    int rc = alloc_block(h, out, (long long)live, rb.k);
    if (rc) return rc;
    k_rows_permute<<<std::min(((int)live + 255) / 256, 4096), 256, 0, h->stream>>>(rb, i_out, out);
    CKL();
  }
  h->launches += 3;
  out.ordered = 1;
  out.k = rb.k;
  pool_free(h, k_in); pool_free(h, k_out); pool_free(h, i_in); pool_free(h, i_out); pool_free(h, tmp);
  free_block(h, rb);
  rb = out;
  return 0;
}

static void free_transposes(lsb200_ctx* h) {
  for (RowTranspose& t : h->transposes) {
    pool_free(h, t.node); pool_free(h, t.off); pool_free(h, t.entry); pool_free(h, t.t);
  }
  h->transposes.clear();
}

// node-major view of every block: stable radix sort of (node, row) keys, run-length encode, scan
static int build_transposes(lsb200_ctx* h) {
  free_transposes(h);
  for (const RowBlock& rb : h->blocks) {
    RowTranspose tr{};
    const long long E = rb.n * rb.k;
    if (E == 0) { h->transposes.push_back(tr); continue; }
    if (E >= 2147483647LL) return fail(LSB200_ERR_ARG, "deterministic mode: block of %lld entries is too large", E);
    unsigned long long *k_in = nullptr, *k_out = nullptr;
    int *v_in = nullptr, *nodes = nullptr, *counts = nullptr, *nruns = nullptr;
    CK(cudaMallocAsync(&k_in, sizeof(unsigned long long) * E, h->stream));
    CK(cudaMallocAsync(&k_out, sizeof(unsigned long long) * E, h->stream));
    CK(cudaMallocAsync(&v_in, sizeof(int) * E, h->stream));
    CK(cudaMallocAsync(&tr.entry, sizeof(int) * E, h->stream));
    CK(cudaMallocAsync(&nodes, sizeof(int) * E, h->stream));
    CK(cudaMallocAsync(&tr.node, sizeof(int) * E, h->stream));
    CK(cudaMallocAsync(&counts, sizeof(int) * (E + 1), h->stream));
    CK(cudaMallocAsync(&tr.off, sizeof(int) * (E + 2), h->stream));
    CK(cudaMallocAsync(&nruns, sizeof(int), h->stream));
    CK(cudaMallocAsync(&tr.t, sizeof(double) * rb.n, h->stream));
    const int nb = (int)std::min<long long>((E + 255) / 256, 4096);
    k_rows_keys<<<nb, 256, 0, h->stream>>>(rb, k_in, v_in);
    CKL();
    size_t tmp_bytes = 0, need = 0;
    void* tmp = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, need, k_in, k_out, v_in, tr.entry, (int)E, 0, 64, h->stream);
    tmp_bytes = need;
    cub::DeviceRunLengthEncode::Encode(nullptr, need, nodes, tr.node, counts, nruns, (int)E, h->stream);
    tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(nullptr, need, counts, tr.off, (int)E + 1, h->stream);
    tmp_bytes = std::max(tmp_bytes, need);
    CK(cudaMallocAsync(&tmp, tmp_bytes, h->stream));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, tr.entry, (int)E, 0, 64, h->stream));
    k_rows_node_of_key<<<nb, 256, 0, h->stream>>>(E, k_out, nodes);
    CKL();
    CK(cub::DeviceRunLengthEncode::Encode(tmp, tmp_bytes, nodes, tr.node, counts, nruns, (int)E, h->stream));
    int runs = 0;
    CK(cudaMemcpyAsync(&runs, nruns, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemsetAsync(counts + runs, 0, sizeof(int), h->stream));
    CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, tr.off, runs + 1, h->stream));
    // the padding run (node -1) sorts last: drop it
    int last_node = 0;
    if (runs > 0) CK(cudaMemcpyAsync(&last_node, tr.node + runs - 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    tr.n_nodes = (runs > 0 && last_node < 0) ? runs - 1 : runs;
    h->launches += 5;
    pool_free(h, k_in); pool_free(h, k_out); pool_free(h, v_in); pool_free(h, nodes); pool_free(h, counts); pool_free(h, nruns); pool_free(h, tmp);
    h->transposes.push_back(tr);
  }
  return 0;
}

// The planner's cost model (csrc/fdi_plan.cuh) ranks plans well on one grid and badly on the next: at 512^3 the plan it
// prefers with strips costed at 1.0 / 1.4 box steps runs in 409 us against 489 us for its default, at 256^3 the same
// constants lose 76 us against 72 (profiles/r2q).  So the model only proposes: the plans it picks under a handful of
// constant sets are each launched a few times on the handle's own vectors (q = N p, both scratch before a solve) and the
// fastest stays.  Once per handle, a few ms; skipped in deterministic mode (the plan fixes the order of the p.q sums)
// and whenever an environment switch pins the plan.
static void plan_small(lsb200_ctx* h);
static int launch_small(lsb200_ctx* h, int iters);

static int tune_fused(lsb200_ctx* h) {
  static const double models[][3] = {{12, 1.45, 2.0}, {24, 1.45, 2.0}, {24, 1.8, 2.4}, {24, 1.0, 1.4}, {12, 1.0, 1.4}, {24, 0.7, 1.0}};
  const double keep[3] = {FUSED_FACE_COST, FUSED_XS_REL, FUSED_YS_REL};
  std::vector<FusedChoice> cands;
  for (const auto& m : models) {
    FUSED_FACE_COST = m[0]; FUSED_XS_REL = m[1]; FUSED_YS_REL = m[2];
    FusedChoice c;
    if (!plan_fused(h->g, h->sm_count, MAX_PARTIALS, 0, 0, 0, &c)) continue;
    bool seen = false;
    for (const FusedChoice& o : cands) seen = seen || (o.tyt == c.tyt && memcmp(&o.pl, &c.pl, sizeof(FusedPlan)) == 0);
    if (!seen) cands.push_back(c);
  }
  FUSED_FACE_COST = keep[0]; FUSED_XS_REL = keep[1]; FUSED_YS_REL = keep[2];
  // ... the model's favourite of every tile shape ...
#define X(TYT, ST, MINB)                                                                                          \
  {                                                                                                               \
    FusedChoice c;                                                                                                \
    if (plan_fused(h->g, h->sm_count, MAX_PARTIALS, TYT, 0, 0, &c)) {                                             \
      bool seen = false;                                                                                          \
      for (const FusedChoice& o : cands) seen = seen || (o.tyt == c.tyt && memcmp(&o.pl, &c.pl, sizeof(FusedPlan)) == 0); \
      if (!seen) cands.push_back(c);                                                                              \
    }                                                                                                             \
  }
  LSB_FUSED_CFGS(X)
#undef X
  // ... and the neighbours of the default plan: one chunk fewer to two more, strips shorter and longer
  {
    const FusedChoice base = h->fz;
    const int minb = base.minb;
    // (a grid whose plan leaves CTA slots empty is bound by the latency of a march, not by bandwidth: chunks down to
    // 2 planes, twice and four times as many as the model likes)
    const long long slots = (long long)h->sm_count * minb;
    const bool small = base.ctas < slots;
    std::vector<std::pair<int, double>> var;
    for (int dn = -1; dn <= 2; ++dn)
      for (const double fs : {1.0, 0.7, 1.4, 2.0}) var.push_back({base.pl.chunks + dn, fs});
    if (small)
      for (const int mul : {2, 4, 8})
        for (const double fs : {1.0, 0.5, 0.25}) var.push_back({base.pl.chunks * mul, fs});
    for (const auto& v : var) {
      if (cands.size() >= 32) break;
      {
        const int nch = v.first, scz = std::max(small ? 3 : 6, (int)(base.pl.scz * v.second + 0.5));
        if (nch < 1 || nch > std::max(1, h->g.nzl / (small ? 2 : 8))) continue;
        const FusedChoice c = fused_candidate(h->g, base.tyt, minb, nch, scz, h->sm_count);
        if (c.ctas > MAX_PARTIALS) continue;
        bool seen = false;
        for (const FusedChoice& o : cands) seen = seen || (o.tyt == c.tyt && memcmp(&o.pl, &c.pl, sizeof(FusedPlan)) == 0);
        if (!seen) cands.push_back(c);
      }
    }
  }
  h->tuned = true;
  if (cands.size() < 2) return 0;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaMemsetAsync(h->p, 0, sizeof(double) * (size_t)h->g.nall, h->stream));
  const FusedChoice first = h->fz;
  float best_ms = 0.f, first_ms = 0.f;
  int best = -1, rc = 0;
  for (size_t i = 0; i < cands.size() && rc == 0; ++i) {
    h->fz = cands[i];
    h->tma_vec[0] = h->tma_vec[1] = nullptr;  // the TMA box follows the tile shape
    for (int rep = -1; rep < 3 && rc == 0; ++rep) {
      if (rep == 0) CK(cudaEventRecord(e0, h->stream));
      rc = launch_stencil(h, h->p, h->q, 0);
    }
    if (rc) break;
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (i == 0) first_ms = ms;
    if (best < 0 || ms < best_ms) { best = (int)i; best_ms = ms; }
  }
  // three launches resolve ~2 %: a candidate must beat the model's favourite by more than that to replace it
  if (best > 0 && best_ms > 0.98f * first_ms) best = 0;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  h->fz = (rc == 0 && best >= 0) ? cands[best] : first;
  h->tma_vec[0] = h->tma_vec[1] = nullptr;
  return rc;
}

int lsb200_finalize_system(lsb200_handle h) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (h->finalized) return 0;
  CK(cudaSetDevice(h->device));
  if (h->cg_graph) {
    cudaGraphExecDestroy(h->cg_graph);
    h->cg_graph = nullptr;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, h->stream));
  int rc = ensure_vectors(h);
  if (rc) return rc;
  if (h->kind == LSB200_SUPPORT_GRID && h->fused && !h->tuned && !h->plan_forced && !h->deterministic && (rc = tune_fused(h))) return rc;
  rc = merge_blocks(h);
  if (rc) return rc;
  for (RowBlock& rb : h->blocks)
    if ((rc = compact_and_order(h, rb))) return rc;
  const size_t bytes = sizeof(double) * (size_t)h->g.nall;
  CK(cudaMemsetAsync(h->rhs, 0, bytes, h->stream));
  CK(cudaMemsetAsync(h->diag, 0, bytes, h->stream));
  // regularisation diagonal (+ ridge^2)
  if (h->kind == LSB200_SUPPORT_GRID) {
    k_fdi_diag<<<vec_blocks(h, h->g.nloc, 1), VEC_NT, 0, h->stream>>>(h->g, make_weights(h), h->diag);
    CKL();
    h->launches++;
  } else {
    if (!h->p1_table) return fail(LSB200_ERR_STATE, "tet-mesh handle without regularisation tables (lsb200_set_p1_tables)");
    const P1Tables tb{{h->p1_m[0], h->p1_m[1], h->p1_m[2]}, h->p1_table, h->p1_off, h->p1_diag};
    k_p1_diag<<<vec_blocks(h, h->g.nloc, 1), VEC_NT, 0, h->stream>>>(h->g, tb, h->p1_w * h->p1_w,
                                                                      h->add_ridge ? h->ridge * h->ridge : 0.0, h->diag);
    CKL();
    h->launches++;
  }
  if (h->deterministic) {
    if ((rc = build_transposes(h))) return rc;
  } else {
    free_transposes(h);
  }
  for (size_t bi = 0; bi < h->blocks.size(); ++bi) {
    const RowBlock& rb = h->blocks[bi];
    if (rb.n == 0) continue;
    if (h->deterministic) {
      const RowTranspose& tr = h->transposes[bi];
      const int ng = std::max(1, std::min((tr.n_nodes + 255) / 256, h->sm_count * 16));
      k_rows_gather<<<ng, 256, 0, h->stream>>>(h->g, rb, tr, h->rhs, 1, h->sc, 0);
      CKL();
      k_rows_gather<<<ng, 256, 0, h->stream>>>(h->g, rb, tr, h->diag, 2, h->sc, 0);
      CKL();
    } else {
      const int nb = (int)std::min<long long>((rb.n + 255) / 256, 8192);
      k_rows_rhs<<<nb, 256, 0, h->stream>>>(h->g, rb, h->rhs);
      CKL();
      k_rows_diag<<<nb, 256, 0, h->stream>>>(h->g, rb, h->diag);
      CKL();
    }
    h->launches += 2;
  }
  if (h->apply_scaling) {
    k_scaling_from_diag<<<vec_blocks(h, h->g.nloc, 1), VEC_NT, 0, h->stream>>>(h->g.nloc, owned_ptr(h, h->diag), owned_ptr(h, h->s2));
    CKL();
    h->launches++;
  }
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  h->setup_seconds = ms * 1e-3;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  plan_small(h);
  h->finalized = true;
  return 0;
}

int lsb200_comm_unique_id(unsigned char id_out[128]) {
  if (!id_out) return fail(LSB200_ERR_ARG, "null argument");
  int rc = nccl::load();
  if (rc) return rc;
  nccl::unique_id id;
  NCK(nccl::GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return 0;
}

int lsb200_comm_init(lsb200_handle h, const unsigned char nccl_id[128], int rank, int nranks) {
  if (!h || !nccl_id) return fail(LSB200_ERR_ARG, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(LSB200_ERR_ARG, "bad rank %d of %d", rank, nranks);
  if (nranks == 1) { h->rank = 0; h->nranks = 1; return 0; }
  int rc = nccl::load();
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  nccl::unique_id id;
  memcpy(id.internal, nccl_id, 128);
  NCK(nccl::CommInitRank(&h->comm, nranks, id, rank));
  h->rank = rank;
  h->nranks = nranks;
  // a slab needs HALO owned planes to feed its neighbours' halos, and the slabs must tile the grid in rank order
  if (h->g.nzl < HALO) return fail(LSB200_ERR_ARG, "rank %d owns %d planes; a z slab needs at least %d", rank, h->g.nzl, HALO);
  {
    int* d_k = nullptr;
    std::vector<int> hk(2 * (size_t)nranks);
    const int mine[2] = {h->g.k0, h->g.k1};
    CK(cudaMalloc(&d_k, sizeof(int) * 2 * (nranks + 1)));
    CK(cudaMemcpyAsync(d_k + 2 * nranks, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    NCK(nccl::AllGather(d_k + 2 * nranks, d_k, sizeof(mine), nccl::Int8, h->comm, h->stream));
    CK(cudaMemcpyAsync(hk.data(), d_k, sizeof(int) * 2 * nranks, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_k);
    h->slab_k = hk;
    for (int r = 0; r < nranks; ++r) {
      const int want = r == 0 ? 0 : hk[2 * (r - 1) + 1];
      if (hk[2 * r] != want || (r == nranks - 1 && hk[2 * r + 1] != h->g.nz))
        return fail(LSB200_ERR_ARG, "the slabs of the %d ranks do not tile the %d planes in rank order (rank %d owns [%d,%d))", nranks,
                    h->g.nz, r, hk[2 * r], hk[2 * r + 1]);
    }
  }
  // peer mailboxes for the two scalar reductions of an iteration; LSB200_PEER_SCALARS=0 (or any failure to map a peer,
  // e.g. no P2P path between the devices) keeps ncclAllReduce
  static const bool want_peers = !(getenv("LSB200_PEER_SCALARS") && atoi(getenv("LSB200_PEER_SCALARS")) == 0);
  if (want_peers && nranks <= PEER_MAX_RANKS) {
    bool ok = true;
    cudaIpcMemHandle_t mine;
    std::vector<cudaIpcMemHandle_t> all((size_t)nranks);
    char* d_h = nullptr;
    ok = ok && cudaMalloc(&h->mailbox, sizeof(PeerMailbox)) == cudaSuccess;
    ok = ok && cudaMemset(h->mailbox, 0, sizeof(PeerMailbox)) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine, h->mailbox) == cudaSuccess;
    ok = ok && cudaMalloc(&d_h, sizeof(mine) * (nranks + 1)) == cudaSuccess;
    // every rank takes part in the collective even when its own set-up failed (its handle is then all zeros)
    if (!ok) memset(&mine, 0, sizeof(mine));
    if (d_h) {
      CK(cudaMemcpyAsync(d_h + sizeof(mine) * nranks, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
      NCK(nccl::AllGather(d_h + sizeof(mine) * nranks, d_h, sizeof(mine), nccl::Int8, h->comm, h->stream));
      CK(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * nranks, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      cudaFree(d_h);
    } else {
      ok = false;
    }
    PeerTable pt{};
    pt.rank = rank;
    pt.nranks = nranks;
    for (int r = 0; r < nranks && ok; ++r) {
      bool zero = true;
      for (size_t t = 0; t < sizeof(mine); ++t) zero = zero && reinterpret_cast<const char*>(&all[r])[t] == 0;
      if (zero) { ok = false; break; }
      if (r == rank) { pt.box[r] = h->mailbox; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      pt.box[r] = (PeerMailbox*)ptr;
    }
    // all ranks must agree: one rank on NCCL and its peers on mailboxes would wait for each other for ever
    int* d_ok = nullptr;
    CK(cudaMalloc(&d_ok, sizeof(int) * (nranks + 1)));
    const int my_ok = ok ? 1 : 0;
    std::vector<int> oks((size_t)nranks);
    CK(cudaMemcpyAsync(d_ok + nranks, &my_ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NCK(nccl::AllGather(d_ok + nranks, d_ok, sizeof(int), nccl::Int8, h->comm, h->stream));
    CK(cudaMemcpyAsync(oks.data(), d_ok, sizeof(int) * nranks, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_ok);
    for (int r = 0; r < nranks; ++r) ok = ok && oks[r] == 1;
    if (ok) {
      h->peer_host = pt;
      CK(cudaMalloc(&h->peers, sizeof(PeerTable)));
      CK(cudaMemcpy(h->peers, &pt, sizeof(PeerTable), cudaMemcpyHostToDevice));
      // the neighbours' p vectors, for the halo planes (LSB200_PEER_HALO=0 keeps ncclSend / ncclRecv).  Collective: every
      // rank allocates its vectors now and shares the handle of p; all must succeed or none uses it.
      static const bool want_halo = !(getenv("LSB200_PEER_HALO") && atoi(getenv("LSB200_PEER_HALO")) == 0);
      bool hok = want_halo && h->g.nzl >= 3 * HALO && h->g.plane % 2 == 0 && ensure_vectors(h) == 0;
      cudaIpcMemHandle_t pm;
      memset(&pm, 0, sizeof(pm));
      hok = hok && cudaIpcGetMemHandle(&pm, h->p - h->guard) == cudaSuccess;
      if (!hok) memset(&pm, 0, sizeof(pm));
      std::vector<cudaIpcMemHandle_t> pall((size_t)nranks);
      char* d_p = nullptr;
      CK(cudaMalloc(&d_p, sizeof(pm) * (nranks + 1)));
      CK(cudaMemcpyAsync(d_p + sizeof(pm) * nranks, &pm, sizeof(pm), cudaMemcpyHostToDevice, h->stream));
      NCK(nccl::AllGather(d_p + sizeof(pm) * nranks, d_p, sizeof(pm), nccl::Int8, h->comm, h->stream));
      CK(cudaMemcpyAsync(pall.data(), d_p, sizeof(pm) * nranks, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      cudaFree(d_p);
      for (int r = 0; r < nranks; ++r) {
        bool zero = true;
        for (size_t t = 0; t < sizeof(pm); ++t) zero = zero && reinterpret_cast<const char*>(&pall[r])[t] == 0;
        hok = hok && !zero;
      }
      if (hok) {
        for (int side = 0; side < 2 && hok; ++side) {
          const int nb = side == 0 ? rank - 1 : rank + 1;
          if (nb < 0 || nb >= nranks) continue;
          void* ptr = nullptr;
          if (cudaIpcOpenMemHandle(&ptr, pall[nb], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); hok = false; break; }
          h->nb_p_raw[side] = (double*)ptr;
          // same plane size on every rank -> same guard; the rank below receives its UPPER halo planes (stored planes
          // nzl' + 2, nzl' + 3), the rank above its LOWER ones (stored planes 0, 1)
          const long long nb_nzl = h->slab_k[2 * nb + 1] - h->slab_k[2 * nb];
          h->nb_halo[side] = (double*)ptr + h->guard + (side == 0 ? (nb_nzl + HALO) * h->g.plane : 0);
        }
      }
      int* d_ok2 = nullptr;
      CK(cudaMalloc(&d_ok2, sizeof(int) * (nranks + 1)));
      const int my_ok2 = hok ? 1 : 0;
      std::vector<int> oks2((size_t)nranks);
      CK(cudaMemcpyAsync(d_ok2 + nranks, &my_ok2, sizeof(int), cudaMemcpyHostToDevice, h->stream));
      NCK(nccl::AllGather(d_ok2 + nranks, d_ok2, sizeof(int), nccl::Int8, h->comm, h->stream));
      CK(cudaMemcpyAsync(oks2.data(), d_ok2, sizeof(int) * nranks, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      cudaFree(d_ok2);
      for (int r = 0; r < nranks; ++r) hok = hok && oks2[r] == 1;
      h->peer_halo = hok;
    } else {
      for (int r = 0; r < nranks; ++r)
        if (r != rank && pt.box[r]) cudaIpcCloseMemHandle(pt.box[r]);
      cudaFree(h->mailbox);
      h->mailbox = nullptr;
    }
  }
  return 0;
}

int lsb200_comm_uses_peer_memory(lsb200_handle h) { return (h && h->peers) ? 1 : 0; }

// q = N p on device vectors (p must have its halo planes filled on slab handles)
static int apply_normal_dev(lsb200_ctx* h, const double* p, double* q, int parts, int cg_mode) {
  int rc;
  if (parts & 1) {
    if ((rc = launch_regularisation(h, p, q, cg_mode))) return rc;
  } else if (!cg_mode) {
    CK(cudaMemsetAsync(q, 0, sizeof(double) * (size_t)h->g.nall, h->stream));
  }
  if (parts & 2)
    if ((rc = launch_rows(h, p, q, cg_mode))) return rc;
  return 0;
}

int lsb200_apply_normal(lsb200_handle h, const double* p, double* q, int parts) {
  if (!h || !p || !q) return fail(LSB200_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  int rc = ensure_vectors(h);
  if (rc) return rc;
  if ((rc = merge_blocks(h))) return rc;
  if ((rc = copy_in(h, h->p, p))) return rc;
  if ((rc = halo_exchange(h, h->p))) return rc;
  // ridge is applied inside the regularisation kernel; switch it off unless asked for
  if (!(parts & 1) && (parts & 4))
    return fail(LSB200_ERR_ARG, "the ridge part is applied together with the regularisation part (parts bit0)");
  const int keep = h->add_ridge;
  if (!(parts & 4)) h->add_ridge = 0;
  rc = apply_normal_dev(h, h->p, h->q, parts & 3, 0);
  h->add_ridge = keep;
  if (rc) return rc;
  return copy_out(h, q, h->q);
}

int lsb200_get_rhs(lsb200_handle h, double* out) {
  if (!h || !out) return fail(LSB200_ERR_ARG, "null argument");
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  return copy_out(h, out, h->rhs);
}
int lsb200_set_rhs(lsb200_handle h, const double* rhs) {
  if (!h || !rhs) return fail(LSB200_ERR_ARG, "null argument");
  int rc = lsb200_finalize_system(h);  // diagonal / scaling of the rows as they are; their own rhs is replaced below
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  return copy_in(h, h->rhs, rhs);
}
int lsb200_get_diagonal(lsb200_handle h, double* out) {
  if (!h || !out) return fail(LSB200_ERR_ARG, "null argument");
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  return copy_out(h, out, h->diag);
}
int lsb200_get_scaling(lsb200_handle h, double* out) {
  if (!h || !out) return fail(LSB200_ERR_ARG, "null argument");
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  if ((rc = copy_out(h, out, h->s2))) return rc;
  const long long n = (long long)h->g.nx * h->g.ny * h->g.nzl;
  for (long long i = 0; i < n; ++i) out[i] = h->apply_scaling ? sqrt(out[i]) : 1.0;
  return 0;
}

int lsb200_get_c(lsb200_handle h, double* out) {
  if (!h || !out) return fail(LSB200_ERR_ARG, "null argument");
  if (!h->have_c) return fail(LSB200_ERR_STATE, "no solution on the device yet (solve or set_c first)");
  CK(cudaSetDevice(h->device));
  return copy_out(h, out, h->x);
}
int lsb200_set_c(lsb200_handle h, const double* c) {
  if (!h || !c) return fail(LSB200_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  int rc = ensure_vectors(h, true);
  if (rc) return rc;
  if ((rc = copy_in(h, h->x, c))) return rc;
  h->have_c = true;
  return 0;
}

// device <-> device variants: `c_dev` holds the owned planes contiguously (nx * ny * nzl doubles) on the handle's GPU.
// Asynchronous on `stream`, ordered after the handle's own work (the solve / set_c that produced x has completed: both
// synchronise before they return).  Used to replicate a z-slab solution on every GPU before evaluation.
int lsb200_get_c_dev(lsb200_handle h, double* c_dev, void* stream) {
  if (!h || !c_dev) return fail(LSB200_ERR_ARG, "null argument");
  if (!h->have_c) return fail(LSB200_ERR_STATE, "no solution on the device yet (solve or set_c first)");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpy2DAsync(c_dev, sizeof(double) * h->g.nx, owned_ptr(h, h->x), sizeof(double) * h->g.pitch, sizeof(double) * h->g.nx,
                       (size_t)h->g.ny * h->g.nzl, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int lsb200_set_c_dev(lsb200_handle h, const double* c_dev, void* stream) {
  if (!h || !c_dev) return fail(LSB200_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  int rc = ensure_vectors(h, true);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));  // ensure_vectors zeroes the vectors on the handle's stream
  CK(cudaMemcpy2DAsync(owned_ptr(h, h->x), sizeof(double) * h->g.pitch, c_dev, sizeof(double) * h->g.nx, sizeof(double) * h->g.nx,
                       (size_t)h->g.ny * h->g.nzl, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  h->have_c = true;
  return 0;
}

// one CG iteration enqueued on the stream (no host synchronisation)
// ---- small systems: `iters` CG iterations in one persistent cooperative launch (csrc/cg_small.cuh) -----------------------
constexpr long long SMALL_MAX_NODES = 3000000;  // above: an iteration is bandwidth bound and the four launches stream better
// tet meshes: the marching operator fills an SM with 2 CTAs of 8 warps, too few threads for the loop's streaming phases --
// config 3 (1.64 M nodes): 66.0 us per iteration in the loop, 60.3 launch by launch (profiles/r2am_c3.txt); the loop's
// 8 us of barriers against 16 us of launch gaps put the crossover near 0.9 M nodes
constexpr long long SMALL_MAX_NODES_P1 = 800000;
static TmapArg* small_tmap(lsb200_ctx* h) {     // descriptor of p for the plan's tile shape (as launch_stencil builds it)
  for (int t = 0; t < 2; ++t)
    if (h->tma_vec[t] == h->p) return &h->tma_map[t];
  bool ok = false;
  switch (h->fz.tyt) {
#define X(TYT, ST, MINB) case TYT: ok = encode_plane_tiles(h, h->p, MainCfg<4, TYT, ST>::SX, MainCfg<4, TYT, ST>::SY, &h->tma_map[0]); break;
    LSB_FUSED_CFGS(X)
#undef X
  }
  if (!ok) return nullptr;
  h->tma_vec[0] = h->p;
  return &h->tma_map[0];
}

static int small_grid_of(lsb200_ctx* h, const void* kern, int threads, size_t smem, long long useful) {
  int per_sm = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  long long grid = (long long)per_sm * h->sm_count;
  grid = std::min<long long>(grid, SMALL_MAX_CTAS);
  grid = std::min<long long>(grid, std::max<long long>(useful, h->sm_count));
  return (int)grid;
}

// can this handle run the persistent loop, and with how many CTAs (called at the end of finalize)
static void plan_small(lsb200_ctx* h) {
  h->small_ok = false;
  const char* e = getenv("LSB200_PERSIST");
  const int want = e ? atoi(e) : -1;  // 0: never, 1: whenever possible, default: systems of <= SMALL_MAX_NODES nodes
  if (want == 0 || h->nranks > 1 || h->deterministic || h->blocks.size() > (size_t)SMALL_MAX_BLOCKS) return;
  if (want != 1 && h->g.nloc > (h->kind == LSB200_SUPPORT_GRID || !h->p1_march ? SMALL_MAX_NODES : SMALL_MAX_NODES_P1)) return;
  int coop = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device) != cudaSuccess || !coop) return;
  const long long per_cta_nodes = 4096;
  const long long by_nodes = (h->g.nloc + per_cta_nodes - 1) / per_cta_nodes;
  int grid = 0;
  if (h->kind == LSB200_SUPPORT_GRID) {
    if (!h->fused || !h->want_tma || h->g.pitch % 2 != 0 || !small_tmap(h)) return;
    const long long nvb = (long long)h->fz.pl.n_box + h->fz.pl.n_xs + h->fz.pl.n_ys;
    switch (h->fz.tyt) {
#define X(TYT, ST, MINB) \
  case TYT: grid = small_grid_of(h, (const void*)k_cg_small_fdi<TYT, ST, MINB>, 32 * TYT, sizeof(double) * MainCfg<4, TYT, ST>::PLANE * ST, std::max(nvb, by_nodes)); break;
      LSB_FUSED_CFGS(X)
#undef X
    }
  } else {
    if (h->p1_march) {
      const long long nvb = (long long)h->p1_mplan.n_cta + h->p1_mplan.b.n_bnd_cta;
      switch (h->p1_march) {
#define X(ID, NW, NB, MINB)                                                                                              \
  case ID:                                                                                                               \
    grid = small_grid_of(h, (const void*)k_cg_small_p1m<NW, NB, MINB>, 32 * (NW + NB),                                   \
                         sizeof(double) * P1MarchCfg<NW, NB>::PLANE * P1MarchCfg<NW, NB>::ST, std::max(nvb, by_nodes));  \
    break;
        LSB_P1_MARCH_CFGS(X)
#undef X
      }
    } else {
      if (!h->p1_tiles || h->p1_tile_cfg != 1) return;
      const long long nvb = (long long)h->p1_plan.n_bnd_cta + (long long)h->p1_plan.tiles_x * h->p1_plan.tiles_y * h->p1_plan.tiles_z;
      grid = small_grid_of(h, (const void*)k_cg_small_p1<8, 4, 8>, 256, sizeof(double) * P1TileGeom<8, 4>::ELEMS, std::max(nvb, by_nodes));
    }
  }
  if (grid < 1) return;
  if (!h->small_bar) {
    if (cudaMalloc(&h->small_bar, 64 * sizeof(unsigned int)) != cudaSuccess || cudaMemset(h->small_bar, 0, 64 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(&h->small_slots, sizeof(double) * 2 * SMALL_MAX_CTAS) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
  }
  h->small_grid = grid;
  h->small_ok = true;
}

static int launch_small(lsb200_ctx* h, int iters) {
  SmallCgArgs a{};
  a.n = h->g.nloc;
  a.x = owned_ptr(h, h->x); a.r = owned_ptr(h, h->r); a.p = owned_ptr(h, h->p); a.q = owned_ptr(h, h->q);
  a.s2 = h->apply_scaling ? owned_ptr(h, h->s2) : nullptr;
  a.p_all = h->p; a.q_all = h->q;
  a.nblocks = 0;
  for (const RowBlock& rb : h->blocks)
    if (rb.n > 0) a.blocks[a.nblocks++] = rb;
  a.iters = iters;
  a.bar = h->small_bar;
  CK(cudaMemsetAsync(h->small_bar, 0, sizeof(unsigned int), h->stream));  // the arrival counter of the launch's barriers
  a.slots = h->small_slots;
  a.sc = h->sc;
  if (h->kind == LSB200_SUPPORT_GRID) {
    SmallFdi f{h->g, make_uni_consts(make_weights(h)), make_weights(h), h->fz.pl, h->ztab};
    TmapArg* tm = small_tmap(h);
    if (!tm) return fail(LSB200_ERR_STATE, "no TMA descriptor for the persistent CG loop");
    void* args[] = {&a, &f, tm};
    switch (h->fz.tyt) {
#define X(TYT, ST, MINB)                                                                                                   \
  case TYT:                                                                                                                \
    CK(cudaLaunchCooperativeKernel((void*)k_cg_small_fdi<TYT, ST, MINB>, dim3(h->small_grid), dim3(32, TYT), args,          \
                                   sizeof(double) * MainCfg<4, TYT, ST>::PLANE * ST, h->stream));                          \
    break;
      LSB_FUSED_CFGS(X)
#undef X
      default: return fail(LSB200_ERR_STATE, "no persistent CG kernel for tile shape %d", h->fz.tyt);
    }
  } else {
    const P1Tables tb{{h->p1_m[0], h->p1_m[1], h->p1_m[2]}, h->p1_table, h->p1_off, h->p1_diag};
    if (h->p1_march) {
      SmallP1M f{h->g, tb, h->p1_mplan, h->p1_w * h->p1_w, h->add_ridge ? h->ridge * h->ridge : 0.0};
      void* args[] = {&a, &f, &h->p1_mtaps};
      switch (h->p1_march) {
#define X(ID, NW, NB, MINB)                                                                                              \
  case ID:                                                                                                               \
    CK(cudaLaunchCooperativeKernel((void*)k_cg_small_p1m<NW, NB, MINB>, dim3(h->small_grid), dim3(32 * (NW + NB)), args,  \
                                   sizeof(double) * P1MarchCfg<NW, NB>::PLANE * P1MarchCfg<NW, NB>::ST, h->stream));     \
    break;
        LSB_P1_MARCH_CFGS(X)
#undef X
      }
    } else {
      SmallP1 f{h->g, tb, h->p1_plan, h->p1_w * h->p1_w, h->add_ridge ? h->ridge * h->ridge : 0.0};
      void* args[] = {&a, &f, &h->p1_cls};
      CK(cudaLaunchCooperativeKernel((void*)k_cg_small_p1<8, 4, 8>, dim3(h->small_grid), dim3(256), args,
                                     sizeof(double) * P1TileGeom<8, 4>::ELEMS, h->stream));
    }
  }
  h->launches++;
  return 0;
}

static int enqueue_iteration(lsb200_ctx* h) {
  const long long n = h->g.nloc;
  const double* s2 = h->apply_scaling ? owned_ptr(h, h->s2) : nullptr;
  int rc;
  const long long edge = (long long)HALO * h->g.plane;
  if (h->nranks > 1 && h->overlap_shells && h->g.nzl >= 3 * HALO && h->g.plane % 2 == 0) {  // even: double2 alignment
    // z slabs: update the two planes at either end first, ship them to the neighbours on the communication
    // stream, and update the planes in between while they travel
    auto upd = [&](long long off, long long cnt) {
      k_cg_update_p<VEC_NT><<<vec_blocks(h, cnt), VEC_NT, 0, h->stream>>>(cnt, s2 ? s2 + off : nullptr, owned_ptr(h, h->r) + off,
                                                                          owned_ptr(h, h->p) + off, h->sc);
      h->launches++;
    };
    if (h->peer_halo) {
      // the edge planes go straight into the neighbours' halo planes (cg.cuh), the wait sits in front of the matvec
      auto upd_edge = [&](long long off, int side) {
        if (!h->nb_halo[side]) { upd(off, edge); return; }
        k_cg_update_p_edge<VEC_NT><<<vec_blocks(h, edge), VEC_NT, 0, h->stream>>>(
            edge, s2 ? s2 + off : nullptr, owned_ptr(h, h->r) + off, owned_ptr(h, h->p) + off, h->nb_halo[side],
            h->peer_host.box[side == 0 ? h->rank - 1 : h->rank + 1], side == 0 ? 1 : 0, h->counters + 5 + side, h->sc);
        h->launches++;
      };
      upd_edge(0, 0);         // my lowest planes are the upper halo of the rank below: it receives them "from above"
      upd_edge(n - edge, 1);  // my highest planes are the lower halo of the rank above: "from below"
      upd(edge, n - 2 * edge);
      k_cg_halo_wait<<<1, 32, 0, h->stream>>>(h->sc, h->mailbox, h->rank > 0 ? 1 : 0, h->rank < h->nranks - 1 ? 1 : 0);
      CKL();
      h->launches++;
    } else {
      upd(0, edge);
      upd(n - edge, edge);
      CKL();
      CK(cudaEventRecord(h->ev_fork, h->stream));
      CK(cudaStreamWaitEvent(h->side[1], h->ev_fork, 0));
      if ((rc = halo_exchange(h, h->p, h->side[1]))) return rc;
      CK(cudaEventRecord(h->ev_join[1], h->side[1]));
      upd(edge, n - 2 * edge);
      CKL();
      CK(cudaStreamWaitEvent(h->stream, h->ev_join[1], 0));
    }
  } else {
    k_cg_update_p<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, s2, owned_ptr(h, h->r), owned_ptr(h, h->p), h->sc);
    CKL();
    h->launches++;
    if ((rc = halo_exchange(h, h->p))) return rc;
  }
  if ((rc = apply_normal_dev(h, h->p, h->q, 3, 1))) return rc;
  if (h->nranks > 1) {
    if (h->peers) {
      k_cg_sync_pq<<<1, 32, 0, h->stream>>>(h->sc, h->peer_host);
      CKL();
      h->launches++;
    } else {
      k_cg_pack_pq<<<1, 1, 0, h->stream>>>(h->sc);
      CKL();
      h->launches++;
      if ((rc = allreduce_scalar(h, &h->sc->pq))) return rc;
    }
  }
  k_cg_update_xr<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, owned_ptr(h, h->p), owned_ptr(h, h->q), owned_ptr(h, h->x),
                                                                      owned_ptr(h, h->r), s2, h->partials + 2 * MAX_PARTIALS,
                                                                      h->counters + 2, h->sc, h->peers);
  CKL();
  h->launches++;
  if (h->nranks > 1) {
    if (h->peers) {
      k_cg_pull_rho<<<1, 32, 0, h->stream>>>(h->sc, h->peer_host);
    } else {
      if ((rc = allreduce_scalar(h, &h->sc->rho_new))) return rc;
      k_cg_scalar_step<<<1, 1, 0, h->stream>>>(h->sc);
    }
    CKL();
    h->launches++;
  }
  return 0;
}

int lsb200_solve_cg(lsb200_handle h, double rtol, double atol, int64_t maxiter, const double* x0, double* c_out,
                    lsb200_solve_info* info) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  if (!(rtol >= 0.0) || !(atol >= 0.0)) return fail(LSB200_ERR_ARG, "rtol/atol must be non-negative");
  CK(cudaSetDevice(h->device));
  const long long launches0 = h->launches;
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  const long long n = h->g.nloc;
  const size_t bytes = sizeof(double) * (size_t)h->g.nall;
  const double* s2 = h->apply_scaling ? owned_ptr(h, h->s2) : nullptr;
  if (!h->blocks.empty() || true) {
    // pq_part[1] must read 0 when there are no data rows
    CK(cudaMemsetAsync(h->sc, 0, sizeof(CgScalars), h->stream));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, h->stream));

  // ||S b||^2
  CgScalars* sc = h->sc;
  k_cg_residual<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, owned_ptr(h, h->rhs), nullptr, nullptr, s2,
                                                                     h->partials + 2 * MAX_PARTIALS, h->counters + 2, &sc->rho_new);
  CKL();
  h->launches++;
  if ((rc = allreduce_scalar(h, &sc->rho_new))) return rc;
  CK(cudaMemcpyAsync(h->sc_host, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const double bnorm = sqrt(h->sc_host->rho_new);
  const double thr = std::max(atol, rtol * bnorm);
  const long long dof_global = (long long)h->g.nx * h->g.ny * h->g.nz;
  const long long maxit = maxiter > 0 ? (long long)maxiter : 10 * dof_global;

  lsb200_solve_info res;
  memset(&res, 0, sizeof(res));
  res.rhs_norm = bnorm;
  res.threshold = thr;
  res.setup_seconds = h->setup_seconds;

  CK(cudaMemsetAsync(h->p, 0, bytes, h->stream));
  if (bnorm == 0.0) {
    // scipy returns b (all zeros) immediately
    CK(cudaMemsetAsync(h->x, 0, bytes, h->stream));
  } else {
    // x = x0 (or 0), r = b - N x0
    if (x0) {
      if ((rc = copy_in(h, h->x, x0))) return rc;
      if ((rc = halo_exchange(h, h->x))) return rc;
      if ((rc = apply_normal_dev(h, h->x, h->q, 3, 0))) return rc;
    } else {
      CK(cudaMemsetAsync(h->x, 0, bytes, h->stream));
    }
    k_cg_residual<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, owned_ptr(h, h->rhs), x0 ? owned_ptr(h, h->q) : nullptr,
                                                                       owned_ptr(h, h->r), s2, h->partials + 2 * MAX_PARTIALS,
                                                                       h->counters + 2, &sc->rho);
    CKL();
    h->launches++;
    if ((rc = allreduce_scalar(h, &sc->rho))) return rc;
    CK(cudaMemcpyAsync(h->sc_host, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CgScalars init = *h->sc_host;
    init.beta = 0.0;
    init.thr2 = thr * thr;
    init.iter = 0;
    init.maxiter = maxit;
    init.nranks = h->nranks;
    init.epoch = (++h->solve_epoch) << 32;
    init.info = 0;
    for (int t = 0; t < 5; ++t) init.pq_part[t] = 0.0;
    init.done = (sqrt(init.rho) < thr) ? 1 : 0;
    *h->sc_host = init;
    CK(cudaMemcpyAsync(sc, h->sc_host, sizeof(CgScalars), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));

    // enqueue batches of iterations; poll the device scalars once per batch, one batch behind
    const int batch = 16;
    cudaEvent_t ev[2];
    CK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    CgScalars* snap = nullptr;
    CK(cudaMallocHost(&snap, 2 * sizeof(CgScalars)));
    bool done = init.done != 0;
    long long enq = 0;
    int slot = 0;
    bool pending[2] = {false, false};
    // After the first batch (which also sets every kernel attribute) a full batch is captured once into a CUDA
    // graph and replayed: the iteration is 5-7 short launches plus the fork/join of the shell stream, and on
    // small grids the per-launch latency, not the kernels, sets the pace (profiles/r01q: 39 us / iteration on the
    // Claudius-size grid).  LSB200_GRAPH=0 keeps plain launches.
    static const bool want_graph = !(getenv("LSB200_GRAPH") && atoi(getenv("LSB200_GRAPH")) == 0);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = want_graph ? h->cg_graph : nullptr;  // a previous solve of the same system left its graph
    bool graph_tried = gexec != nullptr;
    long long launches_per_batch = h->cg_graph_launches;
    while (!done) {
      const long long todo = std::min<long long>(batch, maxit - enq);
      if (h->small_ok) {
        if ((rc = launch_small(h, (int)todo))) return rc;
      } else if (want_graph && !graph_tried && enq >= batch && todo == batch) {
        graph_tried = true;
        const long long l0 = h->launches;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
          int crc = 0;
          for (long long it = 0; it < batch && !crc; ++it) crc = enqueue_iteration(h);
          const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
          if (crc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            gexec = nullptr;
          }
        } else {
          cudaGetLastError();
        }
        launches_per_batch = h->launches - l0;
        h->launches = l0;  // nothing ran during the capture
      }
      if (h->small_ok) {
        // enqueued above
      } else if (gexec && todo == batch) {
        CK(cudaGraphLaunch(gexec, h->stream));
        h->launches += launches_per_batch;
      } else {
        for (long long it = 0; it < todo; ++it)
          if ((rc = enqueue_iteration(h))) return rc;
      }
      enq += todo;
      CK(cudaMemcpyAsync(&snap[slot], sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaEventRecord(ev[slot], h->stream));
      pending[slot] = true;
      const int other = slot ^ 1;
      if (pending[other]) {
        CK(cudaEventSynchronize(ev[other]));
        pending[other] = false;
        if (snap[other].done) done = true;
      }
      if (!done && enq >= maxit) {
        CK(cudaEventSynchronize(ev[slot]));
        pending[slot] = false;
        done = true;  // maxiter iterations enqueued: the device state says how it ended
      }
      slot ^= 1;
    }
    CK(cudaStreamSynchronize(h->stream));
    if (gexec && gexec != h->cg_graph) {
      h->cg_graph = gexec;
      h->cg_graph_launches = launches_per_batch;
    }
    if (graph) cudaGraphDestroy(graph);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    cudaFreeHost(snap);
  }
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaMemcpyAsync(h->sc_host, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  res.solve_seconds = ms * 1e-3;
  if (bnorm != 0.0) {
    res.info = h->sc_host->info;
    res.iterations = h->sc_host->iter;
    res.residual_norm = sqrt(h->sc_host->rho);
  }
  res.kernel_launches = h->launches - launches0;
  h->have_c = true;
  if (info) *info = res;
  if (c_out) return copy_out(h, c_out, h->x);
  return 0;
}

int lsb200_residual_norm(lsb200_handle h, double* scaled_norm, double* plain_norm) {
  if (!h || !scaled_norm || !plain_norm) return fail(LSB200_ERR_ARG, "null argument");
  if (!h->have_c) return fail(LSB200_ERR_STATE, "no solution on the device yet (solve or set_c first)");
  CK(cudaSetDevice(h->device));
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  const long long n = h->g.nloc;
  if ((rc = halo_exchange(h, h->x))) return rc;
  if ((rc = apply_normal_dev(h, h->x, h->q, 3, 0))) return rc;
  CgScalars* sc = h->sc;
  const double* s2 = h->apply_scaling ? owned_ptr(h, h->s2) : nullptr;
  k_cg_residual<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, owned_ptr(h, h->rhs), owned_ptr(h, h->q), owned_ptr(h, h->p), s2,
                                                                     h->partials + 2 * MAX_PARTIALS, h->counters + 2, &sc->rho_new);
  CKL();
  k_cg_residual<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, owned_ptr(h, h->p), nullptr, nullptr, nullptr,
                                                                     h->partials + 2 * MAX_PARTIALS, h->counters + 2, &sc->pq);
  CKL();
  h->launches += 2;
  if ((rc = allreduce_scalar(h, &sc->rho_new))) return rc;
  if ((rc = allreduce_scalar(h, &sc->pq))) return rc;
  CK(cudaMemcpyAsync(h->sc_host, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *scaled_norm = sqrt(h->sc_host->rho_new);
  *plain_norm = sqrt(h->sc_host->pq);
  return 0;
}

static int eval_common(lsb200_ctx* h, int64_t n, const double* xyz_dev, double* out_dev, cudaStream_t st, bool grad) {
  if (h->g.k0 != 0 || h->g.k1 != h->g.nz) return fail(LSB200_ERR_STATE, "evaluation needs a handle that owns every plane (gather c first)");
  if (!h->have_c) return fail(LSB200_ERR_STATE, "no solution on the device yet (solve or set_c first)");
  if (n == 0) return 0;
  const int nb = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 32);
  if (h->kind == LSB200_SUPPORT_GRID) {
    if (grad) k_eval_gradient_grid<<<nb, 256, 0, st>>>(h->g, h->x, n, xyz_dev, out_dev);
    else k_eval_value_grid<<<nb, 256, 0, st>>>(h->g, h->x, n, xyz_dev, out_dev);
  } else {
    if (grad) k_eval_gradient_tet<<<nb, 256, 0, st>>>(h->g, h->x, n, xyz_dev, out_dev);
    else k_eval_value_tet<<<nb, 256, 0, st>>>(h->g, h->x, n, xyz_dev, out_dev);
  }
  CKL();
  h->launches++;
  return 0;
}

int lsb200_eval_value_dev(lsb200_handle h, int64_t n, const double* xyz_dev, double* out_dev, void* stream) {
  if (!h || (n > 0 && (!xyz_dev || !out_dev))) return fail(LSB200_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  return eval_common(h, n, xyz_dev, out_dev, (cudaStream_t)stream, false);
}
int lsb200_eval_gradient_dev(lsb200_handle h, int64_t n, const double* xyz_dev, double* out_dev, void* stream) {
  if (!h || (n > 0 && (!xyz_dev || !out_dev))) return fail(LSB200_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  return eval_common(h, n, xyz_dev, out_dev, (cudaStream_t)stream, true);
}

static int eval_host(lsb200_ctx* h, int64_t n, const double* xyz, double* out, bool grad) {
  if (n < 0) return fail(LSB200_ERR_ARG, "negative point count");
  if (n == 0) return 0;
  if (!xyz || !out) return fail(LSB200_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  const int ow = grad ? 3 : 1;
  cudaError_t ce;
  const int rc = h->pipe.run(n, PIPE_CHUNK, 3 * sizeof(double), xyz, ow * sizeof(double), out, 0, nullptr, 0,
                             [&](cudaStream_t st, long long, long long m, double* dx, double* d0, double*, double*) {
                               return eval_common(h, m, dx, d0, st, grad);
                             }, &ce);
  return pipe_result(rc, ce, "staged evaluation");
}

int lsb200_eval_value(lsb200_handle h, int64_t n, const double* xyz, double* out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  return eval_host(h, n, xyz, out, false);
}
int lsb200_eval_gradient(lsb200_handle h, int64_t n, const double* xyz, double* out) {
  if (!h) return fail(LSB200_ERR_ARG, "null handle");
  return eval_host(h, n, xyz, out, true);
}

// ------------------------------------------------------------------ fault-frame evaluation (frame.cuh)
static_assert(sizeof(lsb200_profile1d) == sizeof(Profile1D), "profile layout");
static_assert(sizeof(lsb200_fault_profile) == sizeof(FaultProfile), "fault profile layout");

// one solved field -> kernel argument: grid kind, all planes, a solution; t = 21 doubles (3 x 4 world -> local,
// 3 x 3 local -> world rotation) or null = identity
static int make_frame_field(const lsb200_ctx* h, const double* t, const char* what, int f, FrameField* F) {
  if (!h) return fail(LSB200_ERR_ARG, "%s %d: null handle", what, f);
  if (h->g.k0 != 0 || h->g.k1 != h->g.nz) return fail(LSB200_ERR_STATE, "%s %d: the handle must own every plane", what, f);
  if (!h->have_c) return fail(LSB200_ERR_STATE, "%s %d: no solution on the device yet (solve or set_c first)", what, f);
  F->kind = h->kind;
  F->pad = 0;
  F->g = h->g;
  F->c = h->x;
  if (t) {
    for (int i = 0; i < 12; ++i) F->w2l[i] = t[i];
    for (int i = 0; i < 9; ++i) F->l2w[i] = t[12 + i];
  } else {
    for (int i = 0; i < 12; ++i) F->w2l[i] = (i % 5 == 0) ? 1.0 : 0.0;  // [I | 0]
    for (int i = 0; i < 9; ++i) F->l2w[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  return 0;
}

// the three handles of a frame -> kernel argument; all on one device
static int make_frame(const lsb200_handle fields[3], const double* transforms, FrameFields* FF) {
  if (!fields) return fail(LSB200_ERR_ARG, "null field array");
  for (int f = 0; f < 3; ++f) {
    const int rc = make_frame_field(fields[f], transforms ? transforms + 21 * f : nullptr, "fault-frame field", f, &FF->f[f]);
    if (rc) return rc;
    if (fields[f]->device != fields[0]->device) return fail(LSB200_ERR_ARG, "fault-frame fields live on different devices");
  }
  return 0;
}

static int check_profile(const lsb200_fault_profile* p, FaultProfile* P) {
  if (!p) return fail(LSB200_ERR_ARG, "null fault profile");
  memcpy(P, p, sizeof(FaultProfile));
  if (!p->has_function) return 0;
  for (int a = 0; a < 3; ++a)
    for (const lsb200_profile1d* q : {&p->pos[a], &p->neg[a]})
      if (q->kind < LSB200_PROFILE_ZEROS || q->kind > LSB200_PROFILE_SMOOTH_PEAK)
        return fail(LSB200_ERR_ARG, "fault profile axis %d: unknown kind %d", a, (int)q->kind);
  return 0;
}

static int frame_blocks(const lsb200_ctx* h, int64_t n) {
  return (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 32);
}

enum { FRAME_EVAL = 0, FRAME_DISPLACE = 1, FRAME_DISPLACEMENT = 2 };
struct FrameCall {
  int what;
  FrameFields FF;
  FaultProfile P;
  int gradient_of, steps;
  double displacement, offset;
};

// launch on device buffers: out0 = values / moved / d, out1 = gradient (FRAME_EVAL only, may be null)
static int frame_launch(lsb200_ctx* h, const FrameCall& fc, int64_t n, const double* xyz_dev, double* out0, double* out1,
                        cudaStream_t st) {
  if (n == 0) return 0;
  const int nb = frame_blocks(h, n);
  bool tet = false;  // any tet-mesh field: the kernels that carry the (register-hungry) tet branch
  for (int f = 0; f < 3; ++f) tet = tet || fc.FF.f[f].kind == LSB200_SUPPORT_TETMESH;
  switch (fc.what) {
    case FRAME_EVAL:
      if (tet) k_frame_eval<true><<<nb, 256, 0, st>>>(fc.FF, fc.gradient_of, n, xyz_dev, out0, out1);
      else k_frame_eval<false><<<nb, 256, 0, st>>>(fc.FF, fc.gradient_of, n, xyz_dev, out0, out1);
      break;
    case FRAME_DISPLACE:
      if (tet) k_fault_displace<true><<<nb, 256, 0, st>>>(fc.FF, fc.P, fc.displacement, fc.steps, n, xyz_dev, out0);
      else k_fault_displace<false><<<nb, 256, 0, st>>>(fc.FF, fc.P, fc.displacement, fc.steps, n, xyz_dev, out0);
      break;
    default:
      if (tet) k_fault_displacement<true><<<nb, 256, 0, st>>>(fc.FF, fc.P, fc.displacement, fc.offset, n, xyz_dev, out0);
      else k_fault_displacement<false><<<nb, 256, 0, st>>>(fc.FF, fc.P, fc.displacement, fc.offset, n, xyz_dev, out0);
  }
  CKL();
  h->launches++;
  return 0;
}

// host buffers: chunked through a staging buffer on the first field's stream
static int frame_host(lsb200_ctx* h, const FrameCall& fc, int64_t n, const double* xyz, double* out0, double* out1) {
  if (n < 0) return fail(LSB200_ERR_ARG, "negative point count");
  if (n == 0) return 0;
  if (!xyz || !out0) return fail(LSB200_ERR_ARG, "null argument");
  const int w0 = fc.what == FRAME_DISPLACEMENT ? 1 : 3;
  cudaError_t ce;
  const int rc = h->pipe.run(n, PIPE_CHUNK, 3 * sizeof(double), xyz, w0 * sizeof(double), out0, out1 ? 3 * sizeof(double) : 0, out1, 0,
                             [&](cudaStream_t st, long long, long long m, double* dx, double* d0, double* d1, double*) {
                               return frame_launch(h, fc, m, dx, d0, d1, st);
                             }, &ce);
  return pipe_result(rc, ce, "staged frame evaluation");
}

static int frame_prepare(const lsb200_handle fields[3], const double* transforms, const lsb200_fault_profile* profile,
                         bool need_profile, FrameCall* fc) {
  int rc = make_frame(fields, transforms, &fc->FF);
  if (rc) return rc;
  if (need_profile && (rc = check_profile(profile, &fc->P))) return rc;
  CK(cudaSetDevice(fields[0]->device));
  return 0;
}

int lsb200_frame_eval_dev(const lsb200_handle fields[3], const double* transforms, int gradient_of, int64_t n,
                          const double* xyz_dev, double* values_dev, double* gradient_dev, void* stream) {
  FrameCall fc{};
  fc.what = FRAME_EVAL;
  fc.gradient_of = gradient_of;
  if (gradient_dev && (gradient_of < 0 || gradient_of > 2)) return fail(LSB200_ERR_ARG, "gradient_of must be 0, 1 or 2");
  if (n < 0 || (n > 0 && (!xyz_dev || !values_dev))) return fail(LSB200_ERR_ARG, "bad argument");
  const int rc = frame_prepare(fields, transforms, nullptr, false, &fc);
  return rc ? rc : frame_launch(fields[0], fc, n, xyz_dev, values_dev, gradient_dev, (cudaStream_t)stream);
}
int lsb200_frame_eval(const lsb200_handle fields[3], const double* transforms, int gradient_of, int64_t n,
                      const double* xyz, double* values, double* gradient) {
  FrameCall fc{};
  fc.what = FRAME_EVAL;
  fc.gradient_of = gradient_of;
  if (gradient && (gradient_of < 0 || gradient_of > 2)) return fail(LSB200_ERR_ARG, "gradient_of must be 0, 1 or 2");
  const int rc = frame_prepare(fields, transforms, nullptr, false, &fc);
  return rc ? rc : frame_host(fields[0], fc, n, xyz, values, gradient);
}

int lsb200_fault_displace_dev(const lsb200_handle fields[3], const double* transforms,
                              const lsb200_fault_profile* profile, double displacement, int steps, int64_t n,
                              const double* xyz_dev, double* moved_dev, void* stream) {
  FrameCall fc{};
  fc.what = FRAME_DISPLACE;
  fc.displacement = displacement;
  fc.steps = steps;
  if (steps < 0) return fail(LSB200_ERR_ARG, "negative step count");
  if (n < 0 || (n > 0 && (!xyz_dev || !moved_dev || xyz_dev == moved_dev))) return fail(LSB200_ERR_ARG, "bad argument");
  const int rc = frame_prepare(fields, transforms, profile, true, &fc);
  return rc ? rc : frame_launch(fields[0], fc, n, xyz_dev, moved_dev, nullptr, (cudaStream_t)stream);
}
int lsb200_fault_displace(const lsb200_handle fields[3], const double* transforms, const lsb200_fault_profile* profile,
                          double displacement, int steps, int64_t n, const double* xyz, double* moved) {
  FrameCall fc{};
  fc.what = FRAME_DISPLACE;
  fc.displacement = displacement;
  fc.steps = steps;
  if (steps < 0) return fail(LSB200_ERR_ARG, "negative step count");
  const int rc = frame_prepare(fields, transforms, profile, true, &fc);
  return rc ? rc : frame_host(fields[0], fc, n, xyz, moved, nullptr);
}

// A sequence of faults applied one after the other (BaseFeature._apply_faults,
// LoopStructural/modelling/features/_base_geological_feature.py:252-270): the points go to the device once, every fault
// is one k_fault_displace launch between two device buffers, the result comes back once.  With `field` != NULL the moved
// points never come back: the field is evaluated at them on the device (GeologicalFeature.evaluate_value /
// evaluate_gradient of a faulted feature, LS/modelling/features/_geological_feature.py:116-157,231-236) and only the
// values (n) or world gradients (n,3) are returned.
static int fault_chain(int n_faults, const lsb200_handle* fields, const double* transforms, const lsb200_fault_profile* profiles,
                       const double* displacements, const int* steps, lsb200_handle field, const double* field_transform,
                       int gradient, int64_t n, const double* xyz, double* out) {
  if (n_faults < 0 || (n_faults > 0 && (!fields || !profiles || !displacements || !steps))) return fail(LSB200_ERR_ARG, "bad argument");
  if (n < 0) return fail(LSB200_ERR_ARG, "negative point count");
  if (n == 0) return 0;
  if (!xyz || !out) return fail(LSB200_ERR_ARG, "null argument");
  if (n_faults == 0 && !field) {
    memcpy(out, xyz, sizeof(double) * 3 * (size_t)n);
    return 0;
  }
  std::vector<FrameCall> calls((size_t)n_faults);
  for (int f = 0; f < n_faults; ++f) {
    FrameCall& fc = calls[f];
    fc = FrameCall{};
    fc.what = FRAME_DISPLACE;
    fc.displacement = displacements[f];
    fc.steps = steps[f];
    if (steps[f] < 0) return fail(LSB200_ERR_ARG, "fault %d: negative step count", f);
    const int rc = frame_prepare(fields + 3 * f, transforms ? transforms + 63 * f : nullptr, profiles + f, true, &fc);
    if (rc) return rc;
  }
  FrameField F{};
  if (field) {
    const int rc = make_frame_field(field, field_transform, "evaluated field", 0, &F);
    if (rc) return rc;
  }
  lsb200_ctx* h = n_faults > 0 ? fields[0] : field;
  for (int f = 0; f < n_faults; ++f)
    if (fields[3 * f]->device != h->device) return fail(LSB200_ERR_ARG, "the faults of a chain must live on one device");
  if (field && field->device != h->device) return fail(LSB200_ERR_ARG, "the evaluated field lives on another device than the faults");
  CK(cudaSetDevice(h->device));
  const int ow = field ? (gradient ? 3 : 1) : 3;
  // per chunk: the points go up once, every fault is one k_fault_displace launch between the chunk's two device buffers
  // (input buffer and scratch), the field is evaluated at the moved points, one result comes back
  cudaError_t ce;
  const int rc = h->pipe.run(
      n, PIPE_CHUNK, 3 * sizeof(double), xyz, ow * sizeof(double), out, 0, nullptr, 3 * sizeof(double),
      [&](cudaStream_t st, long long, long long m, double* dx, double* d0, double*, double* tmp) -> int {
        double* buf[2] = {dx, tmp};
        int cur = 0;
        for (int f = 0; f < n_faults; ++f) {
          // the last displacement writes straight into the output buffer when no field follows
          double* dst = (!field && f == n_faults - 1) ? d0 : buf[1 - cur];
          const int r = frame_launch(h, calls[f], m, buf[cur], dst, nullptr, st);
          if (r) return r;
          if (dst != d0) cur = 1 - cur;
        }
        if (field) {
          const int nb = frame_blocks(h, m);
          if (F.kind == LSB200_SUPPORT_TETMESH) k_field_eval<true><<<nb, 256, 0, st>>>(F, gradient ? 1 : 0, m, buf[cur], d0);
          else k_field_eval<false><<<nb, 256, 0, st>>>(F, gradient ? 1 : 0, m, buf[cur], d0);
          CKL();
          h->launches++;
        }
        return 0;
      }, &ce);
  return pipe_result(rc, ce, "staged fault chain");
}

int lsb200_fault_displace_chain(int n_faults, const lsb200_handle* fields, const double* transforms,
                                const lsb200_fault_profile* profiles, const double* displacements, const int* steps, int64_t n,
                                const double* xyz, double* moved) {
  return fault_chain(n_faults, fields, transforms, profiles, displacements, steps, nullptr, nullptr, 0, n, xyz, moved);
}

int lsb200_faulted_eval(int n_faults, const lsb200_handle* fields, const double* transforms, const lsb200_fault_profile* profiles,
                        const double* displacements, const int* steps, lsb200_handle field, const double* field_transform,
                        int gradient, int64_t n, const double* xyz, double* out) {
  if (!field) return fail(LSB200_ERR_ARG, "null field handle");
  return fault_chain(n_faults, fields, transforms, profiles, displacements, steps, field, field_transform, gradient, n, xyz, out);
}

int lsb200_fault_displacement_dev(const lsb200_handle fields[3], const double* transforms,
                                  const lsb200_fault_profile* profile, double displacement, double fault_offset,
                                  int64_t n, const double* xyz_dev, double* d_dev, void* stream) {
  FrameCall fc{};
  fc.what = FRAME_DISPLACEMENT;
  fc.displacement = displacement;
  fc.offset = fault_offset;
  if (n < 0 || (n > 0 && (!xyz_dev || !d_dev))) return fail(LSB200_ERR_ARG, "bad argument");
  const int rc = frame_prepare(fields, transforms, profile, true, &fc);
  return rc ? rc : frame_launch(fields[0], fc, n, xyz_dev, d_dev, nullptr, (cudaStream_t)stream);
}
int lsb200_fault_displacement(const lsb200_handle fields[3], const double* transforms,
                              const lsb200_fault_profile* profile, double displacement, double fault_offset, int64_t n,
                              const double* xyz, double* d) {
  FrameCall fc{};
  fc.what = FRAME_DISPLACEMENT;
  fc.displacement = displacement;
  fc.offset = fault_offset;
  const int rc = frame_prepare(fields, transforms, profile, true, &fc);
  return rc ? rc : frame_host(fields[0], fc, n, xyz, d, nullptr);
}

// ---- lattice evaluation
static int make_lattice(const double start[3], const double stop[3], const int64_t n[3], int order_f, Lattice* lat,
                        long long* total) {
  if (!start || !stop || !n) return fail(LSB200_ERR_ARG, "null lattice argument");
  long long t = 1;
  for (int a = 0; a < 3; ++a) {
    if (n[a] < 1) return fail(LSB200_ERR_ARG, "lattice axis %d has %lld points", a, (long long)n[a]);
    lat->start[a] = start[a];
    lat->stop[a] = stop[a];
    lat->n[a] = n[a];
    lat->delta[a] = n[a] > 1 ? (stop[a] - start[a]) / (double)(n[a] - 1) : 0.0;  // np.linspace: step = delta / div
    t *= n[a];
  }
  lat->order_f = order_f ? 1 : 0;
  *total = t;
  return 0;
}

int lsb200_eval_lattice_dev(lsb200_handle h, const double* transform, const double start[3], const double stop[3],
                            const int64_t n[3], int order_f, int gradient, double* out_dev, void* stream) {
  FrameField F;
  Lattice lat;
  long long total;
  int rc = make_frame_field(h, transform, "lattice field", 0, &F);
  if (rc || (rc = make_lattice(start, stop, n, order_f, &lat, &total))) return rc;
  if (!out_dev) return fail(LSB200_ERR_ARG, "null output");
  CK(cudaSetDevice(h->device));
  if (F.kind == LSB200_SUPPORT_TETMESH)
    k_eval_lattice<true><<<frame_blocks(h, total), 256, 0, (cudaStream_t)stream>>>(F, lat, 0, total, gradient ? 1 : 0, out_dev);
  else
    k_eval_lattice<false><<<frame_blocks(h, total), 256, 0, (cudaStream_t)stream>>>(F, lat, 0, total, gradient ? 1 : 0, out_dev);
  CKL();
  h->launches++;
  return 0;
}

int lsb200_eval_lattice(lsb200_handle h, const double* transform, const double start[3], const double stop[3],
                        const int64_t n[3], int order_f, int gradient, double* out) {
  FrameField F;
  Lattice lat;
  long long total;
  int rc = make_frame_field(h, transform, "lattice field", 0, &F);
  if (rc || (rc = make_lattice(start, stop, n, order_f, &lat, &total))) return rc;
  if (!out) return fail(LSB200_ERR_ARG, "null output");
  CK(cudaSetDevice(h->device));
  const int ow = gradient ? 3 : 1;
  cudaError_t ce;
  const bool tet = F.kind == LSB200_SUPPORT_TETMESH;
  rc = h->pipe.run(total, 2 * PIPE_CHUNK, 0, nullptr, ow * sizeof(double), out, 0, nullptr, 0,
                   [&](cudaStream_t st, long long o, long long m, double*, double* d0, double*, double*) -> int {
                     if (tet) k_eval_lattice<true><<<frame_blocks(h, m), 256, 0, st>>>(F, lat, o, m, gradient ? 1 : 0, d0);
                     else k_eval_lattice<false><<<frame_blocks(h, m), 256, 0, st>>>(F, lat, o, m, gradient ? 1 : 0, d0);
                     CKL();
                     h->launches++;
                     return 0;
                   }, &ce);
  return pipe_result(rc, ce, "staged lattice evaluation");
}

// ------------------------------------------------------------------ isosurface of a lattice volume (iso.cuh)
struct lsb200_iso_result {
  int device = 0;
  long long nv = 0, nt = 0;
  double* verts = nullptr;
  double* normals = nullptr;
  int* tris = nullptr;
};

int lsb200_isosurface(int device, const int64_t dims[3], const double* volume, double level, const double spacing[3],
                      lsb200_iso* out, int64_t* n_vertices, int64_t* n_triangles) {
  if (!dims || !volume || !spacing || !out || !n_vertices || !n_triangles) return fail(LSB200_ERR_ARG, "null argument");
  *out = nullptr;
  *n_vertices = *n_triangles = 0;
  if (dims[0] < 2 || dims[1] < 2 || dims[2] < 2) return fail(LSB200_ERR_ARG, "an isosurface needs at least 2 samples per axis");
  const long long n = (long long)dims[0] * dims[1] * dims[2];
  if (7 * n >= 2147483647LL) return fail(LSB200_ERR_ARG, "volume of %lld samples: the limit is 306 million", n);
  if (!(level == level)) return fail(LSB200_ERR_ARG, "NaN level");
  int ndev = lsb200_device_count();
  if (ndev <= 0) return fail(LSB200_ERR_CUDA, "no CUDA device available (the engine has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(LSB200_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  CK(cudaSetDevice(device));
  IsoGrid g{dims[0], dims[1], dims[2], {spacing[0], spacing[1], spacing[2]}, level};
  const long long ne = 7 * n, nc = (long long)(dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1);
  cudaStream_t st = nullptr;
  double* dv = nullptr;
  int *flag = nullptr, *id = nullptr, *count = nullptr, *first = nullptr;
  void* tmp = nullptr;
  lsb200_iso_result* res = new lsb200_iso_result();
  res->device = device;
  auto cleanup = [&]() {
    cudaFree(dv); cudaFree(flag); cudaFree(id); cudaFree(count); cudaFree(first); cudaFree(tmp);
    if (st) cudaStreamDestroy(st);
  };
  auto bail = [&](cudaError_t e, const char* what) {
    cleanup();
    cudaFree(res->verts); cudaFree(res->normals); cudaFree(res->tris);
    delete res;
    return fail(e == cudaErrorMemoryAllocation ? LSB200_ERR_ALLOC : LSB200_ERR_CUDA, "isosurface: %s: %s", what, cudaGetErrorString(e));
  };
#define ICK(x, what) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return bail(e_, what); } while (0)
  ICK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "stream");
  ICK(cudaMalloc(&dv, sizeof(double) * n), "volume");
  ICK(cudaMalloc(&flag, sizeof(int) * (ne + 1)), "edge flags");
  ICK(cudaMalloc(&id, sizeof(int) * (ne + 1)), "edge ids");
  ICK(cudaMalloc(&count, sizeof(int) * (nc + 1)), "cell counts");
  ICK(cudaMalloc(&first, sizeof(int) * (nc + 1)), "cell offsets");
  ICK(cudaMemcpyAsync(dv, volume, sizeof(double) * n, cudaMemcpyHostToDevice, st), "upload");
  ICK(cudaMemsetAsync(flag + ne, 0, sizeof(int), st), "flags");
  ICK(cudaMemsetAsync(count + nc, 0, sizeof(int), st), "counts");
  const int nbe = (int)std::min<long long>((ne + 255) / 256, 1 << 20), nbc = (int)std::min<long long>((nc + 255) / 256, 1 << 20);
  k_iso_edge_flags<<<nbe, 256, 0, st>>>(g, dv, flag);
  k_iso_cell_counts<<<nbc, 256, 0, st>>>(g, dv, count);
  ICK(cudaGetLastError(), "launch");
  size_t need1 = 0, need2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need1, flag, id, (int)(ne + 1), st);
  cub::DeviceScan::ExclusiveSum(nullptr, need2, count, first, (int)(nc + 1), st);
  ICK(cudaMalloc(&tmp, std::max(need1, need2)), "scan scratch");
  size_t have = std::max(need1, need2);
  ICK(cub::DeviceScan::ExclusiveSum(tmp, have, flag, id, (int)(ne + 1), st), "scan");
  have = std::max(need1, need2);
  ICK(cub::DeviceScan::ExclusiveSum(tmp, have, count, first, (int)(nc + 1), st), "scan");
  int totals[2] = {0, 0};
  ICK(cudaMemcpyAsync(&totals[0], id + ne, sizeof(int), cudaMemcpyDeviceToHost, st), "totals");
  ICK(cudaMemcpyAsync(&totals[1], first + nc, sizeof(int), cudaMemcpyDeviceToHost, st), "totals");
  ICK(cudaStreamSynchronize(st), "sync");
  res->nv = totals[0];
  res->nt = totals[1];
  if (res->nv > 0) {
    ICK(cudaMalloc(&res->verts, sizeof(double) * 3 * res->nv), "vertices");
    ICK(cudaMalloc(&res->normals, sizeof(double) * 3 * res->nv), "normals");
    k_iso_vertices<<<nbe, 256, 0, st>>>(g, dv, flag, id, res->verts, res->normals);
  }
  if (res->nt > 0) {
    ICK(cudaMalloc(&res->tris, sizeof(int) * 3 * res->nt), "triangles");
    k_iso_triangles<<<nbc, 256, 0, st>>>(g, dv, id, first, res->tris);
  }
  ICK(cudaGetLastError(), "launch");
  ICK(cudaStreamSynchronize(st), "sync");
#undef ICK
  cleanup();
  *out = res;
  *n_vertices = res->nv;
  *n_triangles = res->nt;
  return 0;
}

int lsb200_isosurface_fetch(lsb200_iso iso, double* vertices, int32_t* triangles, double* normals) {
  if (!iso) return fail(LSB200_ERR_ARG, "null isosurface");
  CK(cudaSetDevice(iso->device));
  if (vertices && iso->nv) CK(cudaMemcpy(vertices, iso->verts, sizeof(double) * 3 * iso->nv, cudaMemcpyDeviceToHost));
  if (normals && iso->nv) CK(cudaMemcpy(normals, iso->normals, sizeof(double) * 3 * iso->nv, cudaMemcpyDeviceToHost));
  if (triangles && iso->nt) CK(cudaMemcpy(triangles, iso->tris, sizeof(int) * 3 * iso->nt, cudaMemcpyDeviceToHost));
  return 0;
}

int lsb200_isosurface_free(lsb200_iso iso) {
  if (!iso) return 0;
  cudaSetDevice(iso->device);
  cudaFree(iso->verts);
  cudaFree(iso->normals);
  cudaFree(iso->tris);
  delete iso;
  return 0;
}

int lsb200_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes <= 0) return fail(LSB200_ERR_ARG, "bad argument");
  *out = nullptr;
  if (lsb200_device_count() <= 0) return fail(LSB200_ERR_CUDA, "no CUDA device available");
  cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(LSB200_ERR_ALLOC, "cudaHostAlloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
  return 0;
}
int lsb200_host_free(void* p) {
  if (p) cudaFreeHost(p);
  return 0;
}

// Diagnostics: the launch plan of the regularisation operator as text (after lsb200_finalize_system: the tuned plan)
int lsb200_describe_plan(lsb200_handle h, char* out, int64_t capacity) {
  if (!h || !out || capacity < 1) return fail(LSB200_ERR_ARG, "bad argument");
  if (h->kind == LSB200_SUPPORT_GRID && h->fused) {
    const FusedPlan& pl = h->fz.pl;
    snprintf(out, (size_t)capacity,
             "k_fdi_fused: tiles 64x%d (%d x %d), %d box chunks of %d planes (first %d), %d box CTAs, strips of %d planes "
             "(%d x-strip + %d y-strip CTAs), %d CTAs per SM, %s staging, %s",
             4 * h->fz.tyt, pl.tiles_x, pl.tiles_y, pl.chunks, pl.cz, pl.lo_len, pl.n_box, pl.scz, pl.n_xs, pl.n_ys, h->fz.minb,
             (h->want_tma && h->g.pitch % 2 == 0) ? "TMA" : "cp.async", h->tuned ? "chosen by timing" : "planner default");
  } else if (h->kind == LSB200_SUPPORT_GRID) {
    snprintf(out, (size_t)capacity, "k_fdi_normal: grid %u x %u x %u, z chunks of %d planes", h->sgrid.x, h->sgrid.y, h->sgrid.z, h->cz);
  } else if (h->p1_split && !h->small_ok) {
    const P1MarchPlan& pl = h->p1_splan;
    snprintf(out, (size_t)capacity, "k_p1_march: tiles 64x8 (%d x %d), 4 warps, %d z chunks of %d planes, %d CTAs, beside k_p1_boundary (%lld nodes) on a side stream",
             pl.tiles_x, pl.tiles_y, pl.chunks, pl.cz, pl.n_cta, (long long)(pl.b.nbz + pl.b.nby + pl.b.nbx));
  } else if (h->p1_march) {
    const P1MarchPlan& pl = h->p1_mplan;
    snprintf(out, (size_t)capacity, "k_p1_march: tiles 64x%d (%d x %d), %d + %d warps, %d z chunks of %d planes, %d CTAs (+%d boundary-only), %d CTAs per SM",
             2 * h->p1_march_nw, pl.tiles_x, pl.tiles_y, h->p1_march_nw, h->p1_march_nb, pl.chunks, pl.cz, pl.n_cta, pl.b.n_bnd_cta, h->p1_march_minb);
  } else if (h->p1_tiles) {
    const P1TilePlan& pl = h->p1_plan;
    snprintf(out, (size_t)capacity, "k_p1_tiles: tile configuration %d, %d x %d x %d tiles, %d boundary-only CTAs", h->p1_tile_cfg,
             pl.tiles_x, pl.tiles_y, pl.tiles_z, pl.n_bnd_cta);
  } else {
    snprintf(out, (size_t)capacity, "k_p1_normal: one thread per node");
  }
  return 0;
}

// Diagnostics: one launch of the regularisation operator (q = N p on the handle's own vectors) with the per-CTA trace
// switched on; out[3 b .. 3 b + 2] = {start ns, end ns, SM id} of CTA b (csrc/common.cuh, g_cta_trace).
int lsb200_debug_cta_trace(lsb200_handle h, int64_t max_ctas, uint64_t* out, int64_t* n_ctas) {
  if (!h || !out || !n_ctas || max_ctas < 1) return fail(LSB200_ERR_ARG, "bad argument");
  CK(cudaSetDevice(h->device));
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  long long ctas = 0;
  if (h->kind == LSB200_SUPPORT_GRID) {
    if (!h->fused) return fail(LSB200_ERR_STATE, "the per-CTA trace needs the fused stencil plan");
    ctas = (long long)h->fz.pl.n_box + h->fz.pl.n_xs + h->fz.pl.n_ys;
  } else {
    if (!h->p1_tiles) return fail(LSB200_ERR_STATE, "the per-CTA trace needs the tiled P1 plan");
    ctas = h->p1_split ? (long long)h->p1_splan.n_cta
           : h->p1_march ? (long long)h->p1_mplan.n_cta + h->p1_mplan.b.n_bnd_cta
                       : (long long)h->p1_plan.n_bnd_cta + (long long)h->p1_plan.tiles_x * h->p1_plan.tiles_y * h->p1_plan.tiles_z;
  }
  if (ctas > max_ctas) return fail(LSB200_ERR_ARG, "the launch has %lld CTAs, the buffer holds %lld", ctas, (long long)max_ctas);
  unsigned long long* buf = nullptr;
  CK(cudaMalloc(&buf, sizeof(unsigned long long) * 3 * (size_t)ctas));
  CK(cudaMemset(buf, 0, sizeof(unsigned long long) * 3 * (size_t)ctas));
  CK(cudaMemcpyAsync(h->p, h->rhs, sizeof(double) * (size_t)h->g.nall, cudaMemcpyDeviceToDevice, h->stream));
  for (int rep = 0; rep < 3 && rc == 0; ++rep) rc = launch_regularisation(h, h->p, h->q, 0);  // warm
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpyToSymbol(g_cta_trace, &buf, sizeof(buf)));
  if (rc == 0) rc = launch_regularisation(h, h->p, h->q, 0);
  const cudaError_t se = cudaStreamSynchronize(h->stream);
  unsigned long long* none = nullptr;
  cudaMemcpyToSymbol(g_cta_trace, &none, sizeof(none));
  if (rc == 0 && se == cudaSuccess) cudaMemcpy(out, buf, sizeof(unsigned long long) * 3 * (size_t)ctas, cudaMemcpyDeviceToHost);
  cudaFree(buf);
  h->have_c = false;
  if (se != cudaSuccess) return fail(LSB200_ERR_CUDA, "trace launch: %s", cudaGetErrorString(se));
  *n_ctas = ctas;
  return rc;
}

// Diagnostics: `iters` iterations of the persistent CG loop with CTA 0's phase clock switched on; out_us[0..7] = mean
// microseconds per iteration CTA 0 spent in: p update, barrier, operator, barrier, data rows, barrier + p.q, x/r update,
// barrier + rho (a barrier's figure includes the wait for the slowest CTA of the phase before it).
int lsb200_debug_small_phases(lsb200_handle h, int iters, double* out_us) {
  if (!h || !out_us || iters < 1) return fail(LSB200_ERR_ARG, "bad argument");
  CK(cudaSetDevice(h->device));
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  if (!h->small_ok) return fail(LSB200_ERR_STATE, "the handle does not use the persistent CG loop");
  CgScalars s{};
  s.rho = 1.0; s.pq = 1.0; s.pq_part[0] = 1.0; s.beta = 0.5; s.thr2 = 0.0; s.maxiter = 1LL << 60; s.nranks = h->nranks;
  *h->sc_host = s;
  CK(cudaMemcpyAsync(h->p, h->rhs, sizeof(double) * (size_t)h->g.nall, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->r, h->rhs, sizeof(double) * (size_t)h->g.nall, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->sc, h->sc_host, sizeof(CgScalars), cudaMemcpyHostToDevice, h->stream));
  if ((rc = launch_small(h, 3))) return rc;  // warm
  unsigned long long* buf = nullptr;
  CK(cudaMalloc(&buf, sizeof(unsigned long long) * 8));
  CK(cudaMemset(buf, 0, sizeof(unsigned long long) * 8));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpyToSymbol(g_cta_trace, &buf, sizeof(buf)));
  rc = launch_small(h, iters);
  const cudaError_t se = cudaStreamSynchronize(h->stream);
  unsigned long long* none = nullptr;
  cudaMemcpyToSymbol(g_cta_trace, &none, sizeof(none));
  unsigned long long ns[8] = {0};
  if (rc == 0 && se == cudaSuccess) cudaMemcpy(ns, buf, sizeof(ns), cudaMemcpyDeviceToHost);
  cudaFree(buf);
  h->have_c = false;
  if (se != cudaSuccess) return fail(LSB200_ERR_CUDA, "phase trace launch: %s", cudaGetErrorString(se));
  for (int k = 0; k < 8; ++k) out_us[k] = 1e-3 * (double)ns[k] / iters;
  return rc;
}

int lsb200_time_kernel(lsb200_handle h, int which, int reps, double* seconds_per_launch) {
  if (!h || !seconds_per_launch || reps < 1) return fail(LSB200_ERR_ARG, "bad argument");
  CK(cudaSetDevice(h->device));
  int rc = lsb200_finalize_system(h);
  if (rc) return rc;
  // a state in which every kernel does its full work: done = 0, huge maxiter, rho = pq = 1
  CgScalars s;
  memset(&s, 0, sizeof(s));
  s.rho = 1.0; s.pq = 1.0; s.pq_part[0] = 1.0; s.beta = 0.5; s.thr2 = 0.0; s.maxiter = 1LL << 60; s.nranks = h->nranks;
  s.epoch = (++h->solve_epoch) << 32;
  *h->sc_host = s;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  // p = rhs so that the vectors hold representative data
  CK(cudaMemcpyAsync(h->p, h->rhs, sizeof(double) * (size_t)h->g.nall, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->r, h->rhs, sizeof(double) * (size_t)h->g.nall, cudaMemcpyDeviceToDevice, h->stream));
  float total_ms = 0.f;
  if (which == 1 && h->small_ok) {
    // small systems: the iterations of one persistent launch (csrc/cg_small.cuh), as lsb200_solve_cg enqueues them
    CK(cudaMemcpyAsync(h->sc, h->sc_host, sizeof(CgScalars), cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_small(h, 3))) return rc;
    CK(cudaEventRecord(e0, h->stream));
    if ((rc = launch_small(h, reps))) return rc;
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&total_ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *seconds_per_launch = total_ms * 1e-3 / reps;
    h->have_c = false;
    return 0;
  }
  // three warm-up launches, then `reps` launches back to back between two events, as inside the CG loop
  for (int rep = -3; rep < reps; ++rep) {
    if (rep == 0) CK(cudaEventRecord(e0, h->stream));
    // the whole iteration advances the scalars: reset them every time; a single kernel does not
    if (which == 1 || rep == -3) CK(cudaMemcpyAsync(h->sc, h->sc_host, sizeof(CgScalars), cudaMemcpyHostToDevice, h->stream));
    if (which == 0) rc = launch_regularisation(h, h->p, h->q, 1);
    else if (which == 1) rc = enqueue_iteration(h);
    else if (which == 2) rc = launch_rows(h, h->p, h->q, 1);
    else if (which == 3 || which == 4) {
      const long long n = h->g.nloc;
      const double* s2 = h->apply_scaling ? owned_ptr(h, h->s2) : nullptr;
      if (which == 3)  // alone, without the peers' pushes it would wait for: p.q from the scalar block (peers = null)
        k_cg_update_xr<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, owned_ptr(h, h->p), owned_ptr(h, h->q), owned_ptr(h, h->x),
                                                                            owned_ptr(h, h->r), s2, h->partials + 2 * MAX_PARTIALS,
                                                                            h->counters + 2, h->sc, nullptr);
      else
        k_cg_update_p<VEC_NT><<<vec_blocks(h, n), VEC_NT, 0, h->stream>>>(n, s2, owned_ptr(h, h->r), owned_ptr(h, h->p), h->sc);
      CKL();
      h->launches++;
    } else return fail(LSB200_ERR_ARG, "unknown kernel selector %d", which);
    if (rc) return rc;
  }
  CK(cudaEventRecord(e1, h->stream));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&total_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *seconds_per_launch = total_ms * 1e-3 / reps;
  h->have_c = false;
  return 0;
}

}  // extern "C"

// Host-array entry points (lsb200_eval_value, lsb200_frame_eval, lsb200_faulted_eval, lsb200_eval_lattice ...): the
// caller's pageable arrays go through page-locked staging in chunks, three chunks in flight on three streams, so that the
// host -> device copy of chunk i+1, the kernel of chunk i and the device -> host copy of chunk i-1 overlap, and the
// pageable <-> page-locked copies on the host (the slowest leg: one core moves ~10 GB/s) are spread over a few threads.
// Replaces the one-stream copy / kernel / copy / synchronise loop per 16 M points of round 1 (2 GB/s, profiles/r02d).
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace lsb {

inline void par_memcpy(void* dst, const void* src, size_t bytes) {
  constexpr size_t kMin = 1u << 20;  // bytes per thread below which starting it costs more than it moves
  if (bytes < 4 * kMin) {
    memcpy(dst, src, bytes);
    return;
  }
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int nt = (int)std::min<size_t>(std::min<unsigned>(hw, 16u), bytes / kMin);
  const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) {
    const size_t o = per * t;
    if (o >= bytes) break;
    th.emplace_back([=]() { memcpy((char*)dst + o, (const char*)src + o, std::min(per, bytes - o)); });
  }
  memcpy(dst, src, std::min(per, bytes));
  for (auto& t : th) t.join();
}

struct HostPipe {
  static constexpr int SLOTS = 3;
  struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    char *pin_in = nullptr, *pin_out = nullptr, *dev_in = nullptr, *dev_out = nullptr, *dev_tmp = nullptr;
    bool busy = false;
    long long first = 0, count = 0;  // the chunk whose result waits in pin_out
  } s[SLOTS];
  size_t cap_in = 0, cap_out = 0, cap_tmp = 0;
  bool created = false;

  cudaError_t ensure(size_t in_bytes, size_t out_bytes, size_t tmp_bytes) {
    cudaError_t e;
    if (!created) {
      for (Slot& q : s) {
        if ((e = cudaStreamCreateWithFlags(&q.st, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming)) != cudaSuccess) return e;
      }
      created = true;
    }
    auto grow = [&](size_t need, size_t& cap, bool pinned, char* Slot::*pin, char* Slot::*dev) -> cudaError_t {
      if (need <= cap) return cudaSuccess;
      const size_t ncap = std::max(need, cap + cap / 2);
      for (Slot& q : s) {
        if (pinned && pin) {
          if (q.*pin) cudaFreeHost(q.*pin);
          q.*pin = nullptr;
          if ((e = cudaHostAlloc((void**)&(q.*pin), ncap, cudaHostAllocDefault)) != cudaSuccess) return e;
        }
        if (q.*dev) cudaFree(q.*dev);
        q.*dev = nullptr;
        if ((e = cudaMalloc((void**)&(q.*dev), ncap)) != cudaSuccess) return e;
      }
      cap = ncap;
      return cudaSuccess;
    };
    if ((e = grow(in_bytes, cap_in, true, &Slot::pin_in, &Slot::dev_in)) != cudaSuccess) return e;
    if ((e = grow(out_bytes, cap_out, true, &Slot::pin_out, &Slot::dev_out)) != cudaSuccess) return e;
    if ((e = grow(tmp_bytes, cap_tmp, false, nullptr, &Slot::dev_tmp)) != cudaSuccess) return e;
    return cudaSuccess;
  }

  void destroy() {
    for (Slot& q : s) {
      if (q.st) cudaStreamSynchronize(q.st);
      if (q.pin_in) cudaFreeHost(q.pin_in);
      if (q.pin_out) cudaFreeHost(q.pin_out);
      if (q.dev_in) cudaFree(q.dev_in);
      if (q.dev_out) cudaFree(q.dev_out);
      if (q.dev_tmp) cudaFree(q.dev_tmp);
      if (q.done) cudaEventDestroy(q.done);
      if (q.st) cudaStreamDestroy(q.st);
      q = Slot{};
    }
    cap_in = cap_out = cap_tmp = 0;
    created = false;
  }

  // n points; per point `in_b` bytes in (0: nothing to upload), `o0_b` bytes to out0 and `o1_b` bytes to out1 (0 / null:
  // unused), `tmp_b` bytes of device scratch.  launch(stream, first, m, dev_in, dev_out0, dev_out1, dev_tmp) enqueues
  // the kernels of one chunk and returns 0 or an error code (the pipe drains and returns it).
  template <class Launch>
  int run(long long n, long long chunk, size_t in_b, const void* in, size_t o0_b, void* out0, size_t o1_b, void* out1,
          size_t tmp_b, Launch launch, cudaError_t* cuda_err) {
    *cuda_err = cudaSuccess;
    if (n <= 0) return 0;
    chunk = std::max<long long>(1, std::min(chunk, n));
    if ((*cuda_err = ensure(in_b * chunk, (o0_b + o1_b) * chunk, tmp_b * chunk)) != cudaSuccess) return -1;
    int rc = 0;
    auto retire = [&](Slot& q) {  // result of the chunk in this slot -> caller's arrays
      if (!q.busy) return;
      const cudaError_t e = cudaEventSynchronize(q.done);
      if (e != cudaSuccess && *cuda_err == cudaSuccess) *cuda_err = e;
      if (e == cudaSuccess && rc == 0) {
        par_memcpy((char*)out0 + o0_b * q.first, q.pin_out, o0_b * q.count);
        if (o1_b && out1) par_memcpy((char*)out1 + o1_b * q.first, q.pin_out + o0_b * q.count, o1_b * q.count);
      }
      q.busy = false;
    };
    int slot = 0;
    for (long long o = 0; o < n && rc == 0 && *cuda_err == cudaSuccess; o += chunk, slot = (slot + 1) % SLOTS) {
      Slot& q = s[slot];
      retire(q);
      if (rc || *cuda_err != cudaSuccess) break;
      const long long m = std::min(chunk, n - o);
      if (in_b) {
        par_memcpy(q.pin_in, (const char*)in + in_b * o, in_b * m);
        if ((*cuda_err = cudaMemcpyAsync(q.dev_in, q.pin_in, in_b * m, cudaMemcpyHostToDevice, q.st)) != cudaSuccess) break;
      }
      rc = launch(q.st, o, m, (double*)q.dev_in, (double*)q.dev_out, o1_b ? (double*)(q.dev_out + o0_b * m) : nullptr,
                  (double*)q.dev_tmp);
      if (rc) break;
      if ((*cuda_err = cudaMemcpyAsync(q.pin_out, q.dev_out, (o0_b + o1_b) * m, cudaMemcpyDeviceToHost, q.st)) != cudaSuccess) break;
      if ((*cuda_err = cudaEventRecord(q.done, q.st)) != cudaSuccess) break;
      q.busy = true;
      q.first = o;
      q.count = m;
    }
    for (int t = 0; t < SLOTS; ++t) retire(s[(slot + t) % SLOTS]);  // oldest first
    if (rc == 0 && *cuda_err != cudaSuccess) rc = -1;
    return rc;
  }
};

}  // namespace lsb

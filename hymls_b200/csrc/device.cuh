// Small device/host helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/hymls_b200.h"
#include "params.hpp"

namespace hymls {

#define HY_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      throw ::hymls::Error(HYMLS_B200_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e__) + \
                                                    " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// RAII device buffer (cudaMallocAsync-free: plain cudaMalloc, sized once per pattern)
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;    // elements in use
  size_t cap = 0;  // elements allocated (grow-only: cudaMalloc/cudaFree of GB-sized buffers stall for 100s of ms)
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cap = 0;
  }
  void alloc(size_t count) {
    if (p && count <= cap) {
      n = count;
      return;
    }
    release();
    n = cap = count;
    if (count) HY_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) {
    alloc(h.size());
    if (!h.empty()) HY_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  size_t bytes() const { return n * sizeof(T); }
};

// "largest opt-in dynamic shared memory size requested so far" of one kernel, tracked PER DEVICE: function
// attributes belong to the device/context, so a process-wide flag is wrong once a process drives two devices
struct PerDeviceLimit {
  size_t v[64];
  PerDeviceLimit() { for (size_t& x : v) x = 48 * 1024; }
  bool raise(size_t want) {  // true: the attribute has to be (re)set on the current device
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (want <= v[d]) return false;
    v[d] = want;
    return true;
  }
};

extern thread_local double g_devBytes;  // running total for statistics

}  // namespace hymls

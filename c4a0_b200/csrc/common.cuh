// common.cuh — host-side helpers shared by the translation units of libc4a0_engine.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/c4a0_engine.h"

namespace c4host {

// Records a thread-local message (returned by c4a0_last_error) and passes `code` through.
int fail(int code, const char* fmt, ...);
const char* last_error();
// 0 when a CUDA device is usable; otherwise C4A0_E_CUDA with a "no CPU fallback" message.
int no_gpu_error();

inline unsigned blocks_for(size_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

template <typename T>
struct DevBuf {
  T* p = nullptr;
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

}  // namespace c4host

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (call);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return c4host::fail(C4A0_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),    \
                          __FILE__, __LINE__);                                                    \
  } while (0)

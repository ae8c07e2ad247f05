// cpf_common.h — error plumbing, device guard and small helpers shared by the translation units of libcpfftlog.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/cpfftlog.h"

namespace cpf {

void set_error(const std::string& msg);

int fail(int code, const char* fmt, ...);

#define CPF_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::cpf::fail(CPF_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

#define CPF_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != CPF_OK) return _rc; \
  } while (0)

// Makes `device` current for the scope and restores the caller's device afterwards (the caller is typically a
// PyTorch process whose current device must not change under it).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      switched = (err == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// Scratch memory comes from a PRIVATE stream-ordered pool per device (never the device's default pool, whose attributes belong to the
// host application: torch / JAX allocate next to this library).  The pool keeps at most CPF_SCRATCH_KEEP_MB (default 2048) MiB of
// unused memory across synchronisations, so that repeated calls recycle their scratch instead of going back to the driver (~100 us per
// allocation, milliseconds for GB-sized scratch) while a large one-off call cannot hold on to tens of GB; cpf_trim() returns it all.
cudaMemPool_t scratch_pool(int device);      // cpf_fftlog.cu; nullptr on failure (alloc then reports the CUDA error)

struct ScratchBuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = scratch_pool(dev);
    if (!pool) return cudaErrorMemoryAllocation;
    return cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, pool, stream);
  }
  ~ScratchBuf() {
    if (p) cudaFreeAsync(p, s);
  }
};

inline bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
inline int ilog2(long long v) {
  int l = 0;
  while ((1LL << (l + 1)) <= v) ++l;
  return l;
}

}  // namespace cpf

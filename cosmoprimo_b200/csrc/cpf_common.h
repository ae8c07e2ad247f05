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

// Stream-ordered scratch buffer: freed (stream-ordered) when it goes out of scope.
struct ScratchBuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    keep_pool_warm();
    return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
  }
  // The default memory pool hands its free blocks back to the driver at every synchronisation (release threshold 0), which
  // makes each cudaMallocAsync after a sync a fresh driver allocation (~100 us, milliseconds for GB-sized scratch).  Raise
  // the threshold once per device so that scratch memory is recycled inside the pool.
  static void keep_pool_warm() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t threshold = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done[dev] = true;
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

// cpf_peak.cu — cpf_measure_fp64_peak: sustained DFMA rate of the device.  MEASURED_PEAKS.json (driver-written) has
// HBM and bf16 numbers only; the FFTLog roofline needs the fp64 FMA peak, so bench.py measures it with this.
#include "cpf_common.h"

namespace cpf {

// 16 independent FMA chains per thread, ITERS rounds: enough ILP to saturate the fp64 pipe at any occupancy
template <int ITERS>
__global__ void __launch_bounds__(256) dfma_chain_kernel(double* out, const double a, const double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace cpf

extern "C" int cpf_measure_fp64_peak(int device, double* flops_per_s) {
  using namespace cpf;
  if (!flops_per_s) return fail(CPF_EINVAL, "cpf_measure_fp64_peak: null pointer");
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_measure_fp64_peak: device %d out of range", device);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CPF_CUDA(cudaGetDeviceProperties(&prop, device));
  double* d_out = nullptr;
  CPF_CUDA(cudaMalloc(&d_out, sizeof(double)));
  cudaEvent_t e0, e1;
  CPF_CUDA(cudaEventCreate(&e0));
  CPF_CUDA(cudaEventCreate(&e1));
  constexpr int ITERS = 4096;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double best = 0.;
  for (int rep = 0; rep < 6; ++rep) {   // first reps are warm-up (clock ramp)
    CPF_CUDA(cudaEventRecord(e0));
    dfma_chain_kernel<ITERS><<<blocks, threads>>>(d_out, 0.999999, 1e-9);
    CPF_CUDA(cudaEventRecord(e1));
    CPF_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CPF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16.0 * ITERS * (double)blocks * threads / (ms * 1e-3);
    if (rep >= 2 && flops > best) best = flops;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *flops_per_s = best;
  return CPF_OK;
}

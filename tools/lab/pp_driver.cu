// pp_driver.cu — A/B check and timing of the FFTLog kernels through the C ABI (libcpfftlog.so): the ping-pong kernel
// (CPF_FFTLOG_KERNEL=pp0|pp1|pp2) against the per-pair kernel on random tables.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pp_driver pp_driver.cu -L../../cosmoprimo_b200 -lcpfftlog
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/cpfftlog.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
#define CPF(x) do { int rc = (x); if (rc != 0) { printf("cpf error %d at %d: %s\n", rc, __LINE__, cpf_last_error()); exit(1);} } while (0)

static double urand() { return (rand() + 0.5) / ((double)RAND_MAX + 1.); }

static const char* g_only = nullptr;   // run only this kernel (plus "fast" as the reference)

static int run_case(int n, int P, long long batch, int in_has_P, int reps) {
  int N = 1; while (N < 2 * n) N *= 2;
  const int npad = N - n, in_left = npad / 2, out_left = npad - npad / 2, nb = N / 2 + 1;
  std::vector<double> pre((size_t)P * N), post((size_t)P * N), u((size_t)P * nb * 2);
  for (auto& x : pre) x = 0.5 + urand();
  for (auto& x : post) x = 0.5 + urand();
  for (auto& x : u) x = 2. * urand() - 1.;
  cpf_plan* plan = nullptr;
  CPF(cpf_plan_create(&plan, n, N, P, in_left, out_left, pre.data(), u.data(), post.data(), nullptr, 0));
  const size_t in_elems = (size_t)batch * (in_has_P ? P : 1) * n, out_elems = (size_t)batch * P * n;
  std::vector<double> h_in(in_elems);
  for (auto& x : h_in) x = 2. * urand() - 1.;
  double *d_in, *d_out;
  CK(cudaMalloc(&d_in, in_elems * 8)); CK(cudaMalloc(&d_out, out_elems * 8));
  CK(cudaMemcpy(d_in, h_in.data(), in_elems * 8, cudaMemcpyHostToDevice));
  std::vector<double> ref(out_elems), got(out_elems);
  const char* kernels[] = {"fast", "pp0", "pp1", "pp2", "persistent", "pp3", "stream"};
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int bad = 0;
  for (int k = 0; k < 7; ++k) {
    if (g_only && k > 0 && strcmp(g_only, kernels[k]) != 0) continue;
    setenv("CPF_FFTLOG_KERNEL", kernels[k], 1);
    setenv("CPF_FFTLOG_PERSISTENT", k == 4 ? "1" : "0", 1);
    CK(cudaMemset(d_out, 0xff, out_elems * 8));
    CPF(cpf_fftlog(plan, d_in, batch, in_has_P, 0, 0., 0, 0., 0, d_out, 1, 1, nullptr));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), d_out, out_elems * 8, cudaMemcpyDeviceToHost));
    double err = 0., scale = 0.;
    if (k == 0) ref = got;
    for (size_t i = 0; i < out_elems; ++i) { double d = fabs(got[i] - ref[i]); if (!(d <= err)) err = d; if (fabs(ref[i]) > scale) scale = fabs(ref[i]); }
    for (int i = 0; i < 3; ++i) CPF(cpf_fftlog(plan, d_in, batch, in_has_P, 0, 0., 0, 0., 0, d_out, 1, 1, nullptr));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) CPF(cpf_fftlog(plan, d_in, batch, in_has_P, 0, 0., 0, 0., 0, d_out, 1, 1, nullptr));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double tr = (double)batch * P;
    const bool ok = err <= 1e-13 * scale;
    if (!ok) ++bad;
    printf("n=%5d P=%d batch=%7lld hasP=%d %-10s %9.3f us %7.2f M tr/s  %7.0f SM-cyc/pair  max|d|/max|ref| = %.2e %s\n", n, P, batch, in_has_P,
           kernels[k], ms * 1e3, tr / (ms * 1e-3) / 1e6, ms * 1e-3 * 1.965e9 * 148 / (tr / 2), err / scale, ok ? "ok" : "MISMATCH");
  }
  CPF(cpf_plan_destroy(plan));
  CK(cudaFree(d_in)); CK(cudaFree(d_out));
  return bad;
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 50;
  int bad = 0;
  if (argc > 2) g_only = argv[2];
  if (argc > 3) {   // single case: n P batch
    bad = run_case(atoi(argv[3]), atoi(argv[4]), atoll(argv[5]), 1, reps);
    return bad != 0;
  }
  bad += run_case(2048, 3, 4096, 1, reps);
  bad += run_case(2048, 1, 100000, 1, reps / 5 + 1);
  bad += run_case(2048, 3, 4095, 0, reps);
  bad += run_case(2000, 3, 4097, 1, reps);
  bad += run_case(1919, 2, 5001, 0, reps);
  bad += run_case(2048, 1, 7, 1, 5);
  bad += run_case(2048, 3, 1, 1, 5);
  bad += run_case(2048, 2, 300, 1, 5);
  bad += run_case(1024, 3, 8192, 1, reps);
  bad += run_case(1000, 1, 8191, 1, reps);
  bad += run_case(512, 3, 16384, 1, reps);
  bad += run_case(512, 1, 5, 1, 5);
  printf(bad ? "FAILED: %d mismatching runs\n" : "all runs match\n", bad);
  return bad != 0;
}

// e2e_lab.cu — host-buffer (pinned) cpf_fftlog call under different staging settings, and the raw copy-only pipeline
// with the same chunking for comparison.  Build: see build_lab.sh; run on the GPU box.
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

int main() {
  const int n = 2048, N = 4096, P = 3;
  const long long batch = 4096;
  const int nb = N / 2 + 1;
  std::vector<double> pre((size_t)P * N), post((size_t)P * N), u((size_t)P * nb * 2);
  for (auto& x : pre) x = 0.5 + urand();
  for (auto& x : post) x = 0.5 + urand();
  for (auto& x : u) x = 2. * urand() - 1.;
  cpf_plan* plan = nullptr;
  CPF(cpf_plan_create(&plan, n, N, P, 1024, 1024, pre.data(), u.data(), post.data(), nullptr, 0));
  const size_t elems = (size_t)batch * P * n, bytes = elems * 8;
  double *h_in, *h_out;
  CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
  for (size_t i = 0; i < elems; ++i) h_in[i] = 2. * urand() - 1.;
  struct Cfg { const char* cap; const char* small; const char* nbuf; const char* path; };
  const Cfg cfgs[] = {{"16384", "1024", "4", "staged"}, {"32768", "32768", "3", "staged"}, {"8192", "1024", "4", "staged"},
                      {"8192", "8192", "4", "staged"},  {"4096", "1024", "8", "staged"},   {"16384", "2048", "8", "staged"},
                      {"16384", "16384", "2", "staged"}, {"65536", "65536", "4", "staged"}, {"16384", "1024", "4", "zerocopy"},
                      {"16384", "1024", "4", "mixed"},  {"4096", "1024", "4", "mixed"},   {"32768", "2048", "4", "mixed"},
                      {"8192", "1024", "8", "mixed"},   {"65536", "4096", "3", "mixed"}};
  for (const Cfg& c : cfgs) {
    setenv("CPF_STAGE_CAP_KB", c.cap, 1); setenv("CPF_STAGE_SMALL_KB", c.small, 1); setenv("CPF_STAGE_NBUF", c.nbuf, 1);
    setenv("CPF_HOST_PATH", c.path, 1);
    double best = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      CPF(cpf_fftlog(plan, h_in, batch, 1, 0, 0., 0, 0., 0, h_out, 0, 0, nullptr));
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    printf("%-8s cap %6s KB small %6s KB nbuf %s : %7.3f ms  %6.2f GB/s each way  %5.2f M tr/s\n", c.path, c.cap, c.small, c.nbuf, best,
           bytes / best / 1e6, batch * P / best / 1e3);
  }
  // copy-only pipelines with 16 MB chunks: (a) independent H2D and D2H streams, (b) per-chunk H2D -> D2H chains on 4 streams
  {
    double *d_a, *d_b;
    CK(cudaMalloc(&d_a, bytes)); CK(cudaMalloc(&d_b, bytes));
    const size_t chunk = 16u << 20;
    const int nch = (int)((bytes + chunk - 1) / chunk);
    cudaStream_t s[8];
    for (int i = 0; i < 8; ++i) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 3; ++mode) {
      double best = 1e30;
      for (int rep = 0; rep < 5; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s[0]));
        for (int i = 1; i < 8; ++i) CK(cudaStreamWaitEvent(s[i], e0, 0));
        for (int c = 0; c < nch; ++c) {
          const size_t off = (size_t)c * chunk, len = bytes - off < chunk ? bytes - off : chunk;
          if (mode == 0) {
            CK(cudaMemcpyAsync((char*)d_a + off, (char*)h_in + off, len, cudaMemcpyHostToDevice, s[0]));
            CK(cudaMemcpyAsync((char*)h_out + off, (char*)d_b + off, len, cudaMemcpyDeviceToHost, s[1]));
          } else {
            const int nst = mode == 1 ? 4 : 8;
            cudaStream_t st = s[c % nst];
            CK(cudaMemcpyAsync((char*)d_a + off, (char*)h_in + off, len, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync((char*)h_out + off, (char*)d_a + off, len, cudaMemcpyDeviceToHost, st));
          }
        }
        for (int i = 1; i < 8; ++i) { cudaEvent_t ev; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); CK(cudaEventRecord(ev, s[i])); CK(cudaStreamWaitEvent(s[0], ev, 0)); CK(cudaEventDestroy(ev)); }
        CK(cudaEventRecord(e1, s[0])); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      printf("copy only, %s: %7.3f ms  %6.2f GB/s each way\n", mode == 0 ? "independent H2D / D2H streams" : mode == 1 ? "H2D->D2H chains on 4 streams" : "H2D->D2H chains on 8 streams",
             best, bytes / best / 1e6);
    }
  }
  return 0;
}

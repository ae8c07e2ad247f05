// CPU emulation of the three-pass register FFT (cpf_fft_core.h): the per-thread phase functions are executed for
// every thread id in turn, with a plain array standing in for shared memory, and the result is compared with an
// O(N^2) long-double DFT.  Catches index-mapping / twiddle / butterfly mistakes without a GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../cosmoprimo_b200/csrc/cpf_fft_core.h"

using namespace cpf;

template <int R1, bool HALF_IN, bool HALF_OUT>
double run() {
  typedef Geo<R1> G;
  const int N = G::N, T = G::T;
  std::vector<double2> x(N), tw1(6 * 256), tw2(6 * 16), S(G::SMEM_ELEMS);
  const long double PI = acosl(-1.0L);
  const int expo[6] = {1, 2, 3, 4, 8, 12};
  for (int s = 0; s < 6; ++s)
    for (int n2 = 0; n2 < 256; ++n2) {
      long double a = -2 * PI * ((long long)expo[s] * n2 % N) / N;
      tw1[s * 256 + n2] = mk2((double)cosl(a), (double)sinl(a));
    }
  for (int s = 0; s < 6; ++s)
    for (int m2 = 0; m2 < 16; ++m2) {
      long double a = -2 * PI * (expo[s] * m2) / 256;
      tw2[s * 16 + m2] = mk2((double)cosl(a), (double)sinl(a));
    }
  srand(1234 + R1);
  for (int j = 0; j < N; ++j) {
    x[j] = mk2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    if (HALF_IN && j >= N / 2) x[j] = mk2(0, 0);
  }
  std::vector<double2> v(T * 16);
  for (int t = 0; t < T; ++t)
    for (int r = 0; r < 16; ++r) v[t * 16 + r] = x[t + T * r];
  for (int t = 0; t < T; ++t) fft_pass1<R1, HALF_IN>(t, *(double2(*)[16]) & v[t * 16], S.data(), tw1.data());
  for (int t = 0; t < T; ++t) fft_pass2<R1>(t, S.data(), tw2.data());
  for (int t = 0; t < T; ++t) fft_pass3<R1, HALF_OUT>(t, *(double2(*)[16]) & v[t * 16], S.data());
  // reference
  std::vector<long double> c(N), s(N);
  for (int j = 0; j < N; ++j) { c[j] = cosl(2 * PI * j / N); s[j] = sinl(2 * PI * j / N); }
  double maxerr = 0, maxabs = 0;
  for (int k = 0; k < N; ++k) {
    if (HALF_OUT && (k / T) >= 8) continue;
    long double re = 0, im = 0;
    for (int j = 0; j < N; ++j) {
      int idx = (int)(((long long)j * k) % N);
      re += x[j].x * c[idx] + x[j].y * s[idx];
      im += x[j].y * c[idx] - x[j].x * s[idx];
    }
    const double2 got = v[(k % T) * 16 + k / T];
    maxerr = fmax(maxerr, fmax(fabs(got.x - (double)re), fabs(got.y - (double)im)));
    maxabs = fmax(maxabs, fmax(fabsl(re), fabsl(im)));
  }
  return maxerr / maxabs;
}

int main() {
  int bad = 0;
#define CHECK(R1, HI, HO) { double e = run<R1, HI, HO>(); printf("R1=%d half_in=%d half_out=%d rel_err=%.3e\n", R1, HI, HO, e); if (!(e < 1e-14)) bad++; }
  CHECK(4, false, false) CHECK(4, true, false) CHECK(4, false, true) CHECK(4, true, true)
  CHECK(8, false, false) CHECK(8, true, false) CHECK(8, false, true) CHECK(8, true, true)
  CHECK(16, false, false) CHECK(16, true, false) CHECK(16, false, true) CHECK(16, true, true)
  printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}

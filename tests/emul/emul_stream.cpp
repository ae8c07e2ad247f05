// CPU emulation of the "stream" FFTLog kernel's data flow (cpf_stream_core.h): the per-thread phases are executed for
// every thread id in turn between the synchronisation points the kernel has (group barrier / warp barrier), with a
// plain array standing in for shared memory and for the tensor-memory tables, and g = FFT(ut .* FFT(z)) on the
// lower half of the output window is compared with O(N^2) long-double DFTs.  The warp-level phases are run in a
// scrambled warp order to check that they only depend on data of their own warp.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../cosmoprimo_b200/csrc/cpf_stream_core.h"

using namespace cpf;

struct HostTables {
  const double* tab;   // this thread's [ST_NTAB][32] doubles
  template <int TABLE, int SET> void issue(int, int) {}
  void wait(int) {}
  template <int SET> double getd(int table, int ch, int, int i) const { return tab[table * 32 + 8 * ch + i]; }
  template <int SET> double2 get(int table, int ch, int, int i) const { return mk2(tab[table * 32 + 8 * ch + 2 * i], tab[table * 32 + 8 * ch + 2 * i + 1]); }
};

static double2 root(long long num, long long den) {
  const long double a = -2.0L * acosl(-1.0L) * (long double)(num % den) / (long double)den;
  return mk2((double)cosl(a), (double)sinl(a));
}

int main() {
  const int N = 4096, T = 256;
  std::vector<double2> z(N), ut(N), S(ST_GROUP_ELEMS);
  std::vector<double> M(512), tw((size_t)3 * 32 * T), tabs((size_t)T * ST_NTAB * 32);
  srand(4321);
  for (int j = 0; j < N; ++j) {
    z[j] = j < N / 2 ? mk2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5) : mk2(0, 0);
    ut[j] = mk2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
  }
  st_build_tables(tw.data(), M.data());          // the very tables the library uploads
  for (int t = 0; t < T; ++t) {
    const int H = t >> 4, L = t & 15;
    double* tb = &tabs[(size_t)t * ST_NTAB * 32];
    const int region_of[3] = {ST_TW1, ST_TW2, ST_TW1B};
    for (int reg = 0; reg < 3; ++reg)
      for (int e = 0; e < 32; ++e) tb[region_of[reg] * 32 + e] = tw[((size_t)reg * 32 + e) * T + t];
    for (int k = 0; k < 16; ++k) {
      tb[ST_UT * 32 + 2 * k] = ut[H + 16 * L + 256 * k].x;
      tb[ST_UT * 32 + 2 * k + 1] = ut[H + 16 * L + 256 * k].y;
    }
  }
  // poison the exchange buffer: reads of slots nobody wrote show up as NaN
  for (auto& s : S) s = mk2(NAN, NAN);
  auto tables = [&](int t) { HostTables h; h.tab = &tabs[(size_t)t * ST_NTAB * 32]; return h; };
  const int warp_order[8] = {5, 2, 7, 0, 3, 6, 1, 4};
  std::vector<double2> out((size_t)T * 16);
  for (int pass = 0; pass < 2; ++pass) {   // two pairs back to back: the second P1 overwrites what P3' just read
    for (int t = 0; t < T; ++t) {
      double2 v[8];
      for (int r = 0; r < 8; ++r) v[r] = z[t + T * r];
      HostTables h = tables(t);
      st_p1(t, v, S.data(), h);
    }
    // group barrier
    for (int wi = 0; wi < 8; ++wi) {         // each warp runs P2 | warp barrier | P3.mul.P1' | warp barrier | P2' alone
      const int w = warp_order[wi];
      for (int t = 32 * w; t < 32 * w + 32; ++t) { HostTables h = tables(t); st_p2(t, S.data(), h); }
      for (int t = 32 * w + 31; t >= 32 * w; --t) { HostTables h = tables(t); st_p3_mul_p1(t, S.data(), h); }
      for (int t = 32 * w; t < 32 * w + 32; ++t) { HostTables h = tables(t); st_p2b(t, S.data(), h, M.data()); }
    }
    // group barrier
    for (int t = T - 1; t >= 0; --t) {
      double2 v[16];
      st_p3b(t, v, S.data(), M.data());
      for (int r = 0; r < 8; ++r) out[(size_t)t * 16 + r] = v[r];
      if (pass == 0) {                      // the next pair's P1 of this thread may run before other threads' P3'
        double2 v8[8];
        for (int r = 0; r < 8; ++r) v8[r] = z[t + T * r];
        HostTables h = tables(t);
        st_p1(t, v8, S.data(), h);
      }
    }
  }
  // reference: g = FFT(ut .* FFT(z)), long double
  const long double PI = acosl(-1.0L);
  std::vector<long double> c(N), s(N);
  for (int j = 0; j < N; ++j) { c[j] = cosl(2 * PI * j / N); s[j] = sinl(2 * PI * j / N); }
  std::vector<long double> yr(N), yi(N);
  for (int k = 0; k < N; ++k) {
    long double re = 0, im = 0;
    for (int j = 0; j < N / 2; ++j) {
      const int idx = (int)(((long long)j * k) % N);
      re += z[j].x * c[idx] + z[j].y * s[idx];
      im += z[j].y * c[idx] - z[j].x * s[idx];
    }
    yr[k] = re * ut[k].x - im * ut[k].y;
    yi[k] = re * ut[k].y + im * ut[k].x;
  }
  double maxerr = 0, maxabs = 0;
  for (int k = 0; k < N / 2; ++k) {
    long double re = 0, im = 0;
    for (int j = 0; j < N; ++j) {
      const int idx = (int)(((long long)j * k) % N);
      re += yr[j] * c[idx] + yi[j] * s[idx];
      im += yi[j] * c[idx] - yr[j] * s[idx];
    }
    const double2 got = out[(size_t)(k % T) * 16 + k / T];
    const double e = fmax(fabs(got.x - (double)re), fabs(got.y - (double)im));
    if (!(e <= maxerr)) maxerr = e;   // NaN-propagating max
    maxabs = fmax(maxabs, fmax(fabsl(re), fabsl(im)));
  }
  printf("stream flow: rel_err=%.3e\n", maxerr / maxabs);
  const bool ok = maxerr / maxabs < 1e-14;
  printf(ok ? "OK\n" : "FAIL\n");
  return ok ? 0 : 1;
}

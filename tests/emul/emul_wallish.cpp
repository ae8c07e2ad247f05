// CPU emulation of the fused Wallish2018 kernel (csrc/cpf_wallish.cu :: wallish_fused_kernel): the per-thread phase
// functions of cpf_wallish_core.h / cpf_fft_core.h are run for all 256 thread ids in turn, phase by phase, with plain
// arrays standing in for shared memory.  Input: a binary file with klin[4096], pk_a[4096], pk_b[4096]; output: a binary
// file with the DST-II coefficients, second derivatives, boxes, cut coefficients and exp(DST-III)/k for both columns.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../cosmoprimo_b200/csrc/cpf_wallish_core.h"

using namespace cpf;
typedef WallishGeo G;

static std::vector<double2> tw1(6 * 256), tw2(6 * 16), twd(G::N);
static double wtab[32];

static void tables() {
  const long double PI = acosl(-1.0L);
  const int expo[6] = {1, 2, 3, 4, 8, 12};
  for (int e = 0; e < 6; ++e) {
    for (int n2 = 0; n2 < 256; ++n2) { long double a = -2 * PI * ((long long)expo[e] * n2 % G::N) / G::N; tw1[e * 256 + n2] = mk2((double)cosl(a), (double)sinl(a)); }
    for (int m2 = 0; m2 < 16; ++m2) { long double a = -2 * PI * (expo[e] * m2) / 256; tw2[e * 16 + m2] = mk2((double)cosl(a), (double)sinl(a)); }
  }
  for (int k = 0; k < G::N; ++k) { long double a = -PI * k / (2.0L * G::N); twd[k] = mk2((double)cosl(a), (double)sinl(a)); }
  wtab[0] = 1.; double c = 0.;
  for (int i = 1; i < 32; ++i) { wtab[i] = 1. / (4. - c); c = wtab[i]; }
}

typedef double2 Regs[16];

static void fft4096(std::vector<double2>& v, std::vector<double2>& S) {
  for (int t = 0; t < 256; ++t) fft_pass1<16, false>(t, *(Regs*)&v[t * 16], S.data(), tw1.data());
  for (int t = 0; t < 256; ++t) fft_pass2<16>(t, S.data(), tw2.data());
  for (int t = 0; t < 256; ++t) fft_pass3<16, false>(t, *(Regs*)&v[t * 16], S.data());
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: emul_wallish in.bin out.bin\n"); return 2; }
  tables();
  std::vector<double> klin(G::N), pa(G::N), pb(G::N);
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  if (fread(klin.data(), 8, G::N, f) != (size_t)G::N || fread(pa.data(), 8, G::N, f) != (size_t)G::N || fread(pb.data(), 8, G::N, f) != (size_t)G::N) return 2;
  fclose(f);
  std::vector<double2> v(256 * 16), S(G::BUF), X(G::BUF), DD(G::BUF);
  std::vector<double> red(256);
  std::vector<int> redi(256), box(8);
  WallishGap gaps[4];
  for (int t = 0; t < 256; ++t)
    for (int r = 0; r < 16; ++r) {
      const int n = t + 256 * r;
      const double sign = n < G::N / 2 ? 1. : -1.;
      const int j = n < G::N / 2 ? 2 * n : 2 * (G::N - 1 - n) + 1;
      v[t * 16 + r] = mk2(sign * log(klin[j] * pa[j]), sign * log(klin[j] * pb[j]));
    }
  fft4096(v, S);
  for (int t = 0; t < 256; ++t) for (int r = 0; r < 16; ++r) S[t + 256 * r] = v[t * 16 + r];
  for (int t = 0; t < 256; ++t) wallish_dst2_post(t, *(Regs*)&v[t * 16], S.data(), X.data(), twd.data());
  std::vector<double> out;
  for (int kk = 0; kk < G::N; ++kk) { out.push_back(X[wpos(kk & 1, kk >> 1)].x); out.push_back(X[wpos(kk & 1, kk >> 1)].y); }
  std::vector<double2> E(256);
  std::vector<double> Mf(256);
  for (int t = 0; t < 256; ++t) wallish_forward_local(t, X.data(), E.data(), Mf.data(), wtab);
  for (int t = 0; t < 256; ++t) wallish_forward_store(t, X.data(), S.data(), E.data(), Mf.data(), wtab);
  for (int t = 0; t < 256; ++t) wallish_backward_local(t, S.data(), E.data(), Mf.data(), wtab);
  for (int t = 0; t < 256; ++t) wallish_backward_dd(t, X.data(), S.data(), DD.data(), E.data(), Mf.data(), wtab);
  for (int h = 0; h < 2; ++h) for (int i = 0; i < G::H; ++i) { out.push_back(DD[wpos(h, i)].x); out.push_back(DD[wpos(h, i)].y); }
  // argmax boxes: per-chunk maxima from the backward pass, merged in thread order (the kernel merges with warp shuffles)
  std::vector<WallishBest> chunk(256), cand(256);
  for (int t = 0; t < 256; ++t) wallish_backward_dd(t, X.data(), S.data(), DD.data(), E.data(), Mf.data(), wtab, &chunk[t]);
  auto merge = [&](const std::vector<WallishBest>& b, int q) {
    double v = 0.; int i = -1;
    for (int t = 128 * (q >> 1); t < 128 * (q >> 1) + 128; ++t) wallish_best_merge(v, i, (q & 1) ? b[t].vy : b[t].vx, (q & 1) ? b[t].iy : b[t].ix);
    return i;
  };
  for (int q = 0; q < 4; ++q) box[2 * q] = merge(chunk, q);
  for (int t = 0; t < 256; ++t) cand[t] = wallish_chunk_candidate(t, DD.data(), box[4 * (t >> 7)] + G::MARGIN_SECOND, box[4 * (t >> 7) + 2] + G::MARGIN_SECOND, chunk[t]);
  for (int t = 0; t < 4; ++t) {
    const int h = t >> 1, col = t & 1;
    const int amax = box[2 * t], bmax = merge(cand, t);
    const int b0 = amax + G::OFF_LO, b1 = bmax < 0 ? G::H : bmax + G::OFF_HI;
    gaps[t] = wallish_gap_solve(X.data(), h, col, b0, b1, wtab);
    out.push_back((double)b0); out.push_back((double)b1);
  }
  for (int e = 0; e < G::N; ++e) {
    const int h = e >> 11, i = e & (G::H - 1);
    const double2 y = X[wpos(h, i)];
    X[wpos(h, i)] = mk2(wallish_fill(y.x, i, gaps[2 * h]), wallish_fill(y.y, i, gaps[2 * h + 1]));
  }
  for (int kk = 0; kk < G::N; ++kk) { out.push_back(X[wpos(kk & 1, kk >> 1)].x); out.push_back(X[wpos(kk & 1, kk >> 1)].y); }
  for (int t = 0; t < 256; ++t) wallish_dst3_pre(t, X.data(), *(Regs*)&v[t * 16], twd.data());
  fft4096(v, S);
  std::vector<double> res(2 * G::N);
  for (int t = 0; t < 256; ++t)
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = wallish_dst3_out_index(t + 256 * r, sign);
      res[2 * j] = exp(sign * v[t * 16 + r].x / G::N) / klin[j];
      res[2 * j + 1] = exp(sign * v[t * 16 + r].y / G::N) / klin[j];
    }
  out.insert(out.end(), res.begin(), res.end());
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  return 0;
}

// CPU emulation of the fused Wallish2018 kernel (csrc/cpf_wallish.cu :: wallish_fused_kernel): the per-thread phase
// functions of cpf_wallish_core.h / cpf_fft_core.h are run for all 256 thread ids in turn, phase by phase, with plain
// arrays standing in for shared memory (one buffer, as in the kernel) and for the register arrays.  Input: a binary file with klin[4096],
// pk_a[4096], pk_b[4096], nk, kout[nk], pkout_a[nk], pkout_b[nk]; output: a binary file with the DST-II coefficients, second derivatives,
// boxes, cut coefficients, exp(DST-III)/k and pknow for both columns.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../cosmoprimo_b200/csrc/cpf_fastmath.h"
#include "../../cosmoprimo_b200/csrc/cpf_wallish_core.h"
#include "../../cosmoprimo_b200/csrc/cpf_wallish_final.h"

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
  // input: klin[4096], pk_a[4096], pk_b[4096], nk (as a double), kout[nk], pkout_a[nk], pkout_b[nk]
  std::vector<double> klin(G::N), pa(G::N), pb(G::N);
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  double nkd = 0.;
  if (fread(klin.data(), 8, G::N, f) != (size_t)G::N || fread(pa.data(), 8, G::N, f) != (size_t)G::N || fread(pb.data(), 8, G::N, f) != (size_t)G::N ||
      fread(&nkd, 8, 1, f) != 1) return 2;
  const int nk = (int)nkd;
  std::vector<double> kout(nk), qa(nk), qb(nk);
  if (fread(kout.data(), 8, nk, f) != (size_t)nk || fread(qa.data(), 8, nk, f) != (size_t)nk || fread(qb.data(), 8, nk, f) != (size_t)nk) return 2;
  fclose(f);
  // ONE buffer, as in the kernel; every loop over t below is one barrier-delimited phase
  std::vector<double2> v(256 * 16), d(256 * 16), B(G::BUF);
  std::vector<double2> E(256), Eb(256);
  std::vector<double> Mf(256), Mb(256);
  std::vector<int> box(8);
  WallishGap gaps[4];
  for (int t = 0; t < 256; ++t)
    for (int r = 0; r < 16; ++r) {
      const int n = t + 256 * r;
      const double sign = n < G::N / 2 ? 1. : -1.;
      const int j = n < G::N / 2 ? 2 * n : 2 * (G::N - 1 - n) + 1;
      v[t * 16 + r] = mk2(sign * fast_log(klin[j] * pa[j]), sign * fast_log(klin[j] * pb[j]));
    }
  fft4096(v, B);
  for (int t = 0; t < 256; ++t) for (int r = 0; r < 16; ++r) B[t + 256 * r] = v[t * 16 + r];
  for (int t = 0; t < 256; ++t) wallish_dst2_coef(t, *(Regs*)&v[t * 16], B.data(), twd.data());
  for (int t = 0; t < 256; ++t) wallish_dst2_store(t, *(Regs*)&v[t * 16], B.data());
  std::vector<double> out;
  for (int kk = 0; kk < G::N; ++kk) { out.push_back(B[wpos(kk & 1, kk >> 1)].x); out.push_back(B[wpos(kk & 1, kk >> 1)].y); }
  std::vector<WallishBest> chunk(256), cand(256);
  for (int t = 0; t < 256; ++t) wallish_forward_local(t, B.data(), *(Regs*)&d[t * 16], E.data(), Mf.data(), wtab);
  for (int t = 0; t < 256; ++t) wallish_forward_fix_backward_local(t, *(Regs*)&d[t * 16], E.data(), Mf.data(), Eb.data(), Mb.data(), wtab);
  for (int t = 0; t < 256; ++t) chunk[t] = wallish_backward_dd(t, B.data(), *(Regs*)&d[t * 16], Eb.data(), Mb.data(), wtab);
  for (int h = 0; h < 2; ++h) for (int i = 0; i < G::H; ++i) { const double2 dd = d[(128 * h + i / 16) * 16 + i % 16]; out.push_back(dd.x); out.push_back(dd.y); }
  // argmax boxes: per-chunk maxima from the backward pass, merged in thread order (the kernel merges with warp shuffles)
  auto merge = [&](const std::vector<WallishBest>& b, int q) {
    double bv = 0.; int i = -1;
    for (int t = 128 * (q >> 1); t < 128 * (q >> 1) + 128; ++t) wallish_best_merge(bv, i, (q & 1) ? b[t].vy : b[t].vx, (q & 1) ? b[t].iy : b[t].ix);
    return i;
  };
  for (int q = 0; q < 4; ++q) box[2 * q] = merge(chunk, q);
  for (int t = 0; t < 256; ++t) cand[t] = wallish_chunk_candidate(t, *(Regs*)&d[t * 16], box[4 * (t >> 7)] + G::MARGIN_SECOND, box[4 * (t >> 7) + 2] + G::MARGIN_SECOND, chunk[t]);
  for (int t = 0; t < 4; ++t) {
    const int h = t >> 1, col = t & 1;
    const int amax = box[2 * t], bmax = merge(cand, t);
    const int b0 = amax + G::OFF_LO, b1 = bmax < 0 ? G::H : bmax + G::OFF_HI;
    gaps[t] = wallish_gap_solve(B.data(), h, col, b0, b1, wtab);
    out.push_back((double)b0); out.push_back((double)b1);
  }
  for (int e = 0; e < G::N; ++e) {
    const int h = e >> 11, i = e & (G::H - 1);
    const double2 y = B[wpos(h, i)];
    B[wpos(h, i)] = mk2(wallish_fill(y.x, i, gaps[2 * h]), wallish_fill(y.y, i, gaps[2 * h + 1]));
  }
  for (int kk = 0; kk < G::N; ++kk) { out.push_back(B[wpos(kk & 1, kk >> 1)].x); out.push_back(B[wpos(kk & 1, kk >> 1)].y); }
  for (int t = 0; t < 256; ++t) wallish_dst3_pre(t, B.data(), *(Regs*)&v[t * 16], twd.data());
  fft4096(v, B);
  std::vector<double> res(2 * G::N);
  for (int t = 0; t < 256; ++t)
    for (int r = 0; r < 16; ++r) {
      double sign;
      const int j = wallish_dst3_out_index(t + 256 * r, sign);
      const double rk = 1. / klin[j];
      res[2 * j] = fast_exp(sign * v[t * 16 + r].x / G::N) * rk;
      res[2 * j + 1] = fast_exp(sign * v[t * 16 + r].y / G::N) * rk;
    }
  out.insert(out.end(), res.begin(), res.end());
  // final stage: spliced clamped spline solved in the buffer, evaluated at kout, blended
  WallishFinalPlan fp;
  const std::string why = wallish_final_plan(klin.data(), G::N, kout.data(), nk, &fp, argc > 3 ? atoi(argv[3]) : 0);
  if (!why.empty()) { fprintf(stderr, "plan: %s\n", why.c_str()); return 3; }
  for (auto& e : B) e = mk2(nan(""), nan(""));             // nothing of the earlier phases may be read
  for (int j = fp.i0; j < fp.i1; ++j) B[ypos(fp.lz + j - fp.i0)] = mk2(res[2 * j], res[2 * j + 1]);
  for (int c = 0; c < fp.lz + fp.rz; ++c) {
    const int row = c < fp.lz ? fp.nl - fp.lz + c : nk - fp.nr + (c - fp.lz);
    B[ypos(c < fp.lz ? c : fp.nmid + c)] = mk2(qa[row], qb[row]);
  }
  WallishFinFac fc;
  fc.facT = fp.facT.data(); fc.t0 = fp.ut0; fc.t1 = fp.ut1; fc.Lw = fp.uLw; fc.cp = fp.ucp; fc.P = fp.uP; fc.Q = fp.uQ;
  for (int t = 0; t < 256; ++t) wallish_fin_forward_local(t, fp.nc, B.data(), fc, *(Regs*)&d[t * 16], E.data(), Mf.data());
  for (int t = 0; t < 256; ++t) wallish_fin_fix_backward_local(t, fc, *(Regs*)&d[t * 16], E.data(), Mf.data(), Eb.data(), Mb.data());
  for (int t = 0; t < 256; ++t) wallish_fin_backward_fix(t, fc, *(Regs*)&d[t * 16], Eb.data(), Mb.data());
  std::vector<double> pknow(2 * nk);
  for (int round = 0; round < fp.nrounds; ++round) {
    for (int t = 0; t < 256; ++t) wallish_fin_scatter(t, round, fp.slotT.data(), *(Regs*)&d[t * 16], B.data() + fp.slbase);
    for (int q = fp.qstart[round]; q < fp.qstart[round + 1]; ++q) {
      const double k = kout[q], th = k > 1. ? exp(-400. * (k - 1.) * (k - 1.)) : 1.;
      const double2 r2 = wallish_fin_eval(q, fp.qinfo.data(), fp.qh.data(), B.data(), B.data() + fp.slbase, mk2(qa[q], qb[q]), th);
      pknow[2 * q] = r2.x; pknow[2 * q + 1] = r2.y;
    }
  }
  out.insert(out.end(), pknow.begin(), pknow.end());
  out.push_back((double)(fp.ut1 - fp.ut0)); out.push_back((double)fp.nrounds); out.push_back((double)fp.nc); out.push_back((double)fp.lz); out.push_back((double)fp.rz);
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  return 0;
}

// cpf_wallish_final.h — host-side tables of the final stage of the fused Wallish2018 kernel (bao_filter.py:415-423): which knots the
// spliced clamped spline keeps, their elimination factors in the kernel's thread-major order, and, per output wavenumber, the
// interval / Hermite factors / slope slots of its evaluation.  Everything here depends on the two wavenumber grids only (not on the
// spectra).  Plain C++ (no CUDA): cpf_wallish.cu builds the tables per call, tests/emul/emul_wallish.cpp runs the same code on the CPU.
#pragma once

#include <math.h>
#include <string>
#include <vector>

#include "cpf_spline_core.h"
#include "cpf_wallish_core.h"

namespace cpf {

struct WallishFinalPlan {
  int i0 = 0, i1 = 0;          // rows of klin kept: 1e-2 < k < 1.5                       (:415)
  int nl = 0, nr = 0;          // output wavenumbers below 5e-4 / above 2                 (:417)
  int lz = 0, rz = 0;          // ... of which the nearest lz / rz stay in the solve
  int nmid = 0, nc = 0;        // filtered knots, kept knots
  int nrounds = 0;             // the needed slopes go through the slot array in this many rounds (1 unless nk is huge)
  int slbase = 0, cap = 0;     // first element of the slot array in the shared buffer, its capacity
  std::vector<double> knots;   // [nc]
  std::vector<double> facT;    // [16][4][256]
  int ut0 = 0, ut1 = 0;        // threads whose 16 knots all have the converged uniform-grid factors uLw, ucp, uP, uQ
  double uLw = 0., ucp = 0., uP = 0., uQ = 0.;
  std::vector<int> slotT;      // [nrounds][16][256]
  std::vector<int> qstart;     // [nrounds + 1] output wavenumbers of a round
  std::vector<int> qinfo;      // [nk][4]
  std::vector<double> qh;      // [nk][4]
};

// returns an empty string, or what is wrong with the grids
// cap_limit > 0 shrinks the slot array (tests of the multi-round evaluation)
inline std::string wallish_final_plan(const double* klin, const int nlin, const double* kout, const int nk, WallishFinalPlan* out, const int cap_limit = 0) {
  typedef WallishGeo G;
  WallishFinalPlan& p = *out;
  p = WallishFinalPlan();
  p.i0 = 0; p.i1 = nlin;
  while (p.i0 < nlin && !(klin[p.i0] > 1e-2)) ++p.i0;                    // mask = (k > 1e-2) & (k < 1.5)   (:415)
  while (p.i1 > p.i0 && !(klin[p.i1 - 1] < 1.5)) --p.i1;
  while (p.nl < nk && kout[p.nl] < 5e-4) ++p.nl;                         // mask_left = self.k < 5e-4        (:417)
  while (p.nr < nk - p.nl && kout[nk - 1 - p.nr] > 2.) ++p.nr;           // mask_right = self.k > 2
  p.nmid = p.i1 - p.i0;
  if (p.nl + p.nmid + p.nr < 2) return "fewer than two knots survive the k cuts";
  const int room = G::T * G::CH - p.nmid;                                // knots the kernel can hold beside the filtered ones
  if (room < 0) return "more filtered knots than the kernel holds";
  p.lz = p.nl < 64 ? p.nl : 64;
  p.rz = p.nr < 64 ? p.nr : 64;
  while (p.lz + p.rz > room) { if (p.lz >= p.rz) --p.lz; else --p.rz; }
  if ((p.lz < p.nl && p.lz < 24) || (p.rz < p.nr && p.rz < 24)) return "too many filtered knots to keep enough unfiltered ones beside them";
  p.nc = p.lz + p.nmid + p.rz;
  p.knots.resize(p.nc);
  for (int i = 0; i < p.lz; ++i) p.knots[i] = kout[p.nl - p.lz + i];
  for (int i = 0; i < p.nmid; ++i) p.knots[p.lz + i] = klin[p.i0 + i];
  for (int i = 0; i < p.rz; ++i) p.knots[p.lz + p.nmid + i] = kout[nk - p.nr + i];
  for (int i = 1; i < p.nc; ++i)
    if (!(p.knots[i] > p.knots[i - 1])) return "spliced knots are not increasing";
  if (p.nc < 2) return "fewer than two knots survive the k cuts";
  // elimination factors of the clamped spline on the kept knots, thread-major
  p.facT.assign((size_t)G::CH * 4 * G::T, 0.);
  double cprev = 0.;
  for (int i = 0; i < p.nc; ++i) {
    double Lw, P, Q;
    spline_factor_step(p.knots.data(), p.nc, 1, i, cprev, Lw, P, Q);
    const int t = i / G::CH, j = i % G::CH;
    p.facT[(size_t)(4 * j + 0) * G::T + t] = Lw;
    p.facT[(size_t)(4 * j + 1) * G::T + t] = cprev;
    p.facT[(size_t)(4 * j + 2) * G::T + t] = P;
    p.facT[(size_t)(4 * j + 3) * G::T + t] = Q;
  }
  // uniform stretch: the factors at the middle filtered knot, and the largest run of whole chunks around it that agree with them to 1e-12
  // (the filtered knots are a linspace, bao_filter.py:364: their spacing jitters by rounding only, ~1e-13 relative)
  if (p.nmid >= 3 * G::CH) {
    const int mid = p.lz + p.nmid / 2;
    const double ref[4] = {p.facT[(size_t)(4 * (mid % G::CH) + 0) * G::T + mid / G::CH], p.facT[(size_t)(4 * (mid % G::CH) + 1) * G::T + mid / G::CH],
                           p.facT[(size_t)(4 * (mid % G::CH) + 2) * G::T + mid / G::CH], p.facT[(size_t)(4 * (mid % G::CH) + 3) * G::T + mid / G::CH]};
    auto chunk_ok = [&](int t) {
      if (t < 0 || (t + 1) * G::CH > p.nc) return false;
      for (int j = 0; j < G::CH; ++j)
        for (int f = 0; f < 4; ++f)
          if (!(fabs(p.facT[(size_t)(4 * j + f) * G::T + t] - ref[f]) <= 1e-12 * fabs(ref[f]))) return false;
      return true;
    };
    int t0 = mid / G::CH, t1 = t0;
    if (chunk_ok(t0)) {
      t1 = t0 + 1;
      while (chunk_ok(t0 - 1)) --t0;
      while (chunk_ok(t1)) ++t1;
      p.ut0 = t0; p.ut1 = t1;
      p.uLw = ref[0]; p.ucp = ref[1]; p.uP = ref[2]; p.uQ = ref[3];
    }
  }
  // evaluation of the output wavenumbers
  p.slbase = ypos(p.nc) + 1;
  p.cap = G::BUF - p.slbase;
  if (cap_limit > 0 && cap_limit < p.cap) p.cap = cap_limit;
  if (p.cap < 4) return "no room for the slope slots";
  p.qinfo.assign((size_t)nk * 4, 0);
  p.qh.assign((size_t)nk * 4, 0.);
  p.qstart.assign(1, 0);
  std::vector<int> slot(p.nc, -1);
  std::vector<int> round_slots;       // flattened [round][nc]
  int used = 0;
  auto close_round = [&](int qend) {
    round_slots.insert(round_slots.end(), slot.begin(), slot.end());
    p.qstart.push_back(qend);
    std::fill(slot.begin(), slot.end(), -1);
    used = 0;
  };
  for (int q = 0; q < nk; ++q) {
    int* qi = &p.qinfo[(size_t)4 * q];
    if (q < p.nl || q >= nk - p.nr) { qi[0] = -1; continue; }            // a knot of the spliced spline: it returns pk itself
    const double kq = kout[q];
    // the full knot set starts at kout[0] when nl > 0 (klin[i0] otherwise) and ends at kout[nk-1] when nr > 0 (klin[i1-1] otherwise)
    if (kq < p.knots[0] || kq > p.knots[p.nc - 1]) { qi[0] = -2; continue; }   // CubicSpline(extrapolate=False): NaN (:420)
    const int i = spline_interval(p.knots.data(), p.nc, kq);
    int need = (slot[i] < 0) + (slot[i + 1] < 0);
    if (used + need > p.cap) { close_round(q); need = 2; }
    if (slot[i] < 0) slot[i] = used++;
    if (slot[i + 1] < 0) slot[i + 1] = used++;
    const double dx = p.knots[i + 1] - p.knots[i], u = (kq - p.knots[i]) / dx;
    qi[0] = ypos(i); qi[1] = ypos(i + 1); qi[2] = slot[i]; qi[3] = slot[i + 1];
    double* h = &p.qh[(size_t)4 * q];
    h[0] = (1. + 2. * u) * (1. - u) * (1. - u);
    h[1] = u * u * (3. - 2. * u);
    h[2] = u * (1. - u) * (1. - u) * dx;
    h[3] = -u * u * (1. - u) * dx;
  }
  close_round(nk);
  p.nrounds = (int)p.qstart.size() - 1;
  p.slotT.assign((size_t)p.nrounds * G::CH * G::T, -1);
  for (int r = 0; r < p.nrounds; ++r)
    for (int i = 0; i < p.nc; ++i)
      p.slotT[((size_t)G::CH * r + i % G::CH) * G::T + i / G::CH] = round_slots[(size_t)r * p.nc + i];
  return std::string();
}

}  // namespace cpf

// CPU run of the Eisenstein & Hu point functions the CUDA generator is built from (csrc/cpf_eh_core.h).
// in.bin: B, nk, T_cmb, omega_r, k_pivot (doubles), params [B,5], z [B], k [nk];  out.bin: pk [B, nk], derived [B, 4]
#include <cstdio>
#include <vector>
#include "../../cosmoprimo_b200/csrc/cpf_eh_core.h"

using namespace cpf;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  double head[5];
  if (fread(head, 8, 5, f) != 5) return 2;
  const int B = (int)head[0], nk = (int)head[1];
  std::vector<double> params(5 * B), z(B), k(nk), out((size_t)B * nk + 4 * B);
  if (fread(params.data(), 8, params.size(), f) != params.size() || fread(z.data(), 8, B, f) != (size_t)B || fread(k.data(), 8, nk, f) != (size_t)nk) return 2;
  fclose(f);
  for (int b = 0; b < B; ++b) {
    const double* p = &params[5 * b];
    const EHCoeffs c = eh_coeffs(p[0], p[1], p[2], p[3], p[4], z[b], head[2], head[3], head[4]);
    for (int j = 0; j < nk; ++j) out[(size_t)b * nk + j] = eh_pk_point(c, k[j], log(k[j]));
    double* d = &out[(size_t)B * nk + 4 * b];
    d[0] = c.rs_drag * c.h; d[1] = c.z_drag; d[2] = c.growth_sq; d[3] = c.growth_rate;
  }
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  return 0;
}

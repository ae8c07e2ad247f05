// cpf_eh.cu — on-device Eisenstein & Hu linear P(k, z) rows: the input generator of every FFTLog workload
// (cosmoprimo/eisenstein_hu.py, see cpf_eh_core.h), written directly in the (rows, nk) layout cpf_fftlog reads, optionally
// as Kaiser multipoles ell = 0, 2, 4.  One CTA per cosmology: thread 0 derives the ~20 fitting coefficients, all threads
// then evaluate the transfer function on the shared k grid (2 log, 5 exp, 1 cbrt, 1 sin per point).
#include "cpf_common.h"
#include "cpf_eh_core.h"

namespace cpf {

__global__ void __launch_bounds__(256) eh_pk_kernel(const double* __restrict__ params, const double* __restrict__ z, const long long B,
                                                    const double* __restrict__ k, const int nk, const double T_cmb, const double omega_r,
                                                    const double k_pivot, const int kaiser, double* __restrict__ out,
                                                    double* __restrict__ derived) {
  __shared__ EHCoeffs sc;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const double* p = params + 5 * b;
      sc = eh_coeffs(p[0], p[1], p[2], p[3], p[4], z ? z[b] : 0., T_cmb, omega_r, k_pivot);
      if (derived) {
        double* d = derived + 4 * b;
        d[0] = sc.rs_drag * sc.h;     // Thermodynamics.rs_drag, Mpc/h (:163)
        d[1] = sc.z_drag;
        d[2] = sc.growth_sq;
        d[3] = sc.growth_rate;
      }
    }
    __syncthreads();
    const EHCoeffs c = sc;
    const double f = c.growth_rate;
    const double m0 = 1. + 2. * f / 3. + f * f / 5., m2 = 4. * f / 3. + 4. * f * f / 7., m4 = 8. * f * f / 35.;
    for (int j = threadIdx.x; j < nk; j += blockDim.x) {
      const double kj = k[j];
      const double pk = eh_pk_point(c, kj, log(kj));
      if (kaiser) {
        double* o = out + (b * 3) * nk + j;
        __stcs(o, m0 * pk);
        __stcs(o + nk, m2 * pk);
        __stcs(o + 2 * (long long)nk, m4 * pk);
      } else {
        __stcs(out + b * nk + j, pk);
      }
    }
  }
}

}  // namespace cpf

using namespace cpf;

extern "C" {

int cpf_eh_pk(const double* params, const double* z, int64_t B, const double* k, int nk, double T_cmb, double omega_r,
              double k_pivot, int kaiser, double* out, double* derived, int on_device, int device, void* stream_) {
  if (B < 0 || nk < 0) return fail(CPF_EINVAL, "cpf_eh_pk: negative size");
  if (B == 0 || nk == 0) return CPF_OK;
  if (!params || !k || !out) return fail(CPF_EINVAL, "cpf_eh_pk: null buffer");
  if (!(T_cmb > 0.) || !(omega_r >= 0.) || !(k_pivot > 0.)) return fail(CPF_EINVAL, "cpf_eh_pk: T_cmb, k_pivot must be > 0 and omega_r >= 0");
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_eh_pk: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int P = kaiser ? 3 : 1;
  const size_t ocells = (size_t)B * P * nk;
  ScratchBuf dp, dz, dk, dout, dder;
  const double *p_params = params, *p_z = z, *p_k = k;
  double *p_out = out, *p_der = derived;
  if (!on_device) {
    CPF_CUDA(dp.alloc((size_t)B * 5 * sizeof(double), stream));
    CPF_CUDA(dk.alloc((size_t)nk * sizeof(double), stream));
    CPF_CUDA(dout.alloc(ocells * sizeof(double), stream));
    CPF_CUDA(cudaMemcpyAsync(dp.p, params, (size_t)B * 5 * sizeof(double), cudaMemcpyHostToDevice, stream));
    CPF_CUDA(cudaMemcpyAsync(dk.p, k, (size_t)nk * sizeof(double), cudaMemcpyHostToDevice, stream));
    p_params = (const double*)dp.p; p_k = (const double*)dk.p; p_out = (double*)dout.p;
    if (z) {
      CPF_CUDA(dz.alloc((size_t)B * sizeof(double), stream));
      CPF_CUDA(cudaMemcpyAsync(dz.p, z, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, stream));
      p_z = (const double*)dz.p;
    }
    if (derived) {
      CPF_CUDA(dder.alloc((size_t)B * 4 * sizeof(double), stream));
      p_der = (double*)dder.p;
    }
  }
  const unsigned grid = (unsigned)(B < 148LL * 64 ? B : 148LL * 64);
  eh_pk_kernel<<<grid, 256, 0, stream>>>(p_params, p_z, B, p_k, nk, T_cmb, omega_r, k_pivot, kaiser ? 1 : 0, p_out, p_der);
  CPF_CUDA(cudaGetLastError());
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(out, p_out, ocells * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (derived) CPF_CUDA(cudaMemcpyAsync(derived, p_der, (size_t)B * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  }
  return CPF_OK;
}

}  // extern "C"

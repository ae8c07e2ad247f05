// cpf_eh.cu — on-device Eisenstein & Hu linear P(k, z) rows: the input generator of every FFTLog workload
// (cosmoprimo/eisenstein_hu.py, see cpf_eh_core.h), written directly in the (rows, nk) layout cpf_fftlog reads, optionally
// as Kaiser multipoles ell = 0, 2, 4.  The ~20 fitting coefficients are derived by one thread per cosmology (eh_coeffs_kernel); then one
// CTA per cosmology: all threads evaluate the transfer function on the shared k grid once and write one row per redshift of that cosmology (the
// z dependence is the growth factor only).
#include "cpf_common.h"
#include "cpf_eh_core.h"

namespace cpf {

#define EH_MAX_NZ 1024

// fitting coefficients: one thread per cosmology (~25 dependent pow / log / sqrt calls: 20 us of latency that a whole CTA used
// to sit through while its thread 0 worked)
__global__ void eh_coeffs_kernel(const double* __restrict__ params, const long long B, const double T_cmb, const double omega_r,
                                 const double k_pivot, EHCoeffs* __restrict__ coeffs) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* p = params + 5 * b;
  coeffs[b] = eh_coeffs(p[0], p[1], p[2], p[3], p[4], 0., T_cmb, omega_r, k_pivot);
}

__global__ void __launch_bounds__(256, 4) eh_pk_kernel(const EHCoeffs* __restrict__ coeffs, const double* __restrict__ z, const long long B,
                                                    const int nz, const double* __restrict__ k, const int nk, const int kaiser,
                                                    double* __restrict__ out, double* __restrict__ derived) {
  __shared__ double g2[EH_MAX_NZ], gf[EH_MAX_NZ];
  const int P = kaiser ? 3 : 1;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();                       // g2 / gf of the previous cosmology are no longer read
    const EHCoeffs c = coeffs[b];          // the same 23 doubles for every thread: broadcast loads
    for (int iz = threadIdx.x; iz < nz; iz += blockDim.x) {
      double gs, fr;
      eh_growth(c, z ? z[b * nz + iz] : 0., gs, fr);
      g2[iz] = gs;
      gf[iz] = fr;
      if (derived) {
        double* d = derived + 4 * (b * nz + iz);
        d[0] = c.rs_drag * c.h;       // Thermodynamics.rs_drag, Mpc/h (:163)
        d[1] = c.z_drag;
        d[2] = gs;
        d[3] = fr;
      }
    }
    __syncthreads();
    // the transfer function does not depend on z: one evaluation per k feeds all nz rows of the cosmology
    for (int j = threadIdx.x; j < nk; j += blockDim.x) {
      const double kj = k[j];
      const double pk0 = eh_pk0_point(c, kj, log(kj));
      for (int iz = 0; iz < nz; ++iz) {
        const double pk = pk0 * g2[iz];
        double* o = out + ((b * nz + iz) * P) * nk + j;
        if (kaiser) {
          const double f = gf[iz];
          __stcs(o, (1. + 2. * f / 3. + f * f / 5.) * pk);
          __stcs(o + nk, (4. * f / 3. + 4. * f * f / 7.) * pk);
          __stcs(o + 2 * (long long)nk, (8. * f * f / 35.) * pk);
        } else {
          __stcs(o, pk);
        }
      }
    }
  }
}

}  // namespace cpf

using namespace cpf;

extern "C" {

int cpf_eh_pk(const double* params, const double* z, int64_t B, int nz, const double* k, int nk, double T_cmb, double omega_r,
              double k_pivot, int kaiser, double* out, double* derived, int on_device, int device, void* stream_) {
  if (B < 0 || nk < 0) return fail(CPF_EINVAL, "cpf_eh_pk: negative size");
  if (nz < 1 || nz > EH_MAX_NZ) return fail(CPF_EINVAL, "cpf_eh_pk: nz = %d must be in 1..%d", nz, EH_MAX_NZ);
  if (!z && nz != 1) return fail(CPF_EINVAL, "cpf_eh_pk: z = NULL (redshift 0) needs nz = 1");
  if (B == 0 || nk == 0) return CPF_OK;
  if (!params || !k || !out) return fail(CPF_EINVAL, "cpf_eh_pk: null buffer");
  if (!(T_cmb > 0.) || !(omega_r >= 0.) || !(k_pivot > 0.)) return fail(CPF_EINVAL, "cpf_eh_pk: T_cmb, k_pivot must be > 0 and omega_r >= 0");
  int ndev = 0;
  CPF_TRY(cpf_device_count(&ndev));
  if (device < 0 || device >= ndev) return fail(CPF_EINVAL, "cpf_eh_pk: device %d out of range (%d visible)", device, ndev);
  DeviceGuard guard(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int P = kaiser ? 3 : 1;
  const size_t ocells = (size_t)B * nz * P * nk;
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
      CPF_CUDA(dz.alloc((size_t)B * nz * sizeof(double), stream));
      CPF_CUDA(cudaMemcpyAsync(dz.p, z, (size_t)B * nz * sizeof(double), cudaMemcpyHostToDevice, stream));
      p_z = (const double*)dz.p;
    }
    if (derived) {
      CPF_CUDA(dder.alloc((size_t)B * nz * 4 * sizeof(double), stream));
      p_der = (double*)dder.p;
    }
  }
  ScratchBuf dcoef;
  CPF_CUDA(dcoef.alloc((size_t)B * sizeof(EHCoeffs), stream));
  eh_coeffs_kernel<<<(unsigned)((B + 63) / 64), 64, 0, stream>>>(p_params, B, T_cmb, omega_r, k_pivot, (EHCoeffs*)dcoef.p);
  const unsigned grid = (unsigned)(B < 148LL * 64 ? B : 148LL * 64);
  eh_pk_kernel<<<grid, 256, 0, stream>>>((const EHCoeffs*)dcoef.p, p_z, B, nz, p_k, nk, kaiser ? 1 : 0, p_out, p_der);
  CPF_CUDA(cudaGetLastError());
  if (!on_device) {
    CPF_CUDA(cudaMemcpyAsync(out, p_out, ocells * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (derived) CPF_CUDA(cudaMemcpyAsync(derived, p_der, (size_t)B * nz * 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CPF_CUDA(cudaStreamSynchronize(stream));
  }
  return CPF_OK;
}

}  // extern "C"

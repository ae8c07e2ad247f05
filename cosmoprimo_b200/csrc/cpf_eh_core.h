// cpf_eh_core.h — Eisenstein & Hu (1998) linear matter power spectrum, scalar building blocks (host + device).
// Follows cosmoprimo/eisenstein_hu.py: per-cosmology coefficients (_set_rsdrag :34-63, compute :65-92), transfer function
// (Transfer.transfer_k :241-283), primordial spectrum (Primordial.pk_k :189-214 with alpha_s = beta_s = 0), P(k)
// (Fourier.pk_interpolator :321-324), growth factor and growth rate (Background :115-153) on the reference's background
// for flat LCDM without massive neutrinos (cosmology.py:1675-1760: matter + photons + massless neutrinos + Lambda).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define CPF_EHD __host__ __device__ __forceinline__
#else
#define CPF_EHD inline
inline double rcbrt(double x) { return 1. / cbrt(x); }
#endif

namespace cpf {

struct EHCoeffs {
  double h, frac_b, k_eq, k_silk, rs_drag, z_drag, alpha_c, beta_c, alpha_b, beta_b, beta_node;
  double Omega0_m, Omega0_r;
  double ln_q_scale;        // ln(h / (13.41 k_eq)): q = k [h/Mpc] * exp(ln_q_scale)
  double ln_silk_scale;     // ln(h / k_silk)
  double amp;               // P(k) = T^2 * amp * k^(n_s) ... see eh_pk_point
  double nsm1, ln_kp;       // n_s - 1, ln(k_pivot / h)
  double growth_sq;         // D(z)^2, znorm = 0 (eisenstein_hu.py:319)
  double growth_rate;       // f(z) = Omega_m(z)^0.55 (eisenstein_hu.py:141-153, w = -1)
};

// growth factor squared (znorm = 0, :132-136, :319) and growth rate Omega_m(z)^0.55 (:141-153, w = -1) on the reference's
// flat background: comoving densities in units of the critical density today (cosmology.py:1675-1745)
CPF_EHD void eh_growth(const EHCoeffs& c, const double z, double& growth_sq, double& growth_rate) {
  const double Omega0_de = 1. - c.Omega0_m - c.Omega0_r;
  const double zp1 = 1. + z;
  const double crit = c.Omega0_m + c.Omega0_r * zp1 + Omega0_de / (zp1 * zp1 * zp1);
  const double Om = c.Omega0_m / crit, Ode = Omega0_de / (zp1 * zp1 * zp1) / crit;
  const double D = 1. / zp1 * 5. * Om / 2. / (pow(Om, 4. / 7.) - Ode + (1. + Om / 2.) * (1. + Ode / 70.));
  growth_sq = D * D;
  growth_rate = pow(Om, 0.55);
}

// params = (h, omega_b, omega_cdm, n_s, A_s); omega_r = Omega0_r h^2 (photons + massless neutrinos)
CPF_EHD EHCoeffs eh_coeffs(const double h, const double omega_b, const double omega_cdm, const double n_s, const double A_s,
                           const double z, const double T_cmb, const double omega_r, const double k_pivot) {
  EHCoeffs c;
  const double omega_m = omega_cdm + omega_b;
  const double theta = T_cmb / 2.7;
  c.h = h;
  c.frac_b = omega_b / omega_m;
  const double z_eq = 2.5e4 * omega_m * pow(theta, -4.) - 1.;                                       // EH eq. 2
  c.k_eq = 0.0746 * omega_m * pow(theta, -2.);                                                       // EH eq. 3, 1/Mpc
  const double b1 = 0.313 * pow(omega_m, -0.419) * (1. + 0.607 * pow(omega_m, 0.674));
  const double b2 = 0.238 * pow(omega_m, 0.223);
  c.z_drag = 1345. * pow(omega_m, 0.251) / (1. + 0.659 * pow(omega_m, 0.828)) * (1. + b1 * pow(omega_b, b2));   // :53
  const double r_drag = 31.5 * omega_b * pow(theta, -4.) * (1000. / (1. + c.z_drag));               // EH eq. 5
  const double r_eq = 31.5 * omega_b * pow(theta, -4.) * (1000. / (1. + z_eq));
  c.rs_drag = 2. / (3. * c.k_eq) * sqrt(6. / r_eq) * log((sqrt(1. + r_drag) + sqrt(r_drag + r_eq)) / (1. + sqrt(r_eq)));   // EH eq. 6
  c.k_silk = 1.6 * pow(omega_b, 0.52) * pow(omega_m, 0.73) * (1. + pow(10.4 * omega_m, -0.95));     // EH eq. 7
  const double a1 = pow(46.9 * omega_m, 0.670) * (1. + pow(32.1 * omega_m, -0.532));                 // EH eq. 11
  const double a2 = pow(12.0 * omega_m, 0.424) * (1. + pow(45.0 * omega_m, -0.582));
  c.alpha_c = pow(a1, -c.frac_b) * pow(a2, -(c.frac_b * c.frac_b * c.frac_b));
  const double bb1 = 0.944 / (1. + pow(458. * omega_m, -0.708));                                     // EH eq. 12
  const double bb2 = 0.395 * pow(omega_m, -0.0266);
  c.beta_c = 1. / ((1. + bb1 * pow(1. - c.frac_b, bb2)) - 1.);   // as written at :84 (the -1 sits outside the bracket)
  const double yd = (1. + z_eq) / (1. + c.z_drag);
  const double sq = sqrt(1. + yd);
  const double G = yd * (-6. * sq + (2. + 3. * yd) * log((sq + 1.) / (sq - 1.)));                    // EH eq. 15
  c.alpha_b = 2.07 * c.k_eq * c.rs_drag * pow(1. + r_drag, -0.75) * G;
  c.beta_node = 8.41 * pow(omega_m, 0.435);                                                          // EH eq. 23
  c.beta_b = 0.5 + c.frac_b + (3. - 2. * c.frac_b) * sqrt((17.2 * omega_m) * (17.2 * omega_m) + 1.); // EH eq. 24
  c.Omega0_m = omega_m / (h * h);
  c.ln_q_scale = log(h / (13.41 * c.k_eq));
  c.ln_silk_scale = log(h / c.k_silk);
  // P = T^2 * (3 Om 100^2 / (2 c^2 k^2))^-2 * 9/25 * 2 pi^2 / k^3 / h^3 * h^3 A_s (k / kp)^(n_s-1)       (:321-324, :213)
  const double ckms = 299792.458, PI = 3.14159265358979323846;
  const double p2d = 3. * c.Omega0_m * 100. * 100. / (2. * ckms * ckms);                             // times k^-2, to the power -2
  c.amp = 1. / (p2d * p2d) * (9. / 25. * 2. * PI * PI) * A_s;                                        // times k^4 / k^3 = k
  c.nsm1 = n_s - 1.;
  c.ln_kp = log(k_pivot / h);
  c.Omega0_r = omega_r / (h * h);
  eh_growth(c, z, c.growth_sq, c.growth_rate);
  return c;
}

// Transfer function T(k) (:241-283) for k in h/Mpc, lnk = ln(k).  Powers are taken as exp(p ln x) on the shared ln k
// (2 log, 3 exp, 1 rcbrt, 1 sin, 8 divisions per point); agrees with the reference's pow() calls to a few ulp.
CPF_EHD double eh_transfer_point(const EHCoeffs& c, const double k, const double lnk) {
  const double E = 2.718281828459045, PI = 3.14159265358979323846;
  const double kk = k * c.h;                                       // 1/Mpc
  const double q = kk / (13.41 * c.k_eq);                          // EH eq. 10
  const double ks = kk * c.rs_drag, inv_ks = 1. / ks;
  const double ln_beta = log(E + 1.8 * c.beta_c * q), ln_nobeta = log(E + 1.8 * q);
  const double q108 = exp(1.08 * (lnk + c.ln_q_scale));
  const double cc = 386. / (1. + 69.9 * q108);
  const double C_alpha = 14.2 / c.alpha_c + cc, C_noalpha = 14.2 + cc;
  const double ks54 = ks / 5.4, ks54sq = ks54 * ks54;
  const double f = 1. / (1. + ks54sq * ks54sq);                    // EH eq. 18
  const double q2 = q * q;
  const double T0_noalpha = ln_beta / (ln_beta + C_noalpha * q2), T0_alpha = ln_beta / (ln_beta + C_alpha * q2);
  const double Tc = f * T0_noalpha + (1. - f) * T0_alpha;
  const double bn = c.beta_node * inv_ks;
  const double s_tilde = c.rs_drag * rcbrt(1. + bn * bn * bn);     // EH eq. 22
  const double x = kk * s_tilde / PI, y = PI * x;                  // numpy.sinc(x) = sin(pi x) / (pi x)
  const double sinc = (y == 0.) ? 1. : sin(y) / y;
  const double ks52 = ks / 5.2;
  const double Tb1 = ln_nobeta / ((ln_nobeta + C_noalpha * q2) * (1. + ks52 * ks52));               // EH eq. 21
  const double bb = c.beta_b * inv_ks;
  const double Tb2 = c.alpha_b / (1. + bb * bb * bb) * exp(-exp(1.4 * (lnk + c.ln_silk_scale)));
  const double Tb = sinc * (Tb1 + Tb2);
  return c.frac_b * Tb + (1. - c.frac_b) * Tc;                     // EH eq. 16
}

// P(k, z = 0 amplitude): multiply by growth_sq for the redshift wanted
CPF_EHD double eh_pk0_point(const EHCoeffs& c, const double k, const double lnk) {
  const double T = eh_transfer_point(c, k, lnk);
  return T * T * c.amp * k * exp(c.nsm1 * (lnk - c.ln_kp));
}

CPF_EHD double eh_pk_point(const EHCoeffs& c, const double k, const double lnk) { return eh_pk0_point(c, k, lnk) * c.growth_sq; }

}  // namespace cpf

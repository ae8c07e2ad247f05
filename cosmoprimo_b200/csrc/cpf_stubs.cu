// temporary: entry points not implemented yet
#include "cpf_common.h"
extern "C" {
int cpf_spline_fit(const double*, const double*, int, int64_t, int, double*, int, int, void*) { return cpf::fail(CPF_EUNSUPPORTED, "cpf_spline_fit: not implemented yet"); }
int cpf_spline_eval(const double*, const double*, const double*, int, int64_t, const double*, int, int, int, double*, int, int, void*) { return cpf::fail(CPF_EUNSUPPORTED, "cpf_spline_eval: not implemented yet"); }
int cpf_dst(int, const double*, int, int64_t, double*, int, int, void*) { return cpf::fail(CPF_EUNSUPPORTED, "cpf_dst: not implemented yet"); }
int cpf_wallish2018(const double*, const double*, int, const double*, const double*, int, int64_t, double*, int32_t*, int, int, void*) { return cpf::fail(CPF_EUNSUPPORTED, "cpf_wallish2018: not implemented yet"); }
}

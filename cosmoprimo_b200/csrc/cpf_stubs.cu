// temporary: entry points not implemented yet
#include "cpf_common.h"
extern "C" {
int cpf_dst(int, const double*, int, int64_t, double*, int, int, void*) { return cpf::fail(CPF_EUNSUPPORTED, "cpf_dst: not implemented yet"); }
int cpf_wallish2018(const double*, const double*, int, const double*, const double*, int, int64_t, double*, int32_t*, int, int, void*) { return cpf::fail(CPF_EUNSUPPORTED, "cpf_wallish2018: not implemented yet"); }
}

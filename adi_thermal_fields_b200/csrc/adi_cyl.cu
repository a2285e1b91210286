// placeholder, replaced below
#include "adi_ctx.h"
namespace adi { void cyl_release(adi_ctx *) {} }
extern "C" {
int adi_cyl_bind(adi_ctx *, int, int, int, int, double, double, double) { adi::set_error("cyl: not built"); return ADI_ESTATE; }
int adi_cyl_step(adi_ctx *, const double *, double *, const adi_cyl_params *, const uint8_t *, const double *, void *) { adi::set_error("cyl: not built"); return ADI_ESTATE; }
int adi_cyl_step_host(adi_ctx *, const double *, double *, int, const adi_cyl_params *, const uint8_t *, const double *, void *) { adi::set_error("cyl: not built"); return ADI_ESTATE; }
}

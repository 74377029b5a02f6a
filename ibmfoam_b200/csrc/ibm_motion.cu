// Loop 5 — collisions, wall hits, periodic copies, sub-stepped integration (placeholder until K5 lands).
#include "ibm_ctx.h"
namespace ibm {
int motion_advance(ibm_ctx* c, double, ibm_solve_stats*) { return ctx_fail(c, IBM_E_UNSUPPORTED, "ibm_advance: not built yet"); }
int motion_substep(ibm_ctx* c, double) { return ctx_fail(c, IBM_E_UNSUPPORTED, "ibm_substep: not built yet"); }
void motion_destroy(ibm_ctx*) {}
int combine_forces(ibm_ctx* c) { return ctx_fail(c, IBM_E_UNSUPPORTED, "ibm_combine_forces: not built yet"); }
}  // namespace ibm

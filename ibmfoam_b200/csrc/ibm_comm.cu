// Multi-GPU plumbing (placeholder until the slab decomposition lands).
#include "ibm_ctx.h"
namespace ibm { void comm_destroy(ibm_ctx*) {} }
extern "C" {
int ibm_comm_unique_id_bytes(void) { return 128; }
int ibm_comm_get_unique_id(void*) { return IBM_E_UNSUPPORTED; }
int ibm_comm_init(ibm_ctx* c, const void*, int, int) { return ibm::ctx_fail(c, IBM_E_UNSUPPORTED, "ibm_comm_init: not built yet"); }
}

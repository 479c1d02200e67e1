// placeholder until the stepped path lands
#include "host_util.h"
extern "C" {
#define STUB(sig) b200moby_status sig { return b2m_fail(B200MOBY_ERR_UNSUPPORTED, "not implemented yet"); }
STUB(b200moby_create(const b200moby_scene_desc*, int, b200moby_handle*))
STUB(b200moby_destroy(b200moby_handle))
STUB(b200moby_set_state(b200moby_handle, const double*, const double*))
STUB(b200moby_get_state(b200moby_handle, double*, double*))
STUB(b200moby_set_state_dev(b200moby_handle, const double*, const double*, void*))
STUB(b200moby_get_state_dev(b200moby_handle, double*, double*, void*))
STUB(b200moby_step(b200moby_handle, double, int, void*))
STUB(b200moby_get_counters(b200moby_handle, b200moby_counters*))
STUB(b200moby_reset_counters(b200moby_handle))
STUB(b200moby_get_time(b200moby_handle, double*))
STUB(b200moby_get_last_lcp(b200moby_handle, int*, double*, int))
STUB(b200moby_fwd_dyn_batched(b200moby_handle, const double*, double*, double, void*))
STUB(b200moby_find_contacts_batched(b200moby_handle, const double*, const double*, int, int*, double*, double*, double*, double*, int*, double*, void*))
STUB(b200moby_delassus_batched(b200moby_handle, const double*, const double*, int, double*, double*, int*, void*))
}

// placeholder until the tcgen05 path lands
#include "common.cuh"
bool ens_tc_supported(const Net&) { return false; }
int ens_tc_prepare(cmbpo_ctx*, Net&) { return 0; }
int ens_forward_tc(cmbpo_ctx*, Net&, const float*, int64_t, float*, int) {
    cmbpo_set_error("tcgen05 path not built");
    return 1;
}

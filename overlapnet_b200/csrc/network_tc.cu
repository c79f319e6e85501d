// network_tc.cu -- tensor-core (tcgen05) path; filled in below.
#include "common.cuh"
namespace ovn {
int tc_pack_weights(ovn_handle* h) { (void)h; return OVN_OK; }
void tc_free(ovn_handle* h) { (void)h; }
int leg_forward_tc(ovn_handle* h, const float*, int, float*, cudaStream_t) {
  OVN_SET_ERR(h, OVN_ERR_BAD_CONFIG, "tensor-core leg not built yet");
}
int heads_forward_tc(ovn_handle* h, const float*, const float*, const int32_t*, const int32_t*, int, float*,
                     int32_t*, float*, cudaStream_t) {
  OVN_SET_ERR(h, OVN_ERR_BAD_CONFIG, "tensor-core heads not built yet");
}
}  // namespace ovn

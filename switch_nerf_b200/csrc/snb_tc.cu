// placeholder, replaced below
#include "snb_common.cuh"
namespace snb {
bool tc_supported(const Model* m) { return false; }
int tc_pack_weights(Model*, const snb_weights*, cudaStream_t) { return SNB_OK; }
size_t tc_workspace_bytes(const Model*, int64_t, double) { return 0; }
int tc_forward(Model*, const float*, int64_t, const float*, const snb_route_opts*, float*, int32_t*, float*, float*, int32_t*, Arena&, cudaStream_t) { set_error("tc path not built"); return SNB_EUNSUPPORTED; }
}

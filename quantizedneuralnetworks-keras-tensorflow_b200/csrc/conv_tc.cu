// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace qnnb {
bool conv_tc_supported(const qnnb_conv_desc& d, const char** why) { *why = "tcgen05 path not built"; return false; }
int launch_conv_tc(const qnnb_conv_desc& d, const void* x, const void* w, void* y, cudaStream_t st) {
  set_error("tcgen05 path not built"); return QNNB_EUNSUPPORTED;
}
}

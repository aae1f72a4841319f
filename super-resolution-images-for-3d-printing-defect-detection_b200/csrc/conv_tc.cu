// placeholder until the tcgen05 engine lands
#include "common.cuh"
#include "conv_common.cuh"
namespace srb {
bool conv_tc_eligible(const ConvParams&) { return false; }
int conv_tc_launch(const ConvParams&, cudaStream_t) { set_error("tcgen05 engine not built"); return SRB_E_UNSUPPORTED; }
}
extern "C" int srb_conv_tc_set_variant(int) { return 0; }

// placeholder until the tcgen05 path lands (replaced below in this round)
#include "lfsr_common.cuh"
using namespace lfsr;
extern "C" size_t lfsr_conv2d_tc_packed_floats(int, int, int, int) { return 0; }
extern "C" int lfsr_pack_conv_tc(const float*, float*, int, int, int, int) { set_error("tc path not built"); return LFSR_ERR_UNSUPPORTED; }
extern "C" int lfsr_conv2d_tc(const lfsr_tensor*, const float*, const lfsr_tensor*, const lfsr_conv_desc*, void*) { set_error("tc path not built"); return LFSR_ERR_UNSUPPORTED; }
extern "C" int lfsr_conv2d_tc_supported(const lfsr_tensor*, const lfsr_tensor*, const lfsr_conv_desc*) { return 0; }

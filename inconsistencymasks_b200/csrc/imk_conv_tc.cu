// tcgen05 implicit-GEMM convolution engine (placeholder until the UMMA path lands).
#include "imk_unet.cuh"
namespace imk {
bool conv_tc_supported(const ConvLayer &) { return false; }
int conv_tc_pack(ConvLayer &, const float *, std::vector<void *> &) { return IMK_OK; }
int conv_tc_launch(const ConvLayer &, const __half *, const __half *, __half *, __half *, int64_t, int, int, cudaStream_t) {
    set_error("conv_tc_launch: engine not built");
    return IMK_ESTATE;
}
}  // namespace imk

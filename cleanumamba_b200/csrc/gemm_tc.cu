// tcgen05 / TMEM / TMA tap-GEMM (CUM_MATH_TF32X3, CUM_MATH_TF32) -- placeholder until the kernel lands.
#include "common.cuh"
namespace cum {
int gemm_tc_fwd(const cum_gemm_desc&, cudaStream_t) {
    set_error("gemm: tcgen05 path not built yet");
    return CUM_ENOTSUP;
}
}  // namespace cum

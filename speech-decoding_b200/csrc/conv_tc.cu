// tcgen05 / TMEM / TMA implicit-GEMM Conv1d family (bf16) -- placeholder until the kernels land.
#include "common.cuh"

namespace sd {
bool conv_fwd_tc_supported(const sd_conv_args&) { return false; }
bool conv_wgrad_tc_supported(const sd_wgrad_args&) { return false; }
int conv_fwd_tc(const sd_conv_args&, cudaStream_t) { set_error("tcgen05 conv not built"); return 1; }
int conv_wgrad_tc(const sd_wgrad_args&, cudaStream_t) { set_error("tcgen05 wgrad not built"); return 1; }
}  // namespace sd

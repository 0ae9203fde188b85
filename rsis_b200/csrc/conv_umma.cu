// tcgen05 / TMA implicit-GEMM convolution (split-bf16, fp32 TMEM accumulation).  Placeholder until the kernel lands:
// reports "unsupported" so RSIS_IMPL_AUTO resolves to the fp32 CUDA-core path.
#include "common.cuh"

namespace rsis {
extern const bool kHasTcgen05 = false;
bool conv2d_umma_supported(const rsis_tensor*, int, const rsis_conv_weights*, const rsis_tensor*, const rsis_tensor*,
                           const rsis_tensor*, int, int) {
  return false;
}
int conv2d_umma(const rsis_tensor*, int, const rsis_conv_weights*, const rsis_tensor*, const rsis_tensor*,
                const rsis_tensor*, int, int, int, cudaStream_t) {
  return RSIS_ERR_UNSUPPORTED;
}
bool convlstm_cell_umma_supported(const rsis_tensor*, int, const rsis_conv_weights*) { return false; }
int convlstm_cell_umma(const rsis_tensor*, int, const rsis_conv_weights*, const float*, const rsis_tensor*,
                       const rsis_tensor*, const rsis_tensor*, uint32_t*, int, int, cudaStream_t) {
  return RSIS_ERR_UNSUPPORTED;
}
}  // namespace rsis

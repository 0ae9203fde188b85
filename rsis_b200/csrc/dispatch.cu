// C-ABI entry points for the two convolution-shaped primitives; picks the tcgen05 or the CUDA-core implementation.
#include "common.cuh"

namespace rsis {
int conv2d_simt(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, cudaStream_t st);
int convlstm_cell_simt(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const float* c_prev,
                       const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, cudaStream_t st);
// tcgen05 path (conv_umma.cu).  `supported` is a pure host-side shape/format check.
bool conv2d_umma_supported(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                           const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad);
int conv2d_umma(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, void* workspace,
                size_t workspace_bytes, cudaStream_t st);
size_t conv_umma_workspace_bytes();
bool convlstm_cell_umma_supported(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w);
int convlstm_cell_umma(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const float* c_prev,
                       const float* gate_preact, const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, void* workspace, size_t workspace_bytes,
                       int cta_cap, cudaStream_t st);
int set_precision(int mode);
int set_static_weights(int on);
int get_precision();
int convlstm_cell_group_max();
bool convlstm_cell_group_supported(const rsis_cell_args* cells, int n);
int convlstm_cell_group_umma(const rsis_cell_args* cells, int n, cudaStream_t st);
}  // namespace rsis

using namespace rsis;

extern "C" {

size_t rsis_conv_workspace_bytes(void) { return conv_umma_workspace_bytes(); }

int rsis_conv2d(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, int impl,
                void* workspace, size_t workspace_bytes, rsis_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == RSIS_IMPL_SIMT) return conv2d_simt(srcs, n_src, w, residual, y, y2, stride, pad, relu, st);
  const bool ok = srcs && w && y && conv2d_umma_supported(srcs, n_src, w, residual, y, y2, stride, pad);
  if (impl == RSIS_IMPL_TCGEN05) {
    if (!ok) return RSIS_ERR_UNSUPPORTED;
    return conv2d_umma(srcs, n_src, w, residual, y, y2, stride, pad, relu, workspace, workspace_bytes, st);
  }
  if (impl != RSIS_IMPL_AUTO) return RSIS_ERR_BAD_ARG;
  return ok ? conv2d_umma(srcs, n_src, w, residual, y, y2, stride, pad, relu, workspace, workspace_bytes, st)
            : conv2d_simt(srcs, n_src, w, residual, y, y2, stride, pad, relu, st);
}

int rsis_convlstm_cell(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const float* c_prev,
                       const float* gate_preact, const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, int impl, void* workspace,
                       size_t workspace_bytes, rsis_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int cta_cap = (impl >> 8) & 0xffff;  // RSIS_IMPL_CTA_CAP(n)
  impl &= 0xff;
  if (gate_preact && (impl == RSIS_IMPL_SIMT || !(srcs && w && convlstm_cell_umma_supported(srcs, n_src, w))))
    return RSIS_ERR_UNSUPPORTED;  // the hoisted-gates form exists in the tcgen05 kernel only
  if (impl == RSIS_IMPL_SIMT)
    return convlstm_cell_simt(srcs, n_src, w, c_prev, h_out, h_split, c_out, side_max, side_stride, side_offset, st);
  const bool ok = srcs && w && convlstm_cell_umma_supported(srcs, n_src, w);
  if (impl == RSIS_IMPL_TCGEN05) {
    if (!ok) return RSIS_ERR_UNSUPPORTED;
    return convlstm_cell_umma(srcs, n_src, w, c_prev, gate_preact, h_out, h_split, c_out, side_max, side_stride, side_offset,
                              workspace, workspace_bytes, cta_cap, st);
  }
  if (impl != RSIS_IMPL_AUTO) return RSIS_ERR_BAD_ARG;
  return ok ? convlstm_cell_umma(srcs, n_src, w, c_prev, gate_preact, h_out, h_split, c_out, side_max, side_stride, side_offset,
                                 workspace, workspace_bytes, cta_cap, st)
            : convlstm_cell_simt(srcs, n_src, w, c_prev, h_out, h_split, c_out, side_max, side_stride, side_offset,
                                 st);
}

int rsis_set_precision(int mode) {
  if (mode != RSIS_PRECISION_SPLIT_BF16 && mode != RSIS_PRECISION_BF16) return RSIS_ERR_BAD_ARG;
  return set_precision(mode);
}
int rsis_get_precision(void) { return get_precision(); }
int rsis_set_static_weights(int on) { return set_static_weights(on); }

int rsis_convlstm_cell_group_max(void) { return convlstm_cell_group_max(); }

int rsis_convlstm_cell_group(const rsis_cell_args* cells, int n_cells, rsis_stream_t stream) {
  if (!cells || n_cells < 1) return RSIS_ERR_BAD_ARG;
  if (!convlstm_cell_group_supported(cells, n_cells)) return RSIS_ERR_UNSUPPORTED;
  return convlstm_cell_group_umma(cells, n_cells, (cudaStream_t)stream);
}

}  // extern "C"

// Layout / element-format plumbing: the reference's float32 NCHW image batch -> NHWC activations in either
// element format (float32 or split-bf16 hi|lo planes), and NHWC format conversion.
#include "common.cuh"

namespace rsis {

// One thread per output pixel-channel; coalesced on the NHWC side.  Used for the 3-channel image batch
// (/root/reference/src/test.py:35 `encoder(x)`), where the strided NCHW reads hit three planes only.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, void* dst, size_t plane, int fmt, int C,
                                    size_t HW, size_t total) {
  pdl_trigger();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t np = i / C;
    const size_t n = np / HW, p = np % HW;
    store_elem(dst, plane, fmt, i, src[(n * C + c) * HW + p]);
  }
}

// Format conversion / channel-slice copy: src and dst may both be pitched views (pixel pitch scs / dcs elements).
__global__ void convert_kernel(View src, int scs, void* dst, size_t plane, int fmt, int dcs, int C, size_t total) {
  pdl_trigger();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / C;
    const int c = (int)(i - pix * C);
    store_elem(dst, plane, fmt, pix * dcs + c, load_elem(src, pix * scs + c));
  }
}

}  // namespace rsis

using namespace rsis;

static inline int grid_for(size_t total, int block) {
  size_t b = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

extern "C" {

int rsis_nchw_to_nhwc(const float* src_nchw, const rsis_tensor* dst, rsis_stream_t stream) {
  if (!src_nchw || !valid_tensor(dst)) return RSIS_ERR_BAD_ARG;
  if (!is_dense(*dst)) return RSIS_ERR_UNSUPPORTED;
  const size_t total = numel(*dst);
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      src_nchw, dst->data, total, dst->fmt, dst->c, (size_t)dst->h * dst->w, total);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_convert(const rsis_tensor* src, const rsis_tensor* dst, rsis_stream_t stream) {
  if (!valid_tensor(src) || !valid_tensor(dst)) return RSIS_ERR_BAD_ARG;
  if (src->n != dst->n || src->h != dst->h || src->w != dst->w || src->c != dst->c) return RSIS_ERR_BAD_ARG;
  const size_t total = numel(*dst);
  convert_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      make_view(*src), pitch(*src), dst->data, plane_elems(*dst), dst->fmt, pitch(*dst), dst->c, total);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

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

// im2col for the ResNet stem (vision.py:12: 7x7, stride 2, pad 3 on a 3-channel image): y[n,ho,wo,(kh*KW+kw)*C + c] =
// x[n, ho*stride - pad + kh, wo*stride - pad + kw, c] (zero outside the image and for the padding channels beyond
// KH*KW*C), written in the split-bf16 operand format so that the stem becomes a 1x1 tensor-core convolution over
// K = 147 (+5) channels instead of a CUDA-core kernel.  One thread per (output pixel, group of 8 output channels).
__global__ void im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t y_plane, int N, int H,
                              int W, int C, int Ho, int Wo, int KH, int KW, int stride, int pad, int Cy) {
  pdl_trigger();
  const int G = Cy >> 3;
  const size_t total = (size_t)N * Ho * Wo * G;
  const int K = KH * KW * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    size_t r = i / G;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float v = 0.f;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
        const int kh = tap / KW, kw = tap - kh * KW;
        const int hi_ = ho * stride - pad + kh, wi_ = wo * stride - pad + kw;
        if (hi_ >= 0 && hi_ < H && wi_ >= 0 && wi_ < W) v = x[(((size_t)n * H + hi_) * W + wi_) * C + c];
      }
      split_bf16(v, hi[j], lo[j]);
    }
    const size_t o = (r * Wo + wo) * Cy + g * 8;  // r == n*Ho + ho here
    *reinterpret_cast<uint4*>(y + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(y + o + y_plane) = *reinterpret_cast<const uint4*>(lo);
  }
}

// The reference's shape (3-channel image, 7x7 taps, 152 packed channels) with every division by a compile-time constant
// and the image read through the read-only path: the generic kernel above spends ~40 instructions of run-time index
// arithmetic per element (80 us at batch 8, 256x256 -- as long as three layer-1 convolutions).
template <int C, int KH, int KW, int CY>
__global__ void __launch_bounds__(256) im2col_fixed_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                           size_t y_plane, int N, int H, int W, int Ho, int Wo,
                                                           int stride, int pad) {
  pdl_trigger();
  constexpr int G = CY / 8, K = KH * KW * C;
  const size_t total = (size_t)N * Ho * Wo * G;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    size_t r = i / G;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const float* xn = x + (size_t)n * H * W * C;
    __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float v = 0.f;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
        const int kh = tap / KW, kw = tap - kh * KW;
        const int hi_ = h0 + kh, wi_ = w0 + kw;
        if (hi_ >= 0 && hi_ < H && wi_ >= 0 && wi_ < W) v = __ldg(xn + ((size_t)hi_ * W + wi_) * C + c);
      }
      split_bf16(v, hi[j], lo[j]);
    }
    const size_t o = (r * Wo + wo) * CY + g * 8;  // r == n*Ho + ho here
    *reinterpret_cast<uint4*>(y + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(y + o + y_plane) = *reinterpret_cast<const uint4*>(lo);
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

int rsis_im2col(const rsis_tensor* x, int kh, int kw, int stride, int pad, const rsis_tensor* y, rsis_stream_t stream) {
  if (!valid_tensor(x) || !valid_tensor(y) || kh < 1 || kw < 1 || stride < 1 || pad < 0) return RSIS_ERR_BAD_ARG;
  if (x->fmt != RSIS_FMT_F32 || y->fmt != RSIS_FMT_SPLIT_BF16 || !is_dense(*x) || !is_dense(*y)) return RSIS_ERR_UNSUPPORTED;
  const int Ho = (x->h + 2 * pad - kh) / stride + 1, Wo = (x->w + 2 * pad - kw) / stride + 1;
  if (y->n != x->n || y->h != Ho || y->w != Wo || y->c < kh * kw * x->c || y->c % 8 != 0) return RSIS_ERR_BAD_ARG;
  if (!aligned16(y->data)) return RSIS_ERR_ALIGN;
  const size_t total = (size_t)y->n * Ho * Wo * (y->c / 8);
  if (x->c == 3 && kh == 7 && kw == 7 && y->c == 152) {
    const size_t b = (total + 255) / 256;
    im2col_fixed_kernel<3, 7, 7, 152><<<(unsigned)(b < 148 * 32 ? b : 148 * 32), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float*>(x->data), reinterpret_cast<__nv_bfloat16*>(y->data), plane_elems(*y), x->n, x->h,
        x->w, Ho, Wo, stride, pad);
    RSIS_CHECK_LAUNCH();
    return RSIS_OK;
  }
  im2col_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float*>(x->data), reinterpret_cast<__nv_bfloat16*>(y->data), plane_elems(*y), x->n, x->h,
      x->w, x->c, Ho, Wo, kh, kw, stride, pad, y->c);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

// Train-mode BatchNorm (batch statistics) for the encoder -- the `encoder.train()` forward of
// /root/reference/src/train.py:71-77 (nn.BatchNorm2d in training mode inside torchvision's Bottleneck and the skip
// heads model.py:59-63).  The convolution kernels fold EVAL-mode statistics into their epilogue; in training mode the
// statistics depend on the convolution's own output, so the pipeline is: conv (raw fp32 output) -> bn_stats (per-channel
// sum / sum of squares over N*H*W) -> bn_finalize (scale/shift of this batch + the running-statistics update with
// momentum, unbiased variance, num_batches_tracked) -> affine_act (normalise + residual + ReLU, written once in the
// next consumer's element format).  All HBM-streaming: one read of the conv output per kernel.
#include "common.cuh"

namespace rsis {

// acc: [2][C] doubles (sum, sum of squares), zero on entry; x: dense float32 NHWC, M = N*H*W pixels.
__global__ void bn_stats_kernel(const float* __restrict__ x, size_t M, int C, double* __restrict__ acc) {
  pdl_trigger();
  extern __shared__ float red[];  // [256][8]
  const int C4 = C >> 2;
  const int lanes_c = C4 < 256 ? C4 : 256;
  const int rows_per_iter = 256 / lanes_c;
  const size_t per_block = (M + gridDim.x - 1) / gridDim.x;
  const size_t p0 = (size_t)blockIdx.x * per_block;
  const size_t p1 = p0 + per_block < M ? p0 + per_block : M;
  // uniform trip count (the body holds block barriers): threads whose channel group falls beyond C4 in the last
  // round run it with an empty pixel range
  for (int c4base = 0; c4base < C4; c4base += lanes_c) {
    const int c4 = c4base + threadIdx.x % lanes_c;
    const bool live = c4 < C4 && threadIdx.x < rows_per_iter * lanes_c;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (size_t p = live ? p0 + threadIdx.x / lanes_c : p1; p < p1; p += rows_per_iter) {
      const float4 v = *reinterpret_cast<const float4*>(x + p * C + 4 * c4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
    float* mine = red + threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mine[j] = s[j];
      mine[4 + j] = q[j];
    }
    __syncthreads();
    if (threadIdx.x < lanes_c && c4 < C4) {  // one thread per channel group combines the rows of the block in double
      double ds[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
      for (int r = 0; r < rows_per_iter; ++r) {
        const float* o = red + (r * lanes_c + threadIdx.x) * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ds[j] += (double)o[j];
          dq[j] += (double)o[4 + j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(acc + 4 * c4 + j, ds[j]);
        atomicAdd(acc + C + 4 * c4 + j, dq[j]);
      }
    }
    __syncthreads();
  }
}

// scale = w / sqrt(var_biased + eps), shift = b - mean * scale; running stats as nn.BatchNorm2d does in training mode
// (momentum < 0 means cumulative moving average, i.e. momentum=None); re-zeroes acc for the next use.
__global__ void bn_finalize_kernel(double* __restrict__ acc, double count, int C, const float* __restrict__ w,
                                   const float* __restrict__ b, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  pdl_trigger();
  // ONE block (so that num_batches_tracked is read by every thread before thread 0 bumps it)
  double factor = momentum;
  if (num_batches) {
    const long long nb = *num_batches + 1;
    if (momentum < 0.f) factor = 1.0 / (double)nb;
    __syncthreads();
    if (threadIdx.x == 0) *num_batches = nb;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = acc[c] / count;
    double var = acc[C + c] / count - mean * mean;
    if (var < 0) var = 0;
    acc[c] = 0;
    acc[C + c] = 0;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double g = w ? (double)w[c] : 1.0, beta = b ? (double)b[c] : 0.0;
    scale[c] = (float)(g * invstd);
    shift[c] = (float)(beta - mean * g * invstd);
    if (mean_out) mean_out[c] = (float)mean;
    if (invstd_out) invstd_out[c] = (float)invstd;
    if (running_mean && running_var) {
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (float)((1.0 - factor) * running_mean[c] + factor * mean);
      running_var[c] = (float)((1.0 - factor) * running_var[c] + factor * unbiased);
    }
  }
}

__device__ __forceinline__ void ld4(const View& v, size_t idx, float out[4]) {
  if (v.fmt == RSIS_FMT_F32) {
    const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.p) + idx);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else {
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(v.p);
    const uint2 h = *reinterpret_cast<const uint2*>(q + idx);
    const uint2 l = *reinterpret_cast<const uint2*>(q + idx + v.plane);
    out[0] = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
    out[1] = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
    out[2] = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
    out[3] = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  }
}
__device__ __forceinline__ void st4(void* p, size_t plane, int fmt, size_t idx, const float v[4]) {
  if (fmt == RSIS_FMT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(p);
    *reinterpret_cast<uint2*>(b + idx) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(b + idx + plane) = *reinterpret_cast<uint2*>(lo);
  }
}

// y = [relu]( x * scale[c] + shift[c] [+ residual] ), dense NHWC, any element formats.
__global__ void affine_act_kernel(View x, const float* __restrict__ scale, const float* __restrict__ shift, View res,
                                  int has_res, int relu, void* y, size_t y_plane, int y_fmt, void* y2, size_t y2_plane,
                                  int y2_fmt, int C, size_t total4) {
  pdl_trigger();
  const int C4 = C >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t idx = i * 4;
    float v[4], r[4];
    ld4(x, idx, v);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + c));
    v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y); v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
    if (has_res) {
      ld4(res, idx, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    st4(y, y_plane, y_fmt, idx, v);
    if (y2) st4(y2, y2_plane, y2_fmt, idx, v);
  }
}

}  // namespace rsis

using namespace rsis;

extern "C" {

size_t rsis_bn_workspace_bytes(int channels) { return channels > 0 ? (size_t)2 * channels * sizeof(double) : 0; }

int rsis_bn_train_stats(const rsis_tensor* x, const float* weight, const float* bias, float eps, float momentum,
                        float* running_mean, float* running_var, int64_t* num_batches_tracked, double* workspace,
                        float* scale, float* shift, float* batch_mean, float* batch_invstd, rsis_stream_t stream) {
  if (!valid_tensor(x) || !workspace || !scale || !shift) return RSIS_ERR_BAD_ARG;
  if (x->fmt != RSIS_FMT_F32 || !is_dense(*x) || x->c % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(x->data) || !aligned16(scale) || !aligned16(shift)) return RSIS_ERR_ALIGN;
  if ((running_mean == nullptr) != (running_var == nullptr)) return RSIS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t M = (size_t)x->n * x->h * x->w;
  int blocks = (int)((M + 15) / 16);
  if (blocks > 592) blocks = 592;
  bn_stats_kernel<<<blocks, 256, 256 * 8 * sizeof(float), st>>>(reinterpret_cast<const float*>(x->data), M, x->c,
                                                               workspace);
  RSIS_CHECK_LAUNCH();
  bn_finalize_kernel<<<1, 1024, 0, st>>>(workspace, (double)M, x->c, weight, bias, eps, momentum,
                                                          running_mean, running_var,
                                                          reinterpret_cast<long long*>(num_batches_tracked), scale,
                                                          shift, batch_mean, batch_invstd);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_affine_act(const rsis_tensor* x, const float* scale, const float* shift, const rsis_tensor* residual, int relu,
                    const rsis_tensor* y, const rsis_tensor* y2, rsis_stream_t stream) {
  if (!valid_tensor(x) || !valid_tensor(y) || !scale || !shift) return RSIS_ERR_BAD_ARG;
  auto same = [&](const rsis_tensor* t) {
    return valid_tensor(t) && t->n == x->n && t->h == x->h && t->w == x->w && t->c == x->c && is_dense(*t) &&
           aligned16(t->data);
  };
  if (!same(x) || !same(y) || (y2 && !same(y2)) || (residual && !same(residual))) return RSIS_ERR_UNSUPPORTED;
  if (x->c % 4 != 0 || !aligned16(scale) || !aligned16(shift)) return RSIS_ERR_UNSUPPORTED;
  const size_t total4 = numel(*x) / 4;
  size_t b = (total4 + 255) / 256;
  const int blocks = (int)(b < 148 * 16 ? (b ? b : 1) : 148 * 16);
  View rv = residual ? make_view(*residual) : make_view(*x);
  affine_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      make_view(*x), scale, shift, rv, residual ? 1 : 0, relu ? 1 : 0, y->data, plane_elems(*y), y->fmt,
      y2 ? y2->data : nullptr, y2 ? plane_elems(*y2) : 0, y2 ? y2->fmt : 0, x->c, total4);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

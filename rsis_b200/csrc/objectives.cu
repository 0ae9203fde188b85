// Soft-IoU cost / loss of the training loop, fused:
//   * the per-step cost matrix of /root/reference/src/train.py:96-110 -- sigmoid(out_mask) of one decoder step against
//     ALL gt_maxseqlen ground-truth masks of the image: the reference materialises `y_pred_i.repeat(1, gtT, 1)`
//     ([B*gtT, HW]) and runs softIoU (utils/hungarian.py:64-90) as ~8 elementwise / reduction kernels, then copies the
//     [B, gtT] result to the host every step (train.py:110);
//   * the final softIoULoss rows (utils/objectives.py:27-34) and their gradient w.r.t. the mask logits.
// softIoU(target, out, e): s = sigmoid(out); num = sum(s*y); den = sum(s + y - s*y) + e; cost = 1 - num/den.
//
// One pass over HBM: a CTA owns a pixel slice of one image, loads each logit ONCE (sigmoid in registers) and walks
// the G ground-truth rows of that image with per-thread partial sums for all of them; 16-byte loads, G independent
// loads in flight per thread.  Algorithmic bytes per call: 4*B*HW (logits) + B*G*HW*{4 float | 1 uint8} (masks).
// Ground-truth masks may be float32 (the reference's own format, dataset.py:142-146) or uint8 (4x fewer bytes).
#include <math.h>

#include "common.cuh"

namespace rsis {

constexpr int kIouMaxG = 32;  // ground-truth rows walked per pass (gt_maxseqlen is 20 in the reference, args.py)

__device__ __forceinline__ void load_gt4(const float* p, float out[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
__device__ __forceinline__ void load_gt4(const uint8_t* p, float out[4]) {
  const uchar4 t = *reinterpret_cast<const uchar4*>(p);
  out[0] = (float)t.x; out[1] = (float)t.y; out[2] = (float)t.z; out[3] = (float)t.w;
}

// acc: [B][G][2] (num, sum y) followed by [B] (sum sigmoid); zero on entry.  grid = (pixel slices, B).
template <typename GT>
__global__ void __launch_bounds__(256) soft_iou_partial_kernel(const float* __restrict__ logits, const GT* __restrict__ gt,
                                                               int G, int g0, long long HW, float* __restrict__ acc,
                                                               int B, int Gtot) {
  const int b = blockIdx.y;
  const long long per = ((HW / 4 + gridDim.x - 1) / gridDim.x) * 4;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < HW ? p0 + per : HW;
  float num[kIouMaxG], sy[kIouMaxG];
#pragma unroll
  for (int g = 0; g < kIouMaxG; ++g) num[g] = sy[g] = 0.f;
  float ssig = 0.f;
  const float* lrow = logits + (size_t)b * HW;
  const GT* grow = gt + ((size_t)b * Gtot + g0) * HW;
  for (long long p = p0 + 4LL * threadIdx.x; p < p1; p += 4LL * blockDim.x) {
    const float4 m = *reinterpret_cast<const float4*>(lrow + p);
    float s[4] = {sigmoidf_acc(m.x), sigmoidf_acc(m.y), sigmoidf_acc(m.z), sigmoidf_acc(m.w)};
    ssig += (s[0] + s[1]) + (s[2] + s[3]);
#pragma unroll
    for (int g = 0; g < kIouMaxG; ++g) {
      if (g < G) {
        float y[4];
        load_gt4(grow + (size_t)g * HW + p, y);
        num[g] += (s[0] * y[0] + s[1] * y[1]) + (s[2] * y[2] + s[3] * y[3]);
        sy[g] += (y[0] + y[1]) + (y[2] + y[3]);
      }
    }
  }
  // block reduction: shuffles inside a warp, shared memory across the 8 warps, one atomic per value and block
  __shared__ float red[8][2 * kIouMaxG + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < kIouMaxG; ++g) {
    if (g < G) {
      float a = num[g], c = sy[g];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, sft);
        c += __shfl_xor_sync(0xffffffffu, c, sft);
      }
      if (lane == 0) {
        red[warp][2 * g] = a;
        red[warp][2 * g + 1] = c;
      }
    }
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) ssig += __shfl_xor_sync(0xffffffffu, ssig, sft);
  if (lane == 0) red[warp][2 * kIouMaxG] = ssig;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][i];
    atomicAdd(acc + ((size_t)b * Gtot + g0) * 2 + i, v);
  }
  if (threadIdx.x == 0 && g0 == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][2 * kIouMaxG];
    atomicAdd(acc + (size_t)B * Gtot * 2 + b, v);
  }
}

// cost[b*sb + g*sg] = weight * (1 - num / den), den = sum(s) + sum(y) - num + eps; re-zeroes acc.
__global__ void soft_iou_finish_kernel(float* __restrict__ acc, int B, int G, float eps, float weight,
                                       float* __restrict__ cost, long long sb, long long sg, float* __restrict__ num_out,
                                       float* __restrict__ den_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * G) return;
  const int b = i / G, g = i - b * G;
  const float num = acc[2 * (size_t)i], sy = acc[2 * (size_t)i + 1];
  const float ssig = acc[(size_t)B * G * 2 + b];
  const float den = ssig + sy - num + eps;
  cost[b * sb + g * sg] = weight * (1.f - num / den);
  if (num_out) num_out[i] = num;
  if (den_out) den_out[i] = den;
  acc[2 * (size_t)i] = 0.f;
  acc[2 * (size_t)i + 1] = 0.f;
}
__global__ void soft_iou_zero_sig_kernel(float* __restrict__ acc, int B, int G) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) acc[(size_t)B * G * 2 + b] = 0.f;
}

// d cost / d logit = -weight * (y*den - num*(1-y)) / den^2 * s*(1-s), times the incoming dcost of the row.
template <typename GT>
__global__ void soft_iou_bwd_kernel(const float* __restrict__ logits, const GT* __restrict__ gt, long long HW,
                                    const float* __restrict__ num, const float* __restrict__ den,
                                    const float* __restrict__ dcost, float weight, float* __restrict__ dlogits,
                                    size_t total4) {
  const long long HW4 = HW / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / HW4;
    const float n = num[r], d = den[r];
    const float k = -weight * dcost[r] / (d * d);
    const float4 m = *reinterpret_cast<const float4*>(logits + i * 4);
    float y[4];
    load_gt4(gt + i * 4, y);
    const float mv[4] = {m.x, m.y, m.z, m.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = sigmoidf_acc(mv[j]);
      o[j] = k * (y[j] * d - n * (1.f - y[j])) * s * (1.f - s);
    }
    *reinterpret_cast<float4*>(dlogits + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace rsis

using namespace rsis;

extern "C" {

size_t rsis_soft_iou_workspace_bytes(int b, int g) {
  return (b > 0 && g > 0) ? ((size_t)b * g * 2 + b) * sizeof(float) : 0;
}

int rsis_soft_iou_cost(const float* logits, const void* gt, int gt_is_u8, int b, int g, int64_t hw, float eps,
                       float weight, float* workspace, float* cost, int64_t cost_stride_b, int64_t cost_stride_g,
                       float* num_out, float* den_out, rsis_stream_t stream) {
  if (!logits || !gt || !workspace || !cost || b < 1 || g < 1 || hw < 4) return RSIS_ERR_BAD_ARG;
  if (hw % 4 != 0 || (long long)b * g > 0x7fffffffLL || b > 65535) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(logits) || (gt_is_u8 ? (reinterpret_cast<uintptr_t>(gt) & 3u) != 0 : !aligned16(gt)))
    return RSIS_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  // pixel slices: ~4 CTAs per SM, at least 1024 pixels (one float4 per thread) each
  long long slices = (4LL * 148 + b - 1) / b;
  const long long max_slices = (hw + 1023) / 1024;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  for (int g0 = 0; g0 < g; g0 += kIouMaxG) {
    const int gn = g - g0 < kIouMaxG ? g - g0 : kIouMaxG;
    const dim3 grid((unsigned)slices, (unsigned)b);
    if (gt_is_u8)
      soft_iou_partial_kernel<uint8_t><<<grid, 256, 0, st>>>(logits, reinterpret_cast<const uint8_t*>(gt), gn, g0,
                                                            (long long)hw, workspace, b, g);
    else
      soft_iou_partial_kernel<float><<<grid, 256, 0, st>>>(logits, reinterpret_cast<const float*>(gt), gn, g0,
                                                          (long long)hw, workspace, b, g);
    RSIS_CHECK_LAUNCH();
  }
  soft_iou_finish_kernel<<<ceil_div(b * g, 256), 256, 0, st>>>(workspace, b, g, eps, weight, cost,
                                                               (long long)cost_stride_b, (long long)cost_stride_g,
                                                               num_out, den_out);
  RSIS_CHECK_LAUNCH();
  soft_iou_zero_sig_kernel<<<ceil_div(b, 256), 256, 0, st>>>(workspace, b, g);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_soft_iou_bwd(const float* logits, const void* gt, int gt_is_u8, int rows, int64_t hw, const float* num,
                      const float* den, const float* dcost, float weight, float* dlogits, rsis_stream_t stream) {
  if (!logits || !gt || !num || !den || !dcost || !dlogits || rows < 1 || hw < 4) return RSIS_ERR_BAD_ARG;
  if (hw % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(logits) || !aligned16(dlogits) || (gt_is_u8 ? (reinterpret_cast<uintptr_t>(gt) & 3u) != 0 : !aligned16(gt)))
    return RSIS_ERR_ALIGN;
  const size_t total4 = (size_t)rows * (size_t)hw / 4;
  size_t blocks = (total4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (gt_is_u8)
    soft_iou_bwd_kernel<uint8_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        logits, reinterpret_cast<const uint8_t*>(gt), (long long)hw, num, den, dcost, weight, dlogits, total4);
  else
    soft_iou_bwd_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        logits, reinterpret_cast<const float*>(gt), (long long)hw, num, den, dcost, weight, dlogits, total4);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

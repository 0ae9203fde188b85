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

constexpr int kIouMaxG = 8;  // ground-truth rows per CTA: 16 accumulators keep the kernel at 4 CTAs per SM (one wave); the
                              // logits (1/gtT of the bytes) are re-read by each group of 8 rows

__device__ __forceinline__ void load_gt4(const float* p, float out[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
}
__device__ __forceinline__ void load_gt4(const uint8_t* p, float out[4]) {
  const uchar4 t = *reinterpret_cast<const uchar4*>(p);
  out[0] = (float)t.x; out[1] = (float)t.y; out[2] = (float)t.z; out[3] = (float)t.w;
}

// acc: [B][G][2] (num, sum y) followed by [B] (sum sigmoid); zero on entry.
// grid = (pixel slices, B, groups of kIouMaxG ground-truth rows).
struct IouFinish {
  float eps, weight;
  float* cost;
  long long sb, sg;
  float* num_out;
  float* den_out;
};

// The LAST CTA to arrive (ticket counter behind the sums) turns the sums into costs and re-zeroes the workspace, so
// one call is one launch.
template <typename GT>
__global__ void __launch_bounds__(256, 4) soft_iou_partial_kernel(const float* __restrict__ logits,
                                                                  const GT* __restrict__ gt, long long HW,
                                                                  float* __restrict__ acc, int B, int Gtot,
                                                                  const IouFinish fin) {
  const int b = blockIdx.y;
  const int g0 = blockIdx.z * kIouMaxG;
  const int G = Gtot - g0 < kIouMaxG ? Gtot - g0 : kIouMaxG;
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
  if (threadIdx.x == 0 && blockIdx.z == 0) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][2 * kIouMaxG];
    atomicAdd(acc + (size_t)B * Gtot * 2 + b, v);
  }
  // ---- last CTA: costs + workspace reset ----
  __shared__ unsigned s_last;
  __threadfence();
  __syncthreads();
  unsigned* counter = reinterpret_cast<unsigned*>(acc + (size_t)B * Gtot * 2 + B);
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    s_last = atomicAdd(counter, 1u) == total - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int n = B * Gtot;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int bb = i / Gtot, g = i - bb * Gtot;
    const float nm = __ldcg(acc + 2 * (size_t)i), sy_ = __ldcg(acc + 2 * (size_t)i + 1);
    const float ss = __ldcg(acc + (size_t)n * 2 + bb);
    const float den = ss + sy_ - nm + fin.eps;
    fin.cost[bb * fin.sb + g * fin.sg] = fin.weight * (1.f - nm / den);
    if (fin.num_out) fin.num_out[i] = nm;
    if (fin.den_out) fin.den_out[i] = den;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * n + B; i += blockDim.x) acc[i] = 0.f;
  if (threadIdx.x == 0) *counter = 0u;
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


// ---------------------------------------------------------------------------------------------------------------
// Hungarian matching on the device: utils/hungarian.py:91-125 runs `Munkres().compute`
// per image on the HOST over the [gtT, T] cost matrix (after a D2H copy of the costs AND of all ground-truth masks).
// Here: one CTA per image; the cost matrix is staged in shared memory and one thread runs the O(n^2 m) shortest
// augmenting path algorithm with dual potentials (Kuhn-Munkres, rectangular form) in double precision.  The problem is
// tiny (<= 32 x 32, gt_maxseqlen = 20, maxseqlen = 10): what matters is that nothing leaves the device.
// Output convention = the reference's `permute_indices`: perm[b][column] = the row assigned to that column for
// column < min(rows, cols); every other entry stays 0 (np.zeros initialisation of hungarian.py:113).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHungMax = 32;

__global__ void hungarian_kernel(const float* __restrict__ cost, long long sb, long long sr, long long sc, int R, int C,
                                 int32_t* __restrict__ perm, int perm_len, float* __restrict__ total_out) {
  __shared__ double a[kHungMax][kHungMax + 1];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
    const int r = i / C, c = i - r * C;
    // non-finite costs (a diverged step: NaN logits -> NaN soft-IoU) would leave `way[]` unset and the augmenting
    // loop without a minimum: map them to a large finite cost so the solver always terminates with a valid assignment
    const float cv = cost[b * sb + r * sr + c * sc];
    a[r][c] = isfinite(cv) ? (double)cv : (cv < 0.f ? -1e30 : 1e30);
  }
  for (int i = threadIdx.x; i < perm_len; i += blockDim.x) perm[(size_t)b * perm_len + i] = 0;
  __syncthreads();
  if (threadIdx.x != 0) return;
  // assign every index of the SMALLER side to a distinct index of the larger side: n "workers" x m "jobs", n <= m
  const bool cols_are_workers = C <= R;
  const int n = cols_are_workers ? C : R, m = cols_are_workers ? R : C;
  auto w = [&](int i, int j) { return cols_are_workers ? a[j][i] : a[i][j]; };  // cost of worker i doing job j
  double u[kHungMax + 1], v[kHungMax + 1], minv[kHungMax + 1];
  int p[kHungMax + 1], way[kHungMax + 1];
  bool used[kHungMax + 1];
  for (int j = 0; j <= m; ++j) {
    v[j] = 0;
    p[j] = 0;
    way[j] = 0;
  }
  for (int i = 0; i <= n; ++i) u[i] = 0;
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    for (int j = 0; j <= m; ++j) {
      minv[j] = 1e300;
      used[j] = false;
    }
    do {
      used[j0] = true;
      const int i0 = p[j0];
      double delta = 1e300;
      int j1 = 0;
      for (int j = 1; j <= m; ++j) {
        if (used[j]) continue;
        const double cur = w(i0 - 1, j - 1) - u[i0] - v[j];
        if (cur < minv[j]) {
          minv[j] = cur;
          way[j] = j0;
        }
        if (minv[j] < delta) {
          delta = minv[j];
          j1 = j;
        }
      }
      for (int j = 0; j <= m; ++j) {
        if (used[j]) {
          u[p[j]] += delta;
          v[j] -= delta;
        } else {
          minv[j] -= delta;
        }
      }
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  double total = 0;
  for (int j = 1; j <= m; ++j) {
    if (p[j] == 0) continue;
    const int worker = p[j] - 1, job = j - 1;
    const int row = cols_are_workers ? job : worker, col = cols_are_workers ? worker : job;
    total += a[row][col];
    if (col < perm_len) perm[(size_t)b * perm_len + col] = row;
  }
  if (total_out) total_out[b] = (float)total;
}


// ---------------------------------------------------------------------------------------------------------------
// Masked class / stop losses (utils/objectives.py:6-25 over utils/hungarian.py:10-59).  Tiny tensors ([B*T, C] and
// [B, T]): one CTA, one launch each way; the selected-row sum and count are produced on the device so that the mean of
// train.py:161,168 needs no masked_select (no host synchronisation, capturable).
// A row is selected when (uint8)sw != 0 -- the reference's `sw.byte()` mask.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sw_selected(float sw) { return (unsigned char)sw != 0; }

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(256) masked_nll_fwd_kernel(const float* __restrict__ probs,
                                                             const long long* __restrict__ target,
                                                             const float* __restrict__ sw,
                                                             const float* __restrict__ balance, int rows, int C,
                                                             float* __restrict__ cost_rows, float* __restrict__ sum_count) {
  __shared__ float red[32];
  float s = 0.f, n = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const long long t = target[r];
    float cost = 0.f;
    const bool sel = sw_selected(sw[r]);
    if (t >= 0 && t < C) {
      cost = -logf(probs[(size_t)r * C + t]);
      if (balance) cost *= balance[t];
    }
    if (cost_rows) cost_rows[r] = sel ? cost : 0.f;
    if (sel) {
      s += cost;
      n += 1.f;
    }
  }
  s = block_sum_256(s, red);
  n = block_sum_256(n, red);
  if (threadIdx.x == 0) {
    sum_count[0] = s;
    sum_count[1] = n;
  }
}

__global__ void __launch_bounds__(256) masked_nll_bwd_kernel(const float* __restrict__ probs,
                                                             const long long* __restrict__ target,
                                                             const float* __restrict__ sw,
                                                             const float* __restrict__ balance,
                                                             const float* __restrict__ dcost, long long dstride, int rows,
                                                             int C, float* __restrict__ dprobs) {
  const size_t total = (size_t)rows * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i - (size_t)r * C);
    float g = 0.f;
    if (sw_selected(sw[r]) && (long long)c == target[r])
      g = -dcost[(size_t)r * dstride] * (balance ? balance[c] : 1.f) / probs[i];
    dprobs[i] = g;
  }
}

// StableBalancedMaskedBCE (hungarian.py:34-59): lv = out - out*t + max(-out,0) + log(exp(-max) + exp(-out-max));
// cost = (1-bw)*lv*t + bw*lv*(1-t);  bw = balance_weight, or sum(t) / n when none is given (computed here, pass 1).
__global__ void __launch_bounds__(256) masked_bce_fwd_kernel(const float* __restrict__ target,
                                                             const float* __restrict__ logits,
                                                             const float* __restrict__ sw, float balance_weight,
                                                             long long n, float* __restrict__ cost_rows,
                                                             float* __restrict__ out3) {
  __shared__ float red[32];
  float bw = balance_weight;
  if (bw < 0.f) {
    float pos = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) pos += target[i];
    pos = block_sum_256(pos, red);
    bw = pos / (float)n;  // num_positive / (num_positive + num_negative), the latter sum being n exactly
  }
  float s = 0.f, cnt = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float o = logits[i], t = target[i];
    const float mx = fmaxf(-o, 0.f);
    const float lv = o - o * t + mx + logf(expf(-mx) + expf(-o - mx));
    const float cost = (1.f - bw) * lv * t + bw * lv * (1.f - t);
    const bool sel = sw_selected(sw[i]);
    if (cost_rows) cost_rows[i] = sel ? cost : 0.f;
    if (sel) {
      s += cost;
      cnt += 1.f;
    }
  }
  s = block_sum_256(s, red);
  cnt = block_sum_256(cnt, red);
  if (threadIdx.x == 0) {
    out3[0] = s;
    out3[1] = cnt;
    out3[2] = bw;
  }
}

__global__ void __launch_bounds__(256) masked_bce_bwd_kernel(const float* __restrict__ target,
                                                             const float* __restrict__ logits,
                                                             const float* __restrict__ sw, const float* __restrict__ bw_ptr,
                                                             const float* __restrict__ dcost, long long dstride,
                                                             long long n, float* __restrict__ dlogits) {
  const float bw = *bw_ptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = 0.f;
    if (sw_selected(sw[i])) {
      const float o = logits[i], t = target[i];
      const float sig = 1.f / (1.f + expf(-o));
      // d lv / d out = sigmoid(out) - t
      g = dcost[i * dstride] * ((1.f - bw) * t + bw * (1.f - t)) * (sig - t);
    }
    dlogits[i] = g;
  }
}

}  // namespace rsis

using namespace rsis;

extern "C" {

size_t rsis_soft_iou_workspace_bytes(int b, int g) {
  return (b > 0 && g > 0) ? ((size_t)b * g * 2 + b + 1) * sizeof(float) : 0;  // sums + one ticket counter
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
  // ONE wave: 4 CTAs per SM over (pixel slices x images x row groups), at least 1024 pixels per slice
  const int groups = ceil_div(g, kIouMaxG);
  if (groups > 65535) return RSIS_ERR_UNSUPPORTED;
  long long slices = (4LL * 148) / ((long long)b * groups);
  const long long max_slices = (hw + 1023) / 1024;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  const dim3 grid((unsigned)slices, (unsigned)b, (unsigned)groups);
  const IouFinish fin{eps, weight, cost, (long long)cost_stride_b, (long long)cost_stride_g, num_out, den_out};
  if (gt_is_u8)
    soft_iou_partial_kernel<uint8_t><<<grid, 256, 0, st>>>(logits, reinterpret_cast<const uint8_t*>(gt), (long long)hw,
                                                          workspace, b, g, fin);
  else
    soft_iou_partial_kernel<float><<<grid, 256, 0, st>>>(logits, reinterpret_cast<const float*>(gt), (long long)hw,
                                                        workspace, b, g, fin);
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

int rsis_hungarian_match(const float* cost, int64_t stride_b, int64_t stride_r, int64_t stride_c, int b, int rows,
                         int cols, int32_t* perm, int perm_len, float* total_cost, rsis_stream_t stream) {
  if (!cost || !perm || b < 1 || rows < 1 || cols < 1 || perm_len < 1) return RSIS_ERR_BAD_ARG;
  if (rows > kHungMax || cols > kHungMax) return RSIS_ERR_UNSUPPORTED;
  hungarian_kernel<<<b, 32, 0, (cudaStream_t)stream>>>(cost, (long long)stride_b, (long long)stride_r,
                                                      (long long)stride_c, rows, cols, perm, perm_len, total_cost);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_masked_nll_fwd(const float* probs, const int64_t* target, const float* sw, const float* balance, int rows,
                        int num_classes, float* cost_rows, float* sum_count, rsis_stream_t stream) {
  if (!probs || !target || !sw || !sum_count || rows < 1 || num_classes < 1) return RSIS_ERR_BAD_ARG;
  masked_nll_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(probs, reinterpret_cast<const long long*>(target), sw,
                                                             balance, rows, num_classes, cost_rows, sum_count);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_masked_nll_bwd(const float* probs, const int64_t* target, const float* sw, const float* balance,
                        const float* dcost, int64_t dcost_stride, int rows, int num_classes, float* dprobs,
                        rsis_stream_t stream) {
  if (!probs || !target || !sw || !dcost || !dprobs || rows < 1 || num_classes < 1 || dcost_stride < 0)
    return RSIS_ERR_BAD_ARG;
  const size_t total = (size_t)rows * num_classes;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  masked_nll_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      probs, reinterpret_cast<const long long*>(target), sw, balance, dcost, (long long)dcost_stride, rows, num_classes,
      dprobs);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_masked_bce_fwd(const float* target, const float* logits, const float* sw, float balance_weight, int64_t n,
                        float* cost_rows, float* sum_count_bw, rsis_stream_t stream) {
  if (!target || !logits || !sw || !sum_count_bw || n < 1) return RSIS_ERR_BAD_ARG;
  masked_bce_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(target, logits, sw, balance_weight, (long long)n, cost_rows,
                                                             sum_count_bw);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_masked_bce_bwd(const float* target, const float* logits, const float* sw, const float* balance_weight,
                        const float* dcost, int64_t dcost_stride, int64_t n, float* dlogits, rsis_stream_t stream) {
  if (!target || !logits || !sw || !balance_weight || !dcost || !dlogits || n < 1 || dcost_stride < 0)
    return RSIS_ERR_BAD_ARG;
  size_t blocks = ((size_t)n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  masked_bce_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(target, logits, sw, balance_weight, dcost,
                                                                            (long long)dcost_stride, (long long)n, dlogits);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

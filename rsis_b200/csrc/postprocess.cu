// Evaluation post-processing on the device: threshold + ignore mask + area + column-major
// run-length encoding of the predicted masks -- what /root/reference/src/eval.py:97-127 does per instance on the host
// (`(pred_mask > th).astype("uint8")`, `segmentation[ignore_pixels==1] = 0`, `np.sum(segmentation)`,
// `mask.encode(np.asfortranarray(...))` = rleEncode of /root/reference/src/coco/common/maskApi.c:32-41) after copying
// the [B,T,H,W] float masks to the host.  Here only the run lengths leave the device.
//
//   rle_bits_kernel   thresholds the row-major float mask and stores it TRANSPOSED (column-major bytes, the scan order
//                     of the COCO format) through a 32x32 shared-memory tile: coalesced on both sides;
//   rle_scan_kernel   one CTA per mask: every thread owns a contiguous chunk of the column-major bytes, counts the value
//                     changes in it (the predecessor of element 0 is 0, maskApi.c:36), a block-wide exclusive scan
//                     gives each chunk its output offset, a second walk writes the change positions, and the run lengths
//                     are their adjacent differences: cnts[0] = leading zeros (0 when the mask starts with a one).
// HBM-bound: 4 bytes read + 1 written + 1 read (L2) per pixel.
#include "common.cuh"

namespace rsis {

__global__ void rle_bits_kernel(const float* __restrict__ masks, const uint8_t* __restrict__ ignore, float th, int H,
                                int W, uint8_t* __restrict__ bits) {
  __shared__ uint8_t tile[32][33];
  const size_t a = (size_t)H * W;
  const float* m = masks + blockIdx.z * a;
  const uint8_t* ig = ignore ? ignore + blockIdx.z * a : nullptr;
  uint8_t* out = bits + blockIdx.z * a;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int y = y0 + r, x = x0 + threadIdx.x;
    uint8_t v = 0;
    if (y < H && x < W) {
      v = m[(size_t)y * W + x] > th ? 1 : 0;
      if (ig && ig[(size_t)y * W + x] == 1) v = 0;
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = x0 + r, y = y0 + threadIdx.x;
    if (x < W && y < H) out[(size_t)x * H + y] = tile[threadIdx.x][r];
  }
}

constexpr int kRleThreads = 1024;

__global__ void __launch_bounds__(kRleThreads) rle_scan_kernel(const uint8_t* __restrict__ bits, long long a,
                                                               uint32_t* __restrict__ pos, uint32_t* __restrict__ counts,
                                                               int max_runs, int32_t* __restrict__ n_runs,
                                                               uint32_t* __restrict__ areas) {
  __shared__ unsigned s_warp[32], s_area[32];
  __shared__ unsigned s_total;
  const int n = blockIdx.x;
  const uint8_t* t = bits + (size_t)n * a;
  uint32_t* mypos = pos + (size_t)n * (a + 1);
  const long long chunk = (a + kRleThreads - 1) / kRleThreads;
  const long long j0 = (long long)threadIdx.x * chunk;
  const long long j1 = j0 + chunk < a ? j0 + chunk : a;
  // pass 1: value changes and ones in my chunk
  unsigned changes = 0, ones = 0;
  uint8_t p = j0 > 0 && j0 < a ? t[j0 - 1] : 0;
  for (long long j = j0; j < j1; ++j) {
    const uint8_t v = t[j];
    changes += v != p;
    ones += v;
    p = v;
  }
  // block-wide exclusive scan of `changes` (and sum of `ones`)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = changes;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const unsigned o = __shfl_up_sync(0xffffffffu, incl, s);
    if (lane >= s) incl += o;
  }
  unsigned asum = ones;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, s);
  if (lane == 31) s_warp[warp] = incl;
  if (lane == 0) s_area[warp] = asum;
  __syncthreads();
  if (warp == 0) {
    unsigned w = s_warp[lane], wi = w;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, wi, s);
      if (lane >= s) wi += o;
    }
    s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    if (lane == 31) s_total = wi;
    unsigned ar = s_area[lane];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ar += __shfl_xor_sync(0xffffffffu, ar, s);
    if (lane == 0 && areas) areas[n] = ar;
  }
  __syncthreads();
  unsigned off = s_warp[warp] + incl - changes;
  const unsigned K = s_total;  // number of value changes; runs = K + 1
  // pass 2: positions of the changes
  p = j0 > 0 && j0 < a ? t[j0 - 1] : 0;
  for (long long j = j0; j < j1; ++j) {
    const uint8_t v = t[j];
    if (v != p) mypos[off++] = (uint32_t)j;
    p = v;
  }
  __syncthreads();  // the positions were written by this block: visible block-wide after the barrier
  if (threadIdx.x == 0) n_runs[n] = (int32_t)(K + 1);
  uint32_t* mycnt = counts + (size_t)n * max_runs;
  for (unsigned i = threadIdx.x; i <= K && i < (unsigned)max_runs; i += blockDim.x) {
    const uint32_t hi = i < K ? mypos[i] : (uint32_t)a;
    const uint32_t lo = i > 0 ? mypos[i - 1] : 0u;
    mycnt[i] = hi - lo;
  }
}

}  // namespace rsis

using namespace rsis;

extern "C" {

size_t rsis_rle_workspace_bytes(int n, int h, int w) {
  if (n < 1 || h < 1 || w < 1) return 0;
  const size_t a = (size_t)h * w;
  return (((size_t)n * a + 15) / 16) * 16 + (size_t)n * (a + 1) * sizeof(uint32_t);
}

int rsis_rle_encode(const float* masks, float threshold, const uint8_t* ignore, int n, int h, int w, void* workspace,
                    uint32_t* counts, int max_runs, int32_t* n_runs, uint32_t* areas, rsis_stream_t stream) {
  if (!masks || !workspace || !counts || !n_runs || n < 1 || h < 1 || w < 1 || max_runs < 1) return RSIS_ERR_BAD_ARG;
  const size_t a = (size_t)h * w;
  if (a >= 0xffffffffull || n > 65535) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(workspace)) return RSIS_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* bits = reinterpret_cast<uint8_t*>(workspace);
  uint32_t* pos = reinterpret_cast<uint32_t*>(bits + (((size_t)n * a + 15) / 16) * 16);
  const dim3 grid(ceil_div(w, 32), ceil_div(h, 32), n);
  if (grid.y > 65535) return RSIS_ERR_UNSUPPORTED;
  rle_bits_kernel<<<grid, dim3(32, 8), 0, st>>>(masks, ignore, threshold, h, w, bits);
  RSIS_CHECK_LAUNCH();
  rle_scan_kernel<<<n, kRleThreads, 0, st>>>(bits, (long long)a, pos, counts, max_runs, n_runs, areas);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

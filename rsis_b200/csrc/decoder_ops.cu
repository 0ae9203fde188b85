// HBM-streaming primitives around the convolutions: max-pool, align-corners bilinear upsampling, the 1-channel
// mask head and the class/stop heads.  All are element-parallel, read NHWC with 16-byte channel vectors and write
// each output once.
#include <math.h>

#include "common.cuh"

namespace rsis {

__device__ __forceinline__ void load4(const View& v, size_t idx, float out[4]) {
  if (v.fmt == RSIS_FMT_F32) {
    const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.p) + idx);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else {
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(v.p);
    const uint2 h = *reinterpret_cast<const uint2*>(b + idx);
    const uint2 l = *reinterpret_cast<const uint2*>(b + idx + v.plane);
    const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&h);
    const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = __bfloat162float(hp[j]) + __bfloat162float(lp[j]);
  }
}

__device__ __forceinline__ void store4v(void* p, size_t plane, int fmt, size_t idx, const float v[4]) {
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

// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) -- /root/reference/src/modules/vision.py:15
__global__ void maxpool3x3s2_kernel(View x, void* y, size_t y_plane, int y_fmt, int N, int H, int W, int C, int Ho,
                                    int Wo) {
  pdl_trigger();
  const int C4 = C >> 2;
  const size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t r = i / C4;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hi = ho * 2 - 1 + dh;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const int wi = wo * 2 - 1 + dw;
        if (wi < 0 || wi >= W) continue;
        float v[4];
        load4(x, (((size_t)n * H + hi) * W + wi) * C + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], v[j]);
      }
    }
    store4v(y, y_plane, y_fmt, (((size_t)n * Ho + ho) * Wo + wo) * C + c, best);
  }
}

// nn.UpsamplingBilinear2d(size) == bilinear with align_corners=True -- /root/reference/src/modules/model.py:149-150,
// 163-164.  Index arithmetic follows ATen's area_pixel_compute_source_index(align_corners=true): src = scale * dst with
// scale = (in - 1) / (out - 1) evaluated in float.
struct UpsampleArgs {
  View x;
  void* y;
  size_t y_plane;
  int y_fmt, y_cs, N, H, W, C, Ho, Wo;
  float sh, sw;
};

__device__ __forceinline__ void upsample_bilinear_body(const UpsampleArgs& a, size_t first, size_t stride) {
  const int C4 = a.C >> 2;
  const size_t total = (size_t)a.N * a.Ho * a.Wo * C4;
  for (size_t i = first; i < total; i += stride) {
    const int c = (int)(i % C4) * 4;
    size_t r = i / C4;
    const int wo = (int)(r % a.Wo);
    r /= a.Wo;
    const int ho = (int)(r % a.Ho);
    const int n = (int)(r / a.Ho);
    const float fh = a.sh * ho, fw = a.sw * wo;
    const int h1 = min((int)fh, a.H - 1), w1 = min((int)fw, a.W - 1);  // ATen guard_index_and_lambda
    const int h1p = h1 < a.H - 1 ? 1 : 0, w1p = w1 < a.W - 1 ? 1 : 0;
    const float h1l = fminf(fmaxf(fh - h1, 0.f), 1.f), h0l = 1.f - h1l;
    const float w1l = fminf(fmaxf(fw - w1, 0.f), 1.f), w0l = 1.f - w1l;
    float v00[4], v01[4], v10[4], v11[4], o[4];
    const size_t base = (((size_t)n * a.H + h1) * a.W + w1) * a.C + c;
    load4(a.x, base, v00);
    load4(a.x, base + (size_t)w1p * a.C, v01);
    load4(a.x, base + (size_t)h1p * a.W * a.C, v10);
    load4(a.x, base + (size_t)h1p * a.W * a.C + (size_t)w1p * a.C, v11);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = h0l * (w0l * v00[j] + w1l * v01[j]) + h1l * (w0l * v10[j] + w1l * v11[j]);
    store4v(a.y, a.y_plane, a.y_fmt, (((size_t)n * a.Ho + ho) * a.Wo + wo) * a.y_cs + c, o);
  }
}

__global__ void upsample_bilinear_kernel(const UpsampleArgs a) {
  pdl_trigger();
  upsample_bilinear_body(a, blockIdx.x * (size_t)blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// Several independent upsamplings in one launch (the decoder's wavefront schedule: the x2 upsamplings feeding levels
// 1..4 at four different time-steps); problem i owns the block range [first[i], first[i+1]).
constexpr int kMaxUpsampleGroup = 4;
struct UpsampleGroup {
  UpsampleArgs a[kMaxUpsampleGroup];
  int first[kMaxUpsampleGroup + 1];
  int n;
};

__global__ void upsample_bilinear_group_kernel(const __grid_constant__ UpsampleGroup g) {
  pdl_trigger();
  int i = 0;
#pragma unroll
  for (int k = 1; k < kMaxUpsampleGroup; ++k)
    if (k < g.n && (int)blockIdx.x >= g.first[k]) i = k;
  const int nb = g.first[i + 1] - g.first[i];
  upsample_bilinear_body(g.a[i], (size_t)(blockIdx.x - g.first[i]) * blockDim.x + threadIdx.x, (size_t)nb * blockDim.x);
}

// conv_out (model.py:167): k x k conv (k = 1 or 3), Cin -> 1, + bias; optional sigmoid copy (test.py:50).
// One thread per output pixel; weights staged in shared memory as [tap][c].
__global__ void mask_head_kernel(const float* __restrict__ x, const float* __restrict__ w_oihw,
                                 const float* __restrict__ bias, float* __restrict__ logits,
                                 float* __restrict__ prob_out, long long prob_stride_n, int N, int H, int W, int C,
                                 int ks) {
  pdl_trigger();
  extern __shared__ float sw[];
  const int taps = ks * ks;
  for (int i = threadIdx.x; i < taps * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;
    sw[i] = w_oihw[c * taps + tap];
  }
  __syncthreads();
  const float b = bias ? bias[0] : 0.f;
  const int pad = ks / 2;
  const size_t HW = (size_t)H * W;
  const size_t total = (size_t)N * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wq = (int)(i % W);
    const int hq = (int)((i / W) % H);
    const size_t n = i / HW;
    float acc = 0.f;
    for (int kh = 0; kh < ks; ++kh) {
      const int hi = hq - pad + kh;
      if (hi < 0 || hi >= H) continue;
      for (int kw = 0; kw < ks; ++kw) {
        const int wi = wq - pad + kw;
        if (wi < 0 || wi >= W) continue;
        const float* px = x + ((n * H + hi) * W + wi) * C;
        const float* pw = sw + (kh * ks + kw) * C;
        for (int c = 0; c < C; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(px + c);
          acc = fmaf(v.x, pw[c], acc);
          acc = fmaf(v.y, pw[c + 1], acc);
          acc = fmaf(v.z, pw[c + 2], acc);
          acc = fmaf(v.w, pw[c + 3], acc);
        }
      }
    }
    acc += b;
    if (logits) logits[i] = acc;
    if (prob_out) prob_out[n * prob_stride_n + (i - n * HW)] = sigmoidf_acc(acc);
  }
}

// Fused tail of the decoder step (model.py:163-167): x2 align-corners bilinear upsampling of the last hidden state
// followed by conv_out (k x k, C -> 1) (+ the sigmoid / stacking of test.py:46-50).  A 32 x 32 output tile's
// (32 + 2*pad)^2 x C window of the UPSAMPLED map is built in shared memory straight from the low-resolution hidden
// state (which stays L2/L1 resident), so the upsampled tensor -- 4x the hidden state -- never exists in HBM.
// Interpolation and accumulation orders equal upsample_bilinear_kernel + mask_head_kernel: bit-identical results.
constexpr int kMaskTile = 32;
constexpr int kMaskSrc = 20;  // source pixels per side a 34 x 34 upsampled tile can touch at scale <= 1/2 (+ the bilinear neighbour)
__global__ void __launch_bounds__(256)
upsample_mask_head_kernel(const float* __restrict__ h, const float* __restrict__ w_oihw, const float* __restrict__ bias,
                          float* __restrict__ logits, float* __restrict__ prob_out, long long prob_stride_n, int H,
                          int W, int C, int Ho, int Wo, int ks, float sh, float sw, int n_inner,
                          long long prob_stride_t) {
  pdl_trigger();
  extern __shared__ __align__(16) float smem_mask[];
  const int pad = ks / 2;
  const int TW = kMaskTile + 2 * pad;
  float* up = smem_mask;                 // [TW][TW][C]
  float* sw_ = smem_mask + TW * TW * C;  // [tap][C]
  const int taps = ks * ks;
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * kMaskTile, ox0 = blockIdx.x * kMaskTile;
  for (int i = threadIdx.x; i < taps * C; i += blockDim.x) {
    const int tap = i / C, c = i % C;
    sw_[i] = w_oihw[c * taps + tap];
  }
  const int C4 = C >> 2;
  const float* hn = h + (size_t)n * H * W * C;
  for (int i = threadIdx.x; i < TW * TW * C4; i += blockDim.x) {
    const int c = (i % C4) * 4;
    const int t = i / C4;
    const int tx = t % TW, ty = t / TW;
    const int oy = oy0 + ty - pad, ox = ox0 + tx - pad;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);  // conv_out's zero padding outside the upsampled map
    if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
      const float fh = sh * oy, fw = sw * ox;
      const int h1 = min((int)fh, H - 1), w1 = min((int)fw, W - 1);
      const int h1p = h1 < H - 1 ? 1 : 0, w1p = w1 < W - 1 ? 1 : 0;
      const float h1l = fminf(fmaxf(fh - h1, 0.f), 1.f), h0l = 1.f - h1l;
      const float w1l = fminf(fmaxf(fw - w1, 0.f), 1.f), w0l = 1.f - w1l;
      const float* b = hn + ((size_t)h1 * W + w1) * C + c;
      const float4 v00 = __ldg(reinterpret_cast<const float4*>(b));
      const float4 v01 = __ldg(reinterpret_cast<const float4*>(b + (size_t)w1p * C));
      const float4 v10 = __ldg(reinterpret_cast<const float4*>(b + (size_t)h1p * W * C));
      const float4 v11 = __ldg(reinterpret_cast<const float4*>(b + (size_t)h1p * W * C + (size_t)w1p * C));
      o.x = h0l * (w0l * v00.x + w1l * v01.x) + h1l * (w0l * v10.x + w1l * v11.x);
      o.y = h0l * (w0l * v00.y + w1l * v01.y) + h1l * (w0l * v10.y + w1l * v11.y);
      o.z = h0l * (w0l * v00.z + w1l * v01.z) + h1l * (w0l * v10.z + w1l * v11.z);
      o.w = h0l * (w0l * v00.w + w1l * v01.w) + h1l * (w0l * v10.w + w1l * v11.w);
    }
    *reinterpret_cast<float4*>(up + (size_t)t * C + c) = o;
  }
  __syncthreads();
  const float bs = bias ? bias[0] : 0.f;
  const int px = threadIdx.x % kMaskTile;
  for (int py = threadIdx.x / kMaskTile; py < kMaskTile; py += 256 / kMaskTile) {
    const int oy = oy0 + py, ox = ox0 + px;
    if (oy >= Ho || ox >= Wo) continue;
    float acc = 0.f;
    for (int kh = 0; kh < ks; ++kh) {
      const int hi = oy - pad + kh;
      if (hi < 0 || hi >= Ho) continue;  // same skipping (hence summation order) as mask_head_kernel
      for (int kw = 0; kw < ks; ++kw) {
        const int wi = ox - pad + kw;
        if (wi < 0 || wi >= Wo) continue;
        const float* pxl = up + ((size_t)(py + kh) * TW + (px + kw)) * C;
        const float* pw = sw_ + (kh * ks + kw) * C;
        for (int c = 0; c < C; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(pxl + c);
          acc = fmaf(v.x, pw[c], acc);
          acc = fmaf(v.y, pw[c + 1], acc);
          acc = fmaf(v.z, pw[c + 2], acc);
          acc = fmaf(v.w, pw[c + 3], acc);
        }
      }
    }
    acc += bs;
    const size_t pix = (size_t)oy * Wo + ox;
    if (logits) logits[(size_t)n * Ho * Wo + pix] = acc;
    if (prob_out) prob_out[(size_t)(n % n_inner) * prob_stride_n + (size_t)(n / n_inner) * prob_stride_t + pix] = sigmoidf_acc(acc);
  }
}

// Compile-time specialisation of the kernel above for the reference's shapes (hidden/16 = 8 channels, 3x3 or 1x1
// conv_out): the generic kernel spends ~5500 instructions per thread, mostly loop / index overhead of its run-time
// (ks, C) loops (ncu: issue slots 65 % busy, 29.6 us); here everything unrolls, the tile geometry divides by
// constants, and a thread owns 4 VERTICALLY ADJACENT outputs so the 6 x KS window rows it needs are read from shared
// memory once.  The staged tile holds zeros outside the map, so out-of-range taps contribute fma(0, w, acc) = acc:
// the accumulation order (kh, kw, c) and hence every result bit equals the generic kernel's.
template <int C, int KS>
__global__ void __launch_bounds__(256, 3)  // (4 blocks per SM = 64 registers: spills, measured slower)
upsample_mask_head_fixed_kernel(const float* __restrict__ h, const float* __restrict__ w_oihw,
                                const float* __restrict__ bias, float* __restrict__ logits,
                                float* __restrict__ prob_out, long long prob_stride_n, int H, int W, int Ho, int Wo,
                                float sh, float sw, int n_inner, long long prob_stride_t) {
  pdl_trigger();
  constexpr int PAD = KS / 2, TW = kMaskTile + 2 * PAD, TAPS = KS * KS, C4 = C / 4;
  extern __shared__ __align__(16) float smem_mask[];
  float* up = smem_mask;                 // [TW][TW][C]
  float* sw_ = smem_mask + TW * TW * C;  // [tap][C]
  float* src = sw_ + TAPS * C;           // [kMaskSrc][kMaskSrc][C]: the source window of this tile
  const int n = blockIdx.z;
  const int oy0 = blockIdx.y * kMaskTile, ox0 = blockIdx.x * kMaskTile;
  for (int i = threadIdx.x; i < TAPS * C; i += 256) sw_[i] = w_oihw[(i % C) * TAPS + i / C];
  const float* hn = h + (size_t)n * H * W * C;
  // The interpolation below reads four source pixels per upsampled value.  Straight from global memory that is nine
  // dependent rounds of loads per thread (~20 us per block when the hidden state comes from HBM: the all-steps launch
  // took 125 us); the tile's source window (<= 20 x 20 pixels for the x2 case) is therefore brought into shared memory
  // first, in ONE round of independent, coalesced loads.  Same values, same formula: bit-identical results.
  const int oyA = max(oy0 - PAD, 0), oyB = min(oy0 + kMaskTile + PAD - 1, Ho - 1);
  const int oxA = max(ox0 - PAD, 0), oxB = min(ox0 + kMaskTile + PAD - 1, Wo - 1);
  const int sy0 = min((int)(sh * oyA), H - 1), sy1 = min(min((int)(sh * oyB), H - 1) + 1, H - 1);
  const int sx0 = min((int)(sw * oxA), W - 1), sx1 = min(min((int)(sw * oxB), W - 1) + 1, W - 1);
  const int nsy = sy1 - sy0 + 1, nsx = sx1 - sx0 + 1;
  const bool staged = nsy <= kMaskSrc && nsx <= kMaskSrc;  // (block-uniform; other scales read global memory directly)
  if (staged) {
    for (int i = threadIdx.x; i < nsy * nsx * C4; i += 256) {
      const int c = (i % C4) * 4, p = i / C4, x = p % nsx, y = p / nsx;
      *reinterpret_cast<float4*>(src + ((size_t)y * kMaskSrc + x) * C + c) =
          __ldg(reinterpret_cast<const float4*>(hn + ((size_t)(sy0 + y) * W + (sx0 + x)) * C + c));
    }
  }
  // Per tile row / column of the upsampled window: source offset (elements; -1 outside the map), step to the bilinear
  // neighbour and its weight -- computed ONCE per block (2 x TW threads) instead of once per interpolated value: the
  // per-value index arithmetic (int <-> float conversions, clamps, divisions) was half of this kernel's instructions.
  int* roff = reinterpret_cast<int*>(src + kMaskSrc * kMaskSrc * C);  // [TW] row offset, [TW] row step
  int* coff = roff + 2 * TW;                                           // [TW] column offset, [TW] column step
  float* rl = reinterpret_cast<float*>(coff + 2 * TW);                 // [TW] h1l, [TW] w1l
  const int rstride = staged ? kMaskSrc * C : W * C;
  if (threadIdx.x < TW) {
    const int oy = oy0 + (int)threadIdx.x - PAD;
    int off = -1, step = 0;
    float l = 0.f;
    if (oy >= 0 && oy < Ho) {
      const float fh = sh * oy;
      const int h1 = min((int)fh, H - 1);
      off = (staged ? h1 - sy0 : h1) * rstride;
      step = h1 < H - 1 ? rstride : 0;
      l = fminf(fmaxf(fh - h1, 0.f), 1.f);
    }
    roff[threadIdx.x] = off;
    roff[TW + threadIdx.x] = step;
    rl[threadIdx.x] = l;
  } else if (threadIdx.x >= 64 && threadIdx.x < 64 + TW) {
    const int tx = (int)threadIdx.x - 64;
    const int ox = ox0 + tx - PAD;
    int off = -1, step = 0;
    float l = 0.f;
    if (ox >= 0 && ox < Wo) {
      const float fw = sw * ox;
      const int w1 = min((int)fw, W - 1);
      off = (staged ? w1 - sx0 : w1) * C;
      step = w1 < W - 1 ? C : 0;
      l = fminf(fmaxf(fw - w1, 0.f), 1.f);
    }
    coff[tx] = off;
    coff[TW + tx] = step;
    rl[TW + tx] = l;
  }
  __syncthreads();
  const float* sbase = staged ? src : hn;
  for (int i = threadIdx.x; i < TW * TW * C4; i += 256) {
    const int c = (i % C4) * 4;
    const int t = i / C4;
    const int tx = t % TW, ty = t / TW;
    const int ro = roff[ty], co = coff[tx];
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ro >= 0 && co >= 0) {
      const float h1l = rl[ty], h0l = 1.f - h1l;
      const float w1l = rl[TW + tx], w0l = 1.f - w1l;
      // one code path for both sources (generic loads), so the interpolation below compiles to ONE instruction sequence
      // -- the one of upsample_bilinear_kernel, whose results this kernel reproduces bit for bit
      const float* b = sbase + ro + co + c;
      const int rs = roff[TW + ty], cs = coff[TW + tx];
      const float4 v00 = *reinterpret_cast<const float4*>(b);
      const float4 v01 = *reinterpret_cast<const float4*>(b + cs);
      const float4 v10 = *reinterpret_cast<const float4*>(b + rs);
      const float4 v11 = *reinterpret_cast<const float4*>(b + rs + cs);
      o.x = h0l * (w0l * v00.x + w1l * v01.x) + h1l * (w0l * v10.x + w1l * v11.x);
      o.y = h0l * (w0l * v00.y + w1l * v01.y) + h1l * (w0l * v10.y + w1l * v11.y);
      o.z = h0l * (w0l * v00.z + w1l * v01.z) + h1l * (w0l * v10.z + w1l * v11.z);
      o.w = h0l * (w0l * v00.w + w1l * v01.w) + h1l * (w0l * v10.w + w1l * v11.w);
    }
    *reinterpret_cast<float4*>(up + ((size_t)(c >> 2) * (TW * TW) + t) * 4) = o;  // [c / 4][pixel][4]: see below
  }
  __syncthreads();
  const float bs = bias ? bias[0] : 0.f;
  const int px = threadIdx.x % kMaskTile;
  const int py0 = (threadIdx.x / kMaskTile) * 4;   // 8 thread rows x 4 outputs = 32 tile rows
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // output q reads window rows py0 + q + kh; every acc[q] still accumulates in (kh, kw, c) order
#pragma unroll
  for (int kh = 0; kh < KS; ++kh) {
#pragma unroll
    for (int kw = 0; kw < KS; ++kw) {
      float wv[C];
#pragma unroll
      for (int c = 0; c < C; c += 4) {
        const float4 t = *reinterpret_cast<const float4*>(sw_ + (kh * KS + kw) * C + c);  // warp-uniform: broadcast
        wv[c] = t.x; wv[c + 1] = t.y; wv[c + 2] = t.z; wv[c + 3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // the window is stored as C / 4 planes of [pixel][4 floats]: the 32 lanes of a warp (32 adjacent pixels) read 512
        // contiguous bytes per float4 load.  With [pixel][C] (32 bytes per pixel) every such load was a 2-way bank
        // conflict, and the kernel ran at the shared-memory bandwidth that left (ncu: 115 us for the ten steps of a pass)
        const float* pxl = up + ((size_t)(py0 + q + kh) * TW + (px + kw)) * 4;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(pxl + (size_t)(c >> 2) * (TW * TW * 4));
          acc[q] = fmaf(v.x, wv[c], acc[q]);
          acc[q] = fmaf(v.y, wv[c + 1], acc[q]);
          acc[q] = fmaf(v.z, wv[c + 2], acc[q]);
          acc[q] = fmaf(v.w, wv[c + 3], acc[q]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oy = oy0 + py0 + q, ox = ox0 + px;
    if (oy >= Ho || ox >= Wo) continue;
    const float a = acc[q] + bs;
    const size_t pix = (size_t)oy * Wo + ox;
    if (logits) logits[(size_t)n * Ho * Wo + pix] = a;
    if (prob_out) prob_out[(size_t)(n % n_inner) * prob_stride_n + (size_t)(n / n_inner) * prob_stride_t + pix] = sigmoidf_acc(a);
  }
}

// fc_class + Softmax + fc_stop on the max-pooled side features (model.py:169-182). One CTA per image.
__global__ void class_stop_heads_kernel(const uint32_t* __restrict__ side_max, int F, const float* __restrict__ w_class,
                                        const float* __restrict__ b_class, int num_classes,
                                        const float* __restrict__ w_stop, const float* __restrict__ b_stop,
                                        float* __restrict__ feat_out, float* __restrict__ class_probs,
                                        long long class_stride, float* __restrict__ stop_logit,
                                        float* __restrict__ stop_prob, long long stop_stride, int n_inner,
                                        long long class_stride_t, long long stop_stride_t) {
  pdl_trigger();
  extern __shared__ float sm[];
  float* feat = sm;            // [F]
  float* logit = sm + F;       // [num_classes + 1]; the last entry is the stop logit
  const int n = blockIdx.x;
  // image n = step (n / n_inner), batch element (n % n_inner): the all-steps form writes [b][t] outputs
  const size_t co = (size_t)(n % n_inner) * class_stride + (size_t)(n / n_inner) * class_stride_t;
  const size_t so = (size_t)(n % n_inner) * stop_stride + (size_t)(n / n_inner) * stop_stride_t;
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    const float v = key_to_float(side_max[(size_t)n * F + i]);
    feat[i] = v;
    if (feat_out) feat_out[(size_t)n * F + i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = warp; o <= num_classes; o += nwarps) {
    const float* wrow = o < num_classes ? w_class + (size_t)o * F : w_stop;
    float acc = 0.f;
    for (int i = lane; i < F; i += 32) acc = fmaf(feat[i], wrow[i], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) logit[o] = acc + (o < num_classes ? b_class[o] : b_stop[0]);
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int o = lane; o < num_classes; o += 32) mx = fmaxf(mx, logit[o]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float sum = 0.f;
    for (int o = lane; o < num_classes; o += 32) sum += expf(logit[o] - mx);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    for (int o = lane; o < num_classes; o += 32) class_probs[co + o] = expf(logit[o] - mx) / sum;
    if (lane == 0) {
      const float s = logit[num_classes];
      if (stop_logit) stop_logit[so] = s;
      if (stop_prob) stop_prob[so] = sigmoidf_acc(s);
    }
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

int rsis_maxpool3x3s2(const rsis_tensor* x, const rsis_tensor* y, rsis_stream_t stream) {
  if (!valid_tensor(x) || !valid_tensor(y)) return RSIS_ERR_BAD_ARG;
  const int Ho = (x->h + 2 - 3) / 2 + 1, Wo = (x->w + 2 - 3) / 2 + 1;
  if (y->n != x->n || y->c != x->c || y->h != Ho || y->w != Wo) return RSIS_ERR_BAD_ARG;
  if (x->c % 4 != 0 || !is_dense(*x) || !is_dense(*y)) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(x->data) || !aligned16(y->data)) return RSIS_ERR_ALIGN;
  const size_t total = numel(*y) / 4;
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(make_view(*x), y->data, numel(*y), y->fmt,
                                                                            x->n, x->h, x->w, x->c, Ho, Wo);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

static int fill_upsample(UpsampleArgs& a, const rsis_tensor* x, const rsis_tensor* y) {
  if (!valid_tensor(x) || !valid_tensor(y) || y->n != x->n || y->c != x->c) return RSIS_ERR_BAD_ARG;
  if (x->c % 4 != 0 || !is_dense(*x) || pitch(*y) % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(x->data) || !aligned16(y->data)) return RSIS_ERR_ALIGN;
  a.x = make_view(*x);
  a.y = y->data;
  a.y_plane = plane_elems(*y);
  a.y_fmt = y->fmt;
  a.y_cs = pitch(*y);
  a.N = x->n;
  a.H = x->h;
  a.W = x->w;
  a.C = x->c;
  a.Ho = y->h;
  a.Wo = y->w;
  a.sh = y->h > 1 ? (float)(x->h - 1) / (float)(y->h - 1) : 0.f;
  a.sw = y->w > 1 ? (float)(x->w - 1) / (float)(y->w - 1) : 0.f;
  return RSIS_OK;
}

int rsis_upsample_bilinear(const rsis_tensor* x, const rsis_tensor* y, rsis_stream_t stream) {
  UpsampleArgs a;
  if (int e = fill_upsample(a, x, y)) return e;
  const size_t total = numel(*y) / 4;
  upsample_bilinear_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_upsample_bilinear_group(const rsis_tensor* xs, const rsis_tensor* ys, int n, rsis_stream_t stream) {
  if (!xs || !ys || n < 1 || n > kMaxUpsampleGroup) return RSIS_ERR_BAD_ARG;
  UpsampleGroup g{};
  g.n = n;
  int first = 0;
  for (int i = 0; i < n; ++i) {
    if (int e = fill_upsample(g.a[i], xs + i, ys + i)) return e;
    g.first[i] = first;
    // blocks in proportion to the output sizes (every problem at least one block, at most four per SM)
    size_t blocks = (numel(ys[i]) / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;  // one 16-byte output vector per thread up to 16 blocks per SM: the kernel
                                               // is latency-bound (4 dependent-free loads per thread), so more threads
                                               // in flight, not longer grid-stride loops
    first += (int)(blocks < 1 ? 1 : blocks);
  }
  for (int i = n; i <= kMaxUpsampleGroup; ++i) g.first[i] = first;
  upsample_bilinear_group_kernel<<<first, 256, 0, (cudaStream_t)stream>>>(g);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_mask_head(const rsis_tensor* x, const float* w_oihw, const float* bias, int ksize, float* logits,
                   float* prob_out, int64_t prob_stride_n, rsis_stream_t stream) {
  if (!valid_tensor(x) || !w_oihw || (!logits && !prob_out)) return RSIS_ERR_BAD_ARG;
  if (x->fmt != RSIS_FMT_F32 || x->c % 4 != 0 || x->c > 256 || (ksize != 1 && ksize != 3) || !is_dense(*x))
    return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(x->data)) return RSIS_ERR_ALIGN;
  const size_t total = (size_t)x->n * x->h * x->w;
  const size_t smem = (size_t)ksize * ksize * x->c * sizeof(float);
  mask_head_kernel<<<grid_for(total, 256), 256, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float*>(x->data), w_oihw, bias, logits, prob_out, (long long)prob_stride_n, x->n, x->h,
      x->w, x->c, ksize);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

static int upsample_mask_head_impl(const rsis_tensor* h, const float* w_oihw, const float* bias, int ksize, int out_h,
                                   int out_w, float* logits, float* prob_out, int64_t prob_stride_n, int n_inner,
                                   int64_t prob_stride_t, rsis_stream_t stream) {
  if (!valid_tensor(h) || !w_oihw || (!logits && !prob_out) || out_h < 1 || out_w < 1) return RSIS_ERR_BAD_ARG;
  if (h->fmt != RSIS_FMT_F32 || !is_dense(*h) || h->c % 4 != 0 || h->c > 16 || (ksize != 1 && ksize != 3))
    return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(h->data)) return RSIS_ERR_ALIGN;
  const int pad = ksize / 2, TW = kMaskTile + 2 * pad;
  const bool fixed = h->c == 8;  // the unrolled kernels: they also stage the tile's source window (kMaskSrc^2 pixels)
  const size_t smem = (size_t)(TW * TW * h->c + ksize * ksize * h->c + (fixed ? kMaskSrc * kMaskSrc * h->c + 6 * TW : 0)) * sizeof(float);
  static bool attr_set = false;  // idempotent; a race only repeats the call
  if (!attr_set) {
    RSIS_CUDA_TRY(cudaFuncSetAttribute(upsample_mask_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (34 * 34 * 16 + 9 * 16) * (int)sizeof(float)));
    const int fixed_bytes = (34 * 34 * 8 + 9 * 8 + kMaskSrc * kMaskSrc * 8 + 6 * 34) * (int)sizeof(float);
    RSIS_CUDA_TRY(cudaFuncSetAttribute(upsample_mask_head_fixed_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fixed_bytes));
    RSIS_CUDA_TRY(cudaFuncSetAttribute(upsample_mask_head_fixed_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fixed_bytes));
    attr_set = true;
  }
  const float sh = out_h > 1 ? (float)(h->h - 1) / (float)(out_h - 1) : 0.f;
  const float sw = out_w > 1 ? (float)(h->w - 1) / (float)(out_w - 1) : 0.f;
  const dim3 grid(ceil_div(out_w, kMaskTile), ceil_div(out_h, kMaskTile), h->n);
  if (grid.y > 65535 || grid.z > 65535) return RSIS_ERR_UNSUPPORTED;
  const float* hp = reinterpret_cast<const float*>(h->data);
  if (h->c == 8 && ksize == 3)
    upsample_mask_head_fixed_kernel<8, 3><<<grid, 256, smem, (cudaStream_t)stream>>>(
        hp, w_oihw, bias, logits, prob_out, (long long)prob_stride_n, h->h, h->w, out_h, out_w, sh, sw, n_inner, (long long)prob_stride_t);
  else if (h->c == 8 && ksize == 1)
    upsample_mask_head_fixed_kernel<8, 1><<<grid, 256, smem, (cudaStream_t)stream>>>(
        hp, w_oihw, bias, logits, prob_out, (long long)prob_stride_n, h->h, h->w, out_h, out_w, sh, sw, n_inner, (long long)prob_stride_t);
  else
    upsample_mask_head_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(hp, w_oihw, bias, logits, prob_out,
                                                                        (long long)prob_stride_n, h->h, h->w, h->c,
                                                                        out_h, out_w, ksize, sh, sw, n_inner, (long long)prob_stride_t);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_upsample_mask_head(const rsis_tensor* h, const float* w_oihw, const float* bias, int ksize, int out_h,
                            int out_w, float* logits, float* prob_out, int64_t prob_stride_n, rsis_stream_t stream) {
  return upsample_mask_head_impl(h, w_oihw, bias, ksize, out_h, out_w, logits, prob_out, prob_stride_n,
                                 h ? (h->n > 0 ? h->n : 1) : 1, 0, stream);
}

int rsis_upsample_mask_head_steps(const rsis_tensor* h, int steps, const float* w_oihw, const float* bias, int ksize,
                                  int out_h, int out_w, float* prob_out, int64_t prob_stride_n, int64_t prob_stride_t,
                                  rsis_stream_t stream) {
  if (!valid_tensor(h) || steps < 1 || h->n % steps != 0 || !prob_out) return RSIS_ERR_BAD_ARG;
  return upsample_mask_head_impl(h, w_oihw, bias, ksize, out_h, out_w, nullptr, prob_out, prob_stride_n, h->n / steps,
                                 prob_stride_t, stream);
}

static int class_stop_heads_impl(const uint32_t* side_max, int n, int f, const float* w_class, const float* b_class,
                                 int num_classes, const float* w_stop, const float* b_stop, float* feat_out,
                                 float* class_probs, int64_t class_stride, float* stop_logit, float* stop_prob,
                                 int64_t stop_stride, int n_inner, int64_t class_stride_t, int64_t stop_stride_t,
                                 rsis_stream_t stream) {
  if (!side_max || !w_class || !b_class || !w_stop || !b_stop || !class_probs) return RSIS_ERR_BAD_ARG;
  if (n < 1 || f < 1 || num_classes < 1) return RSIS_ERR_BAD_ARG;
  if (f + num_classes + 1 > 10000) return RSIS_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(f + num_classes + 1) * sizeof(float);
  class_stop_heads_kernel<<<n, 256, smem, (cudaStream_t)stream>>>(side_max, f, w_class, b_class, num_classes, w_stop,
                                                                 b_stop, feat_out, class_probs,
                                                                 (long long)class_stride, stop_logit, stop_prob,
                                                                 (long long)stop_stride, n_inner, (long long)class_stride_t, (long long)stop_stride_t);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_class_stop_heads(const uint32_t* side_max, int n, int f, const float* w_class, const float* b_class,
                          int num_classes, const float* w_stop, const float* b_stop, float* feat_out,
                          float* class_probs, int64_t class_stride, float* stop_logit, float* stop_prob,
                          int64_t stop_stride, rsis_stream_t stream) {
  return class_stop_heads_impl(side_max, n, f, w_class, b_class, num_classes, w_stop, b_stop, feat_out, class_probs,
                               class_stride, stop_logit, stop_prob, stop_stride, n > 0 ? n : 1, 0, 0, stream);
}

int rsis_class_stop_heads_steps(const uint32_t* side_max, int n, int steps, int f, const float* w_class,
                                const float* b_class, int num_classes, const float* w_stop, const float* b_stop,
                                float* class_probs, int64_t class_stride, int64_t class_stride_t, float* stop_prob,
                                int64_t stop_stride, int64_t stop_stride_t, rsis_stream_t stream) {
  if (n < 1 || steps < 1) return RSIS_ERR_BAD_ARG;
  return class_stop_heads_impl(side_max, n * steps, f, w_class, b_class, num_classes, w_stop, b_stop, nullptr, class_probs,
                               class_stride, nullptr, stop_prob, stop_stride, n, class_stride_t, stop_stride_t, stream);
}

}  // extern "C"

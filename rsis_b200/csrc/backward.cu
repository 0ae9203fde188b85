// Backward primitives of the RSIS hot path -- what `loss.backward()` (/root/reference/src/train.py:184) executes through
// ResNet101 / FeatureExtractor / ConvLSTMCell / RSIS:
//   * convolution weight + bias gradient (implicit GEMM over pixels, fp32 CUDA cores);
//   * the weight transform that turns the DATA gradient of a convolution into a forward convolution (rsis_conv2d);
//   * zero insertion (the data gradient of a stride-2 convolution);
//   * train-mode BatchNorm backward fused with the ReLU mask and the residual branch;
//   * max-pool 3x3/s2 backward (first-maximum tie rule of ATen's CPU kernel);
//   * the ConvLSTM gate non-linearities + state update, forward (with the activated gates kept for backward) and
//     backward (clstm.py:47-58);
//   * global max-pool with arg-max (model.py:143) and its scatter backward;
//   * the adjoint of align-corners bilinear upsampling (model.py:149-150,163-164);
//   * fc_class + softmax + fc_stop backward (model.py:169-182).
// All tensors are NHWC; gradients are float32.
#include <math.h>

#include "common.cuh"

namespace rsis {

// tcgen05 weight gradient (conv_umma.cu)
bool conv_wgrad_umma_supported(const rsis_tensor* x, const rsis_tensor* dy, int kh, int kw, int stride, int pad,
                               size_t workspace_bytes);
int conv_wgrad_umma(const rsis_tensor* x, const rsis_tensor* dy, int ksize, int stride, float* dw_oihw, int accumulate,
                    void* workspace, cudaStream_t st);
size_t conv_wgrad_umma_workspace_bytes();

__device__ __forceinline__ void bw_ld4(const View& v, size_t idx, float out[4]) {
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
__device__ __forceinline__ void bw_st4(void* p, size_t plane, int fmt, size_t idx, const float v[4]) {
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


// ---------------------------------------------------------------------------------------------------------------
// conv weight gradient: dw[co][ci][kh][kw] += sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s-pad+kh, wo*s-pad+kw, ci]
// GEMM view: rows = co, columns k = (kh*KW+kw)*Cin + ci, reduction over the P = N*Ho*Wo output pixels.
// CTA tile 64 x 64 x 16 pixels, 256 threads with a 4x4 register tile each; the pixel range is split across
// gridDim.y CTAs whose partial tiles are combined with float atomics.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  View x;
  View dy;
  float* dw;
  int H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
  int K;            // KH*KW*Cin
  long long P;      // N*Ho*Wo
  long long p_per;  // pixels per gridDim.y slice (multiple of 16)
  int tiles_k;
};

constexpr int kWgT = 64;   // tile edge (co and k)
constexpr int kWgP = 16;   // pixels per chunk
constexpr int kWgLd = kWgT + 4;

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradParams p) {
  __shared__ __align__(16) float Ds[2][kWgP][kWgLd];  // dy tile   [pixel][co]
  __shared__ __align__(16) float Xs[2][kWgP][kWgLd];  // x gather  [pixel][k]
  const int tid = threadIdx.x;
  const int tile_k = blockIdx.x % p.tiles_k, tile_co = blockIdx.x / p.tiles_k;
  const int k0 = tile_k * kWgT, co0 = tile_co * kWgT;
  const long long p_begin = (long long)blockIdx.y * p.p_per;
  long long p_end = p_begin + p.p_per;
  if (p_end > p.P) p_end = p.P;
  if (p_begin >= p_end) return;

  // loader mapping: this thread fills pixel row lp, columns lc + 16*j (j = 0..3) of both tiles
  const int lp = tid >> 4, lc = tid & 15;
  int kk_kh[4], kk_kw[4], kk_ci[4];
  bool kk_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + lc + 16 * j;
    kk_ok[j] = k < p.K;
    const int tap = kk_ok[j] ? k / p.Cin : 0;
    kk_ci[j] = kk_ok[j] ? k - tap * p.Cin : 0;
    kk_kh[j] = tap / p.KW;
    kk_kw[j] = tap - kk_kh[j] * p.KW;
  }
  const int HoWo = p.Ho * p.Wo;
  float d_reg[4], x_reg[4];
  auto gather = [&](long long pc) {
    const long long pix = pc + lp;
    if (pix < p_end) {
      const int n = (int)(pix / HoWo);
      const int r = (int)(pix - (long long)n * HoWo);
      const int ho = r / p.Wo, wo = r - ho * p.Wo;
      const size_t dyr = (size_t)pix * p.Cout;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = co0 + lc + 16 * j;
        d_reg[j] = co < p.Cout ? load_elem(p.dy, dyr + co) : 0.f;
        const int hi = ho * p.stride - p.pad + kk_kh[j], wi = wo * p.stride - p.pad + kk_kw[j];
        x_reg[j] = (kk_ok[j] && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
                       ? load_elem(p.x, (((size_t)n * p.H + hi) * p.W + wi) * p.Cin + kk_ci[j])
                       : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) d_reg[j] = x_reg[j] = 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Ds[buf][lp][lc + 16 * j] = d_reg[j];
      Xs[buf][lp][lc + 16 * j] = x_reg[j];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;  // register tile: co rows ty*4.., k columns tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gather(p_begin);
  stash(0);
  __syncthreads();
  int cur = 0;
  for (long long pc = p_begin; pc < p_end; pc += kWgP) {
    const bool more = pc + kWgP < p_end;
    if (more) gather(pc + kWgP);
#pragma unroll
    for (int pp = 0; pp < kWgP; ++pp) {
      const float4 a = *reinterpret_cast<const float4*>(&Ds[cur][pp][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[cur][pp][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }

  const int taps = p.KH * p.KW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + tx * 4 + j;
    if (k >= p.K) continue;
    const int tap = k / p.Cin, ci = k - tap * p.Cin;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + ty * 4 + i;
      if (co >= p.Cout) continue;
      atomicAdd(p.dw + ((size_t)co * p.Cin + ci) * taps + tap, acc[i][j]);
    }
  }
}

// Weight (+ bias) gradient of a convolution with ONE output channel and few input channels (conv_out, model.py:107:
// hidden/16 -> 1): every thread walks pixels with its CI*KS*KS partial sums in registers; warp shuffles + shared
// memory combine a block, one atomicAdd per weight and block.  dw index = c * KS*KS + tap (OIHW with O = 1).
template <int CI, int KS>
__global__ void __launch_bounds__(256) wgrad_cout1_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, float* __restrict__ dbias, int N,
                                                          int H, int W) {
  constexpr int TAPS = KS * KS, NW = TAPS * CI, PAD = KS / 2;
  float acc[NW];
  float accb = 0.f;
#pragma unroll
  for (int i = 0; i < NW; ++i) acc[i] = 0.f;
  const size_t total = (size_t)N * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int wq = (int)(i % W);
    const int hq = (int)((i / W) % H);
    const size_t n = i / ((size_t)H * W);
    const float g = dy[i];
    accb += g;
#pragma unroll
    for (int kh = 0; kh < KS; ++kh) {
      const int hi = hq - PAD + kh;
      if (hi < 0 || hi >= H) continue;
#pragma unroll
      for (int kw = 0; kw < KS; ++kw) {
        const int wi = wq - PAD + kw;
        if (wi < 0 || wi >= W) continue;
        const float* px = x + ((n * H + hi) * W + wi) * CI;
#pragma unroll
        for (int c = 0; c < CI; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(px + c);
          acc[(kh * KS + kw) * CI + c] = fmaf(g, v.x, acc[(kh * KS + kw) * CI + c]);
          acc[(kh * KS + kw) * CI + c + 1] = fmaf(g, v.y, acc[(kh * KS + kw) * CI + c + 1]);
          acc[(kh * KS + kw) * CI + c + 2] = fmaf(g, v.z, acc[(kh * KS + kw) * CI + c + 2]);
          acc[(kh * KS + kw) * CI + c + 3] = fmaf(g, v.w, acc[(kh * KS + kw) * CI + c + 3]);
        }
      }
    }
  }
  __shared__ float red[8][NW + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    float v = acc[i];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
    if (lane == 0) red[warp][i] = v;
  }
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) accb += __shfl_xor_sync(0xffffffffu, accb, sft);
  if (lane == 0) red[warp][NW] = accb;
  __syncthreads();
  for (int i = threadIdx.x; i <= NW; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][i];
    if (i < NW) {
      if (dw) atomicAdd(dw + (i % CI) * TAPS + i / CI, v);
    } else if (dbias) {
      atomicAdd(dbias, v);
    }
  }
}

// dbias[c] += sum over pixels of dy[p][c]  (float atomics across gridDim.x slices)
__global__ void channel_sum_kernel(View dy, long long P, int C, float* __restrict__ out) {
  __shared__ float red[256];
  const int lanes_c = C < 256 ? C : 256;
  const int rows = 256 / lanes_c;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < P ? p0 + per : P;
  for (int c = threadIdx.x % lanes_c; c < C; c += lanes_c) {
    float s = 0.f;
    if (threadIdx.x < rows * lanes_c) {
      long long q = p0 + threadIdx.x / lanes_c;
      for (; q + 3LL * rows < p1; q += 4LL * rows) {  // four independent loads in flight
        const float v0 = load_elem(dy, (size_t)q * C + c), v1 = load_elem(dy, (size_t)(q + rows) * C + c);
        const float v2 = load_elem(dy, (size_t)(q + 2LL * rows) * C + c);
        const float v3 = load_elem(dy, (size_t)(q + 3LL * rows) * C + c);
        s += (v0 + v1) + (v2 + v3);
      }
      for (; q < p1; q += rows) s += load_elem(dy, (size_t)q * C + c);
    }
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < lanes_c) {
      float t = 0.f;
      for (int r = 0; r < rows; ++r) t += red[r * lanes_c + threadIdx.x];
      atomicAdd(out + c, t);
    }
    __syncthreads();
  }
}

// out[ci - ci0][co][KH-1-kh][KW-1-kw] = w[co][ci][kh][kw] for ci in [ci0, ci0 + nci): the OIHW weights of the
// convolution that computes the data gradient (a correlation with the 180-degree rotated, in/out swapped kernel).
__global__ void dgrad_weights_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw, int ci0, int nci,
                                     float* __restrict__ out) {
  const int taps = kh * kw;
  const size_t total = (size_t)nci * cout * taps;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const size_t r = i / taps;
    const int co = (int)(r % cout);
    const int ci = (int)(r / cout);
    out[i] = w[((size_t)co * cin + ci0 + ci) * taps + (taps - 1 - tap)];
  }
}

// y[n, 2i, 2j, :] = x[n, i, j, :], every other element of y is zero.
__global__ void dilate2x_kernel(View x, void* y, size_t y_plane, int y_fmt, int N, int H, int W, int C, int Ho,
                                int Wo) {
  const int C4 = C >> 2;
  const size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t r = i / C4;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (!(ho & 1) && !(wo & 1) && (ho >> 1) < H && (wo >> 1) < W)
      bw_ld4(x, (((size_t)n * H + (ho >> 1)) * W + (wo >> 1)) * C + c, o);
    bw_st4(y, y_plane, y_fmt, i * 4, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// train-mode BatchNorm backward (+ ReLU mask, + residual branch)
//   g = dy * (y > 0)            (y = the block's post-activation output; absent: no ReLU)
//   dbias = sum g, dweight = sum g * xhat, xhat = (x - mean) * invstd
//   dx = weight * invstd * (g - dbias / M - xhat * dweight / M);   dres = g
// ---------------------------------------------------------------------------------------------------------------
// acc: [2][C] doubles (sum g, sum g*xhat), zero on entry
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ x, View y, int has_y, const float* __restrict__ dy,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, size_t M, int C,
                                     double* __restrict__ acc) {
  extern __shared__ float red[];  // [256][8]
  const int C4 = C >> 2;
  const int lanes_c = C4 < 256 ? C4 : 256;
  const int rows_per_iter = 256 / lanes_c;
  const size_t per_block = (M + gridDim.x - 1) / gridDim.x;
  const size_t p0 = (size_t)blockIdx.x * per_block;
  const size_t p1 = p0 + per_block < M ? p0 + per_block : M;
  // uniform trip count (the body holds block barriers); threads past C4 in the last round carry no pixels
  for (int c4base = 0; c4base < C4; c4base += lanes_c) {
    const int c4 = c4base + threadIdx.x % lanes_c;
    const bool live = c4 < C4;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 mu = live ? *reinterpret_cast<const float4*>(mean + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 is = live ? *reinterpret_cast<const float4*>(invstd + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, isv[4] = {is.x, is.y, is.z, is.w};
    if (live && threadIdx.x < rows_per_iter * lanes_c) {
#pragma unroll 4
      for (size_t p = p0 + threadIdx.x / lanes_c; p < p1; p += rows_per_iter) {
        const size_t idx = p * C + 4 * c4;
        const float4 xv = *reinterpret_cast<const float4*>(x + idx);
        const float4 gv = *reinterpret_cast<const float4*>(dy + idx);
        float g[4] = {gv.x, gv.y, gv.z, gv.w};
        const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
        if (has_y) {
          float yv[4];
          bw_ld4(y, idx, yv);
#pragma unroll
          for (int j = 0; j < 4; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[j] += g[j];
          q[j] = fmaf(g[j], (xr[j] - muv[j]) * isv[j], q[j]);
        }
      }
    }
    float* mine = red + threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mine[j] = s[j];
      mine[4 + j] = q[j];
    }
    __syncthreads();
    if (threadIdx.x < lanes_c && live) {
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

__global__ void bn_bwd_finalize_kernel(double* __restrict__ acc, int C, float* __restrict__ dweight,
                                       float* __restrict__ dbias, float* __restrict__ dweight_acc,
                                       float* __restrict__ dbias_acc) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    dbias[c] = (float)acc[c];
    dweight[c] = (float)acc[C + c];
    if (dbias_acc) dbias_acc[c] += (float)acc[c];
    if (dweight_acc) dweight_acc[c] += (float)acc[C + c];
    acc[c] = 0;
    acc[C + c] = 0;
  }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, View y, int has_y, const float* __restrict__ dy,
                                    const float* __restrict__ weight, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ dweight,
                                    const float* __restrict__ dbias, float inv_m, void* dx, size_t dx_plane, int dx_fmt,
                                    float* __restrict__ dres, int C, size_t total4) {
  const int C4 = C >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t idx = i * 4;
    const float4 xv = *reinterpret_cast<const float4*>(x + idx);
    const float4 gv = *reinterpret_cast<const float4*>(dy + idx);
    float g[4] = {gv.x, gv.y, gv.z, gv.w};
    const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
    if (has_y) {
      float yv[4];
      bw_ld4(y, idx, yv);
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float is = invstd[c + j];
      const float xh = (xr[j] - mean[c + j]) * is;
      const float w = weight ? weight[c + j] : 1.f;
      o[j] = w * is * (g[j] - dbias[c + j] * inv_m - xh * dweight[c + j] * inv_m);
    }
    bw_st4(dx, dx_plane, dx_fmt, idx, o);
    if (dres) *reinterpret_cast<float4*>(dres + idx) = make_float4(g[0], g[1], g[2], g[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// nn.MaxPool2d(3, 2, 1) backward.  Gather form (deterministic): an input pixel receives dy of every window whose
// arg-max it is; the arg-max of a window is its FIRST maximum in (kh, kw) scan order (ATen cpu max_pool2d: a later
// element wins only if it is strictly greater).
// ---------------------------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_bwd_kernel(View x, const float* __restrict__ dy, float* __restrict__ dx, int N, int H,
                                        int W, int C, int Ho, int Wo) {
  const int C4 = C >> 2;
  const size_t total = (size_t)N * H * W * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t r = i / C4;
    const int wi = (int)(r % W);
    r /= W;
    const int hi = (int)(r % H);
    const int n = (int)(r / H);
    float mine[4];
    bw_ld4(x, (((size_t)n * H + hi) * W + wi) * C + c, mine);
    float out[4] = {0.f, 0.f, 0.f, 0.f};
    // windows (ho, wo) containing (hi, wi): 2*ho - 1 <= hi <= 2*ho + 1
    const int ho_lo = hi >> 1, ho_hi = (hi + 1) >> 1;  // ho with 2ho-1<=hi<=2ho+1  <=>  (hi-1)/2 <= ho <= (hi+1)/2
    const int wo_lo = wi >> 1, wo_hi = (wi + 1) >> 1;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      if (ho >= Ho) continue;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        if (wo >= Wo) continue;
        // is (hi, wi) the first maximum of this window?
        bool win[4] = {true, true, true, true};
        for (int dh = 0; dh < 3; ++dh) {
          const int h2 = ho * 2 - 1 + dh;
          if (h2 < 0 || h2 >= H) continue;
          for (int dw = 0; dw < 3; ++dw) {
            const int w2 = wo * 2 - 1 + dw;
            if (w2 < 0 || w2 >= W) continue;
            if (h2 == hi && w2 == wi) continue;
            float v[4];
            bw_ld4(x, (((size_t)n * H + h2) * W + w2) * C + c, v);
            const bool before = (h2 < hi) || (h2 == hi && w2 < wi);
#pragma unroll
            for (int j = 0; j < 4; ++j) win[j] = win[j] && (before ? (mine[j] > v[j]) : (mine[j] >= v[j]));
          }
        }
        const float4 g = *reinterpret_cast<const float4*>(dy + (((size_t)n * Ho + ho) * Wo + wo) * C + c);
        if (win[0]) out[0] += g.x;
        if (win[1]) out[1] += g.y;
        if (win[2]) out[2] += g.z;
        if (win[3]) out[3] += g.w;
      }
    }
    *reinterpret_cast<float4*>(dx + (((size_t)n * H + hi) * W + wi) * C + c) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ConvLSTM gate non-linearities and state update (clstm.py:47-58).  `gates` holds the pre-activations
// [N,H,W,4*Ch] in the reference's block order [in | remember | out | cell] and is overwritten with the activated
// gates (kept for backward).
// ---------------------------------------------------------------------------------------------------------------
__global__ void lstm_gates_fwd_kernel(float* __restrict__ gates, const float* __restrict__ c_prev,
                                      float* __restrict__ h_out, float* __restrict__ c_out, void* h2, size_t h2_plane,
                                      int h2_fmt, int h2_cs, int Ch, size_t total4) {
  const int Ch4 = Ch >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ch4) * 4;
    const size_t pix = i / Ch4;
    float* gp = gates + pix * 4 * Ch + c;
    const float4 a_i = *reinterpret_cast<const float4*>(gp);
    const float4 a_f = *reinterpret_cast<const float4*>(gp + Ch);
    const float4 a_o = *reinterpret_cast<const float4*>(gp + 2 * Ch);
    const float4 a_g = *reinterpret_cast<const float4*>(gp + 3 * Ch);
    const float pi[4] = {a_i.x, a_i.y, a_i.z, a_i.w}, pf[4] = {a_f.x, a_f.y, a_f.z, a_f.w};
    const float po[4] = {a_o.x, a_o.y, a_o.z, a_o.w}, pg[4] = {a_g.x, a_g.y, a_g.z, a_g.w};
    float cp[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t idx = pix * Ch + c;
    if (c_prev) {
      const float4 t = *reinterpret_cast<const float4*>(c_prev + idx);
      cp[0] = t.x; cp[1] = t.y; cp[2] = t.z; cp[3] = t.w;
    }
    float gi[4], gf[4], go[4], gg[4], cn[4], hn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gi[j] = sigmoidf_acc(pi[j]);
      gf[j] = sigmoidf_acc(pf[j]);
      go[j] = sigmoidf_acc(po[j]);
      gg[j] = tanhf(pg[j]);
      cn[j] = gf[j] * cp[j] + gi[j] * gg[j];
      hn[j] = go[j] * tanhf(cn[j]);
    }
    *reinterpret_cast<float4*>(gp) = make_float4(gi[0], gi[1], gi[2], gi[3]);
    *reinterpret_cast<float4*>(gp + Ch) = make_float4(gf[0], gf[1], gf[2], gf[3]);
    *reinterpret_cast<float4*>(gp + 2 * Ch) = make_float4(go[0], go[1], go[2], go[3]);
    *reinterpret_cast<float4*>(gp + 3 * Ch) = make_float4(gg[0], gg[1], gg[2], gg[3]);
    *reinterpret_cast<float4*>(c_out + idx) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    *reinterpret_cast<float4*>(h_out + idx) = make_float4(hn[0], hn[1], hn[2], hn[3]);
    if (h2) bw_st4(h2, h2_plane, h2_fmt, pix * h2_cs + c, hn);
  }
}

// Appendix B of SURVEY.md.  dh = dh_a (+ dh_b); dc_total = dc_next + dh * o * (1 - tanh(c)^2).
__global__ void lstm_gates_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                      const float* __restrict__ c_new, const float* __restrict__ dh_a, int dha_cs,
                                      const float* __restrict__ dh_b, int dhb_cs, const float* __restrict__ dc_next,
                                      int dcn_cs, void* dgates, size_t dg_plane, int dg_fmt, float* __restrict__ dc_prev,
                                      int Ch, size_t total4) {
  const int Ch4 = Ch >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ch4) * 4;
    const size_t pix = i / Ch4;
    const float* gp = gates + pix * 4 * Ch + c;
    const float4 a_i = *reinterpret_cast<const float4*>(gp);
    const float4 a_f = *reinterpret_cast<const float4*>(gp + Ch);
    const float4 a_o = *reinterpret_cast<const float4*>(gp + 2 * Ch);
    const float4 a_g = *reinterpret_cast<const float4*>(gp + 3 * Ch);
    const float gi[4] = {a_i.x, a_i.y, a_i.z, a_i.w}, gf[4] = {a_f.x, a_f.y, a_f.z, a_f.w};
    const float go[4] = {a_o.x, a_o.y, a_o.z, a_o.w}, gg[4] = {a_g.x, a_g.y, a_g.z, a_g.w};
    const size_t idx = pix * Ch + c;
    float cp[4] = {0.f, 0.f, 0.f, 0.f}, dh[4] = {0.f, 0.f, 0.f, 0.f}, dcn[4] = {0.f, 0.f, 0.f, 0.f};
    if (c_prev) {
      const float4 t = *reinterpret_cast<const float4*>(c_prev + idx);
      cp[0] = t.x; cp[1] = t.y; cp[2] = t.z; cp[3] = t.w;
    }
    if (dh_a) {
      const float4 t = *reinterpret_cast<const float4*>(dh_a + pix * dha_cs + c);
      dh[0] = t.x; dh[1] = t.y; dh[2] = t.z; dh[3] = t.w;
    }
    if (dh_b) {
      const float4 t = *reinterpret_cast<const float4*>(dh_b + pix * dhb_cs + c);
      dh[0] += t.x; dh[1] += t.y; dh[2] += t.z; dh[3] += t.w;
    }
    if (dc_next) {
      const float4 t = *reinterpret_cast<const float4*>(dc_next + pix * dcn_cs + c);
      dcn[0] = t.x; dcn[1] = t.y; dcn[2] = t.z; dcn[3] = t.w;
    }
    const float4 cv = *reinterpret_cast<const float4*>(c_new + idx);
    const float cn[4] = {cv.x, cv.y, cv.z, cv.w};
    float di[4], df[4], dO[4], dg[4], dcp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tc = tanhf(cn[j]);
      const float d_o = dh[j] * tc;
      const float dc = dcn[j] + dh[j] * go[j] * (1.f - tc * tc);
      di[j] = dc * gg[j] * gi[j] * (1.f - gi[j]);
      df[j] = dc * cp[j] * gf[j] * (1.f - gf[j]);
      dO[j] = d_o * go[j] * (1.f - go[j]);
      dg[j] = dc * gi[j] * (1.f - gg[j] * gg[j]);
      dcp[j] = dc * gf[j];
    }
    const size_t dp = pix * 4 * Ch + c;
    bw_st4(dgates, dg_plane, dg_fmt, dp, di);
    bw_st4(dgates, dg_plane, dg_fmt, dp + Ch, df);
    bw_st4(dgates, dg_plane, dg_fmt, dp + 2 * Ch, dO);
    bw_st4(dgates, dg_plane, dg_fmt, dp + 3 * Ch, dg);
    *reinterpret_cast<float4*>(dc_prev + idx) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// global max-pool with arg-max (model.py:143).  CTA = (image, group of <= 32 channels, slice of the pixels); the
// slices are combined with a 64-bit atomicMax on (order-preserving key << 32 | ~pixel index), so the FIRST maximum
// wins ties; a finishing kernel unpacks keys (what rsis_class_stop_heads reads) and pixel indices.
// ---------------------------------------------------------------------------------------------------------------
__global__ void global_maxpool_kernel(const float* __restrict__ h, int HW, int C, unsigned long long* __restrict__ packed,
                                      int side_stride, int side_offset) {
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  const int CG = C < 32 ? C : 32;
  const int R = 256 / CG;
  const int n = blockIdx.x;
  const int c = blockIdx.y * CG + threadIdx.x % CG;
  const int row = threadIdx.x / CG;
  const int per = (HW + gridDim.z - 1) / gridDim.z;
  const int p0 = blockIdx.z * per;
  const int p1 = p0 + per < HW ? p0 + per : HW;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  if (row < R && c < C) {
    const float* base = h + (size_t)n * HW * C + c;
#pragma unroll 4
    for (int p = p0 + row; p < p1; p += R) {
      const float v = base[(size_t)p * C];
      if (v > best || best_i == 0x7fffffff) {
        best = v;
        best_i = p;
      }
    }
  }
  s_val[threadIdx.x] = best;
  s_idx[threadIdx.x] = best_i;
  __syncthreads();
  if (threadIdx.x < CG && c < C) {
    float b = s_val[threadIdx.x];
    int bi = s_idx[threadIdx.x];
    for (int r = 1; r < R; ++r) {
      const float v = s_val[r * CG + threadIdx.x];
      const int vi = s_idx[r * CG + threadIdx.x];
      if (vi != 0x7fffffff && (bi == 0x7fffffff || v > b || (v == b && vi < bi))) {
        b = v;
        bi = vi;
      }
    }
    if (bi != 0x7fffffff)
      atomicMax(packed + (size_t)n * side_stride + side_offset + c,
                ((unsigned long long)float_to_key(b) << 32) | (unsigned long long)(0xffffffffu - (unsigned)bi));
  }
}

__global__ void global_maxpool_finish_kernel(const unsigned long long* __restrict__ packed, int total,
                                             uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned long long v = packed[i];
  keys[i] = (uint32_t)(v >> 32);
  idx[i] = (int32_t)(0xffffffffu - (uint32_t)(v & 0xffffffffu));
}

// dh[n, idx[n][c], c] += dside[n][c]
__global__ void global_maxpool_bwd_kernel(const float* __restrict__ dside, const int32_t* __restrict__ idx,
                                          int side_stride, int side_offset, float* __restrict__ dh, int N, int HW,
                                          int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const size_t s = (size_t)n * side_stride + side_offset + c;
  dh[((size_t)n * HW + idx[s]) * C + c] += dside[s];
}

// ---------------------------------------------------------------------------------------------------------------
// adjoint of upsample_bilinear_kernel (decoder_ops.cu): dx[n,hi,wi,:] = sum over the output pixels whose 2x2
// footprint contains (hi, wi) of weight * dy.  Gather form; each candidate re-evaluates the forward's own index /
// weight arithmetic so the pair is an exact adjoint.
// ---------------------------------------------------------------------------------------------------------------
__global__ void upsample_bilinear_bwd_kernel(const float* __restrict__ dy, int dy_cs, float* __restrict__ dx, int N,
                                             int H, int W, int C, int Ho, int Wo, float sh, float sw) {
  const int C4 = C >> 2;
  const size_t total = (size_t)N * H * W * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t r = i / C4;
    const int wi = (int)(r % W);
    r /= W;
    const int hi = (int)(r % H);
    const int n = (int)(r / H);
    int ho_lo = 0, ho_hi = Ho - 1, wo_lo = 0, wo_hi = Wo - 1;
    if (sh > 0.f) {
      ho_lo = max(0, (int)floorf((hi - 1) / sh) - 1);
      ho_hi = min(Ho - 1, (int)ceilf((hi + 1) / sh) + 1);
    }
    if (sw > 0.f) {
      wo_lo = max(0, (int)floorf((wi - 1) / sw) - 1);
      wo_hi = min(Wo - 1, (int)ceilf((wi + 1) / sw) + 1);
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const float fh = sh * ho;
      const int h1 = min((int)fh, H - 1);
      const int h1p = h1 < H - 1 ? 1 : 0;
      const float h1l = fminf(fmaxf(fh - h1, 0.f), 1.f), h0l = 1.f - h1l;
      float wh = 0.f;
      if (h1 == hi) wh += h0l;
      if (h1 + h1p == hi) wh += h1l;
      if (h1 != hi && h1 + h1p != hi) continue;
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const float fw = sw * wo;
        const int w1 = min((int)fw, W - 1);
        const int w1p = w1 < W - 1 ? 1 : 0;
        if (w1 != wi && w1 + w1p != wi) continue;
        const float w1l = fminf(fmaxf(fw - w1, 0.f), 1.f), w0l = 1.f - w1l;
        float ww = 0.f;
        if (w1 == wi) ww += w0l;
        if (w1 + w1p == wi) ww += w1l;
        const float4 g = *reinterpret_cast<const float4*>(dy + (((size_t)n * Ho + ho) * Wo + wo) * dy_cs + c);
        const float k = wh * ww;
        acc[0] = fmaf(k, g.x, acc[0]);
        acc[1] = fmaf(k, g.y, acc[1]);
        acc[2] = fmaf(k, g.z, acc[2]);
        acc[3] = fmaf(k, g.w, acc[3]);
      }
    }
    *reinterpret_cast<float4*>(dx + (((size_t)n * H + hi) * W + wi) * C + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fc_class + softmax + fc_stop backward (model.py:169-182).  Kernel 1 (one CTA per image): dlogit and dfeat;
// kernel 2: parameter gradients (sum over the batch), accumulated into the given buffers.
// ---------------------------------------------------------------------------------------------------------------
__global__ void heads_bwd_data_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs,
                                      const float* __restrict__ dstop, int F, int NC, const float* __restrict__ w_class,
                                      const float* __restrict__ w_stop, float* __restrict__ dlogit,
                                      float* __restrict__ dfeat) {
  extern __shared__ float sm[];  // [NC + 1]
  const int n = blockIdx.x;
  if (threadIdx.x == 0) {
    float dot = 0.f;
    for (int k = 0; k < NC; ++k) dot = fmaf(probs[(size_t)n * NC + k], dprobs ? dprobs[(size_t)n * NC + k] : 0.f, dot);
    for (int k = 0; k < NC; ++k)
      sm[k] = probs[(size_t)n * NC + k] * ((dprobs ? dprobs[(size_t)n * NC + k] : 0.f) - dot);
    sm[NC] = dstop ? dstop[n] : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= NC; k += blockDim.x) dlogit[(size_t)n * (NC + 1) + k] = sm[k];
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < NC; ++k) acc = fmaf(sm[k], w_class[(size_t)k * F + f], acc);
    acc = fmaf(sm[NC], w_stop[f], acc);
    dfeat[(size_t)n * F + f] = acc;
  }
}

__global__ void heads_bwd_param_kernel(const float* __restrict__ dlogit, const float* __restrict__ feat, int N, int F,
                                       int NC, float* __restrict__ dw_class, float* __restrict__ db_class,
                                       float* __restrict__ dw_stop, float* __restrict__ db_stop) {
  const int k = blockIdx.x;  // 0..NC (NC = the stop row)
  float* dw = k < NC ? dw_class + (size_t)k * F : dw_stop;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc = fmaf(dlogit[(size_t)n * (NC + 1) + k], feat[(size_t)n * F + f], acc);
    dw[f] += acc;
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int n = 0; n < N; ++n) acc += dlogit[(size_t)n * (NC + 1) + k];
    if (k < NC) db_class[k] += acc; else db_stop[0] += acc;
  }
}

static inline int grid_for(size_t total, int block) {
  size_t b = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace rsis

using namespace rsis;

static inline bool f32_dense(const rsis_tensor* t) {
  return valid_tensor(t) && t->fmt == RSIS_FMT_F32 && is_dense(*t) && aligned16(t->data);
}
static inline bool same_shape(const rsis_tensor* a, const rsis_tensor* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}

extern "C" {

int rsis_conv_dgrad_weights(const float* w_oihw, int cout, int cin, int kh, int kw, int ci0, int nci, float* out_oihw,
                            rsis_stream_t stream) {
  if (!w_oihw || !out_oihw || cout < 1 || cin < 1 || kh < 1 || kw < 1 || ci0 < 0 || nci < 1 || ci0 + nci > cin)
    return RSIS_ERR_BAD_ARG;
  const size_t total = (size_t)nci * cout * kh * kw;
  dgrad_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, cout, cin, kh, kw, ci0, nci,
                                                                               out_oihw);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

size_t rsis_wgrad_workspace_bytes(void) { return conv_wgrad_umma_workspace_bytes(); }

int rsis_conv2d_wgrad(const rsis_tensor* x, const rsis_tensor* dy, int kh, int kw, int stride, int pad, float* dw_oihw,
                      float* dbias, int accumulate, int impl, void* workspace, size_t workspace_bytes,
                      rsis_stream_t stream) {
  if (!valid_tensor(x) || !valid_tensor(dy) || (!dw_oihw && !dbias) || kh < 1 || kw < 1 || stride < 1 || pad < 0)
    return RSIS_ERR_BAD_ARG;
  if (!is_dense(*x) || !is_dense(*dy)) return RSIS_ERR_UNSUPPORTED;
  const int Ho = (x->h + 2 * pad - kh) / stride + 1, Wo = (x->w + 2 * pad - kw) / stride + 1;
  if (dy->n != x->n || dy->h != Ho || dy->w != Wo) return RSIS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = (long long)x->n * Ho * Wo;
  if (impl != RSIS_IMPL_AUTO && impl != RSIS_IMPL_SIMT && impl != RSIS_IMPL_TCGEN05) return RSIS_ERR_BAD_ARG;
  const bool umma_ok = dw_oihw && impl != RSIS_IMPL_SIMT && workspace && aligned16(workspace) &&
                       conv_wgrad_umma_supported(x, dy, kh, kw, stride, pad, workspace_bytes);
  if (impl == RSIS_IMPL_TCGEN05 && dw_oihw && !umma_ok) return RSIS_ERR_UNSUPPORTED;
  if (dy->c == 1 && x->c == 8 && kh == kw && (kh == 1 || kh == 3) && stride == 1 && pad == kh / 2 &&
      x->fmt == RSIS_FMT_F32 && dy->fmt == RSIS_FMT_F32 && aligned16(x->data)) {
    // conv_out: both gradients in one pass over the pixels
    if (!accumulate) {
      if (dw_oihw) RSIS_CUDA_TRY(cudaMemsetAsync(dw_oihw, 0, (size_t)x->c * kh * kw * sizeof(float), st));
      if (dbias) RSIS_CUDA_TRY(cudaMemsetAsync(dbias, 0, sizeof(float), st));
    }
    long long blocks = (P + 255) / 256;
    if (blocks > 592) blocks = 592;
    const float* xp = reinterpret_cast<const float*>(x->data);
    const float* gp = reinterpret_cast<const float*>(dy->data);
    if (kh == 3)
      wgrad_cout1_kernel<8, 3><<<(unsigned)blocks, 256, 0, st>>>(xp, gp, dw_oihw, dbias, x->n, x->h, x->w);
    else
      wgrad_cout1_kernel<8, 1><<<(unsigned)blocks, 256, 0, st>>>(xp, gp, dw_oihw, dbias, x->n, x->h, x->w);
    RSIS_CHECK_LAUNCH();
    return RSIS_OK;
  }
  if (umma_ok) {
    if (int e = conv_wgrad_umma(x, dy, kh, stride, dw_oihw, accumulate, workspace, st)) return e;
  } else if (dw_oihw) {
    WgradParams p{};
    p.x = make_view(*x);
    p.dy = make_view(*dy);
    p.dw = dw_oihw;
    p.H = x->h; p.W = x->w; p.Cin = x->c; p.Ho = Ho; p.Wo = Wo; p.Cout = dy->c;
    p.KH = kh; p.KW = kw; p.stride = stride; p.pad = pad;
    p.K = kh * kw * x->c;
    p.P = P;
    p.tiles_k = ceil_div(p.K, kWgT);
    const long long tiles = (long long)p.tiles_k * ceil_div(p.Cout, kWgT);
    if (tiles > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
    long long split = (4 * 148 + tiles - 1) / tiles;
    const long long max_split = (P + 63) / 64;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    if (split > 65535) split = 65535;
    p.p_per = ((P + split - 1) / split + kWgP - 1) / kWgP * kWgP;
    split = (P + p.p_per - 1) / p.p_per;
    if (!accumulate)
      RSIS_CUDA_TRY(cudaMemsetAsync(dw_oihw, 0, (size_t)dy->c * x->c * kh * kw * sizeof(float), st));
    conv_wgrad_kernel<<<dim3((unsigned)tiles, (unsigned)split), 256, 0, st>>>(p);
    RSIS_CHECK_LAUNCH();
  }
  if (dbias) {
    if (!accumulate) RSIS_CUDA_TRY(cudaMemsetAsync(dbias, 0, (size_t)dy->c * sizeof(float), st));
    long long blocks = (P + 63) / 64;
    if (blocks > 1184) blocks = 1184;
    channel_sum_kernel<<<(unsigned)blocks, 256, 0, st>>>(make_view(*dy), P, dy->c, dbias);
    RSIS_CHECK_LAUNCH();
  }
  return RSIS_OK;
}

int rsis_dilate2x(const rsis_tensor* x, const rsis_tensor* y, rsis_stream_t stream) {
  if (!valid_tensor(x) || !is_dense(*x) || !aligned16(x->data) || !valid_tensor(y) || !is_dense(*y) ||
      !aligned16(y->data) || x->n != y->n || x->c != y->c)
    return RSIS_ERR_BAD_ARG;
  if (x->c % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  if ((y->h + 1) / 2 != x->h || (y->w + 1) / 2 != x->w) return RSIS_ERR_BAD_ARG;
  const size_t total = numel(*y) / 4;
  dilate2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      make_view(*x), y->data, plane_elems(*y), y->fmt, x->n, x->h, x->w, x->c, y->h, y->w);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_bn_train_bwd(const rsis_tensor* x_raw, const rsis_tensor* y_act, const rsis_tensor* dy, const float* weight,
                      const float* mean, const float* invstd, double* workspace, float* dweight, float* dbias,
                      float* dweight_acc, float* dbias_acc, const rsis_tensor* dx, const rsis_tensor* dres,
                      rsis_stream_t stream) {
  if (!f32_dense(x_raw) || !f32_dense(dy) || !valid_tensor(dx) || !mean || !invstd || !workspace || !dweight || !dbias)
    return RSIS_ERR_BAD_ARG;
  if (!same_shape(x_raw, dy) || !same_shape(x_raw, dx) || !is_dense(*dx) || !aligned16(dx->data)) return RSIS_ERR_BAD_ARG;
  if (y_act && (!valid_tensor(y_act) || !same_shape(x_raw, y_act) || !is_dense(*y_act) || !aligned16(y_act->data)))
    return RSIS_ERR_BAD_ARG;
  if (dres && (!f32_dense(dres) || !same_shape(x_raw, dres))) return RSIS_ERR_BAD_ARG;
  if (x_raw->c % 4 != 0 || !aligned16(mean) || !aligned16(invstd)) return RSIS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t M = (size_t)x_raw->n * x_raw->h * x_raw->w;
  const int C = x_raw->c;
  const float* xp = reinterpret_cast<const float*>(x_raw->data);
  const float* gp = reinterpret_cast<const float*>(dy->data);
  View yv = y_act ? make_view(*y_act) : make_view(*x_raw);
  int blocks = (int)((M + 15) / 16);
  if (blocks > 592) blocks = 592;
  bn_bwd_reduce_kernel<<<blocks, 256, 256 * 8 * sizeof(float), st>>>(xp, yv, y_act ? 1 : 0, gp, mean, invstd, M, C,
                                                                    workspace);
  RSIS_CHECK_LAUNCH();
  bn_bwd_finalize_kernel<<<ceil_div(C, 256), 256, 0, st>>>(workspace, C, dweight, dbias, dweight_acc, dbias_acc);
  RSIS_CHECK_LAUNCH();
  const size_t total4 = M * C / 4;
  bn_bwd_apply_kernel<<<grid_for(total4, 256), 256, 0, st>>>(xp, yv, y_act ? 1 : 0, gp, weight, mean, invstd, dweight,
                                                             dbias, 1.0f / (float)M, dx->data, plane_elems(*dx), dx->fmt,
                                                             dres ? reinterpret_cast<float*>(dres->data) : nullptr, C,
                                                             total4);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_maxpool3x3s2_bwd(const rsis_tensor* x, const rsis_tensor* dy, const rsis_tensor* dx, rsis_stream_t stream) {
  if (!valid_tensor(x) || !f32_dense(dy) || !f32_dense(dx) || !same_shape(x, dx)) return RSIS_ERR_BAD_ARG;
  const int Ho = (x->h + 2 - 3) / 2 + 1, Wo = (x->w + 2 - 3) / 2 + 1;
  if (dy->n != x->n || dy->c != x->c || dy->h != Ho || dy->w != Wo) return RSIS_ERR_BAD_ARG;
  if (x->c % 4 != 0 || !is_dense(*x) || !aligned16(x->data)) return RSIS_ERR_UNSUPPORTED;
  const size_t total = numel(*x) / 4;
  maxpool3x3s2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      make_view(*x), reinterpret_cast<const float*>(dy->data), reinterpret_cast<float*>(dx->data), x->n, x->h, x->w,
      x->c, Ho, Wo);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_lstm_gates_fwd(const rsis_tensor* gates, const float* c_prev, const rsis_tensor* h_out, const rsis_tensor* h_out2,
                        const rsis_tensor* c_out, rsis_stream_t stream) {
  if (!f32_dense(gates) || !f32_dense(h_out) || !f32_dense(c_out) || gates->c % 4 != 0) return RSIS_ERR_BAD_ARG;
  const int Ch = gates->c / 4;
  if (Ch % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  auto ok = [&](const rsis_tensor* t) { return t->n == gates->n && t->h == gates->h && t->w == gates->w && t->c == Ch; };
  if (!ok(h_out) || !ok(c_out)) return RSIS_ERR_BAD_ARG;
  if (c_prev && !aligned16(c_prev)) return RSIS_ERR_ALIGN;
  if (h_out2 && (!valid_tensor(h_out2) || !ok(h_out2) || !aligned16(h_out2->data) || pitch(*h_out2) % 4 != 0))
    return RSIS_ERR_BAD_ARG;
  const size_t total4 = (size_t)gates->n * gates->h * gates->w * (Ch / 4);
  lstm_gates_fwd_kernel<<<grid_for(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float*>(gates->data), c_prev, reinterpret_cast<float*>(h_out->data),
      reinterpret_cast<float*>(c_out->data), h_out2 ? h_out2->data : nullptr, h_out2 ? plane_elems(*h_out2) : 0,
      h_out2 ? h_out2->fmt : 0, h_out2 ? pitch(*h_out2) : 0, Ch, total4);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_lstm_gates_bwd(const rsis_tensor* gates, const float* c_prev, const float* c_new, const rsis_tensor* dh_a,
                        const rsis_tensor* dh_b, const rsis_tensor* dc_next, const rsis_tensor* dgates, float* dc_prev,
                        rsis_stream_t stream) {
  if (!f32_dense(gates) || !valid_tensor(dgates) || !is_dense(*dgates) || !aligned16(dgates->data) ||
      !same_shape(gates, dgates) || !c_new || !dc_prev || gates->c % 16 != 0)
    return RSIS_ERR_BAD_ARG;
  const int Ch = gates->c / 4;
  auto ok = [&](const rsis_tensor* t) {
    return valid_tensor(t) && t->fmt == RSIS_FMT_F32 && t->n == gates->n && t->h == gates->h && t->w == gates->w &&
           t->c == Ch && aligned16(t->data) && pitch(*t) % 4 == 0;
  };
  if ((dh_a && !ok(dh_a)) || (dh_b && !ok(dh_b)) || (dc_next && !ok(dc_next))) return RSIS_ERR_BAD_ARG;
  if (!aligned16(c_new) || !aligned16(dc_prev) || (c_prev && !aligned16(c_prev))) return RSIS_ERR_ALIGN;
  const size_t total4 = (size_t)gates->n * gates->h * gates->w * (Ch / 4);
  lstm_gates_bwd_kernel<<<grid_for(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float*>(gates->data), c_prev, c_new,
      dh_a ? reinterpret_cast<const float*>(dh_a->data) : nullptr, dh_a ? pitch(*dh_a) : 0,
      dh_b ? reinterpret_cast<const float*>(dh_b->data) : nullptr, dh_b ? pitch(*dh_b) : 0,
      dc_next ? reinterpret_cast<const float*>(dc_next->data) : nullptr, dc_next ? pitch(*dc_next) : 0,
      dgates->data, plane_elems(*dgates), dgates->fmt, dc_prev, Ch, total4);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_global_maxpool(const rsis_tensor* h, uint64_t* side_packed, int side_stride, int side_offset,
                        rsis_stream_t stream) {
  if (!f32_dense(h) || !side_packed || side_offset < 0 || side_stride < side_offset + h->c) return RSIS_ERR_BAD_ARG;
  const long long HW = (long long)h->h * h->w;
  if (HW > 0x7ffffff0LL) return RSIS_ERR_UNSUPPORTED;
  const int CG = h->c < 32 ? h->c : 32;
  if (256 % CG != 0) return RSIS_ERR_UNSUPPORTED;
  const int groups = ceil_div(h->c, CG);
  long long split = 296 / ((long long)h->n * groups);
  const long long max_split = (HW + 63) / 64;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (h->n > 65535 || groups > 65535) return RSIS_ERR_UNSUPPORTED;
  global_maxpool_kernel<<<dim3(h->n, groups, (unsigned)split), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float*>(h->data), (int)HW, h->c, reinterpret_cast<unsigned long long*>(side_packed),
      side_stride, side_offset);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_global_maxpool_finish(const uint64_t* side_packed, int n, int side_stride, uint32_t* side_keys,
                               int32_t* side_idx, rsis_stream_t stream) {
  if (!side_packed || !side_keys || !side_idx || n < 1 || side_stride < 1) return RSIS_ERR_BAD_ARG;
  const int total = n * side_stride;
  global_maxpool_finish_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned long long*>(side_packed), total, side_keys, side_idx);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_global_maxpool_bwd(const float* dside, const int32_t* side_idx, int side_stride, int side_offset,
                            const rsis_tensor* dh, rsis_stream_t stream) {
  if (!dside || !side_idx || !f32_dense(dh) || side_offset < 0 || side_stride < side_offset + dh->c)
    return RSIS_ERR_BAD_ARG;
  const int total = dh->n * dh->c;
  global_maxpool_bwd_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      dside, side_idx, side_stride, side_offset, reinterpret_cast<float*>(dh->data), dh->n, dh->h * dh->w, dh->c);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_upsample_bilinear_bwd(const rsis_tensor* dy, const rsis_tensor* dx, rsis_stream_t stream) {
  if (!valid_tensor(dy) || dy->fmt != RSIS_FMT_F32 || !f32_dense(dx) || dy->n != dx->n || dy->c != dx->c)
    return RSIS_ERR_BAD_ARG;
  if (dx->c % 4 != 0 || pitch(*dy) % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(dy->data)) return RSIS_ERR_ALIGN;
  const float sh = dy->h > 1 ? (float)(dx->h - 1) / (float)(dy->h - 1) : 0.f;
  const float sw = dy->w > 1 ? (float)(dx->w - 1) / (float)(dy->w - 1) : 0.f;
  const size_t total = numel(*dx) / 4;
  upsample_bilinear_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float*>(dy->data), pitch(*dy), reinterpret_cast<float*>(dx->data), dx->n, dx->h, dx->w,
      dx->c, dy->h, dy->w, sh, sw);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_class_stop_heads_bwd(const float* feat, const float* class_probs, const float* dclass, const float* dstop,
                              int n, int f, const float* w_class, int num_classes, const float* w_stop,
                              float* dlogit_scratch, float* dfeat, float* dw_class, float* db_class, float* dw_stop,
                              float* db_stop, rsis_stream_t stream) {
  if (!feat || !class_probs || !w_class || !w_stop || !dlogit_scratch || !dfeat || !dw_class || !db_class || !dw_stop ||
      !db_stop || n < 1 || f < 1 || num_classes < 1)
    return RSIS_ERR_BAD_ARG;
  if (num_classes + 1 > 10000) return RSIS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  heads_bwd_data_kernel<<<n, 256, (size_t)(num_classes + 1) * sizeof(float), st>>>(class_probs, dclass, dstop, f,
                                                                                  num_classes, w_class, w_stop,
                                                                                  dlogit_scratch, dfeat);
  RSIS_CHECK_LAUNCH();
  heads_bwd_param_kernel<<<num_classes + 1, 256, 0, st>>>(dlogit_scratch, feat, n, f, num_classes, dw_class, db_class,
                                                          dw_stop, db_stop);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

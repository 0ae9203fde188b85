// fp32 CUDA-core implicit-GEMM convolution with two fused epilogues:
//   * folded BatchNorm/bias (+ residual) (+ ReLU)      -- nn.Conv2d + nn.BatchNorm2d (+ReLU, `out += identity`) of
//                                                         /root/reference/src/modules/vision.py:12-19 (torchvision
//                                                         Bottleneck.forward) and model.py:59-63 (skip heads);
//   * the ConvLSTM cell update                           -- /root/reference/src/modules/clstm.py:43-58, plus the
//                                                         global max-pool side feature of model.py:143.
// This is the exact-fp32 path: it serves every shape (any kernel size / stride / channel count, up to three inputs
// concatenated along C without materialising the concat) and is the on-device cross-check for the tcgen05 path.
//
// GEMM view: M = N*Ho*Wo output pixels, N = Cout, K = KH*KW*Cin with k = (kh*KW + kw)*Cin + c.
// CTA tile BM x BN x 16, 256 threads, each thread an (BM/TY) x 4 register tile; the A tile is gathered with
// 64-byte-contiguous channel runs (NHWC), transposed into shared memory; one __syncthreads per K chunk
// (register prefetch of chunk k+1 overlaps the FMAs of chunk k).
#include "common.cuh"

namespace rsis {

constexpr int kBK = 16;
constexpr int kThreads = 256;

struct SrcList {
  View v[3];
  int c_begin[4];  // channel range of source s is [c_begin[s], c_begin[s+1])
  int n;
};

struct ConvParams {
  SrcList src;
  const float* w_kc;
  const float* scale;
  const float* shift;
  int H, W, Cin;
  int Ho, Wo, Cout, cout_pad;
  int KH, KW, stride, pad;
  int K, M;
  // conv epilogue
  View res;
  int has_res, relu;
  void* y;
  size_t y_plane;
  int y_fmt;
  void* y2;
  size_t y2_plane;
  int y2_fmt;
  // cell epilogue
  const float* c_prev;
  float* h_out;
  float* c_out;
  __nv_bfloat16* h_split;  // hi plane; lo plane at + M*Ch
  uint32_t* side_max;
  int side_stride, side_offset;
};

__device__ __forceinline__ void store4(void* p, size_t plane, int fmt, size_t idx, const float v[4]) {
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

template <int BM, int BN, bool CELL>
__global__ void __launch_bounds__(kThreads) conv_simt_kernel(const ConvParams p) {
  pdl_trigger();
  constexpr int TX = BN / 4;        // threads along N, 4 columns each
  constexpr int TY = kThreads / TX; // threads along M
  constexpr int RM = BM / TY;       // rows per thread
  constexpr int AR = BM / 16;       // A rows gathered per thread per chunk
  constexpr int LDA = BM + 4;
  static_assert(RM >= 1 && (RM == 2 || RM % 4 == 0), "row tile");

  __shared__ __align__(16) float As[2][kBK][LDA];
  __shared__ __align__(16) float Bs[2][kBK][BN];
  __shared__ uint32_t s_key[TX];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HoWo = p.Ho * p.Wo;

  // ---- per-thread gather rows (fixed for the whole K loop) ----
  const int lk = tid % kBK;
  const int lr0 = tid / kBK;
  int row_n[AR], row_hw[AR];  // image index (or -1), packed (hi0 << 16 | wi0 & 0xffff)
#pragma unroll
  for (int i = 0; i < AR; ++i) {
    const int m = m0 + lr0 + 16 * i;
    if (m < p.M) {
      const int n = m / HoWo;
      const int r = m - n * HoWo;
      const int ho = r / p.Wo;
      const int wo = r - ho * p.Wo;
      row_n[i] = n;
      row_hw[i] = ((ho * p.stride - p.pad) << 16) | ((wo * p.stride - p.pad) & 0xffff);
    } else {
      row_n[i] = -1;
      row_hw[i] = 0;
    }
  }
  // B tile loader mapping
  constexpr int B_F4 = kBK * BN / 4;  // float4 per chunk
  const int b_r = tid / (BN / 4), b_c = (tid % (BN / 4)) * 4;
  const bool b_active = tid < B_F4;

  float a_reg[AR];
  float4 b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

  auto gather = [&](int kc) {
    const int k = kc * kBK + lk;
    if (k < p.K) {
      const int tap = k / p.Cin;
      const int c = k - tap * p.Cin;
      const int kh = tap / p.KW;
      const int kw = tap - kh * p.KW;
      int s = 0;
      if (p.src.n > 1 && c >= p.src.c_begin[1]) s = 1;
      if (p.src.n > 2 && c >= p.src.c_begin[2]) s = 2;
      const View v = p.src.v[s];
      const int cl = c - p.src.c_begin[s];
      const bool live = c < p.src.c_begin[p.src.n];  // channels past the last source: omitted zero state
#pragma unroll
      for (int i = 0; i < AR; ++i) {
        float val = 0.f;
        const int hi = (row_hw[i] >> 16) + kh;
        const int wi = (int)(short)(row_hw[i] & 0xffff) + kw;
        if (live && row_n[i] >= 0 && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)
          val = load_elem(v, (((size_t)row_n[i] * p.H + hi) * p.W + wi) * v.c + cl);
        a_reg[i] = val;
      }
    } else {
#pragma unroll
      for (int i = 0; i < AR; ++i) a_reg[i] = 0.f;
    }
    if (b_active) {
      const int kr = kc * kBK + b_r;
      b_reg = kr < p.K ? *reinterpret_cast<const float4*>(p.w_kc + (size_t)kr * p.cout_pad + n0 + b_c)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < AR; ++i) As[buf][lk][lr0 + 16 * i] = a_reg[i];
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_r][b_c]) = b_reg;
  };

  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + kBK - 1) / kBK;
  gather(0);
  stash(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    const int cur = kc & 1;
    if (kc + 1 < nk) gather(kc + 1);
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[RM];
      if constexpr (RM == 2) {
        const float2 t = *reinterpret_cast<const float2*>(&As[cur][kk][ty * RM]);
        a[0] = t.x;
        a[1] = t.y;
      } else {
#pragma unroll
        for (int q = 0; q < RM / 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(&As[cur][kk][ty * RM + 4 * q]);
          a[4 * q + 0] = t.x;
          a[4 * q + 1] = t.y;
          a[4 * q + 2] = t.z;
          a[4 * q + 3] = t.w;
        }
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    if (kc + 1 < nk) stash(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const int col = n0 + tx * 4;
  if constexpr (!CELL) {
    if (col >= p.Cout) return;
    const float4 sc = *reinterpret_cast<const float4*>(p.scale + col);
    const float4 sh = *reinterpret_cast<const float4*>(p.shift + col);
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int m = m0 + ty * RM + i;
      if (m >= p.M) continue;
      const size_t idx = (size_t)m * p.Cout + col;
      float v[4];
      v[0] = fmaf(acc[i][0], sc.x, sh.x);
      v[1] = fmaf(acc[i][1], sc.y, sh.y);
      v[2] = fmaf(acc[i][2], sc.z, sh.z);
      v[3] = fmaf(acc[i][3], sc.w, sh.w);
      if (p.has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += load_elem(p.res, idx + j);
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      store4(p.y, p.y_plane, p.y_fmt, idx, v);
      if (p.y2) store4(p.y2, p.y2_plane, p.y2_fmt, idx, v);
    }
  } else {
    // columns are gate-interleaved: this thread's 4 accumulators are (in, remember, out, cell) of one hidden channel
    const int Ch = p.Cout >> 2;
    const int ch = col >> 2;
    const int m_last = (m0 + BM < p.M ? m0 + BM : p.M) - 1;
    const bool one_image = p.side_max && (m0 / HoWo == m_last / HoWo);
    if (one_image) {
      if (tid < TX) s_key[tid] = 0u;
      __syncthreads();
    }
    if (ch < Ch) {
      const float4 sc = *reinterpret_cast<const float4*>(p.scale + col);
      const float4 sh = *reinterpret_cast<const float4*>(p.shift + col);
      const size_t MCh = (size_t)p.M * Ch;
      uint32_t run_key = 0u;
      int run_n = -1;
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty * RM + i;
        if (m >= p.M) continue;
        const size_t idx = (size_t)m * Ch + ch;
        const float gi = sigmoidf_acc(fmaf(acc[i][0], sc.x, sh.x));
        const float gf = sigmoidf_acc(fmaf(acc[i][1], sc.y, sh.y));
        const float go = sigmoidf_acc(fmaf(acc[i][2], sc.z, sh.z));
        const float gg = tanhf(fmaf(acc[i][3], sc.w, sh.w));
        const float cp = p.c_prev ? p.c_prev[idx] : 0.f;
        const float c = gf * cp + gi * gg;
        const float h = go * tanhf(c);
        p.c_out[idx] = c;
        p.h_out[idx] = h;
        if (p.h_split) {
          __nv_bfloat16 hi, lo;
          split_bf16(h, hi, lo);
          p.h_split[idx] = hi;
          p.h_split[idx + MCh] = lo;
        }
        if (p.side_max) {
          const uint32_t key = float_to_key(h);
          if (one_image) {
            run_key = key > run_key ? key : run_key;
          } else {
            const int n = m / HoWo;
            if (n != run_n) {
              if (run_n >= 0) atomicMax(p.side_max + (size_t)run_n * p.side_stride + p.side_offset + ch, run_key);
              run_n = n;
              run_key = key;
            } else {
              run_key = key > run_key ? key : run_key;
            }
          }
        }
      }
      if (p.side_max) {
        if (one_image) {
          if (run_key) atomicMax(&s_key[tx], run_key);
        } else if (run_n >= 0) {
          atomicMax(p.side_max + (size_t)run_n * p.side_stride + p.side_offset + ch, run_key);
        }
      }
    }
    if (one_image) {
      __syncthreads();
      if (tid < TX && (n0 >> 2) + tid < Ch && s_key[tid])
        atomicMax(p.side_max + (size_t)(m0 / HoWo) * p.side_stride + p.side_offset + (n0 >> 2) + tid, s_key[tid]);
    }
  }
}

template <bool CELL>
static int launch_conv_simt(const ConvParams& p, cudaStream_t st) {
  const bool narrow = p.Cout <= 32;
  const long tiles128 = (long)ceil_div(p.M, 128) * ceil_div(p.Cout, narrow ? 32 : 64);
  const bool small = tiles128 < 2 * 148;
  if (narrow) {
    if (small)
      conv_simt_kernel<64, 32, CELL><<<dim3(ceil_div(p.M, 64), ceil_div(p.Cout, 32)), kThreads, 0, st>>>(p);
    else
      conv_simt_kernel<128, 32, CELL><<<dim3(ceil_div(p.M, 128), ceil_div(p.Cout, 32)), kThreads, 0, st>>>(p);
  } else {
    if (small)
      conv_simt_kernel<64, 64, CELL><<<dim3(ceil_div(p.M, 64), ceil_div(p.Cout, 64)), kThreads, 0, st>>>(p);
    else
      conv_simt_kernel<128, 64, CELL><<<dim3(ceil_div(p.M, 128), ceil_div(p.Cout, 64)), kThreads, 0, st>>>(p);
  }
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

// Fills the source list / geometry shared by both entry points.  Returns RSIS_OK or an error.
static int fill_common(ConvParams& p, const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, int stride,
                       int pad, bool allow_missing_tail) {
  if (!srcs || n_src < 1 || n_src > 3 || !w || !w->w_kc || !w->scale || !w->shift) return RSIS_ERR_BAD_ARG;
  if (stride < 1 || pad < 0 || w->kh < 1 || w->kw < 1 || w->cout < 1) return RSIS_ERR_BAD_ARG;
  int c = 0;
  for (int s = 0; s < n_src; ++s) {
    if (!valid_tensor(&srcs[s])) return RSIS_ERR_BAD_ARG;
    if (!is_dense(srcs[s])) return RSIS_ERR_UNSUPPORTED;  // the CUDA-core path reads dense NHWC only
    if (srcs[s].n != srcs[0].n || srcs[s].h != srcs[0].h || srcs[s].w != srcs[0].w) return RSIS_ERR_BAD_ARG;
    if (!aligned16(srcs[s].data)) return RSIS_ERR_ALIGN;
    p.src.v[s] = make_view(srcs[s]);
    p.src.c_begin[s] = c;
    c += srcs[s].c;
  }
  for (int s = n_src; s < 4; ++s) p.src.c_begin[s] = c;
  for (int s = n_src; s < 3; ++s) p.src.v[s] = p.src.v[0];
  p.src.n = n_src;
  // The weight's K layout always spans w->cin channels; a ConvLSTM step whose state is None omits the trailing
  // prev_hidden source (clstm.py:26-37 materialises zeros) and those channels contribute nothing.
  if (c != w->cin && !(allow_missing_tail && c < w->cin)) return RSIS_ERR_BAD_ARG;
  if (w->cout % 4 != 0) return RSIS_ERR_UNSUPPORTED;
  if (!aligned16(w->w_kc) || !aligned16(w->scale) || !aligned16(w->shift)) return RSIS_ERR_ALIGN;
  p.w_kc = w->w_kc;
  p.scale = w->scale;
  p.shift = w->shift;
  p.H = srcs[0].h;
  p.W = srcs[0].w;
  p.Cin = w->cin;
  p.KH = w->kh;
  p.KW = w->kw;
  p.stride = stride;
  p.pad = pad;
  p.Ho = (p.H + 2 * pad - w->kh) / stride + 1;
  p.Wo = (p.W + 2 * pad - w->kw) / stride + 1;
  if (p.Ho < 1 || p.Wo < 1 || p.H > 32000 || p.W > 32000) return RSIS_ERR_UNSUPPORTED;
  p.Cout = w->cout;
  p.cout_pad = round_up(w->cout, 64);
  p.K = w->kh * w->kw * w->cin;
  const long long M = (long long)srcs[0].n * p.Ho * p.Wo;
  if (M > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
  p.M = (int)M;
  return RSIS_OK;
}

int conv2d_simt(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, cudaStream_t st) {
  ConvParams p{};
  if (int e = fill_common(p, srcs, n_src, w, stride, pad, false)) return e;
  if (w->gate_interleaved) return RSIS_ERR_BAD_ARG;
  if (!valid_tensor(y) || y->n != srcs[0].n || y->h != p.Ho || y->w != p.Wo || y->c != p.Cout) return RSIS_ERR_BAD_ARG;
  if (!aligned16(y->data)) return RSIS_ERR_ALIGN;
  if (!is_dense(*y) || (y2 && valid_tensor(y2) && !is_dense(*y2)) || (residual && valid_tensor(residual) && !is_dense(*residual)))
    return RSIS_ERR_UNSUPPORTED;
  p.y = y->data;
  p.y_plane = numel(*y);
  p.y_fmt = y->fmt;
  if (y2) {
    if (!valid_tensor(y2) || numel(*y2) != numel(*y) || y2->c != y->c) return RSIS_ERR_BAD_ARG;
    if (!aligned16(y2->data)) return RSIS_ERR_ALIGN;
    p.y2 = y2->data;
    p.y2_plane = numel(*y2);
    p.y2_fmt = y2->fmt;
  }
  if (residual) {
    if (!valid_tensor(residual) || numel(*residual) != numel(*y) || residual->c != y->c) return RSIS_ERR_BAD_ARG;
    p.res = make_view(*residual);
    p.has_res = 1;
  }
  p.relu = relu ? 1 : 0;
  return launch_conv_simt<false>(p, st);
}

int convlstm_cell_simt(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const float* c_prev,
                       const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, cudaStream_t st) {
  ConvParams p{};
  if (!w || w->kh != w->kw || (w->kh != 1 && w->kh != 3)) return RSIS_ERR_UNSUPPORTED;
  if (int e = fill_common(p, srcs, n_src, w, 1, w->kh / 2, c_prev == nullptr)) return e;
  if (!w->gate_interleaved) return RSIS_ERR_BAD_ARG;
  const int Ch = p.Cout / 4;
  auto ok = [&](const rsis_tensor* t, int fmt) {
    return valid_tensor(t) && t->fmt == fmt && t->n == srcs[0].n && t->h == p.Ho && t->w == p.Wo && t->c == Ch;
  };
  if (!ok(h_out, RSIS_FMT_F32) || !ok(c_out, RSIS_FMT_F32)) return RSIS_ERR_BAD_ARG;
  if (h_split && !ok(h_split, RSIS_FMT_SPLIT_BF16)) return RSIS_ERR_BAD_ARG;
  if (!is_dense(*h_out) || !is_dense(*c_out) || (h_split && !is_dense(*h_split))) return RSIS_ERR_UNSUPPORTED;
  if (side_max && (side_stride < side_offset + Ch || side_offset < 0)) return RSIS_ERR_BAD_ARG;
  p.c_prev = c_prev;
  p.h_out = reinterpret_cast<float*>(h_out->data);
  p.c_out = reinterpret_cast<float*>(c_out->data);
  p.h_split = h_split ? reinterpret_cast<__nv_bfloat16*>(h_split->data) : nullptr;
  p.side_max = side_max;
  p.side_stride = side_stride;
  p.side_offset = side_offset;
  return launch_conv_simt<true>(p, st);
}

}  // namespace rsis

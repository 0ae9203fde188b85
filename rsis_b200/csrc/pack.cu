// Weight packing: the reference's OIHW float32 nn.Conv2d parameters (+ eval-mode nn.BatchNorm2d statistics)
// -> K-major packed weights and a folded per-channel affine.  Runs once per load_state_dict.
//
// Replaces the parameter handling of nn.Conv2d / nn.BatchNorm2d at /root/reference/src/modules/model.py:43-54,
// /root/reference/src/modules/clstm.py:17 and torchvision's Bottleneck (conv -> bn, bias-free).
#include "common.cuh"

namespace rsis {

// packed output channel j -> reference output channel
__device__ __forceinline__ int ref_cout(int j, int cout, int gate_interleave) {
  if (!gate_interleave) return j;
  const int ch = cout >> 2;
  return (j & 3) * ch + (j >> 2);  // packed (channel, gate) <- reference [in|remember|out|cell] blocks (clstm.py:47)
}

__global__ void pack_simt_kernel(const float* __restrict__ w, float* __restrict__ w_kc, int cout, int cin, int kh,
                                 int kw, int cout_pad, int gate_interleave) {
  const int K = kh * kw * cin;
  const size_t total = (size_t)K * cout_pad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % cout_pad);
    const int k = (int)(i / cout_pad);
    float v = 0.f;
    if (j < cout) {
      const int co = ref_cout(j, cout, gate_interleave);
      const int c = k % cin;
      const int tap = k / cin;
      v = w[((size_t)co * cin + c) * (kh * kw) + tap];
    }
    w_kc[i] = v;
  }
}

__global__ void pack_affine_kernel(const float* __restrict__ bias, const float* __restrict__ bn_w,
                                   const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                   const float* __restrict__ bn_var, float eps, int cout, int cout_pad,
                                   int gate_interleave, float* __restrict__ scale, float* __restrict__ shift) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cout_pad) return;
  float s = 0.f, t = 0.f;
  if (j < cout) {
    const int co = ref_cout(j, cout, gate_interleave);
    const float b = bias ? bias[co] : 0.f;
    if (bn_w) {
      // y = (conv + b - mean) / sqrt(var + eps) * gamma + beta
      s = bn_w[co] / sqrtf(bn_var[co] + eps);
      t = (b - bn_mean[co]) * s + bn_b[co];
    } else {
      s = 1.f;
      t = b;
    }
  }
  scale[j] = s;
  shift[j] = t;
}

// tcgen05 pack: two bf16 planes (hi | lo) of shape [cout_pad_umma][k_pad], K-major.  K is laid out as
// [tap][source][64-channel chunk], each chunk zero padded to 64, so that one (tap, source, chunk) of the activation
// TMA box lines up with 64 consecutive weight columns.
__global__ void pack_umma_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_umma, int cout, int cin,
                                 int taps, int cout_pad, int k_pad, int n_src, int c0, int c1, int c2,
                                 int gate_interleave) {
  const size_t plane = (size_t)cout_pad * k_pad;
  const int cs[3] = {c0, c1, c2};
  int chunks[3], chunk_base[4];
  chunk_base[0] = 0;
  for (int s = 0; s < 3; ++s) {
    chunks[s] = s < n_src ? (cs[s] + 63) / 64 : 0;
    chunk_base[s + 1] = chunk_base[s] + chunks[s];
  }
  const int chunks_per_tap = chunk_base[3];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % k_pad);
    const int j = (int)(i / k_pad);
    float v = 0.f;
    const int chunk = kk / 64, within = kk % 64;
    const int tap = chunk / chunks_per_tap;
    const int cidx = chunk % chunks_per_tap;
    if (j < cout && tap < taps) {
      int s = 0;
      while (s < 2 && cidx >= chunk_base[s + 1]) ++s;
      const int cl = (cidx - chunk_base[s]) * 64 + within;  // channel within source s
      if (cl < cs[s]) {
        int c = cl;
        for (int q = 0; q < s; ++q) c += cs[q];
        const int co = ref_cout(j, cout, gate_interleave);
        v = w[((size_t)co * cin + c) * taps + tap];
      }
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    w_umma[i] = hi;
    w_umma[i + plane] = lo;
  }
}


// ---- one-launch pack: CUDA-core pack + folded affine + tcgen05 pack, optionally of the DATA-GRADIENT weights ---------
// `dgrad` != 0: the logical weight tensor being packed is w'[co'][ci'][tap'] = w[ci'][ci0 + co'][taps - 1 - tap'] (the
// rotated, in/out swapped kernel of rsis_conv_dgrad_weights) read straight from the OIHW parameter -- no intermediate.
struct PackSrc {
  const float* w;
  int cin_src;   // input channels of the stored OIHW tensor
  int taps;
  int dgrad, ci0;
};
__device__ __forceinline__ float pack_src(const PackSrc& s, int co, int c, int tap) {
  return s.dgrad ? s.w[((size_t)c * s.cin_src + s.ci0 + co) * s.taps + (s.taps - 1 - tap)]
                 : s.w[((size_t)co * s.cin_src + c) * s.taps + tap];
}

__global__ void pack_all_kernel(PackSrc src, const float* __restrict__ bias, const float* __restrict__ bn_w,
                                const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                const float* __restrict__ bn_var, float eps, int cout, int cin, int cout_pad,
                                int gate_interleave, float* __restrict__ w_kc, float* __restrict__ scale,
                                float* __restrict__ shift, __nv_bfloat16* __restrict__ w_umma, int cout_pad_umma,
                                int k_pad, int n_src, int c0, int c1, int c2) {
  const int taps = src.taps;
  const size_t total_simt = w_kc ? (size_t)taps * cin * cout_pad : 0;
  const size_t plane = w_umma ? (size_t)cout_pad_umma * k_pad : 0;
  const size_t total = total_simt > plane ? total_simt : plane;
  const int cs[3] = {c0, c1, c2};
  int chunk_base[4];
  chunk_base[0] = 0;
  for (int q = 0; q < 3; ++q) chunk_base[q + 1] = chunk_base[q] + (q < n_src ? (cs[q] + 63) / 64 : 0);
  const int chunks_per_tap = chunk_base[3];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total || i < (size_t)cout_pad;
       i += (size_t)gridDim.x * blockDim.x) {
    if (i < (size_t)cout_pad) {  // folded affine (pack_affine_kernel)
      const int j = (int)i;
      float sc = 0.f, sh = 0.f;
      if (j < cout) {
        const int co = ref_cout(j, cout, gate_interleave);
        const float b = bias ? bias[co] : 0.f;
        if (bn_w) {
          sc = bn_w[co] / sqrtf(bn_var[co] + eps);
          sh = (b - bn_mean[co]) * sc + bn_b[co];
        } else {
          sc = 1.f;
          sh = b;
        }
      }
      scale[j] = sc;
      shift[j] = sh;
    }
    if (i < total_simt) {  // pack_simt_kernel
      const int j = (int)(i % cout_pad);
      const int k = (int)(i / cout_pad);
      float v = 0.f;
      if (j < cout) v = pack_src(src, ref_cout(j, cout, gate_interleave), k % cin, k / cin);
      w_kc[i] = v;
    }
    if (i < plane) {  // pack_umma_kernel
      const int kk = (int)(i % k_pad);
      const int j = (int)(i / k_pad);
      float v = 0.f;
      const int chunk = kk / 64, within = kk % 64;
      const int tap = chunk / chunks_per_tap;
      const int cidx = chunk % chunks_per_tap;
      if (j < cout && tap < taps) {
        int q = 0;
        while (q < 2 && cidx >= chunk_base[q + 1]) ++q;
        const int cl = (cidx - chunk_base[q]) * 64 + within;
        if (cl < cs[q]) {
          int c = cl;
          for (int r = 0; r < q; ++r) c += cs[r];
          v = pack_src(src, ref_cout(j, cout, gate_interleave), c, tap);
        }
      }
      __nv_bfloat16 hi, lo;
      split_bf16(v, hi, lo);
      w_umma[i] = hi;
      w_umma[i + plane] = lo;
    }
  }
}

}  // namespace rsis

using namespace rsis;

extern "C" {

static inline int cout_pad_of(int cout) { return round_up(cout, 64); }

size_t rsis_conv_pack_bytes_simt(int cout, int cin, int kh, int kw) {
  if (cout <= 0 || cin <= 0 || kh <= 0 || kw <= 0) return 0;
  return (size_t)kh * kw * cin * cout_pad_of(cout) * sizeof(float);
}

size_t rsis_conv_pack_bytes_affine(int cout) { return cout > 0 ? (size_t)cout_pad_of(cout) * sizeof(float) : 0; }

int rsis_conv_umma_kpad(int kh, int kw, int n_src, const int32_t* src_c) {
  if (kh <= 0 || kw <= 0 || n_src < 1 || n_src > 3 || !src_c) return 0;
  int chunks = 0;
  for (int s = 0; s < n_src; ++s) {
    if (src_c[s] <= 0) return 0;
    chunks += (src_c[s] + 63) / 64;
  }
  return kh * kw * chunks * 64;
}

int rsis_conv_umma_coutpad(int cout) { return cout > 0 ? round_up(cout, 16) : 0; }

size_t rsis_conv_pack_bytes_umma(int cout, int kh, int kw, int n_src, const int32_t* src_c) {
  const int kp = rsis_conv_umma_kpad(kh, kw, n_src, src_c);
  if (kp == 0 || cout <= 0) return 0;
  return (size_t)2 * rsis_conv_umma_coutpad(cout) * kp * sizeof(__nv_bfloat16);
}

int rsis_conv_pack(const float* w_oihw, const float* bias, const float* bn_weight, const float* bn_bias,
                   const float* bn_mean, const float* bn_var, float bn_eps, int cout, int cin, int kh, int kw,
                   int gate_interleave, float* w_kc, float* scale, float* shift, rsis_stream_t stream) {
  if (!w_oihw || !scale || !shift || cout <= 0 || cin <= 0 || kh <= 0 || kw <= 0) return RSIS_ERR_BAD_ARG;
  if (gate_interleave && (cout % 4) != 0) return RSIS_ERR_BAD_ARG;
  if (bn_weight && (!bn_bias || !bn_mean || !bn_var)) return RSIS_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int cp = cout_pad_of(cout);
  if (w_kc) {
    const size_t total = (size_t)kh * kw * cin * cp;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_simt_kernel<<<blocks, 256, 0, st>>>(w_oihw, w_kc, cout, cin, kh, kw, cp, gate_interleave);
    RSIS_CHECK_LAUNCH();
  }
  pack_affine_kernel<<<ceil_div(cp, 128), 128, 0, st>>>(bias, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, cout, cp,
                                                       gate_interleave, scale, shift);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_conv_pack_umma(const float* w_oihw, int cout, int cin, int kh, int kw, int n_src, const int32_t* src_c,
                        int gate_interleave, void* w_umma, rsis_stream_t stream) {
  if (!w_oihw || !w_umma || cout <= 0 || cin <= 0) return RSIS_ERR_BAD_ARG;
  const int kp = rsis_conv_umma_kpad(kh, kw, n_src, src_c);
  if (kp == 0) return RSIS_ERR_BAD_ARG;
  int csum = 0;
  for (int s = 0; s < n_src; ++s) csum += src_c[s];
  if (csum != cin) return RSIS_ERR_BAD_ARG;
  if (gate_interleave && (cout % 4) != 0) return RSIS_ERR_BAD_ARG;
  const int cp = rsis_conv_umma_coutpad(cout);
  const size_t plane = (size_t)cp * kp;
  const int blocks = (int)((plane + 255) / 256 < 4096 ? (plane + 255) / 256 : 4096);
  pack_umma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, reinterpret_cast<__nv_bfloat16*>(w_umma), cout, cin, kh * kw, cp, kp, n_src, src_c[0],
      n_src > 1 ? src_c[1] : 0, n_src > 2 ? src_c[2] : 0, gate_interleave);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

int rsis_conv_pack_all(const float* w_oihw, int w_cout, int w_cin, int kh, int kw, int dgrad, int ci0, int nci,
                       const float* bias, const float* bn_weight, const float* bn_bias, const float* bn_mean,
                       const float* bn_var, float bn_eps, int gate_interleave, int n_src, const int32_t* src_c,
                       float* w_kc, float* scale, float* shift, void* w_umma, rsis_stream_t stream) {
  if (!w_oihw || !scale || !shift || w_cout <= 0 || w_cin <= 0 || kh <= 0 || kw <= 0) return RSIS_ERR_BAD_ARG;
  if (bn_weight && (!bn_bias || !bn_mean || !bn_var)) return RSIS_ERR_BAD_ARG;
  if (dgrad && (ci0 < 0 || nci < 1 || ci0 + nci > w_cin || bias || bn_weight || gate_interleave)) return RSIS_ERR_BAD_ARG;
  // logical convolution being packed
  const int cout = dgrad ? nci : w_cout;
  const int cin = dgrad ? w_cout : w_cin;
  if (gate_interleave && (cout % 4) != 0) return RSIS_ERR_BAD_ARG;
  int kp = 0, cpu = 0, cs[3] = {0, 0, 0};
  if (w_umma) {
    if (!src_c || n_src < 1 || n_src > 3) return RSIS_ERR_BAD_ARG;
    kp = rsis_conv_umma_kpad(kh, kw, n_src, src_c);
    if (kp == 0) return RSIS_ERR_BAD_ARG;
    int csum = 0;
    for (int q = 0; q < n_src; ++q) {
      cs[q] = src_c[q];
      csum += src_c[q];
    }
    if (csum != cin) return RSIS_ERR_BAD_ARG;
    cpu = rsis_conv_umma_coutpad(cout);
  }
  const int cp = cout_pad_of(cout);
  const size_t total_simt = w_kc ? (size_t)kh * kw * cin * cp : 0;
  const size_t plane = (size_t)cpu * kp;
  size_t total = total_simt > plane ? total_simt : plane;
  if (total < (size_t)cp) total = cp;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  const PackSrc src{w_oihw, w_cin, kh * kw, dgrad ? 1 : 0, ci0};
  pack_all_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, bias, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, cout,
                                                           cin, cp, gate_interleave, w_kc, scale, shift,
                                                           reinterpret_cast<__nv_bfloat16*>(w_umma), cpu, kp, n_src,
                                                           cs[0], cs[1], cs[2]);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

// Fused Adam over a flat parameter / gradient buffer: the step after `loss.backward()`
// (/root/reference/src/train.py:185-187: `dec_opt.step()`, `enc_opt.step()`), i.e. torch.optim.Adam as
// utils/utils.py:72-83 builds it (L2 weight decay folded into the gradient, betas (0.9, 0.999), eps 1e-8), in ONE
// elementwise pass per run of parameters that share (lr, weight_decay, repeats) instead of ~10 small kernels per
// parameter tensor.
// `repeats`: utils/utils.py:34-52 (`get_base_params`) yields every backbone parameter once per enclosing module
// (3x for a Bottleneck convolution, 4x inside `downsample`), and torch's per-parameter loop then applies the update
// that many times per `step()` -- sequentially, each time with its own step count and the already-updated value.
// The kernel reproduces exactly that sequence per element.
#include <math.h>

#include "common.cuh"

namespace rsis {

constexpr int kAdamMaxRepeats = 8;

struct AdamCoef {
  float step_size[kAdamMaxRepeats];   // lr / (1 - beta1^t)
  float bc2_sqrt[kAdamMaxRepeats];    // sqrt(1 - beta2^t)
};

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float beta1, float beta2, float eps, float wd, int repeats,
                            const AdamCoef coef) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    const float gi = g[i];
    for (int r = 0; r < repeats; ++r) {
      const float gg = fmaf(wd, pi, gi);                 // grad.add(param, alpha=weight_decay)
      mi = fmaf(gg - mi, 1.f - beta1, mi);               // exp_avg.lerp_(grad, 1 - beta1)
      vi = fmaf(1.f - beta2, gg * gg, vi * beta2);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vi) / coef.bc2_sqrt[r] + eps;
      pi -= coef.step_size[r] * (mi / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace rsis

using namespace rsis;

extern "C" {

int rsis_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int64_t step0, int repeats, rsis_stream_t stream) {
  if (!p || !g || !m || !v || n < 1 || repeats < 1 || step0 < 0) return RSIS_ERR_BAD_ARG;
  if (repeats > kAdamMaxRepeats) return RSIS_ERR_UNSUPPORTED;
  AdamCoef coef{};
  for (int r = 0; r < repeats; ++r) {
    const double t = (double)(step0 + r + 1);
    coef.step_size[r] = (float)((double)lr / (1.0 - pow((double)beta1, t)));
    coef.bc2_sqrt[r] = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  size_t blocks = ((size_t)n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (size_t)n, beta1, beta2, eps, weight_decay,
                                                                 repeats, coef);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // extern "C"

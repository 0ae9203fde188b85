// Shared device/host helpers for the rsis_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rsis_b200.h"

namespace rsis {

// ---- status plumbing ------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e);  // records text for rsis_last_cuda_error()

#define RSIS_CUDA_TRY(expr)                      \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) {                     \
      ::rsis::set_cuda_error(_e);                \
      return RSIS_ERR_CUDA;                      \
    }                                            \
  } while (0)

#define RSIS_CHECK_LAUNCH() RSIS_CUDA_TRY(cudaGetLastError())

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
static inline size_t numel(const rsis_tensor& t) { return (size_t)t.n * t.h * t.w * t.c; }
static inline int pitch(const rsis_tensor& t) { return t.cstride > 0 ? t.cstride : t.c; }  // elements per pixel
static inline bool is_dense(const rsis_tensor& t) { return pitch(t) == t.c; }
static inline size_t plane_elems(const rsis_tensor& t) { return (size_t)t.n * t.h * t.w * pitch(t); }

static inline bool valid_tensor(const rsis_tensor* t) {
  return t && t->data && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && (t->cstride == 0 || t->cstride >= t->c) &&
         (t->fmt == RSIS_FMT_F32 || t->fmt == RSIS_FMT_SPLIT_BF16);
}

// ---- element access for the two activation formats ---------------------------------------------------------
struct View {  // device-side view of an NHWC activation
  const void* p;
  size_t plane;  // elements between the hi and lo planes (split format)
  int fmt;
  int c;
};

static inline View make_view(const rsis_tensor& t) { return View{t.data, plane_elems(t), t.fmt, t.c}; }

__device__ __forceinline__ float load_elem(const View& v, size_t idx) {
  if (v.fmt == RSIS_FMT_F32) return reinterpret_cast<const float*>(v.p)[idx];
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(v.p);
  return __bfloat162float(b[idx]) + __bfloat162float(b[idx + v.plane]);
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ void store_elem(void* p, size_t plane, int fmt, size_t idx, float x) {
  if (fmt == RSIS_FMT_F32) {
    reinterpret_cast<float*>(p)[idx] = x;
  } else {
    __nv_bfloat16 hi, lo;
    split_bf16(x, hi, lo);
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(p);
    b[idx] = hi;
    b[idx + plane] = lo;
  }
}

// ---- order-preserving float <-> uint32 key (for atomicMax-based global max pooling) --------------------------
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Programmatic dependent launch (PDL).  `pdl_trigger` lets the NEXT kernel of the stream begin launching its CTAs
// (prologue only) while this one is still running; a kernel launched with the programmatic-serialization attribute
// calls `pdl_wait` before it touches anything an earlier kernel produced.  Both are no-ops in plain launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace rsis

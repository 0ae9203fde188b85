// Library-level entry points of the C ABI: version, status text, device check.
#include <string.h>

#include "common.cuh"

namespace rsis {
extern const bool kHasTcgen05;  // conv_umma.cu

static thread_local char g_cuda_err[256] = "";

void set_cuda_error(cudaError_t e) {
  const char* s = cudaGetErrorString(e);
  strncpy(g_cuda_err, s ? s : "unknown CUDA error", sizeof(g_cuda_err) - 1);
  g_cuda_err[sizeof(g_cuda_err) - 1] = 0;
}

}  // namespace rsis

extern "C" {

int rsis_abi_version(void) { return RSIS_ABI_VERSION; }

const char* rsis_strerror(int status) {
  switch (status) {
    case RSIS_OK: return "ok";
    case RSIS_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or inconsistent shapes)";
    case RSIS_ERR_UNSUPPORTED: return "shape / format combination not implemented by this build";
    case RSIS_ERR_CUDA: return "CUDA runtime call or kernel launch failed (see rsis_last_cuda_error)";
    case RSIS_ERR_ARCH: return "current device is not compute capability 10.x (B200)";
    case RSIS_ERR_ALIGN: return "pointer is not 16-byte aligned";
    default: return "unknown rsis status";
  }
}

const char* rsis_last_cuda_error(void) { return rsis::g_cuda_err; }

int rsis_has_tcgen05(void) { return rsis::kHasTcgen05 ? 1 : 0; }

int rsis_device_check(void) {
  int dev = 0;
  RSIS_CUDA_TRY(cudaGetDevice(&dev));
  int major = 0;
  RSIS_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? RSIS_OK : RSIS_ERR_ARCH;
}

}  // extern "C"

// api.cu -- library-wide entry points and error plumbing of libslcl.so.
#include "common.cuh"

#include <string.h>

namespace slcl {

static thread_local char g_cuda_error[256] = "";

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();       // launch-configuration errors only; never synchronises
  if (e == cudaSuccess) return SLCL_OK;
  set_cuda_error(e, where);
  return SLCL_ERR_CUDA;
}

void ensure_context_on_this_thread() {
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;                // cudaSetDevice (CUDA >= 12) initialises the primary context and makes it current on this
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);      // thread; unlike cudaFree(0) it is legal during stream capture
    bound = true;
  }
}

int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}

int sm_count() {
  // cached per device (the only global state of the library)
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsDefault;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsDefault;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace slcl

extern "C" int slcl_version(void) { return SLCL_VERSION; }

extern "C" const char* slcl_strerror(int status) {
  switch (status) {
    case SLCL_OK: return "ok";
    case SLCL_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SLCL_ERR_UNSUPPORTED: return "unsupported shape or stride";
    case SLCL_ERR_WORKSPACE: return "workspace too small or misaligned";
    case SLCL_ERR_CUDA: return "CUDA launch error";
    default: return "unknown status";
  }
}

extern "C" const char* slcl_last_cuda_error(void) { return slcl::g_cuda_error; }

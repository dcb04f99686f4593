// Library-level entry points of include/v3d_b200.h: version, status strings, device check.
#include "common.cuh"

namespace v3d {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
}  // namespace v3d

extern "C" int v3d_abi_version(void) { return 1; }

extern "C" const char* v3d_status_string(int status) {
  switch (status) {
    case V3D_OK: return "ok";
    case V3D_ERR_INVALID_ARGUMENT: return "invalid argument";
    case V3D_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case V3D_ERR_CUDA: return "CUDA error";
    case V3D_ERR_UNSUPPORTED_DEVICE: return "unsupported device (need compute capability 10.x)";
    default: return "unknown status";
  }
}

extern "C" const char* v3d_last_cuda_error(void) { return cudaGetErrorString(v3d::g_last_cuda_error); }

// replaces get_cudart_version() (vision3d/ops/csrc/cuda_version.cu) behind get_cuda_version()
extern "C" int v3d_cudart_version(void) { return CUDART_VERSION; }

extern "C" int v3d_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    v3d::set_cuda_error(e);
    return V3D_ERR_CUDA;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    v3d::set_cuda_error(e);
    return V3D_ERR_CUDA;
  }
  return major == 10 ? V3D_OK : V3D_ERR_UNSUPPORTED_DEVICE;
}

// Diagnostics and shared host-side state of the C ABI.
#include "common.cuh"
#include <string.h>

static thread_local char g_last_cuda_error[256] = "";

int rb2_set_cuda_error(cudaError_t e) {
  strncpy(g_last_cuda_error, cudaGetErrorString(e), sizeof(g_last_cuda_error) - 1);
  return RB2_ERR_CUDA;
}

extern "C" int rb2_abi_version(void) { return RB2_ABI_VERSION; }

extern "C" const char* rb2_status_string(int s) {
  switch (s) {
    case RB2_OK: return "ok";
    case RB2_ERR_CUDA: return "CUDA runtime error";
    case RB2_ERR_ARG: return "bad argument";
    case RB2_ERR_CAPACITY: return "device-side capacity exceeded";
    case RB2_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

extern "C" const char* rb2_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int rb2_sm_count(int* out) {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RB2_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return RB2_ERR_NO_DEVICE;
    cached = n;
  }
  if (out) *out = cached;
  return RB2_OK;
}

// sizeof() of the ABI structs, so a binding can verify its own layout.
extern "C" int rb2_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(rb2_hough_desc);
    case 1: return (int)sizeof(rb2_hough_stats);
    case 2: return (int)sizeof(rb2_vote_desc);
    case 3: return (int)sizeof(rb2_ransac_desc);
    case 4: return (int)sizeof(rb2_ransac_result);
    case 5: return (int)sizeof(rb2_refit_result);
    default: return -1;
  }
}

// Diagnostics and shared host-side state of the C ABI.
#include "common.cuh"
#include <string.h>

static thread_local char g_last_cuda_error[256] = "";

int rb2_set_cuda_error(cudaError_t e) {
  strncpy(g_last_cuda_error, cudaGetErrorString(e), sizeof(g_last_cuda_error) - 1);
  cudaGetLastError();  // reported through the return value; do not leave it for the caller's next CUDA call
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
    case 6: return (int)sizeof(rb2_match_desc);
    case 7: return (int)sizeof(rb2_match_result);
    case 8: return (int)sizeof(rb2_pairs_desc);
    case 9: return (int)sizeof(rb2_prepare_desc);
    case 10: return (int)sizeof(rb2_findpeaks_desc);
    case 11: return (int)sizeof(rb2_refine_desc);
    default: return -1;
  }
}

// FP64 roofline denominator: a DFMA-only kernel (8 independent chains per thread),
// timed by the caller with CUDA events.  2 * 8 * iters * blocks * threads flops.
__global__ void __launch_bounds__(256) dfma_burn_kernel(int iters, double* __restrict__ out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

extern "C" int rb2_dfma_burn(int blocks, int iters, double* d_out, void* stream) {
  if (blocks <= 0 || iters <= 0 || !d_out) return RB2_ERR_ARG;
  dfma_burn_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, d_out);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

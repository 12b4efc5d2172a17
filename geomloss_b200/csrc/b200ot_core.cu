// b200ot — library-wide plumbing: error strings, device queries, pipe micro-benchmarks.
#include <stdio.h>
#include <string.h>

#include "b200ot.h"
#include "common.cuh"
#include "host_util.cuh"

namespace b200ot {

static thread_local char g_last_cuda_error[512] = "";  // per host thread: concurrent callers do not race on it

void set_last_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s (%s)", where, cudaGetErrorName(e),
           cudaGetErrorString(e));
  (void)cudaGetLastError();  // clear the sticky-free error so that later calls start clean
}

int num_sms() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ------------------------------------------------------------------------------------------------
// Pipe-ceiling micro-benchmarks.  Each thread runs `iters` steps of U independent dependency chains;
// the result is written only if it is NaN-free garbage the compiler cannot predict, so nothing is
// optimised away.  bench.py divides (threads * iters * ops) by the CUDA-event time.
// ------------------------------------------------------------------------------------------------
constexpr int kUbU = 8;

__global__ void __launch_bounds__(256) ub_mufu_kernel(int iters, float* sink) {
  float v[kUbU];
#pragma unroll
  for (int u = 0; u < kUbU; ++u) v[u] = -0.001f * (threadIdx.x + u);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kUbU; ++u) v[u] = ex2_approx(v[u]) - 1.0f;  // 1 MUFU + 1 FADD (other pipe)
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kUbU; ++u) s += v[u];
  if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(256) ub_ffma_kernel(int iters, float* sink) {
  float v[kUbU];
  const float a = 1.0f + 1e-7f * threadIdx.x, b = 1e-9f;
#pragma unroll
  for (int u = 0; u < kUbU; ++u) v[u] = 0.001f * (threadIdx.x + u);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kUbU; ++u) v[u] = fmaf(v[u], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kUbU; ++u) s += v[u];
  if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(256) ub_ffma2_kernel(int iters, float* sink) {
  float2 v[kUbU];
  const float2 a = make_float2(1.0f + 1e-7f * threadIdx.x, 1.0f - 1e-7f * threadIdx.x);
  const float2 b = make_float2(1e-9f, 2e-9f);
#pragma unroll
  for (int u = 0; u < kUbU; ++u) v[u] = make_float2(0.001f * (threadIdx.x + u), 0.002f * (threadIdx.x + u));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < kUbU; ++u) v[u] = __ffma2_rn(v[u], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kUbU; ++u) s += v[u].x + v[u].y;
  if (s == 123.456f) sink[0] = s;
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

B200OT_API int b200ot_version(void) { return B200OT_VERSION; }

B200OT_API const char* b200ot_strerror(int code) {
  switch (code) {
    case B200OT_OK: return "ok";
    case B200OT_EINVAL: return "invalid argument";
    case B200OT_ESCRATCH: return "scratch buffer too small";
    case B200OT_ECUDA: return "CUDA runtime error (see b200ot_last_cuda_error)";
    case B200OT_EALIGN: return "pointer not sufficiently aligned";
    default: return "unknown b200ot error code";
  }
}

B200OT_API const char* b200ot_last_cuda_error(void) { return g_last_cuda_error; }

B200OT_API int b200ot_ubench(int32_t which, int32_t iters, int32_t blocks, float* sink, int32_t* ops_per_thread_iter,
                  void* stream) {
  if (!sink || iters <= 0 || blocks <= 0) return B200OT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int ops = 0;
  switch (which) {
    case B200OT_UBENCH_MUFU_EX2:
      ub_mufu_kernel<<<blocks, 256, 0, st>>>(iters, sink);
      ops = kUbU;
      break;
    case B200OT_UBENCH_FFMA:
      ub_ffma_kernel<<<blocks, 256, 0, st>>>(iters, sink);
      ops = kUbU;
      break;
    case B200OT_UBENCH_FFMA2:
      ub_ffma2_kernel<<<blocks, 256, 0, st>>>(iters, sink);
      ops = 2 * kUbU;
      break;
    default: return B200OT_EINVAL;
  }
  if (ops_per_thread_iter) *ops_per_thread_iter = ops;
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

}  // extern "C"

// tools/tc_probe.cu — minimal tcgen05 bring-up test (development harness, not shipped).
// One CTA: bulk-TMA two pre-packed bf16 operand tiles (K-major, no-swizzle "interleaved" core-matrix
// layout) into shared memory, issue K/16 tcgen05.mma (M=128, N=NT, fp32 accumulate in TMEM), read the
// accumulator back with tcgen05.ld and compare D = A * B^T with a CPU reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Igeomloss_b200/csrc -Iinclude tools/tc_probe.cu -o build/tc_probe
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "tc.cuh"

using namespace b200ot;

#define CK(x)                                                                           \
  do {                                                                                  \
    cudaError_t e = (x);                                                                \
    if (e != cudaSuccess) {                                                             \
      fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

template <int NT, int K>
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ a_img,
                                                    const __nv_bfloat16* __restrict__ b_img, float* __restrict__ d) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int A_BYTES = 128 * K * 2, B_BYTES = NT * K * 2;
  unsigned char* sa = smem;
  unsigned char* sb = smem + A_BYTES;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_load, A_BYTES + B_BYTES);
    tma_load_1d(sa, a_img, A_BYTES, bar_load);
    tma_load_1d(sb, b_img, B_BYTES, bar_load);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, NT);
#pragma unroll
    for (int k = 0; k < K / 16; ++k) {
      // K-major interleaved: chunk kc (8 elements = 16 B) of all rows is contiguous -> LBO = rows*16 B, SBO = 128 B
      const uint64_t da = make_smem_desc(smem_u32(sa) + k * 2 * (128 * 16), 128 * 16, 128);
      const uint64_t db = make_smem_desc(smem_u32(sb) + k * 2 * (NT * 16), NT * 16, 128);
      umma_bf16(tmem_d, da, db, idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int c = 0; c < 32; ++c) d[row * NT + c0 + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, NT);
}

// element (row, k) of a [rows x K] operand -> index in the interleaved image
static inline size_t img_index(int rows, int row, int k) { return ((size_t)(k / 8) * rows + row) * 8 + (k % 8); }

template <int NT, int K>
static int run() {
  std::vector<float> A(128 * K), B(NT * K);
  std::vector<__nv_bfloat16> ai(128 * K), bi(NT * K);
  srand(7);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)rand() / RAND_MAX - 0.5f);
      A[r * K + k] = __bfloat162float(h);
      ai[img_index(128, r, k)] = h;
    }
  for (int r = 0; r < NT; ++r)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16((float)rand() / RAND_MAX - 0.5f);
      B[r * K + k] = __bfloat162float(h);
      bi[img_index(NT, r, k)] = h;
    }
  __nv_bfloat16 *da, *db;
  float* dd;
  CK(cudaMalloc(&da, ai.size() * 2));
  CK(cudaMalloc(&db, bi.size() * 2));
  CK(cudaMalloc(&dd, 128 * NT * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, 128 * NT * 4));
  const int smem = 128 * K * 2 + NT * K * 2 + 64;
  auto kern = probe_kernel<NT, K>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<1, 128, smem>>>(da, db, dd);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * NT);
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < NT; ++c) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * B[c * K + k];
      const double e = fabs(ref - D[r * NT + c]);
      if (!(e < 1e-3)) ++bad;
      if (e > maxerr || e != e) maxerr = e;
    }
  printf("{\"tc_probe\": \"M128 N%d K%d\", \"max_abs_err\": %.3e, \"bad\": %d, \"d00\": %f}\n", NT, K, maxerr, bad,
         D[0]);
  return bad;
}

int main() {
  int bad = 0;
  bad += run<128, 16>();
  bad += run<128, 64>();
  bad += run<64, 208>();
  bad += run<256, 32>();
  return bad ? 1 : 0;
}

// tools/explore.cu — development harness (not part of the shipped library).
//
// 1. pipe micro-benchmarks (MUFU.EX2, FFMA, FFMA2, FMNMX3, mixes) -> per-SM-per-clock rates;
// 2. softmin partial-kernel variants on synthetic uniform clouds: CUDA-event timing + fp64 CPU check
//    of a sample of rows.
// Build (after __graft_entry__.build(); links the packing / micro-benchmark entry points from the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Igeomloss_b200/csrc tools/explore.cu \
//        -Lgeomloss_b200 -lb200ot -Xlinker -rpath -Xlinker '$ORIGIN/../geomloss_b200' -o build/explore
// Usage: explore [N] [M] [eps] [reps]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "b200ot.h"
#include "host_util.cuh"
#include "pack.cuh"
#include "softmin.cuh"

using namespace b200ot;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

// ---------------- extra micro-benchmarks ----------------
constexpr int U = 8;
__global__ void __launch_bounds__(256) ub_fmnmx3(int iters, float* sink) {
  float v[U];
  float a = 0.5f * threadIdx.x, b = 0.25f * threadIdx.x;
  for (int u = 0; u < U; ++u) v[u] = 0.001f * (threadIdx.x + u);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = fmax3(v[u], a, b), a += 1.f;
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += v[u];
  if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) ub_fmnmx(int iters, float* sink) {
  float v[U];
  float a = 0.5f * threadIdx.x;
  for (int u = 0; u < U; ++u) v[u] = 0.001f * (threadIdx.x + u);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = fmaxf(v[u], a) * 0.999f;
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += v[u];
  if (s == 123.456f) sink[0] = s;
}
// MUFU + k packed FFMA2 per exp: how many FMA-pipe ops fit "for free" next to a saturated MUFU?
template <int K>
__global__ void __launch_bounds__(256) ub_mufu_ffma2(int iters, float* sink) {
  float v[U];
  float2 w[U];
  const float2 a = make_float2(0.999f, 0.998f), b = make_float2(1e-3f, 2e-3f);
  for (int u = 0; u < U; ++u) v[u] = -0.001f * (threadIdx.x + u), w[u] = make_float2(v[u], v[u]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = ex2_approx(v[u]) - 1.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) w[u] = __ffma2_rn(w[u], a, b);
    }
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += v[u] + w[u].x + w[u].y;
  if (s == 123.456f) sink[0] = s;
}
template <int K>
__global__ void __launch_bounds__(256) ub_mufu_ffma(int iters, float* sink) {
  float v[U];
  float w[U];
  for (int u = 0; u < U; ++u) v[u] = -0.001f * (threadIdx.x + u), w[u] = v[u];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = ex2_approx(v[u]) - 1.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) w[u] = fmaf(w[u], 0.999f, 1e-3f);
    }
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += v[u] + w[u];
  if (s == 123.456f) sink[0] = s;
}
// K integer adds (alu pipe) per FFMA2: does a packed FMA take one issue slot or two?
template <int K>
__global__ void __launch_bounds__(256) ub_ffma2_iadd(int iters, float* sink) {
  float2 w[U];
  int q[U];
  const float2 a = make_float2(0.999f, 0.998f), b = make_float2(1e-3f, 2e-3f);
  for (int u = 0; u < U; ++u) w[u] = make_float2(0.001f * (threadIdx.x + u), 0.5f), q[u] = threadIdx.x + u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      w[u] = __ffma2_rn(w[u], a, b);
#pragma unroll
      for (int k = 0; k < K; ++k) q[u] = (q[u] ^ (it + k)) + 0x9e3779b9;
    }
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += w[u].x + w[u].y + (float)q[u];
  if (s == 123.456f) sink[0] = s;
}
template <int K>
__global__ void __launch_bounds__(256) ub_ffma_iadd(int iters, float* sink) {
  float w[U];
  int q[U];
  const float a = 0.999f + 1e-6f * threadIdx.x, b = 1e-3f * threadIdx.x;
  for (int u = 0; u < U; ++u) w[u] = 0.001f * (threadIdx.x + u), q[u] = threadIdx.x + u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      w[u] = fmaf(w[u], a, b);
#pragma unroll
      for (int k = 0; k < K; ++k) q[u] = (q[u] ^ (it + k)) + 0x9e3779b9;
    }
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += w[u] + (float)q[u];
  if (s == 123.456f) sink[0] = s;
}
// MUFU + K register-operand scalar FFMA
template <int K>
__global__ void __launch_bounds__(256) ub_mufu_ffma_reg(int iters, float* sink) {
  float v[U];
  float w[U];
  const float a = 0.999f + 1e-6f * threadIdx.x, b = 1e-3f * threadIdx.x;
  for (int u = 0; u < U; ++u) v[u] = -0.001f * (threadIdx.x + u), w[u] = v[u];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = ex2_approx(v[u]) - 1.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) w[u] = fmaf(w[u], a, b);
    }
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += v[u] + w[u];
  if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) ub_expoly(int iters, float* sink) {
  float2 v[U];
  for (int u = 0; u < U; ++u) v[u] = make_float2(-0.001f * (threadIdx.x + u), -0.002f * (threadIdx.x + u));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float2 e = ex2_poly2(v[u]);
      v[u] = __fadd2_rn(e, dup2(-1.0f));
    }
  }
  float s = 0;
  for (int u = 0; u < U; ++u) s += v[u].x + v[u].y;
  if (s == 123.456f) sink[0] = s;
}
__global__ void __launch_bounds__(256) ub_ex2_f16x2(int iters, float* sink) {
  unsigned v[U];
  for (int u = 0; u < U; ++u) v[u] = 0xB800B800u + threadIdx.x;  // ~ -0.5 halves
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned r;
      asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(v[u]));
      v[u] = r ^ 0x80008000u;  // negate both halves (alu pipe)
    }
  }
  unsigned s = 0;
  for (int u = 0; u < U; ++u) s += v[u];
  if (s == 123456u) sink[0] = (float)s;
}

template <class F>
static double time_kernel(F launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

static void report_ub(const char* name, double ms, double ops_per_thread_iter, int iters, int blocks, int sms,
                      double clk_mhz) {
  const double total = (double)blocks * 256 * iters * ops_per_thread_iter;
  const double rate = total / (ms * 1e-3);
  printf("{\"ubench\": \"%s\", \"ms\": %.4f, \"Gops\": %.1f, \"ops_per_clk_per_sm@%.0fMHz\": %.2f}\n", name, ms,
         rate * 1e-9, clk_mhz, rate / (sms * clk_mhz * 1e6));
}

// ---------------- softmin variants ----------------
struct Problem {
  int64_t N, M;
  float eps;
  float *x, *y, *h, *center, *cols, *part, *out, *lse2;
  std::vector<float> hx, hy, hh;
};

static const char* g_filter = nullptr;
static int g_nsplit = 0;  // > 0: override the number of column splits (wave-quantisation study)

template <class C>
static void run_variant(const char* name, Problem& P, int reps) {
  if (g_filter && !strstr(name, g_filter)) return;
  const int D = C::D;
  const int p = C::P;
  const int64_t mpad = round_up64(P.M, 1024);
  const int ntiles = (int)(round_up64(P.M, C::TJ) / C::TJ);
  const int64_t row_tiles = ceil_div64(P.N, C::ROWS_PER_CTA);
  int64_t want = ceil_div64((int64_t)148 * 2 * 16, row_tiles);
  if (want > 64) want = 64;
  if (want > ntiles) want = ntiles;
  if (want < 1) want = 1;
  if (g_nsplit > 0) want = g_nsplit < ntiles ? g_nsplit : ntiles;
  const int tps = (int)ceil_div64(ntiles, want);
  const int last_pairs = (int)(round_up64(P.M - (int64_t)(ntiles - 1) * C::TJ, 32) / 2);
  const int nsplit = (int)ceil_div64(ntiles, tps);
  const float scale = p == 2 ? sqrtf(kLog2e / P.eps) : kLog2e / P.eps;
  const float clampq = scale * scale * 1e-8f;

  pack_cols_kernel<<<(unsigned)ceil_div64(mpad, 256), 256>>>(P.y, P.h, nullptr, 0.f, kLog2e, nullptr, P.center,
                                                              scale, C::DIRECT ? 1 : 0, D, C::NF2, P.M, mpad,
                                                              P.cols);
  CK(cudaGetLastError());
  auto kern = softmin_partial_kernel<C, false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT + 32, C::SMEM_BYTES));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  dim3 grid((unsigned)row_tiles, (unsigned)nsplit);
  auto launch = [&]() {
    kern<<<grid, C::NT + 32, C::SMEM_BYTES>>>(P.x, P.center, scale, clampq, P.cols, (float2*)P.part, P.N, ntiles,
                                              tps, last_pairs, (const int4*)nullptr, (const int2*)nullptr);
  };
  const double ms = time_kernel(launch, reps);
  CK(b200ot_softmin_finalize(P.part, nsplit, nullptr, 0.f, 1.f, P.out, P.lse2, P.N, P.eps, nullptr) == 0
         ? cudaSuccess
         : cudaErrorUnknown);
  CK(cudaDeviceSynchronize());

  // fp64 CPU check on a sample of rows
  std::vector<float> out(P.N);
  CK(cudaMemcpy(out.data(), P.out, P.N * 4, cudaMemcpyDeviceToHost));
  double max_abs = 0, max_rel = 0;
  const int nsample = 48;
  for (int sidx = 0; sidx < nsample; ++sidx) {
    const int64_t i = (int64_t)((double)sidx / nsample * P.N);
    double mx = -1e300;
    std::vector<double> t(P.M);
    for (int64_t j = 0; j < P.M; ++j) {
      double q = 0;
      for (int k = 0; k < D; ++k) {
        const double d = (double)P.hx[i * D + k] - (double)P.hy[j * D + k];
        q += d * d;
      }
      const double c = p == 2 ? 0.5 * q : sqrt(fmax(q, 1e-8));
      t[j] = (double)P.hh[j] - c / (double)P.eps;
      if (t[j] > mx) mx = t[j];
    }
    double s = 0;
    for (int64_t j = 0; j < P.M; ++j) s += exp(t[j] - mx);
    const double ref = -(double)P.eps * (mx + log(s));
    const double err = fabs(ref - (double)out[i]);
    if (err > max_abs) max_abs = err;
    const double rel = err / fmax(fabs(ref), 1e-30);
    if (rel > max_rel) max_rel = rel;
  }
  const double pairs = (double)P.N * (double)P.M;
  printf(
      "{\"variant\": \"%s\", \"N\": %lld, \"M\": %lld, \"eps\": %g, \"ms\": %.3f, \"Tpairs_s\": %.3f, \"regs\": %d, "
      "\"occ\": %d, \"grid\": [%u,%u], \"waves\": %.2f, \"max_abs_err\": %.3e, \"max_rel_err\": %.3e}\n",
      name, (long long)P.N, (long long)P.M, P.eps, ms, pairs / (ms * 1e-3) * 1e-12, fa.numRegs, occ, grid.x, grid.y,
      (double)grid.x * grid.y / (148.0 * occ), max_abs, max_rel);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 200000;
  const int64_t M = argc > 2 ? atoll(argv[2]) : 200000;
  const float eps = argc > 3 ? (float)atof(argv[3]) : 1e-4f;
  const int reps = argc > 4 ? atoi(argv[4]) : 3;
  g_filter = argc > 5 ? argv[5] : nullptr;
  g_nsplit = argc > 6 ? atoi(argv[6]) : 0;

  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double clk_mhz = clk_khz / 1000.0;
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f}\n", prop.name, prop.multiProcessorCount,
         clk_mhz);
  const int sms = prop.multiProcessorCount;

  float* sink;
  CK(cudaMalloc(&sink, 1024));
  if (!g_filter) {
    const int iters = 4096, blocks = sms * 8;
    int ops = 0;
    double ms;
    ms = time_kernel([&]() { b200ot_ubench(B200OT_UBENCH_MUFU_EX2, iters, blocks, sink, &ops, nullptr); });
    report_ub("mufu_ex2", ms, ops, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { b200ot_ubench(B200OT_UBENCH_FFMA, iters, blocks, sink, &ops, nullptr); });
    report_ub("ffma", ms, ops, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { b200ot_ubench(B200OT_UBENCH_FFMA2, iters, blocks, sink, &ops, nullptr); });
    report_ub("ffma2(as scalar fma)", ms, ops, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_fmnmx3<<<blocks, 256>>>(iters, sink); });
    report_ub("fmnmx3+fadd", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_fmnmx<<<blocks, 256>>>(iters, sink); });
    report_ub("fmnmx+fmul", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_expoly<<<blocks, 256>>>(iters, sink); });
    report_ub("ex2_poly2 (exps)", ms, 2 * U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_ex2_f16x2<<<blocks, 256>>>(iters, sink); });
    report_ub("ex2.f16x2 (exps)", ms, 2 * U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma2<1><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+1ffma2 (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma2<2><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+2ffma2 (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma2<3><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+3ffma2 (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma2<4><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+4ffma2 (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_ffma2_iadd<1><<<blocks, 256>>>(iters, sink); });
    report_ub("ffma2+1x(xor,add) (ffma2 instrs)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_ffma2_iadd<2><<<blocks, 256>>>(iters, sink); });
    report_ub("ffma2+2x(xor,add) (ffma2 instrs)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_ffma_iadd<1><<<blocks, 256>>>(iters, sink); });
    report_ub("ffma+1x(xor,add) (ffma instrs)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma_reg<2><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+2ffma(reg) (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma_reg<4><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+4ffma(reg) (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma_reg<6><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+6ffma(reg) (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma<4><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+4ffma (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma<6><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+6ffma (exps)", ms, U, iters, blocks, sms, clk_mhz);
    ms = time_kernel([&]() { ub_mufu_ffma<8><<<blocks, 256>>>(iters, sink); });
    report_ub("mufu+8ffma (exps)", ms, U, iters, blocks, sms, clk_mhz);
  }

  Problem P;
  P.N = N;
  P.M = M;
  P.eps = eps;
  const int D = 3;
  P.hx.resize(N * D);
  P.hy.resize(M * D);
  P.hh.resize(M);
  srand(1234);
  for (auto& v : P.hx) v = (float)rand() / RAND_MAX;
  for (auto& v : P.hy) v = (float)rand() / RAND_MAX;
  // h_j = log(1/M) + g_j/eps with g of the size of a typical potential
  for (int64_t j = 0; j < M; ++j) P.hh[j] = logf(1.0f / M) + (0.02f * ((float)rand() / RAND_MAX - 0.5f)) / eps;
  CK(cudaMalloc(&P.x, N * D * 4));
  CK(cudaMalloc(&P.y, M * D * 4));
  CK(cudaMalloc(&P.h, M * 4));
  CK(cudaMalloc(&P.center, 16 * 4));
  CK(cudaMalloc(&P.cols, b200ot_packed_cols_floats(M, D, 1) * 4 + 4096));
  CK(cudaMalloc(&P.part, (size_t)64 * N * 8));  // up to 64 splits
  CK(cudaMalloc(&P.out, N * 4));
  CK(cudaMalloc(&P.lse2, N * 4));
  CK(cudaMemcpy(P.x, P.hx.data(), N * D * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(P.y, P.hy.data(), M * D * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(P.h, P.hh.data(), M * 4, cudaMemcpyHostToDevice));
  const float ctr[16] = {0.5f, 0.5f, 0.5f};
  CK(cudaMemcpy(P.center, ctr, 64, cudaMemcpyHostToDevice));

  // round 2: ring depth / tile width / off-load share around the shipped configuration (per-thread stage release)
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 4, 8, 3, true>>("r2 stages4", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 2, 8, 3, true>>("r2 stages2", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 512, 4, 8, 3, true>>("r2 tj512 stages4", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 2048, 2, 8, 3, true>>("r2 tj2048 stages2", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x0u, 256, 1024, 3, 8, 3, true>>("r2 nopoly", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x11u, 256, 1024, 3, 8, 3, true>>("r2 poly2of8", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 16, 3, true>>("r2 ch16 poly1of16", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x101u, 256, 1024, 3, 16, 3, true>>("r2 ch16 poly2of16", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 8, 3, true>>("r2 shipped", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x0u, 256, 1024, 3, 16, 3, true>>("r3 ch16 nopoly", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 16, 3, true>>("r3 ch16 poly1of16", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x0u, 256, 1024, 3, 32, 3, true>>("r3 ch32 nopoly", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 32, 3, true>>("r3 ch32 poly1of32", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x10001u, 256, 1024, 3, 32, 3, true>>("r3 ch32 poly2of32", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x01010101u, 256, 1024, 3, 32, 3, true>>("r3 ch32 poly4of32", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x0u, 256, 1024, 3, 64, 3, true>>("r3 ch64 nopoly", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 64, 3, true>>("r3 ch64 poly1of64", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 4, 16, 3, true>>("r3 ch16 poly1of16 stages4", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 2048, 2, 16, 3, true>>("r3 ch16 poly1of16 tj2048", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0x1u, 256, 1024, 3, 16, 2, true>>("r3 R4 ch16 poly1of16 occ2", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 128, 1024, 3, 16, 6, true>>("r3 NT128 ch16 poly1of16 occ6", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0u, 256, 1024, 3, 4, 2>>("expand R4 CH4", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0u, 256, 1024, 3, 8, 2>>("expand R4 CH8", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 8, 2>>("expand R2 CH8", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 4, 3>>("expand R2 CH4 occ3", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 4, 3, true>>("expand R2 CH4 occ3 guard", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 8, 3, true>>("expand R2 CH8 occ3 guard", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 4, 4, true>>("expand R2 CH4 occ4 guard", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 256, 1024, 3, 2, 4, true>>("expand R2 CH2 occ4 guard", P, reps);
  run_variant<SoftminCfg<3, 3, 2, false, 0u, 256, 1024, 3, 4, 3, true>>("expand R3 CH4 occ3 guard", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0u, 256, 1024, 3, 4, 2, true>>("expand R4 CH4 occ2 guard", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0u, 128, 1024, 3, 4, 5, true>>("expand R4 NT128 CH4 occ5 guard", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 8, 3, true>>("expand R2 CH8 occ3 guard poly1/8", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x1u, 256, 1024, 3, 4, 3>>("expand R2 CH4 occ3 poly1/4", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 128, 1024, 3, 8, 4>>("expand R2 NT128 CH8 occ4", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 128, 1024, 3, 4, 6>>("expand R2 NT128 CH4 occ6", P, reps);
  run_variant<SoftminCfg<3, 1, 2, false, 0u, 256, 1024, 3, 8, 4>>("expand R1 NT256 CH8 occ4", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0u, 512, 1024, 3, 8, 1>>("expand R2 NT512 CH8 occ1", P, reps);
  run_variant<SoftminCfg<3, 3, 2, false, 0u, 256, 1024, 3, 8, 2>>("expand R3 CH8", P, reps);
  run_variant<SoftminCfg<3, 8, 2, false, 0u, 256, 1024, 3, 4, 1>>("expand R8 CH4 occ1", P, reps);
  run_variant<SoftminCfg<3, 8, 2, false, 0u, 128, 1024, 3, 4, 2>>("expand R8 NT128 CH4", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0u, 512, 1024, 3, 8, 1>>("expand R4 NT512 CH8 occ1", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0x01u, 256, 1024, 3, 8, 2>>("expand R4 CH8 poly1/8", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0x11u, 256, 1024, 3, 8, 2>>("expand R4 CH8 poly2/8", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0x0421u, 256, 1024, 3, 16, 2>>("expand R4 CH16 poly3/16", P, reps);
  run_variant<SoftminCfg<3, 8, 2, false, 0x0421u, 128, 1024, 3, 16, 2>>("expand R8 NT128 CH16 poly3/16", P, reps);
  run_variant<SoftminCfg<3, 4, 2, false, 0x1111u, 256, 1024, 3, 16, 2>>("expand R4 CH16 poly4/16", P, reps);
  run_variant<SoftminCfg<3, 2, 2, false, 0x0421u, 256, 1024, 3, 16, 3>>("expand R2 CH16 poly3/16 occ3", P, reps);
  run_variant<SoftminCfg<3, 4, 2, true, 0u, 256, 1024, 3, 4, 2>>("direct R4 CH4", P, reps);
  run_variant<SoftminCfg<3, 2, 1, true, 0u, 256, 1024, 3, 4, 2>>("p1 direct R2 CH4", P, reps);
  run_variant<SoftminCfg<3, 4, 1, true, 0u, 256, 1024, 3, 4, 2>>("p1 direct R4 CH4", P, reps);
  run_variant<SoftminCfg<3, 1, 2, false, 0u, 128, 256, 3, 4, 4>>("expand small R1", P, reps);
  return 0;
}

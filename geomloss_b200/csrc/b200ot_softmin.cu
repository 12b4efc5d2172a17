// b200ot — softmin: column packing, partial-reduction dispatch, finalize, and the C ABI around them.
// Interface contract: include/b200ot.h.  Reference semantics: softmin_tensorized
// (src/geomloss/_legacy/sinkhorn_samples.py:32-71) with the cost routines of :26-29 /
// src/geomloss/_legacy/utils.py:26-61.
#include "b200ot.h"
#include "host_util.cuh"
#include "pack.cuh"
#include "softmin.cuh"

namespace b200ot {

// -------------------------------------------------------------------------------------------------
// finalize: merge partial (m, s) pairs and apply the Sinkhorn epilogue
//   lse2 = M + log2( sum_s s_s 2^(m_s - M) ),  val = -eps ln2 lse2,  out = alpha_old*old + beta*val
// -------------------------------------------------------------------------------------------------
__global__ void softmin_finalize_kernel(const float2* __restrict__ part, int n_part,
                                        const float* __restrict__ out_old, float alpha_old, float beta,
                                        float* __restrict__ out, float* __restrict__ lse2_out, int64_t N,
                                        float neg_eps_ln2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float mm = kNegBig;
  for (int s = 0; s < n_part; ++s) mm = fmaxf(mm, part[(int64_t)s * N + i].x);
  float ss = 0.f;
  for (int s = 0; s < n_part; ++s) {
    const float2 p = part[(int64_t)s * N + i];
    ss += p.y * exp2f(p.x - mm);
  }
  const float lse2 = mm + log2f(ss);
  if (lse2_out) lse2_out[i] = lse2;
  if (out) {
    float v = neg_eps_ln2 * lse2 * beta;
    if (out_old) v = fmaf(alpha_old, out_old[i], v);
    out[i] = v;
  }
}

// -------------------------------------------------------------------------------------------------
// dispatch
// -------------------------------------------------------------------------------------------------
struct SoftminPlan {
  int tj;        // columns per tile
  int rows_cta;  // rows per CTA
  int ntiles;
  int tiles_per_split;
  int n_split;
  int64_t row_tiles;
  int64_t mpad;
  bool small;
};

// Kernel variant selection is a pure function of (N, M, D): big problems use the R=4 register tile,
// small ones a 128x1 tile so that a few thousand points still spread over the 148 SMs.
using CfgBigD3 = SoftminCfg<3, 4, 2, false, 0, 256, 1024, 3, 4, 2>;

template <int D, int P, bool DIRECT>
struct Variants {
  using Big = SoftminCfg<D, (D <= 4 ? 4 : 2), P, DIRECT, 0, 256, 1024, 3, 4, (D <= 4 ? 2 : 2)>;
  using Small = SoftminCfg<D, 1, P, DIRECT, 0, 128, 256, 3, 4, 4>;
};

static SoftminPlan make_plan(int64_t N, int64_t M, int D) {
  SoftminPlan p;
  const int big_rows = 256 * (D <= 4 ? 4 : 2);
  p.small = (N < 4 * (int64_t)big_rows) || (M < 4096);
  p.tj = p.small ? 256 : 1024;
  p.rows_cta = p.small ? 128 : big_rows;
  p.mpad = round_up64(M, p.tj);
  p.ntiles = (int)(p.mpad / p.tj);
  p.row_tiles = ceil_div64(N, p.rows_cta);
  // aim for >= ~16 waves of 2 CTAs/SM when the problem is large enough, never split below one tile
  const int64_t target_ctas = (int64_t)num_sms() * 2 * 16;
  int64_t want = ceil_div64(target_ctas, p.row_tiles);
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  if (want > p.ntiles) want = p.ntiles;
  p.tiles_per_split = (int)ceil_div64(p.ntiles, want);
  p.n_split = (int)ceil_div64(p.ntiles, p.tiles_per_split);
  return p;
}

template <class C>
static int launch_partial(const float* x, const float* center, float scale, float clampq, const float* cols,
                          float* part, const SoftminPlan& pl, int64_t N, cudaStream_t st) {
  auto kern = softmin_partial_kernel<C>;
  static bool attr_done = false;  // per-instantiation
  if (!attr_done) {
    B200OT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid((unsigned)pl.row_tiles, (unsigned)pl.n_split);
  kern<<<grid, C::NT + 32, C::SMEM_BYTES, st>>>(x, center, scale, clampq, cols, reinterpret_cast<float2*>(part), N,
                                                pl.ntiles, pl.tiles_per_split);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

template <int D>
static int partial_for_d(int p, const float* x, const float* center, float scale, float clampq, const float* cols,
                         float* part, const SoftminPlan& pl, int64_t N, cudaStream_t st) {
  if (p == 2) {
    using V = Variants<D, 2, false>;
    return pl.small ? launch_partial<typename V::Small>(x, center, scale, clampq, cols, part, pl, N, st)
                    : launch_partial<typename V::Big>(x, center, scale, clampq, cols, part, pl, N, st);
  } else {
    using V = Variants<D, 1, true>;
    return pl.small ? launch_partial<typename V::Small>(x, center, scale, clampq, cols, part, pl, N, st)
                    : launch_partial<typename V::Big>(x, center, scale, clampq, cols, part, pl, N, st);
  }
}

// scale of the coordinates such that the log2-domain exponent is  H - |X - Y|^2 / 2  (p = 2)
// or  H - |X - Y|  (p = 1)
static inline float coord_scale(int p, float eps) {
  return p == 2 ? sqrtf(kLog2e / eps) : kLog2e / eps;
}

static int softmin_pack_impl(const float* y, const float* h_a, const float* h_b, float h_scale_b,
                             const float* center, int64_t M, int D, int p, float eps, float* cols_out,
                             cudaStream_t st, int tj) {
  const int nf2 = colfmt_nf2(D + 0);  // D coords + 1 extra
  const int64_t mpad = round_up64(M, tj);
  const float scale = coord_scale(p, eps);
  const int threads = 256;
  pack_cols_kernel<<<(unsigned)ceil_div64(mpad, threads), threads, 0, st>>>(
      y, h_a, h_b, h_scale_b, kLog2e, nullptr, center, scale, p == 1 ? 1 : 0, D, 1, nf2, M, mpad, cols_out);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

static int softmin_partial_impl(const float* x, const float* center, const float* cols, float* part,
                                const SoftminPlan& pl, int64_t N, int D, int p, float eps, cudaStream_t st) {
  const float scale = coord_scale(p, eps);
  const float clampq = scale * scale * 1e-8f;  // reference clamp on |x-y|^2 (utils.py:61), in scaled units
  switch (D) {
    case 1: return partial_for_d<1>(p, x, center, scale, clampq, cols, part, pl, N, st);
    case 2: return partial_for_d<2>(p, x, center, scale, clampq, cols, part, pl, N, st);
    case 3: return partial_for_d<3>(p, x, center, scale, clampq, cols, part, pl, N, st);
    default: return B200OT_EINVAL;
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int64_t b200ot_packed_cols_floats(int64_t M, int32_t D, int32_t extra) {
  if (M <= 0 || D <= 0 || extra < 1) return 0;
  const int nf2 = ((D + extra + 1) / 2) * 2;
  return round_up64(M, 1024) / 2 * nf2 * 2;
}

int32_t b200ot_softmin_num_splits(int64_t N, int64_t M, int32_t D) {
  if (N <= 0 || M <= 0) return 0;
  return make_plan(N, M, D).n_split;
}

int64_t b200ot_softmin_scratch_bytes(int64_t N, int64_t M, int32_t D) {
  if (N <= 0 || M <= 0 || D <= 0) return 0;
  const SoftminPlan pl = make_plan(N, M, D);
  const int64_t cols = b200ot_packed_cols_floats(M, D, 1) * 4;
  const int64_t part = (int64_t)pl.n_split * N * 8;
  return round_up64(cols, 256) + round_up64(part, 256);
}

int b200ot_softmin_pack(const float* y, const float* h_a, const float* h_b, float h_scale_b, const float* center,
                        int64_t M, int32_t D, int32_t p, float eps, float* cols_out, void* stream) {
  if (!y || !h_a || !cols_out || M <= 0 || D <= 0 || D > B200OT_MAX_D || (p != 1 && p != 2) || !(eps > 0.f))
    return B200OT_EINVAL;
  if (((uintptr_t)cols_out) & 15) return B200OT_EALIGN;
  // the tile size only affects padding; pad to the largest tile so that either kernel variant can read it
  return softmin_pack_impl(y, h_a, h_b, h_scale_b, center, M, D, p, eps, cols_out, (cudaStream_t)stream, 1024);
}

int b200ot_softmin_partial(const float* x, const float* center, const float* cols, float* part, int32_t n_split,
                           int64_t N, int64_t M, int32_t D, int32_t p, float eps, void* stream) {
  if (!x || !cols || !part || N <= 0 || M <= 0 || D <= 0 || D > B200OT_MAX_D || (p != 1 && p != 2) ||
      !(eps > 0.f))
    return B200OT_EINVAL;
  if ((((uintptr_t)cols) & 15) || (((uintptr_t)part) & 7)) return B200OT_EALIGN;
  const SoftminPlan pl = make_plan(N, M, D);
  if (n_split != pl.n_split) return B200OT_EINVAL;
  return softmin_partial_impl(x, center, cols, part, pl, N, D, p, eps, (cudaStream_t)stream);
}

int b200ot_softmin_finalize(const float* part, int32_t n_part, const float* out_old, float alpha_old, float beta,
                            float* out, float* lse2_out, int64_t N, float eps, void* stream) {
  if (!part || n_part <= 0 || N <= 0 || (!out && !lse2_out)) return B200OT_EINVAL;
  const int threads = 256;
  softmin_finalize_kernel<<<(unsigned)ceil_div64(N, threads), threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(part), n_part, out_old, alpha_old, beta, out, lse2_out, N, -eps * kLn2);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

int b200ot_softmin_fwd(const float* x, const float* y, const float* h_a, const float* h_b, float h_scale_b,
                       const float* center, const float* out_old, float alpha_old, float beta, float* out,
                       float* lse2_out, int64_t N, int64_t M, int32_t D, int32_t p, float eps, void* scratch,
                       int64_t scratch_bytes, void* stream) {
  if (!x || !y || !h_a || !scratch || N <= 0 || M <= 0 || D <= 0 || D > B200OT_MAX_D || (p != 1 && p != 2) ||
      !(eps > 0.f) || (!out && !lse2_out))
    return B200OT_EINVAL;
  if (((uintptr_t)scratch) & 15) return B200OT_EALIGN;
  if (scratch_bytes < b200ot_softmin_scratch_bytes(N, M, D)) return B200OT_ESCRATCH;
  const SoftminPlan pl = make_plan(N, M, D);
  float* cols = reinterpret_cast<float*>(scratch);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) +
                                         round_up64(b200ot_packed_cols_floats(M, D, 1) * 4, 256));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = softmin_pack_impl(y, h_a, h_b, h_scale_b, center, M, D, p, eps, cols, st, 1024);
  if (rc) return rc;
  rc = softmin_partial_impl(x, center, cols, part, pl, N, D, p, eps, st);
  if (rc) return rc;
  return b200ot_softmin_finalize(part, pl.n_split, out_old, alpha_old, beta, out, lse2_out, N, eps, stream);
}

}  // extern "C"

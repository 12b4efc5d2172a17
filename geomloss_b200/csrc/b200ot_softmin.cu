// b200ot — softmin forward: column packing, partial-reduction dispatch, merge / finalize, C ABI.
// Interface contract: include/b200ot.h.  Reference semantics: softmin_tensorized
// (src/geomloss/_legacy/sinkhorn_samples.py:32-71) with the cost routines of :26-29 /
// src/geomloss/_legacy/utils.py:26-61.
#include "b200ot.h"
#include "host_util.cuh"
#include "pack.cuh"
#include "plan.cuh"
#include "softmin.cuh"

namespace b200ot {

// -------------------------------------------------------------------------------------------------
// merge / finalize of partial (m, s) pairs
//   lse2 = M + log2( sum_s s_s 2^(m_s - M) ),  val = -eps ln2 lse2,  out = alpha_old*old + beta*val
// -------------------------------------------------------------------------------------------------
// The merge and the final log run in fp64: with the lazy running max a partial sum can be as large as
// M * 2^64, and log2f of such a value is only accurate to ulp(64) ~ 8e-6 — visible at large eps.
// This is N elements per softmin (elementwise), so the fp64 rate of the part is irrelevant.
struct MergedMS {
  float m;
  double s;
};
__device__ __forceinline__ MergedMS merge_ms(const float2* __restrict__ part, int n_part, int64_t N, int64_t i) {
  float mm = kNegBig;
  for (int s = 0; s < n_part; ++s) mm = fmaxf(mm, part[(int64_t)s * N + i].x);
  double ss = 0.0;
  for (int s = 0; s < n_part; ++s) {
    const float2 p = part[(int64_t)s * N + i];
    ss += (double)p.y * exp2((double)p.x - (double)mm);
  }
  MergedMS r;
  r.m = mm;
  r.s = ss;
  return r;
}

__global__ void softmin_merge_kernel(const float2* __restrict__ part, int n_part, float2* __restrict__ out,
                                     int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  // renormalise so that the merged pair is again (max-like reference, O(1)..O(M) sum) in fp32
  const MergedMS ms = merge_ms(part, n_part, N, i);
  int e = 0;
  const double frac = frexp(ms.s, &e);  // s = frac * 2^e, frac in [0.5, 1)
  out[i] = ms.s > 0.0 ? make_float2(ms.m + (float)e, (float)frac) : make_float2(ms.m, 0.f);
}

__global__ void softmin_finalize_kernel(const float2* __restrict__ part, int n_part,
                                        const float* __restrict__ out_old, float alpha_old, float beta,
                                        float* __restrict__ out, float* __restrict__ lse2_out, int64_t N,
                                        float neg_eps_ln2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const MergedMS ms = merge_ms(part, n_part, N, i);
  const double lse2 = (double)ms.m + log2(ms.s);
  if (lse2_out) lse2_out[i] = (float)lse2;
  if (out) {
    double v = (double)neg_eps_ln2 * lse2 * (double)beta;
    if (out_old) v += (double)alpha_old * (double)out_old[i];
    out[i] = (float)v;
  }
}

// -------------------------------------------------------------------------------------------------
// dispatch
// -------------------------------------------------------------------------------------------------
// Measured on B200 (profiles/r02_explore_variants.jsonl: N = M = 1e6, D = 3; profiles/r02_ab_ops.jsonl: every D, p at
// N = M = 4e5): 2 rows/thread, 3 CTAs/SM (2 for D > 4), sum-guarded lazy max.  What matters is the length of the
// straight-line chunk between two max checks.  D = 3: 8 column pairs 243.4 ms, 16 pairs 232.5 ms (4.30e12 pairs/s,
// 0.93 of the MUFU roofline), 32 pairs 230.5 ms, 64 pairs 229.7 ms — but on a 125k-column shard (multi-GPU) 16 pairs
// is the fastest (30.6 / 31.0 / 32.4 ms), so 16 it is.  With 16-pair chunks the FMA-pipe exp2 off-load no longer pays
// at D = 3, 4 (1 pair in 16: 232.0 ms, 2 in 16: 242.1 ms): the scheduler has enough independent work to keep the MUFU
// queue full on its own.  Per (D, p), 16- vs 8-pair chunks: p = 1 wins 1-5 % for every D; p = 2 wins 5-8 % at D = 3, 4
// but LOSES 4-9 % at D = 1, 2 (so little FMA work per pair that the off-load of one pair in eight still pays there)
// and 3-7 % at D >= 5 (2 CTAs/SM: the longer chunk costs registers the occupancy cannot spare).
// Ranges mode keeps 8-pair chunks: its pieces must be multiples of 2*CH columns, and 16-slot cluster padding
// (kRangesAlign) costs less than the longer chunk gains on cluster-sized pieces.
template <int D, int P, bool DIRECT>
struct Variants {
#ifdef B200OT_BIG_CH  // A/B builds (tools/ab_ops.py)
#ifndef B200OT_BIG_POLY
#define B200OT_BIG_POLY 0u
#endif
  static constexpr int kCH = B200OT_BIG_CH;
  static constexpr unsigned kPoly = (D <= 4 && P == 2 && !DIRECT) ? B200OT_BIG_POLY : 0u;
#else
  static constexpr int kCH = (DIRECT || D == 3 || D == 4) ? 16 : 8;
  static constexpr unsigned kPoly = (D <= 2 && P == 2 && !DIRECT) ? 0x01u : 0u;
#endif
  using Big = SoftminCfg<D, kBigR, P, DIRECT, kPoly, kBigNT, kBigTJ, 3, kCH, (D <= 4 ? 3 : 2), true>;
  using BigRanges = SoftminCfg<D, kBigR, P, DIRECT, 0u, kBigNT, kBigTJ, 3, 8, (D <= 4 ? 3 : 2), true>;
  using Small = SoftminCfg<D, kSmallR, P, DIRECT, 0u, kSmallNT, kSmallTJ, 3, 4, 4, true>;
  static_assert(2 * BigRanges::CH <= kRangesAlign && 2 * Small::CH <= kRangesAlign,
                "ranges pieces are kRangesAlign-aligned");
};

template <class C, class CR = C>
static int launch_partial(const float* x, const float* center, float scale, float clampq, const float* cols,
                          float* part, const ReducePlan& pl, int64_t N, cudaStream_t st, const int4* seg,
                          const int2* pieces) {
  static_assert(C::SMEM_BYTES == CR::SMEM_BYTES && C::NT == CR::NT && C::ROWS_PER_CTA == CR::ROWS_PER_CTA,
                "dense and ranges instantiations share one ReducePlan");
  if (seg != nullptr)
    return launch_reduce<CR>(softmin_partial_kernel<CR, true>, pl, st, x, center, scale, clampq, cols,
                             reinterpret_cast<float2*>(part), N, pl.ntiles, pl.tiles_per_split, 0, seg, pieces);
  return launch_reduce<C>(softmin_partial_kernel<C, false>, pl, st, x, center, scale, clampq, cols,
                          reinterpret_cast<float2*>(part), N, pl.ntiles, pl.tiles_per_split, pl.last_pairs, seg,
                          pieces);
}

template <int D>
static int partial_for_d(int p, const float* x, const float* center, float scale, float clampq, const float* cols,
                         float* part, const ReducePlan& pl, int64_t N, cudaStream_t st, const int4* seg,
                         const int2* pieces) {
  if (p == 2) {
    using V = Variants<D, 2, false>;
    return pl.small ? launch_partial<typename V::Small>(x, center, scale, clampq, cols, part, pl, N, st, seg, pieces)
                    : launch_partial<typename V::Big, typename V::BigRanges>(x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
  } else {
    using V = Variants<D, 1, true>;
    return pl.small ? launch_partial<typename V::Small>(x, center, scale, clampq, cols, part, pl, N, st, seg, pieces)
                    : launch_partial<typename V::Big, typename V::BigRanges>(x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
  }
}

// `p` may carry B200OT_P_UNCLAMPED; M = number of column slots (gather mode: src maps slots to columns)
int softmin_pack_impl(const float* y, const float* h_a, const float* h_b, float h_scale_b, const float* center,
                      int64_t M, int D, int p, float eps, float* cols_out, cudaStream_t st, const int* src) {
  const int pe = p_exponent(p);
  const int nf2 = colfmt_nf2(D, 1);
  const int64_t mpad = round_up64(M, kPackPad);
  const int threads = 256;
  pack_cols_kernel<<<(unsigned)ceil_div64(mpad, threads), threads, 0, st>>>(
      y, h_a, h_b, h_scale_b, kLog2e, nullptr, center, softmin_coord_scale(pe, eps), pe == 1 ? 1 : 0, D, nf2, M,
      mpad, cols_out, src);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

static int softmin_partial_impl(const float* x, const float* center, const float* cols, float* part,
                                const ReducePlan& pl, int64_t N, int D, int p, float eps, cudaStream_t st,
                                const int4* seg = nullptr, const int2* pieces = nullptr) {
  const int pe = p_exponent(p);
  const float scale = softmin_coord_scale(pe, eps);
  // reference clamp on |x-y|^2 (utils.py:61) — or none (pykeops' Norm2) — in scaled units
  const float clampq = scale * scale * cost_clamp(p);
  switch (D) {
    case 1: return partial_for_d<1>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 2: return partial_for_d<2>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 3: return partial_for_d<3>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 4: return partial_for_d<4>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 5: return partial_for_d<5>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 6: return partial_for_d<6>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 7: return partial_for_d<7>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    case 8: return partial_for_d<8>(pe, x, center, scale, clampq, cols, part, pl, N, st, seg, pieces);
    default: return B200OT_EINVAL;
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

B200OT_API int64_t b200ot_packed_cols_floats(int64_t M, int32_t D, int32_t extra) {
  if (M <= 0 || D <= 0 || extra < 1) return 0;
  return round_up64(M, kPackPad) / 2 * colfmt_nf2(D, extra) * 2;
}

B200OT_API int32_t b200ot_softmin_num_splits(int64_t N, int64_t M, int32_t D) {
  (void)D;
  if (N <= 0 || M <= 0) return 0;
  return make_plan(N, M, D).n_split;
}

B200OT_API int64_t b200ot_softmin_scratch_bytes(int64_t N, int64_t M, int32_t D) {
  if (N <= 0 || M <= 0 || D <= 0) return 0;
  if (D > B200OT_MAX_D) return tc_capable_dim(D) ? tc_scratch_bytes(N, M, D) : 0;
  const ReducePlan pl = make_plan(N, M, D);
  const int64_t cols = b200ot_packed_cols_floats(M, D, 1) * 4;
  // forward partials are (m, s) pairs; the backward pass keeps D+1 sums per (split, row)
  const int64_t part = (int64_t)pl.n_split * N * 4 * (D + 1 > 2 ? D + 1 : 2);
  const int64_t simt = round_up64(cols, 256) + round_up64(part, 256);
  return tc_any_routed(D, N, M) ? (simt > tc_scratch_bytes(N, M, D) ? simt : tc_scratch_bytes(N, M, D)) : simt;
}

B200OT_API int b200ot_softmin_pack(const float* y, const float* h_a, const float* h_b, float h_scale_b,
                                   const float* center, int64_t M, int32_t D, int32_t p, float eps,
                                   float* cols_out, void* stream) {
  if (!y || !h_a || !cols_out || M <= 0 || !supported_simt_dim(D) || !valid_p(p) || !(eps > 0.f))
    return B200OT_EINVAL;
  if (((uintptr_t)cols_out) & 15) return B200OT_EALIGN;
  return softmin_pack_impl(y, h_a, h_b, h_scale_b, center, M, D, p, eps, cols_out, (cudaStream_t)stream, nullptr);
}

B200OT_API int b200ot_softmin_partial(const float* x, const float* center, const float* cols, float* part,
                                      int32_t n_split, int64_t N, int64_t M, int32_t D, int32_t p, float eps,
                                      void* stream) {
  if (!x || !cols || !part || N <= 0 || M <= 0 || !supported_simt_dim(D) || !valid_p(p) || !(eps > 0.f))
    return B200OT_EINVAL;
  if ((((uintptr_t)cols) & 15) || (((uintptr_t)part) & 7)) return B200OT_EALIGN;
  const ReducePlan pl = make_plan(N, M, D);
  if (n_split != pl.n_split) return B200OT_EINVAL;
  return softmin_partial_impl(x, center, cols, part, pl, N, D, p, eps, (cudaStream_t)stream);
}

B200OT_API void b200ot_ranges_shape(int32_t variant, int32_t* max_rows_per_segment, int32_t* max_cols_per_piece,
                                    int32_t* col_align) {
  const bool small = (variant == B200OT_RANGES_SMALL);
  if (max_rows_per_segment) *max_rows_per_segment = small ? kSmallNT * kSmallR : kBigNT * kBigR;
  if (max_cols_per_piece) *max_cols_per_piece = small ? kSmallTJ : kBigTJ;
  if (col_align) *col_align = kRangesAlign;
}

B200OT_API int b200ot_softmin_pack_gather(const float* y, const float* h_a, const float* h_b, float h_scale_b,
                                          const float* center, const int32_t* src_index, int64_t n_slots, int32_t D,
                                          int32_t p, float eps, float* cols_out, void* stream) {
  if (!y || !h_a || !src_index || !cols_out || n_slots <= 0 || !supported_simt_dim(D) || !valid_p(p) || !(eps > 0.f))
    return B200OT_EINVAL;
  if (((uintptr_t)cols_out) & 15) return B200OT_EALIGN;
  return softmin_pack_impl(y, h_a, h_b, h_scale_b, center, n_slots, D, p, eps, cols_out, (cudaStream_t)stream,
                           reinterpret_cast<const int*>(src_index));
}

B200OT_API int b200ot_softmin_partial_ranges(const float* x, const float* center, const float* cols,
                                             const b200ot_segment* seg, int64_t n_seg, const b200ot_piece* pieces,
                                             float* part, int64_t N, int32_t D, int32_t p, float eps,
                                             int32_t variant, void* stream) {
  if (!x || !cols || !seg || !pieces || !part || N <= 0 || n_seg <= 0 || n_seg > 0x7fffffff ||
      !supported_simt_dim(D) || !valid_p(p) || !(eps > 0.f) ||
      (variant != B200OT_RANGES_BIG && variant != B200OT_RANGES_SMALL))
    return B200OT_EINVAL;
  if ((((uintptr_t)cols) & 15) || (((uintptr_t)part) & 7) || (((uintptr_t)seg) & 15) || (((uintptr_t)pieces) & 7))
    return B200OT_EALIGN;
  const ReducePlan pl = ranges_plan(variant, n_seg);
  return softmin_partial_impl(x, center, cols, part, pl, N, D, p, eps, (cudaStream_t)stream,
                              reinterpret_cast<const int4*>(seg), reinterpret_cast<const int2*>(pieces));
}

B200OT_API int b200ot_softmin_merge(const float* part, int32_t n_part, float* merged, int64_t N, void* stream) {
  if (!part || !merged || n_part <= 0 || N <= 0) return B200OT_EINVAL;
  if ((((uintptr_t)part) & 7) || (((uintptr_t)merged) & 7)) return B200OT_EALIGN;
  const int threads = 256;
  softmin_merge_kernel<<<(unsigned)ceil_div64(N, threads), threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(part), n_part, reinterpret_cast<float2*>(merged), N);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

B200OT_API int b200ot_softmin_finalize(const float* part, int32_t n_part, const float* out_old, float alpha_old,
                                       float beta, float* out, float* lse2_out, int64_t N, float eps,
                                       void* stream) {
  if (!part || n_part <= 0 || N <= 0 || (!out && !lse2_out) || !(eps > 0.f)) return B200OT_EINVAL;
  if (((uintptr_t)part) & 7) return B200OT_EALIGN;
  const int threads = 256;
  softmin_finalize_kernel<<<(unsigned)ceil_div64(N, threads), threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(part), n_part, out_old, alpha_old, beta, out, lse2_out, N, -eps * kLn2);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

B200OT_API int b200ot_softmin_fwd(const float* x, const float* y, const float* h_a, const float* h_b,
                                  float h_scale_b, const float* center, const float* out_old, float alpha_old,
                                  float beta, float* out, float* lse2_out, int64_t N, int64_t M, int32_t D,
                                  int32_t p, float eps, void* scratch, int64_t scratch_bytes, void* stream) {
  const bool tc = p_exponent(p) == 2 && tc_routed(kTcSoftminFwd, D, N, M);  // exponent from the tensor cores (tcconv.cuh)
  if (!x || !y || !h_a || !scratch || N <= 0 || M <= 0 || (!supported_simt_dim(D) && !tc) || !valid_p(p) ||
      !(eps > 0.f) || (!out && !lse2_out))
    return B200OT_EINVAL;
  if (((uintptr_t)scratch) & 15) return B200OT_EALIGN;
  if (scratch_bytes < b200ot_softmin_scratch_bytes(N, M, D)) return B200OT_ESCRATCH;
  if (tc) {
    float* tc_part = nullptr;
    int n_part = 0;
    const int rc = softmin_partial_tc(x, y, h_a, h_b, h_scale_b, center, N, M, D, eps, scratch, &tc_part, &n_part,
                                      (cudaStream_t)stream);
    if (rc) return rc;
    return b200ot_softmin_finalize(tc_part, n_part, out_old, alpha_old, beta, out, lse2_out, N, eps, stream);
  }
  const ReducePlan pl = make_plan(N, M, D);
  float* cols = reinterpret_cast<float*>(scratch);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) +
                                         round_up64(b200ot_packed_cols_floats(M, D, 1) * 4, 256));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = softmin_pack_impl(y, h_a, h_b, h_scale_b, center, M, D, p, eps, cols, st, nullptr);
  if (rc) return rc;
  rc = softmin_partial_impl(x, center, cols, part, pl, N, D, p, eps, st);
  if (rc) return rc;
  return b200ot_softmin_finalize(part, pl.n_split, out_old, alpha_old, beta, out, lse2_out, N, eps, stream);
}

}  // extern "C"

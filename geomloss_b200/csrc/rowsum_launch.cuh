// b200ot — launch helper of rowsum_partial_kernel: picks the dense or the ranges instantiation.
#pragma once
#include "plan.cuh"
#include "rowsum.cuh"

namespace b200ot {

template <class C>
inline int launch_rowsum_kernel(const ReducePlan& pl, cudaStream_t st, const float* x, const float* center,
                                float scale, float clampq, const float* cols, const float* lse2, float* part,
                                int64_t N, int ntiles, int tiles_per_split, const int4* seg, const int2* pieces) {
  if (seg != nullptr)
    return launch_reduce<C>(rowsum_partial_kernel<C, true>, pl, st, x, center, scale, clampq, cols, lse2, part, N,
                            ntiles, tiles_per_split, 0, seg, pieces);
  return launch_reduce<C>(rowsum_partial_kernel<C, false>, pl, st, x, center, scale, clampq, cols, lse2, part, N,
                          ntiles, tiles_per_split, pl.last_pairs, seg, pieces);
}

}  // namespace b200ot

// b200ot — work decomposition shared by every N x M reduction kernel (softmin, its backward,
// kernel convolutions): which tile shape, how many column splits.
#pragma once
#include "host_util.cuh"

namespace b200ot {

// Tile shapes (must match the Cfg instantiations in the .cu files):
//   big   : 256 consumer threads x 2 rows, 1024-column tiles, 2 CTAs / SM
//   small : 128 consumer threads x 1 row,   256-column tiles  — a few thousand points still spread
//           over the 148 SMs
constexpr int kBigNT = 256, kBigR = 2, kBigTJ = 1024;
constexpr int kSmallNT = 128, kSmallR = 1, kSmallTJ = 256;
constexpr int kPackPad = 1024;  // packed column buffers are padded to this many columns

struct ReducePlan {
  int tj;        // columns per tile
  int rows_cta;  // rows per CTA
  int ntiles;
  int tiles_per_split;
  int n_split;
  int64_t row_tiles;
  bool small;
};

// Pure function of (N, M, D) and the SM count, so that the scratch-size query, the pack stage and the
// reduction stage of the C ABI always agree.
//
// Column splits.  All CTAs of a launch cost the same (rows x tiles of one split), so they run in near lock-step
// waves of `resident` = SMs x CTAs/SM; a grid that is not a whole number of waves pays a full unit for its last
// partial wave (round 1: 5 862 CTAs over 444 slots = 13.2 waves, up to 5.7 % of the kernel).  The plan therefore
// (i) cuts the columns into enough splits for ~40 waves, which bounds that loss by ~1/40 even in the worst case,
// and (ii) among the neighbouring split counts takes the one whose last wave is fullest.  A split keeps >= 8
// tiles, so the per-CTA prologue (row loads, barrier init, TMA ring fill: a few us) stays < 1 % of its ~0.5 ms.
inline ReducePlan make_plan(int64_t N, int64_t M, int D = 3) {
  ReducePlan p;
  const int big_rows = kBigNT * kBigR;
  p.small = (N < 8 * (int64_t)big_rows) || (M < 4 * kBigTJ);
  p.tj = p.small ? kSmallTJ : kBigTJ;
  p.rows_cta = p.small ? kSmallNT * kSmallR : big_rows;
  p.ntiles = (int)(round_up64(M, p.tj) / p.tj);
  p.row_tiles = ceil_div64(N, p.rows_cta);
  const int cps = p.small ? 4 : (D <= 4 ? 3 : 2);  // CTAs/SM of the softmin instantiations (b200ot_softmin.cu)
  const int64_t resident = (int64_t)num_sms() * cps;
  int64_t cap = p.ntiles / 8;  // >= 8 tiles per split ...
  if (cap < 1) cap = 1;
  int64_t want = ceil_div64(40 * resident, p.row_tiles);
  if (want > cap) want = cap;
  // ... unless the problem is too small to fill the machine once: then parallelism comes first
  int64_t fill = ceil_div64(resident, p.row_tiles);
  if (fill > p.ntiles) fill = p.ntiles;
  if (want < fill) want = fill;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  cap = want + 6 < 64 ? want + 6 : 64;
  if (cap > p.ntiles) cap = p.ntiles;
  double best_eff = -1.0;
  int best_tps = p.ntiles;
  for (int64_t w = want; w <= want + 6 && w <= cap; ++w) {
    const int tps = (int)ceil_div64(p.ntiles, w);
    const int ns = (int)ceil_div64(p.ntiles, tps);
    const double waves = (double)(p.row_tiles * ns) / (double)resident;
    const double eff = waves / ceil(waves);
    if (eff > best_eff + 1e-3) {
      best_eff = eff;
      best_tps = tps;
    }
  }
  p.tiles_per_split = best_tps;
  p.n_split = (int)ceil_div64(p.ntiles, p.tiles_per_split);
  return p;
}

// Ranges mode (block-sparse / batched problems): one CTA per row segment, one column split, the pieces of a
// segment are at most one tile wide and aligned to kRangesAlign columns (a whole number of chunks of either
// tile shape: 2 * CH = 16 columns for the big shape, 8 for the small one).
constexpr int kRangesAlign = 16;
inline ReducePlan ranges_plan(int variant, int64_t n_seg) {
  ReducePlan p;
  p.small = (variant == B200OT_RANGES_SMALL);
  p.tj = p.small ? kSmallTJ : kBigTJ;
  p.rows_cta = p.small ? kSmallNT * kSmallR : kBigNT * kBigR;
  p.ntiles = 0;
  p.tiles_per_split = 0;
  p.n_split = 1;
  p.row_tiles = n_seg;
  return p;
}

// Launch one instantiation of a partial-reduction kernel with its dynamic shared memory opt-in.
template <class C, class Kern, class... Args>
inline int launch_reduce(Kern kern, const ReducePlan& pl, cudaStream_t st, Args... args) {
  // the opt-in is per (kernel, device): set it once per device a process drives, not on every launch
  static bool attr_done[64] = {};
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !attr_done[dev]) {
    B200OT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid((unsigned)pl.row_tiles, (unsigned)pl.n_split);
  kern<<<grid, C::NT + 32, C::SMEM_BYTES, st>>>(args...);
  B200OT_CUDA_TRY(cudaGetLastError());
  return B200OT_OK;
}

}  // namespace b200ot

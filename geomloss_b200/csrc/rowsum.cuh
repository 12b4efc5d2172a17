// b200ot — generic "sum over columns" partial-reduction kernel: same tiling / TMA ring / packed-f32x2
// inner loop as softmin.cuh, but the per-row state is a small vector of plain sums (no running max):
//   * softmin backward   (weights 2^(t_ij - lse2_i) are <= 1 by construction)
//   * kernel convolutions out_i = sum_j k(x_i, y_j) w_j  and their row gradients
// Reference semantics: autograd through softmin_tensorized (sinkhorn_samples.py:32-71) and the matvecs
// of kernel_loss (kernel_samples.py:62-82, :116-137).
#pragma once
#include "common.cuh"

namespace b200ot {

enum RowSumMode : int {
  kSoftminBwdP2 = 0,  // acc0 = sum w, acc_{1+k} = sum w Y_k;        w = 2^(H + X.Y + rowterm)
  kSoftminBwdP1 = 1,  // acc0 = sum w, acc_{1+k} = sum w u_k;        w = 2^(H - |X-Y| + rowterm), u = unit vector
  kGaussFwd = 2,      // acc0 = sum W e;                              e = 2^(F + X.Y + rowterm)
  kGaussBwd = 3,      // acc0 = sum W e, acc_{1+k} = sum W e Y_k
  kLaplaceFwd = 4,    // acc0 = sum W 2^(-|X-Y|)
  kLaplaceBwd = 5,    // acc_k = sum W 2^(-|X-Y|) u_k
  kEnergyFwd = 6,     // acc0 = sum W |X-Y|           (sign applied at finalize)
  kEnergyBwd = 7,     // acc_k = sum W u_k
};

__host__ __device__ constexpr bool mode_direct(int m) {
  return m == kSoftminBwdP1 || m == kLaplaceFwd || m == kLaplaceBwd || m == kEnergyFwd || m == kEnergyBwd;
}
__host__ __device__ constexpr int mode_nextra(int m) { return (m == kGaussFwd || m == kGaussBwd) ? 2 : 1; }
__host__ __device__ constexpr int mode_nacc(int m, int D) {
  return (m == kSoftminBwdP2 || m == kSoftminBwdP1 || m == kGaussBwd)
             ? D + 1
             : ((m == kLaplaceBwd || m == kEnergyBwd) ? D : 1);
}

// column pairs unrolled in the inner loop of rowsum_partial_kernel (A/B builds: tools/ab_ops.py)
#ifndef B200OT_ROWSUM_UNROLL
#define B200OT_ROWSUM_UNROLL 2
#endif
constexpr int kRowsumUnroll = B200OT_ROWSUM_UNROLL;

template <int MODE_, int D_, int R_, int NT_ = 256, int TJ_ = 1024, int STAGES_ = 3, int MINB_ = 2>
struct RowSumCfg {
  static constexpr int MODE = MODE_;
  static constexpr int D = D_;
  static constexpr int R = R_;
  static constexpr int NT = NT_;
  static constexpr int TJ = TJ_;
  static constexpr int STAGES = STAGES_;
  static constexpr int MINB = MINB_;
  static constexpr bool DIRECT = mode_direct(MODE);
  static constexpr int NEXTRA = mode_nextra(MODE);
  static constexpr int NACC = mode_nacc(MODE, D);
  static constexpr int NF2 = ((D + NEXTRA + 1) / 2) * 2;
  static constexpr int TILE_FLOATS = (TJ / 2) * NF2 * 2;
  static constexpr int TILE_BYTES = TILE_FLOATS * 4;
  static constexpr int ROWS_PER_CTA = NT * R;
  static constexpr int SMEM_BYTES = STAGES * TILE_BYTES + 2 * STAGES * 8;
};

// rowterm: per-row additive term of the exponent (softmin bwd: rowc_i - lse2_i; gaussian: rowc_i), may be null.
// If rowterm_is_lse2, the kernel forms  (DIRECT ? 0 : -|X|^2/2) - rowterm[i]  itself.
template <class C, bool RANGES>
__global__ void __launch_bounds__(C::NT + 32, C::MINB)
    rowsum_partial_kernel(const float* __restrict__ x, const float* __restrict__ center, float scale, float clampq,
                          const float* __restrict__ cols, const float* __restrict__ lse2, float* __restrict__ part,
                          int64_t N, int ntiles, int tiles_per_split, int last_pairs, const int4* __restrict__ seg,
                          const int2* __restrict__ pieces) {
  constexpr int D = C::D, R = C::R, NT = C::NT, NF2 = C::NF2, STAGES = C::STAGES, NACC = C::NACC, MODE = C::MODE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * C::TILE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.y;
  constexpr bool sparse = RANGES;  // ranges mode (segments + column pieces): see softmin.cuh
  int t0, t1, nrows;
  int64_t row0;
  if (sparse) {
    const int4 sg = seg[blockIdx.x];
    row0 = sg.x;
    nrows = sg.y;
    t0 = sg.z;
    t1 = sg.w;
  } else {
    row0 = (int64_t)blockIdx.x * C::ROWS_PER_CTA;
    nrows = 0;  // (dense mode bounds its rows with N)
    t0 = split * tiles_per_split;
    t1 = min(ntiles, t0 + tiles_per_split);
  }
  const int nt = t1 - t0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NT);  // every consumer THREAD releases the stage itself (see below)
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    if (lane == 0) {
      for (int k = 0; k < nt; ++k) {
        const int st = k % STAGES;
        if (k >= STAGES) mbar_wait(&empty[st], ((k / STAGES) + 1) & 1);
        if (sparse) {
          const int2 pc = pieces[t0 + k];
          const uint32_t bytes = (uint32_t)pc.y * (NF2 * 4);
          mbar_arrive_expect_tx(&full[st], bytes);
          tma_load_1d(tiles + st * C::TILE_FLOATS, cols + (int64_t)pc.x * NF2, bytes, &full[st]);
        } else {
          mbar_arrive_expect_tx(&full[st], C::TILE_BYTES);
          tma_load_1d(tiles + st * C::TILE_FLOATS, cols + (int64_t)(t0 + k) * C::TILE_FLOATS, C::TILE_BYTES,
                      &full[st]);
        }
      }
    }
    return;
  }

  const int tid = threadIdx.x - 32;
  // local row of slot r: see softmin.cuh (ranges mode: a warp owns R*32 consecutive rows, warps without rows idle)
  const auto local_row = [&](int r) { return sparse ? (lane + 32 * (R * (tid >> 5) + r)) : (tid + r * NT); };
  if constexpr (sparse) {
    if (local_row(0) - lane >= nrows) {
      for (int k = 0; k < nt; ++k) {
        const int st = k % STAGES;
        mbar_wait(&full[st], (k / STAGES) & 1);
        mbar_arrive(&empty[st]);
      }
      return;
    }
  }

  float2 X[R][D];
  float2 rt2[R];  // per-row additive exponent term, duplicated
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int64_t i = row0 + local_row(r);
    if constexpr (sparse) {
      if (local_row(r) >= nrows) i = row0 + nrows - 1;
    } else {
      if (i >= N) i = N - 1;
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float c = center ? center[k] : 0.f;
      const float v = scale * (x[i * D + k] - c);
      X[r][k] = dup2(v);
      acc = fmaf(v, v, acc);
    }
    float rt = C::DIRECT ? 0.f : -0.5f * acc;
    if (lse2 != nullptr) rt -= lse2[i];
    rt2[r] = dup2(rt);
  }

  float2 A[R][NACC];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int a = 0; a < NACC; ++a) A[r][a] = dup2(0.f);

  int np_next = sparse ? (nt > 0 ? (pieces[t0].y >> 1) : 0) : (t0 + 1 == ntiles ? last_pairs : C::TJ / 2);  // softmin.cuh
  for (int k = 0; k < nt; ++k) {
    const int st = k % STAGES;
    const int npairs = np_next;
    if constexpr (sparse) {
      if (k + 1 < nt) np_next = pieces[t0 + k + 1].y >> 1;
    } else {
      if (t0 + k + 2 == ntiles) np_next = last_pairs;
    }
    mbar_wait(&full[st], (k / STAGES) & 1);
    const float4* tp = reinterpret_cast<const float4*>(tiles + st * C::TILE_FLOATS);

    // two-level accumulation (tile sums T, then running sums A): fp32 error ~sqrt(TJ/2)+sqrt(tiles) ulps
    float2 T[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int a = 0; a < NACC; ++a) T[r][a] = dup2(0.f);

#pragma unroll(kRowsumUnroll)
    for (int jp = 0; jp < npairs; ++jp) {
      float2 S[NF2];
#pragma unroll
      for (int q = 0; q < NF2 / 2; ++q) {
        const float4 v = tp[jp * (NF2 / 2) + q];
        S[2 * q] = make_float2(v.x, v.y);
        S[2 * q + 1] = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if constexpr (!C::DIRECT) {
          // expansion modes: exponent = S[D] + X.Y + rowterm
          float2 t = __fadd2_rn(S[D], rt2[r]);
#pragma unroll
          for (int d = 0; d < D; ++d) t = __ffma2_rn(X[r][d], S[d], t);
          float2 e;
          e.x = ex2_approx(t.x);
          e.y = ex2_approx(t.y);
          if constexpr (MODE == kGaussFwd || MODE == kGaussBwd) e = __fmul2_rn(e, S[D + 1]);
          T[r][0] = __fadd2_rn(T[r][0], e);
          if constexpr (MODE == kSoftminBwdP2 || MODE == kGaussBwd) {
#pragma unroll
            for (int d = 0; d < D; ++d) T[r][1 + d] = __ffma2_rn(e, S[d], T[r][1 + d]);
          }
        } else {
          float2 df[D];
          float2 qq;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            df[d] = __fadd2_rn(X[r][d], S[d]);  // packed columns hold -Y
            qq = (d == 0) ? __fmul2_rn(df[d], df[d]) : __ffma2_rn(df[d], df[d], qq);
          }
          // |X - Y| with the reference's clamp (utils.py:61); inside the clamp the gradient is zero
          const bool ina = qq.x < clampq, inb = qq.y < clampq;
          qq.x = fmaxf(qq.x, clampq);
          qq.y = fmaxf(qq.y, clampq);
          float2 rinv;
          rinv.x = rsqrt_approx(qq.x);
          rinv.y = rsqrt_approx(qq.y);
          const float2 dist = __fmul2_rn(qq, rinv);
          float2 wgt;  // weight multiplying the per-pair contribution
          if constexpr (MODE == kSoftminBwdP1) {
            const float2 t = __ffma2_rn(dist, dup2(-1.0f), __fadd2_rn(S[D], rt2[r]));
            wgt.x = ex2_approx(t.x);
            wgt.y = ex2_approx(t.y);
          } else if constexpr (MODE == kLaplaceFwd || MODE == kLaplaceBwd) {
            wgt.x = ex2_approx(-dist.x);
            wgt.y = ex2_approx(-dist.y);
            wgt = __fmul2_rn(wgt, S[D]);
          } else {
            wgt = S[D];
          }
          if constexpr (MODE == kSoftminBwdP1) T[r][0] = __fadd2_rn(T[r][0], wgt);
          if constexpr (MODE == kLaplaceFwd) T[r][0] = __fadd2_rn(T[r][0], wgt);
          if constexpr (MODE == kEnergyFwd) T[r][0] = __ffma2_rn(wgt, dist, T[r][0]);
          if constexpr (MODE == kSoftminBwdP1 || MODE == kLaplaceBwd || MODE == kEnergyBwd) {
            float2 wr = __fmul2_rn(wgt, rinv);
            if (ina) wr.x = 0.f;
            if (inb) wr.y = 0.f;
            constexpr int off = (MODE == kSoftminBwdP1) ? 1 : 0;
#pragma unroll
            for (int d = 0; d < D; ++d) T[r][off + d] = __ffma2_rn(wr, df[d], T[r][off + d]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int a = 0; a < NACC; ++a) A[r][a] = __fadd2_rn(A[r][a], T[r][a]);
    // release: each thread arrives after ITS OWN last read of the stage (an elected lane behind a __syncwarp() is as
    // correct, but compute-sanitizer's racecheck does not follow that hand-over and reports the refill as a WAR hazard)
    mbar_arrive(&empty[st]);
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t i = row0 + local_row(r);
    const bool live = sparse ? (local_row(r) < nrows) : (i < N);
    if (live) {
#pragma unroll
      for (int a = 0; a < NACC; ++a) part[((int64_t)split * N + i) * NACC + a] = A[r][a].x + A[r][a].y;
    }
  }
}

}  // namespace b200ot

// b200ot — softmin partial-reduction kernel (the N x M hot loop), templated on the problem shape.
//
// Replaces the reference's softmin_tensorized / softmin_online
// (src/geomloss/_legacy/sinkhorn_samples.py:32-71, :337-346): for every row i
//     lse2_i = log2 sum_j 2^( t_ij ),   t_ij = log2e * ( h_j - |x_i - y_j|^p / (p eps) )
// is accumulated online as a (running max m_i, sum s_i of 2^(t_ij - m_i)) pair, the cost being
// recomputed on the fly from the coordinates and never stored.
//
// Work decomposition
//   grid  = (row tiles, column splits); a CTA owns NT*R rows and a contiguous run of column tiles;
//   block = 1 producer warp + NT/32 consumer warps.  The producer streams packed column tiles
//           (see ColFmt in common.cuh) global -> shared with 1-D bulk TMA copies into a STAGES-deep
//           ring guarded by full/empty mbarriers; the column cloud (16 B/column at D=3) stays in L2.
//   consumer thread = R rows; for each column pair it reads the packet with broadcast LDS.128 and
//           evaluates two columns at once with packed f32x2 FMAs (FFMA2): at D=3 / p=2 a pair costs
//           3 FFMA2 + 1 FADD2 + 1/2 FMNMX3 + 1 MUFU.EX2 + 1/2 FADD2 per row, i.e. the loop is bound
//           by the 16-lane/SM MUFU unit, not by FP32 issue.  POLY > 0 moves a fraction of the
//           exponentials to a polynomial on the FMA pipe (ex2_poly2) to balance the two units.
//   p = 2 uses the expansion  -|X-Y|^2/2 = X.Y - |Y|^2/2 - |X|^2/2  on centred, pre-scaled
//           coordinates (the -|Y|^2/2 term lives in the packed per-column slot, -|X|^2/2 is a row
//           constant added to m at the end); DIRECT evaluates differences explicitly
//           (used for p = 1 and as the accuracy cross-check for p = 2).
#pragma once
#include "common.cuh"

namespace b200ot {

// 2^x on a pair, FMA pipe only (see ex2_poly in common.cuh for the scalar derivation).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = dup2(12582912.0f);
  float2 t = __fadd2_rn(x, magic);
  float2 n = __fadd2_rn(t, dup2(-12582912.0f));
  float2 f = __ffma2_rn(n, dup2(-1.0f), x);
  float2 p = dup2(1.3264726694e-3f);
  p = __ffma2_rn(p, f, dup2(9.6715127364e-3f));
  p = __ffma2_rn(p, f, dup2(5.5507337449e-2f));
  p = __ffma2_rn(p, f, dup2(2.4022242083e-1f));
  p = __ffma2_rn(p, f, dup2(6.9314697760e-1f));
  p = __ffma2_rn(p, f, dup2(1.0f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

template <int D_, int R_, int P_, bool DIRECT_, int POLY_, int NT_ = 256, int TJ_ = 1024, int STAGES_ = 3,
          int CH_ = 4, int MINB_ = 2>
struct SoftminCfg {
  static constexpr int D = D_;            // ambient dimension
  static constexpr int R = R_;            // rows per consumer thread
  static constexpr int P = P_;            // cost exponent (1 or 2)
  static constexpr bool DIRECT = DIRECT_; // explicit differences instead of the dot-product expansion
  static constexpr int POLY = POLY_;      // 0: all exp2 on MUFU; 1: one column pair per chunk on the FMA pipe;
                                          // 2: one pair every other chunk
  static constexpr int NT = NT_;          // consumer threads
  static constexpr int TJ = TJ_;          // columns per tile
  static constexpr int STAGES = STAGES_;
  static constexpr int CH = CH_;          // column pairs per chunk (max is refreshed once per chunk)
  static constexpr int MINB = MINB_;      // CTAs per SM the register budget is planned for
  static constexpr int NEXTRA = 1;
  static constexpr int NF2 = ((D + NEXTRA + 1) / 2) * 2;
  static constexpr int TILE_FLOATS = (TJ / 2) * NF2 * 2;
  static constexpr int TILE_BYTES = TILE_FLOATS * 4;
  static constexpr int ROWS_PER_CTA = NT * R;
  static constexpr int SMEM_BYTES = STAGES * TILE_BYTES + 2 * STAGES * 8;
  static_assert(P == 2 || DIRECT, "p = 1 needs explicit differences");
  static_assert((TJ / 2) % CH == 0, "tile must hold a whole number of chunks");
};

template <class C>
__global__ void __launch_bounds__(C::NT + 32, C::MINB)
    softmin_partial_kernel(const float* __restrict__ x, const float* __restrict__ center, float scale,
                           float clampq, const float* __restrict__ cols, float2* __restrict__ part, int64_t N,
                           int ntiles, int tiles_per_split) {
  constexpr int D = C::D, R = C::R, NT = C::NT, NF2 = C::NF2, CH = C::CH, STAGES = C::STAGES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * C::TILE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.y;
  const int t0 = split * tiles_per_split;
  const int t1 = min(ntiles, t0 + tiles_per_split);
  const int nt = t1 - t0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NT / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ===== producer: one elected lane streams the column tiles of this split =====
    if (lane == 0) {
      for (int k = 0; k < nt; ++k) {
        const int st = k % STAGES;
        if (k >= STAGES) mbar_wait(&empty[st], ((k / STAGES) + 1) & 1);
        mbar_arrive_expect_tx(&full[st], C::TILE_BYTES);
        tma_load_1d(tiles + st * C::TILE_FLOATS, cols + (int64_t)(t0 + k) * C::TILE_FLOATS, C::TILE_BYTES,
                    &full[st]);
      }
    }
    return;
  }

  // ===== consumers =====
  const int tid = threadIdx.x - 32;
  const int64_t row_base = (int64_t)blockIdx.x * C::ROWS_PER_CTA + tid;

  float2 X[R][D];
  float rowc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int64_t i = row_base + (int64_t)r * NT;
    if (i >= N) i = N - 1;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float c = center ? center[k] : 0.f;
      const float v = scale * (x[i * D + k] - c);
      X[r][k] = dup2(v);
      acc = fmaf(v, v, acc);
    }
    rowc[r] = C::DIRECT ? 0.f : -0.5f * acc;
  }

  float m[R];
  float2 nm2[R], s2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m[r] = kNegBig;
    nm2[r] = dup2(-kNegBig);
    s2[r] = dup2(0.f);
  }

  for (int k = 0; k < nt; ++k) {
    const int st = k % STAGES;
    mbar_wait(&full[st], (k / STAGES) & 1);
    const float4* tp = reinterpret_cast<const float4*>(tiles + st * C::TILE_FLOATS);

#pragma unroll 1
    for (int jp = 0; jp < C::TJ / 2; jp += CH) {
      float2 T[R][CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float2 S[NF2];
#pragma unroll
        for (int q = 0; q < NF2 / 2; ++q) {
          const float4 v = tp[(jp + c) * (NF2 / 2) + q];
          S[2 * q] = make_float2(v.x, v.y);
          S[2 * q + 1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float2 t;
          if constexpr (!C::DIRECT) {
            t = S[D];
#pragma unroll
            for (int d = 0; d < D; ++d) t = __ffma2_rn(X[r][d], S[d], t);
          } else {
            float2 qq;
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const float2 df = __fadd2_rn(X[r][d], S[d]);  // packed columns hold -Y
              qq = (d == 0) ? __fmul2_rn(df, df) : __ffma2_rn(df, df, qq);
            }
            if constexpr (C::P == 2) {
              t = __ffma2_rn(qq, dup2(-0.5f), S[D]);
            } else {
              qq.x = fmaxf(qq.x, clampq);
              qq.y = fmaxf(qq.y, clampq);
              float2 dist;
              dist.x = sqrt_approx(qq.x);
              dist.y = sqrt_approx(qq.y);
              t = __ffma2_rn(dist, dup2(-1.0f), S[D]);
            }
          }
          T[r][c] = t;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float cm = fmaxf(T[r][0].x, T[r][0].y);
#pragma unroll
        for (int c = 1; c < CH; ++c) cm = fmax3(cm, T[r][c].x, T[r][c].y);
        if (cm > m[r]) {
          const float sc = ex2_approx(m[r] - cm);
          s2[r] = __fmul2_rn(s2[r], dup2(sc));
          m[r] = cm;
          nm2[r] = dup2(-cm);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float2 a = __fadd2_rn(T[r][c], nm2[r]);
          float2 e;
          bool poly = false;
          if constexpr (C::POLY == 1) poly = (c == 0);
          if constexpr (C::POLY == 2) poly = (c == 0) && ((jp / CH) & 1);
          if (poly) {
            e = ex2_poly2(a);
          } else {
            e.x = ex2_approx(a.x);
            e.y = ex2_approx(a.y);
          }
          s2[r] = __fadd2_rn(s2[r], e);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t i = row_base + (int64_t)r * NT;
    if (i < N) part[(int64_t)split * N + i] = make_float2(m[r] + rowc[r], s2[r].x + s2[r].y);
  }
}

}  // namespace b200ot

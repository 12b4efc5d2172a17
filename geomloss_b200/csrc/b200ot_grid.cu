// b200ot — separable soft-C-transform on regular grids (images / volumes).
// Reference semantics: softmin_grid (src/geomloss/_legacy/utils.py:190-279), called by the image Sinkhorn
// loop (src/geomloss/_legacy/sinkhorn_images.py:26-202): for a (batch, N, N[, N]) array h and the cost
// |x - y|^p / p on the pixel grid x = arange(N)/N, the softmin over the whole grid factorises into one
// 1-D log-sum-exp per axis,
//     out[.., i, ..] = log sum_j exp( in[.., j, ..] - k(x_i - x_j) ),   k(d) = d^2/(2 eps)  (p = 2),  |d|/eps  (p = 1)
// applied along every axis in turn, followed by a multiplication by -eps.
//
// One pass = one launch of grid_pass_kernel over the tensor viewed as [outer][N][inner]: a CTA loads a
// tile of 32 lines (all N entries along the axis) into shared memory with coalesced accesses — along the
// contiguous direction for the strided passes, transposed for the last axis — then every thread owns 8
// outputs of one line at a time and sweeps the N inputs (conflict-free LDS, lanes = lines): cost terms from
// exact integer offsets on the grid (a 14-entry packed table per 8 x 8 chunk), packed fp32 adds, lazy-max
// log-sum-exp in the log2 domain with the same 2^64 sum guard as softmin.cuh, 1 MUFU.EX2 per (output, input) pair.  Results are staged in a second shared tile
// and written back with the access pattern of the load, so passes run in place.
// Work per pass: N^(dim+1) pairs -> SFU-bound like the point-cloud softmin (256^3: 4.3e9 pairs, ~1 ms);
// HBM traffic 2 x 4 N^dim bytes per pass.
#include <math.h>

#include "b200ot.h"
#include "common.cuh"
#include "host_util.cuh"

namespace b200ot {

constexpr int kGridTW = 32;  // lines per CTA tile (one per lane); 16 for N > 880, where two 32-line tiles of
                             // N x 33 floats no longer fit the 227 KB of shared memory (two lanes then share a
                             // line and split its outputs)
constexpr int kGridR = 8;  // outputs per thread per sweep
constexpr int kGridWarps = 8;

struct GridPassArgs {
  const float* h_a;      // first pass: input = in_scale * (h_a + h_scale_b * h_b); later passes read `out` in place
  const float* h_b;      // nullable
  const float* out_old;  // last pass only, nullable
  float* out;
  float h_scale_b, in_scale, alpha_old, beta, out_scale, xscale;
  int N, first, last;
  int64_t outer, inner;
};

template <int P, int TW>
__global__ void __launch_bounds__(kGridWarps * 32) grid_pass_kernel(GridPassArgs A) {
  constexpr int kGridTW = TW, kGridRS = TW + 1, NSUB = 32 / TW;
  extern __shared__ float smem_f[];
  float* tile = smem_f;                      // [N][33] inputs
  float* otile = smem_f + A.N * kGridRS;     // [N][33] outputs
  const int N = A.N;
  const int lane = (threadIdx.x & 31) % TW, warp = (threadIdx.x >> 5) * NSUB + (threadIdx.x & 31) / TW;
  const bool last_axis = (A.inner == 1);
  int64_t o, w0;
  int nw;  // live lines in this tile
  if (last_axis) {
    o = (int64_t)blockIdx.x * kGridTW;
    w0 = 0;
    nw = (int)min((int64_t)kGridTW, A.outer - o);
  } else {
    const int64_t tiles_in = (A.inner + kGridTW - 1) / kGridTW;
    o = blockIdx.x / tiles_in;
    w0 = (blockIdx.x % tiles_in) * kGridTW;
    nw = (int)min((int64_t)kGridTW, A.inner - w0);
  }

  // ---- load ----
  for (int idx = threadIdx.x; idx < N * kGridTW; idx += blockDim.x) {
    int j, w;
    if (last_axis) {
      w = idx / N;
      j = idx % N;
    } else {
      j = idx / kGridTW;
      w = idx % kGridTW;
    }
    float v = -INFINITY;
    if (w < nw) {
      const int64_t gi = last_axis ? ((o + w) * N + j) : ((o * N + j) * A.inner + w0 + w);
      if (A.first) {
        v = A.h_a[gi];
        if (A.h_b) v = fmaf(A.h_scale_b, A.h_b[gi], v);
        v *= A.in_scale;
      } else {
        v = A.out[gi];
      }
    }
    tile[j * kGridRS + w] = v;
  }
  __syncthreads();

  // ---- sweep: warp g owns outputs [g*per_warp, (g+1)*per_warp) of every line of the tile ----
  // A thread owns kGridR = 8 consecutive outputs of its line and sweeps the inputs 8 at a time.  The cost term of the
  // pair (output ib + r, input j0 + c) depends on n = (ib - j0) + (r - c) only: 14 packed table entries
  //     W[k] = ( -k(n), -k(n - 1) ),  n = (ib - j0) + k,  k = r - 2q in [-6, 7]        k(n) = xs^2 n^2  or  xs |n|
  // per chunk (n and n^2 are exact small integers in fp32: ONE rounding per entry) serve its 64 pairs, which are then
  // three packed instructions per two pairs — t = a + W, t - m, sum += 2^(t - m) — instead of four scalar ones per
  // pair: the loop is bound by the 64 MUFU.EX2, not by issue slots (ncu before: XU 77 %, issue 68 %).
  const float kq = (P == 2) ? -A.xscale * A.xscale : -A.xscale;
  const int per_warp = (N + kGridWarps * NSUB - 1) / (kGridWarps * NSUB);
  const int i_end = min(N, (warp + 1) * per_warp);
  for (int ib = warp * per_warp; ib < i_end; ib += kGridR) {
    float m[kGridR], s[kGridR];
#pragma unroll
    for (int r = 0; r < kGridR; ++r) {
      // reference exponent = the j = i term (a lower bound of the row max, usually within a few units of it): a
      // sweep that starts from -big climbs the whole Gaussian flank and re-bases at EVERY chunk left of i
      // (measured: 1.58 XU ops per pair instead of 1)
      m[r] = (ib + r < N) ? fmaxf(tile[(ib + r) * kGridRS + lane], kNegBig) : kNegBig;
      s[r] = 0.f;
    }
    for (int j0 = 0; j0 < N; j0 += 8) {
      float2 a2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a2[q].x = (j0 + 2 * q < N) ? tile[(j0 + 2 * q) * kGridRS + lane] : -INFINITY;
        a2[q].y = (j0 + 2 * q + 1 < N) ? tile[(j0 + 2 * q + 1) * kGridRS + lane] : -INFINITY;
      }
      const float2 base = dup2((float)(ib - j0));
      float2 W[14];
#pragma unroll
      for (int k = -6; k <= 7; ++k) {
        const float2 n = __fadd2_rn(base, make_float2((float)k, (float)(k - 1)));
        if (P == 2) {
          W[k + 6] = __fmul2_rn(__fmul2_rn(n, n), dup2(kq));
        } else {
          W[k + 6] = __fmul2_rn(make_float2(fabsf(n.x), fabsf(n.y)), dup2(kq));
        }
      }
#pragma unroll
      for (int r = 0; r < kGridR; ++r) {
        const float2 nm = dup2(-m[r]);
        float2 cs2 = dup2(0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 e = __fadd2_rn(__fadd2_rn(a2[q], W[r - 2 * q + 6]), nm);
          cs2 = __fadd2_rn(cs2, make_float2(ex2_approx(e.x), ex2_approx(e.y)));
        }
        float cs = cs2.x + cs2.y;
        if (!(cs <= 1.8446744e19f)) {
          // outdated max (or first chunk): rebase on this chunk's max and redo it
          float t[8];
          float cm = -INFINITY;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float n = (float)(ib + r - j0 - c);
            const float av = (c & 1) ? a2[c >> 1].y : a2[c >> 1].x;
            t[c] = av + kq * ((P == 2) ? n * n : fabsf(n));
            cm = fmaxf(cm, t[c]);
          }
          cs = 0.f;
          if (cm > kNegBig) {  // otherwise every input of the chunk is -inf: nothing to add
            s[r] *= ex2_approx(m[r] - cm);
            m[r] = cm;
#pragma unroll
            for (int c = 0; c < 8; ++c) cs += ex2_approx(t[c] - cm);
          }
        }
        s[r] += cs;
      }
    }
#pragma unroll
    for (int r = 0; r < kGridR; ++r) {
      const int i = ib + r;
      if (i < i_end) {
        int e = 0;
        const float f = frexpf(s[r], &e);  // log2(s) = e + log2(f), f in [0.5, 1): exact exponent, accurate mantissa
        otile[i * kGridRS + lane] = (s[r] > 0.f) ? (m[r] + (float)e) + log2f(f) : -INFINITY;
      }
    }
  }
  __syncthreads();

  // ---- store (same access pattern as the load) ----
  for (int idx = threadIdx.x; idx < N * kGridTW; idx += blockDim.x) {
    int j, w;
    if (last_axis) {
      w = idx / N;
      j = idx % N;
    } else {
      j = idx / kGridTW;
      w = idx % kGridTW;
    }
    if (w >= nw) continue;
    const int64_t gi = last_axis ? ((o + w) * N + j) : ((o * N + j) * A.inner + w0 + w);
    float v = otile[j * kGridRS + w];
    if (A.last) {
      v = A.beta * (A.out_scale * v);
      if (A.out_old) v = fmaf(A.alpha_old, A.out_old[gi], v);
    }
    A.out[gi] = v;
  }
}

}  // namespace b200ot

using namespace b200ot;

extern "C" {

B200OT_API int b200ot_softmin_grid(const float* h_a, const float* h_b, float h_scale_b, const float* out_old,
                                   float alpha_old, float beta, float* out, int64_t batch, int32_t N, int32_t dim,
                                   int32_t p, float eps, void* stream) {
  if (!h_a || !out || batch <= 0 || N <= 0 || N > 1024 || dim < 1 || dim > 3 || (p != 1 && p != 2) || !(eps > 0.f))
    return B200OT_EINVAL;
  if (out == h_a || out == h_b || (out_old && out == out_old)) return B200OT_EINVAL;  // passes run in place on `out`
  cudaStream_t st = (cudaStream_t)stream;
  int64_t total = batch;
  for (int d = 0; d < dim; ++d) total *= N;
  // pixel coordinates x = arange(N)/N, scaled so that the log2-domain exponent is in - (X_i - X_j)^2  (p = 2)
  // or in - |X_i - X_j|  (p = 1)                                                          (utils.py:235-242)
  const float xscale = (p == 2 ? sqrtf(kLog2e / (2.0f * eps)) : kLog2e / eps) / (float)N;
  const int tw = ((size_t)2 * N * (kGridTW + 1) * sizeof(float) <= 227 * 1024) ? kGridTW : 16;
  const size_t smem = (size_t)2 * N * (tw + 1) * sizeof(float);
  auto kern = (tw == kGridTW) ? ((p == 2) ? grid_pass_kernel<2, kGridTW> : grid_pass_kernel<1, kGridTW>)
                              : ((p == 2) ? grid_pass_kernel<2, 16> : grid_pass_kernel<1, 16>);
  B200OT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // pass order of the reference: last axis first, then the others (the passes commute mathematically)
  for (int k = 0; k < dim; ++k) {
    const int axis = dim - 1 - k;  // 0 .. dim-1 within the (N, .., N) block
    GridPassArgs a;
    a.h_a = h_a;
    a.h_b = h_b;
    a.out_old = out_old;
    a.out = out;
    a.h_scale_b = h_scale_b;
    a.in_scale = kLog2e;
    a.alpha_old = alpha_old;
    a.beta = beta;
    a.out_scale = -eps * kLn2;
    a.xscale = xscale;
    a.N = N;
    a.first = (k == 0);
    a.last = (k == dim - 1);
    int64_t inner = 1;
    for (int d = axis + 1; d < dim; ++d) inner *= N;
    a.inner = inner;
    a.outer = total / ((int64_t)N * inner);
    const int64_t blocks = (inner == 1) ? ceil_div64(a.outer, tw) : a.outer * ceil_div64(inner, tw);
    kern<<<(unsigned)blocks, kGridWarps * 32, smem, st>>>(a);
    B200OT_CUDA_TRY(cudaGetLastError());
  }
  return B200OT_OK;
}

}  // extern "C"

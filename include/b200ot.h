/* b200ot.h — C ABI of the B200-native Sinkhorn / kernel-MMD reduction engine.
 *
 * This is the drop-in boundary for the hot path of jeanfeydy/geomloss (reference commit 00e493f).
 * The reference has no FFI: its plugin seams are the Python callables
 *   softmin(eps, C_xy, h_y) -> f_x            (src/geomloss/_legacy/sinkhorn_divergence.py:291-303,
 *                                              implemented by softmin_tensorized / softmin_online,
 *                                              src/geomloss/_legacy/sinkhorn_samples.py:32-71,337-346)
 *   kernel matvecs of kernel_loss             (src/geomloss/_legacy/kernel_samples.py:92-146)
 * which the pykeops "online" backend lowers to one CUDA map-reduce launch per call
 * (sinkhorn_samples.py:322-334: generic_logsumexp("(B - (P * cost))", Vi(1), Vi(D), Vj(D), Vj(1), Pm(1))).
 * The functions below are what a maintainer would bind in place of that pykeops call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (or the stated type), row-major, contiguous;
 *     the caller owns every buffer; nothing is allocated, freed or retained by the library;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and the call returns
 *     immediately; no host synchronisation happens inside;
 *   - return value: 0 on success, a negative B200OT_E* code otherwise (never throws, never aborts);
 *   - "rows" x:(N,D) own the output, "cols" y:(M,D) are reduced over; D <= B200OT_MAX_D;
 *   - a reduction is split into two launches so that the column cloud can be packed once and
 *     reused (and so that partial results of column shards living on other GPUs can be merged):
 *         pack  ->  partial reduction  ->  finalize.
 */
#ifndef B200OT_H_
#define B200OT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200OT_API __attribute__((visibility("default")))
#else
#define B200OT_API
#endif

#define B200OT_VERSION 100 /* 0.1.0 */
#define B200OT_MAX_D 8     /* dimensions served by the CUDA-core (register tile) kernels of this build */

/* error codes */
#define B200OT_OK 0
#define B200OT_EINVAL -1   /* bad argument (null pointer, non-positive size, unsupported p / D / kind) */
#define B200OT_ESCRATCH -2 /* scratch buffer too small (see *_scratch_bytes) */
#define B200OT_ECUDA -3    /* a CUDA runtime call failed; see b200ot_last_cuda_error() */
#define B200OT_EALIGN -4   /* pointer not 16-byte aligned */

/* kernel kinds for b200ot_kernel_conv_* (reference: kernel_samples.py:62-82) */
#define B200OT_KERNEL_GAUSSIAN 0  /* exp(-|x-y|^2 / (2 blur^2)) */
#define B200OT_KERNEL_LAPLACIAN 1 /* exp(-sqrt(max(|x-y|^2/blur^2, 1e-8))) */
#define B200OT_KERNEL_ENERGY 2    /* -sqrt(max(|x-y|^2, 1e-8)) */

B200OT_API const char* b200ot_strerror(int code);
B200OT_API const char* b200ot_last_cuda_error(void);
B200OT_API int b200ot_version(void);

/* ---------------------------------------------------------------------------------------------
 * Softmin  —  out_i = -eps * log sum_j exp( h_j - |x_i - y_j|^p / (p * eps) ),   p in {1, 2}
 * replaces softmin_tensorized (sinkhorn_samples.py:32-71) / softmin_online (:337-346).
 * p = 1 uses sqrt(max(|x-y|^2, 1e-8)) like the reference's `distances` (utils.py:56-61).
 * ------------------------------------------------------------------------------------------- */

/* Bytes of scratch needed by b200ot_softmin_fwd / _bwd_x for (N rows, M cols, D). */
B200OT_API int64_t b200ot_softmin_scratch_bytes(int64_t N, int64_t M, int32_t D);

/* One-call softmin:  out (N) <- alpha_old * out_old + beta * softmin(eps, (x, y), h)
 *   h_a (M), h_b (M, nullable):  h_j = h_a[j] + h_scale_b * h_b[j]
 *       (the Sinkhorn loop's  b_log + g / eps  is h_a = b_log, h_b = g, h_scale_b = 1/eps —
 *        sinkhorn_divergence.py:480-488 — fused here instead of a separate elementwise launch);
 *   out_old (N, nullable), alpha_old, beta:  the damping and the 1/2 (f + ft) averaging of
 *       sinkhorn_divergence.py:480-493 fused into the epilogue (alpha_old=0, beta=1: plain softmin);
 *   lse2_out (N, nullable): log2-domain log-sum-exp per row, saved for the backward pass;
 *   center (D, nullable): device vector subtracted from both clouds before the |x|^2 - 2x.y + |y|^2
 *       expansion (improves fp32 conditioning; any point near the data works, zeros if null);
 *   scratch: >= b200ot_softmin_scratch_bytes(N, M, D) bytes, 16-byte aligned.
 */
B200OT_API int b200ot_softmin_fwd(const float* x, const float* y, const float* h_a, const float* h_b, float h_scale_b,
                       const float* center, const float* out_old, float alpha_old, float beta, float* out,
                       float* lse2_out, int64_t N, int64_t M, int32_t D, int32_t p, float eps, void* scratch,
                       int64_t scratch_bytes, void* stream);

/* Gradient of the softmin w.r.t. the row cloud (the only gradient the reference's autograd
 * contract carries: columns and h are detached, sinkhorn_samples.py:179-185, sinkhorn_divergence.py:616-623):
 *   grad_x[i,:] = grad_out[i] * sum_j softmax_j(h_j - C_ij/eps) * dC(x_i, y_j)/dx_i
 * lse2 is the lse2_out of the matching forward call.  grad_x (N,D) is overwritten. */
B200OT_API int b200ot_softmin_bwd_x(const float* x, const float* y, const float* h_a, const float* h_b, float h_scale_b,
                         const float* center, const float* lse2, const float* grad_out, float* grad_x, int64_t N,
                         int64_t M, int32_t D, int32_t p, float eps, void* scratch, int64_t scratch_bytes,
                         void* stream);

/* --- the three stages of b200ot_softmin_fwd, exposed for column-sharded (multi-GPU) use --- */

/* Number of floats of a packed column buffer for M columns of dimension D with `extra` per-column
 * scalars (softmin: 1 = h; gaussian conv: 2 = -|y|^2/2 and weight; laplacian / energy conv: 1 = weight). */
B200OT_API int64_t b200ot_packed_cols_floats(int64_t M, int32_t D, int32_t extra);

/* Pack the column cloud for a softmin at temperature eps (coordinates pre-scaled, h folded into the
 * per-column term, padding columns neutral).  cols_out: b200ot_packed_cols_floats(M, D, 1) floats. */
B200OT_API int b200ot_softmin_pack(const float* y, const float* h_a, const float* h_b, float h_scale_b, const float* center,
                        int64_t M, int32_t D, int32_t p, float eps, float* cols_out, void* stream);

/* Number of column splits the partial reduction will use for (N, M, D): partials hold n_split * N pairs. */
B200OT_API int32_t b200ot_softmin_num_splits(int64_t N, int64_t M, int32_t D);

/* Partial reduction over the packed columns: part[(s*N + i)*2 + {0,1}] = (running max m, sum of
 * exp2(t - m)) in the log2 domain, for s < n_split.  This is the kernel that does the N x M work. */
B200OT_API int b200ot_softmin_partial(const float* x, const float* center, const float* cols, float* part, int32_t n_split,
                           int64_t N, int64_t M, int32_t D, int32_t p, float eps, void* stream);

/* Block-sparse partial reduction (the reference's softmin_multiscale with `ranges`, sinkhorn_samples.py:445-450,
 * built by kernel_truncation :493-530): row tile r (b200ot_sparse_tile_shape rows of x, in order) reduces only over
 * the packed column tiles tile_list[tile_ptr[r] .. tile_ptr[r+1]) (each b200ot_sparse_tile_shape columns).
 * part: (N, 2) pairs (one split).  D <= 3. */
B200OT_API void b200ot_sparse_tile_shape(int32_t* rows_per_tile, int32_t* cols_per_tile);
B200OT_API int b200ot_softmin_partial_sparse(const float* x, const float* center, const float* cols,
                                             const int32_t* tile_ptr, const int32_t* tile_list, float* part,
                                             int64_t N, int64_t M, int32_t D, int32_t p, float eps, void* stream);

/* Collapse n_part partial (m, s) sets into one per row: merged[i*2 + {0,1}].  A rank calls this on its own
 * splits before exchanging partials with the other column shards (N*8 bytes per rank). */
B200OT_API int b200ot_softmin_merge(const float* part, int32_t n_part, float* merged, int64_t N, void* stream);

/* Merge n_part partial (m, s) sets (own splits and/or other ranks' shards, concatenated along the
 * leading axis) and apply the epilogue of b200ot_softmin_fwd. */
B200OT_API int b200ot_softmin_finalize(const float* part, int32_t n_part, const float* out_old, float alpha_old, float beta,
                            float* out, float* lse2_out, int64_t N, float eps, void* stream);

/* --- stages of b200ot_softmin_bwd_x, for column-sharded use: per-shard partial sums
 *     part[(s*N + i)*(D+1) + {0: sum_j w_ij, 1+k: sum_j w_ij * (coordinate or unit-vector term)}],
 *     which add across shards (all-reduce SUM) before the finalize --- */
B200OT_API int b200ot_softmin_bwd_partial(const float* x, const float* center, const float* cols, const float* lse2,
                                          float* part, int32_t n_split, int64_t N, int64_t M, int32_t D, int32_t p,
                                          float eps, void* stream);
/* block-sparse variant (see b200ot_softmin_partial_sparse): part is (N, D+1), one split */
B200OT_API int b200ot_softmin_bwd_partial_sparse(const float* x, const float* center, const float* cols,
                                                 const float* lse2, const int32_t* tile_ptr,
                                                 const int32_t* tile_list, float* part, int64_t N, int64_t M,
                                                 int32_t D, int32_t p, float eps, void* stream);
/* merged[i*width + a] = sum_s part[(s*N + i)*width + a] */
B200OT_API int b200ot_rowsum_merge(const float* part, int32_t n_part, int32_t width, float* merged, int64_t N,
                                   void* stream);
B200OT_API int b200ot_softmin_bwd_finalize(const float* part, int32_t n_part, const float* x, const float* center,
                                           const float* grad_out, float* grad_x, int64_t N, int32_t D, int32_t p,
                                           float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel convolution  —  out_i = sum_j k(x_i, y_j) * w_j      (kernel_samples.py:128-137)
 * ------------------------------------------------------------------------------------------- */
/* Gaussian forward additionally accepts 8 < D <= 64: that range runs on the tensor cores (tcgen05, bf16x3
 * split operands, exponent accumulated in TMEM); other kinds / the gradients are limited to D <= B200OT_MAX_D. */
B200OT_API int64_t b200ot_kernel_conv_scratch_bytes(int64_t N, int64_t M, int32_t D);

B200OT_API int b200ot_kernel_conv_fwd(const float* x, const float* y, const float* w, const float* center, float* out,
                           int64_t N, int64_t M, int32_t D, int32_t kind, float blur, void* scratch,
                           int64_t scratch_bytes, void* stream);

/* grad_x[i,:] = grad_out[i] * sum_j w_j * d k(x_i, y_j) / d x_i   (rows only; columns via a swapped call) */
B200OT_API int b200ot_kernel_conv_bwd_x(const float* x, const float* y, const float* w, const float* center,
                             const float* grad_out, float* grad_x, int64_t N, int64_t M, int32_t D, int32_t kind,
                             float blur, void* scratch, int64_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Grid softmin  —  the separable soft-C-transform on (batch, N, N[, N]) images / volumes
 * replaces softmin_grid (src/geomloss/_legacy/utils.py:190-279), the operator of the image Sinkhorn loop
 * (src/geomloss/_legacy/sinkhorn_images.py:26-202):
 *   out <- alpha_old * out_old + beta * ( -eps * LSE over the whole grid of  h - |x - y|^p / (p eps) ),
 *   h = h_a + h_scale_b * h_b (h_b nullable), pixel coordinates arange(N)/N, p in {1, 2}, dim in {1, 2, 3}.
 * `out` must not alias an input (the per-axis passes run in place on it).  N <= 1024.
 * ------------------------------------------------------------------------------------------- */
B200OT_API int b200ot_softmin_grid(const float* h_a, const float* h_b, float h_scale_b, const float* out_old,
                                   float alpha_old, float beta, float* out, int64_t batch, int32_t N, int32_t dim,
                                   int32_t p, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pipe-ceiling micro-benchmarks (used by bench.py to measure the MUFU / FP32 roofline of the
 * device it runs on; each launches one kernel doing `iters` dependent steps per thread and
 * returns the number of operations executed per thread-step through *ops_per_thread_iter).
 * ------------------------------------------------------------------------------------------- */
#define B200OT_UBENCH_MUFU_EX2 0
#define B200OT_UBENCH_FFMA 1
#define B200OT_UBENCH_FFMA2 2
B200OT_API int b200ot_ubench(int32_t which, int32_t iters, int32_t blocks, float* sink, int32_t* ops_per_thread_iter,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200OT_H_ */

/*
 * decomp_b200 C ABI -- the drop-in boundary of the B200-native deComP hot path.
 *
 * Plain C: raw DEVICE pointers, sizes, leading dimensions (in elements of double) and a
 * cudaStream_t passed as void*.  No torch types, nothing allocated or freed on behalf of
 * the caller: workspaces are sized by the *_workspace_bytes() queries and supplied by the
 * caller.  Every function returns 0 on success and a negative code on failure; the text
 * is available from decomp_last_error().
 *
 * The reference (fujii-team/deComP) is pure Python over an `xp` array namespace; what a
 * maintainer would bind with ctypes are the array expressions inside its update rules.
 * Each entry point below names the reference lines it replaces (paths relative to the
 * reference root).  All matrices are row-major.  Complex data is passed as interleaved
 * (re, im) doubles, i.e. a complex [r, c] matrix is the real [r, 2c] matrix with the same
 * bytes; the real kernels below then compute complex products exactly through the 2x2
 * real block embedding prepared by decomp_make_rhs_f64().
 *
 * Alignment contract (needed by TMA): base pointers of GEMM operands are 16-byte aligned
 * and their leading dimensions are even.  The Python host pads to satisfy this.
 */
#ifndef DECOMP_B200_H_
#define DECOMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DECOMP_OK 0
#define DECOMP_ERR_INVALID (-1)
#define DECOMP_ERR_CUDA (-2)
#define DECOMP_ERR_UNSUPPORTED (-3)

#define DECOMP_ABI_VERSION 2

/* ---- epilogue fused into the NT GEMM ------------------------------------------------ */
enum decomp_epilogue_kind {
  DECOMP_EPI_STORE = 0,     /* out = acc                                                     */
  DECOMP_EPI_STORE_MASK = 1,/* out = acc * mask[row, col / cwidth]      (grads.py:112,122; lasso.py:260) */
  DECOMP_EPI_MU_NUM = 2,    /* out = x * max(acc,0) / max(other,eps)    (grads.py:84,93: acc is the positive part) */
  DECOMP_EPI_MU_DEN = 3,    /* out = x * max(other,0) / max(acc,eps)    (acc is the negative part) */
  DECOMP_EPI_PROX = 4,      /* ISTA/FISTA step, see decomp_epilogue_t   (lasso.py:244-271, 405-414) */
  DECOMP_EPI_KL_RATIO = 5,  /* out = other / (acc + eps) [* mask]       (grads.py:142-160) */
  DECOMP_EPI_PROXQ = 6      /* unmasked ISTA/FISTA step with the gradient step folded into the GEMM operand:
                               B = (I - G/L)^T, other = (y A^H)/L  ->  z = other + acc = w + (yAh - w G)/L, then as
                               PROX.  Needs flags = DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD or `step`; `x` is unused;
                               `other` and `prev` follow the GEMM operand alignment contract (TMA). */
};

#define DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD 1 /* colvec holds step * alpha (lasso.py:287) ready-made */

enum decomp_shrink_kind {
  DECOMP_SHRINK_REAL = 0,    /* lasso.py:192-207 */
  DECOMP_SHRINK_COMPLEX = 1, /* lasso.py:210-225, column pairs are (re, im) */
  DECOMP_SHRINK_POSITIVE = 2 /* lasso.py:228-241 */
};

typedef struct decomp_epilogue {
  int32_t kind;              /* decomp_epilogue_kind */
  int32_t shrink;            /* decomp_shrink_kind (PROX only) */
  int32_t cwidth;            /* 1 real, 2 complex: mask / per-column vectors are indexed by col / cwidth */
  int32_t check;             /* PROX: 1 -> evaluate the convergence test of lasso.py:293/409 in this launch */
  double* out;               /* primary output [M, N] */
  int64_t ldo;
  double* out2;              /* PROX: extrapolated point w_next (may be NULL for ISTA) */
  int64_t ldo2;
  const double* x;           /* MU: factor being updated; PROX: the point the gradient is taken at (w) */
  int64_t ldx;
  const double* other;       /* MU: the other gradient part; PROX: yAt; KL: y */
  int64_t ldother;
  const double* prev;        /* PROX: previous iterate x_prev */
  int64_t ldprev;
  const double* mask;        /* STORE_MASK / KL_RATIO: mask [M, N / cwidth] (may be NULL for KL) */
  int64_t ldmask;
  const double* colvec;      /* PROX: alpha_k per column [N / cwidth] (flags bit 0: already step * alpha_k);
                                real data: 16-byte aligned and readable up to an even element count */
  const double* colvec2;     /* PROX: tolerance per column  tol * s_k  [N / cwidth], same layout rule */
  const double* rowvec;      /* PROX (full mask): per-row factor sum_j mask[row, j]; NULL -> 1 */
  const double* step;        /* PROX: device scalar 1 / L */
  double momentum;           /* PROX: (beta - 1) / beta_next, or i / (i + 3), or 0 */
  int32_t* latch;            /* PROX+check: set to `latch_value` by the last CTA if converged */
  int32_t* scratch;          /* PROX+check: two int32 (violation flag, CTA ticket), zero-initialised */
  int32_t latch_value;
  int32_t flags;             /* DECOMP_EPI_FLAG_* */
} decomp_epilogue_t;

const char* decomp_last_error(void);
int decomp_abi_version(void);
/* Measurement aid (bench.py only): issue rate of the FP64 tensor instruction (DMMA.8x8x4 from registers, no
 * memory traffic) on the current device in TFLOP/s -- the roofline denominator of the FP64 GEMM kernels.
 * Synchronises the device. */
int decomp_probe_dmma_tflops(double* tflops_out);

/* acc[m][n] = sum_k A[m*lda + k] * B[n*ldb + k]   (A: [M,K], B: [N,K], both K-contiguous),
 * followed by the fused epilogue.  TMA-fed, DMMA.8x8x4 mainloop.
 * Replaces: y.dot(d.T), f.dot(d.T), x.dot(d) (grads.py:108-125), xp.dot(A, At), xp.tensordot(y, At),
 * xp.tensordot(x0, AAt) (lasso.py:245,285-289) together with the elementwise expressions named above.
 * If `skip_if` is non-NULL and *skip_if != 0 on the device the launch is a no-op (convergence latch). */
int decomp_gemm_nt_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N,
                       int64_t K, const decomp_epilogue_t* epi, const int32_t* skip_if, void* stream);

/* Back-to-back GEMM of the masked updates without the [M, F] intermediate:
 *   acc[m][n] = sum_j ((sum_k W[m][k] R[j][k]) * mask[m][j / cwidth]) * R[j][n]        W [M, K1], R [F, K1]
 * followed by the epilogue `epi` (DECOMP_EPI_STORE or DECOMP_EPI_PROX; epi->mask is the [M, F / cwidth] mask).
 *   masked ISTA / FISTA  ((w A) * M) A^H   lasso.py:259-271   W = w, R = decomp_make_rhs(A, conj_transpose = 0)
 *   masked NMF           ((x D) * M) D^T   grads.py:112-115   W = x, R = D^T
 * (for complex data the real embedding of A^H is the transpose of that of A, so one operand serves both products).
 * K1 must be 32, 64 or 128 (decomp_gemm_b2b_masked_supported); other widths use two decomp_gemm_nt_f64 launches. */
int decomp_gemm_b2b_masked_supported(int64_t K1);
int decomp_gemm_b2b_masked_f64(const double* W, int64_t ldw, const double* R, int64_t ldr, int64_t M, int64_t K1,
                               int64_t F, const decomp_epilogue_t* epi, const int32_t* skip_if, void* stream);

/* acc[m][n] = sum_k A[k*lda + m] * B[k*ldb + n]   (A: [K,M], B: [K,N]; contraction over rows = samples),
 * split along K across CTAs with a deterministic two-stage reduction through `workspace`.
 *   combine = 0: out = acc             (x.T.dot(y), x.T.dot(f), grads.py:120-125)
 *   combine = 1: out = beta*out + acc  (A = beta*A + xT.x, B = beta*B + xT.y, dictionary_learning.py:151-152)
 *   combine = 2/3: as 0/1 but A,B are interleaved complex and out[i][j] = sum_k conj(a_ki) b_kj is written as
 *                  interleaved complex [M/2, N/2] (ldo in doubles)       (dictionary_learning.py:147-152) */
size_t decomp_gemm_tn_workspace_bytes(int64_t M, int64_t N, int64_t K);
int decomp_gemm_tn_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N,
                       int64_t K, double* out, int64_t ldo, int32_t combine, double beta, void* workspace,
                       size_t workspace_bytes, const int32_t* skip_if, void* stream);

/* Right-hand operand preparation for X . S or X . S^H with a small matrix S [p, q] (complex: interleaved):
 * writes the [N, K] K-contiguous operand B of decomp_gemm_nt_f64 such that X_real . B^T equals the product.
 *   conj_transpose = 0: product X . S    -> B is [q*cw, p*cw]
 *   conj_transpose = 1: product X . S^H  -> B is [p*cw, q*cw]  (for real data B == S, copied) */
int decomp_make_rhs_f64(const double* S, int64_t lds, int64_t p, int64_t q, int32_t is_complex,
                        int32_t conj_transpose, double* B, int64_t ldb, const int32_t* skip_if, void* stream);

/* ---- vector / reduction kernels ------------------------------------------------------- */
/* out[i] = sqrt(sum_j |A[i][j]|^2)                                  (lasso.py:124-127; normalize.py:17-20) */
int decomp_row_norms_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, int32_t is_complex,
                         double* out, void* stream);
/* out[i][j] = A[i][j] * (rowscale ? (invert_row ? 1/rowscale[i] : rowscale[i]) : 1)
 *                     * (colscale ? (invert_col ? 1/colscale[j/cw] : colscale[j/cw]) : 1)
 * (A / s[:, None], x * s, x / s, y * mask1d, A * mean(mask): lasso.py:121-131,189,317) */
int decomp_scale_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, int32_t cwidth,
                     const double* rowscale, int32_t invert_row, const double* colscale, int32_t invert_col,
                     double* out, int64_t ldo, void* stream);
/* out = A * Mask (elementwise, mask indexed by col / cwidth)         (grads.py:113,123; lasso.py:321,433) */
int decomp_mask_mul_f64(const double* A, int64_t lda, const double* mask, int64_t ldm, int64_t rows, int64_t cols,
                        int32_t cwidth, double* out, int64_t ldo, void* stream);
/* out[j] = sum_i A[i][j] * scale   (column sums: mean over the batch of the mask, lasso.py:300-303;
 * with rows<->cols swapped by the caller it is also the per-row mask count of lasso.py:163) */
size_t decomp_col_sums_workspace_bytes(int64_t rows, int64_t cols);
int decomp_col_sums_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, double scale, double* out,
                        void* workspace, size_t workspace_bytes, void* stream);
int decomp_row_sums_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, double scale, double* out,
                        void* stream);
/* Lipschitz bound: *step_out = 1 / max_j sum_i |G[i][j]| for a [k,k] (complex: interleaved) matrix
 * (math_utils/eigen.py:9-20; lasso.py:286), and thr[j] = step * alpha_scaled[j] (lasso.py:287). */
int decomp_gershgorin_step_f64(const double* G, int64_t ldg, int64_t k, int32_t is_complex,
                               const double* alpha_scaled, double* step_out, double* thr_out, void* stream);
/* D_out[i] = D_in[i] / ||D_in[i]|| (strict) or / sqrt(max(||.||^2, 1)) (soft); *maxdiff = max |D_ref - D_out|;
 * if tol_latch != NULL and maxdiff < tol: *tol_latch = latch_value  (normalize.py:2-21; batch_mu.py:21-23).
 * `scratch`: two zero-initialised int32 (a ticket counter and a call counter); with latch_value < 0 the value written
 * is the running number of the call (1, 2, ...), for launches replayed from a CUDA graph. */
int decomp_normalize_rows_f64(const double* D_in, int64_t ldi, int64_t rows, int64_t cols, int32_t is_complex,
                              int32_t strict, double* D_out, int64_t ldo, const double* D_ref, int64_t ldr,
                              double tol, int32_t* tol_latch, int32_t latch_value, double* maxdiff,
                              int32_t* scratch, const int32_t* skip_if, void* stream);
/* out[r] = in[index[r]] row gather (MinibatchData.shuffle / .array, utils/data.py:147-156) */
int decomp_gather_rows_f64(const double* in, int64_t ldi, const int64_t* index, int64_t rows, int64_t cols,
                           double* out, int64_t ldo, void* stream);
/* out[index[r]] = in[r]: the codes of a minibatch back into their rows (x[perm[...]] = x_minibatch, data.py:124-156) */
int decomp_scatter_rows_f64(const double* in, int64_t ldi, const int64_t* index, int64_t rows, int64_t cols,
                            double* out, int64_t ldo, void* stream);

/* Lasso prologue vectors: alpha_out[j] = (alpha / s[j]) * mult, tol_out[j] = tol * s[j]; `mult` is read from the
 * device scalar mult_dev when that is non-NULL (sum of a 1-D mask)   (lasso.py:129-130, 136-138) */
int decomp_lasso_vectors_f64(const double* s, int64_t k, double alpha, double tol, double mult, const double* mult_dev,
                             double* alpha_out, double* tol_out, void* stream);
/* out = a * X + b * Y elementwise (A = beta*A + sum over ranks of the local statistics, dictionary_learning.py:151-152,
 * when the minibatch rows are sharded over GPUs) */
int decomp_axpby_f64(double a, const double* X, int64_t ldx, double b, const double* Y, int64_t ldy, int64_t rows,
                     int64_t cols, double* out, int64_t ldo, void* stream);
/* out = max(D * ((1 - alpha) + alpha * P / max(Q, eps)), 0): the variance-reduced basis step of
 * nmf_methods/kasai.py:75-77 (minibatch NMF, "next" row of the scope table) */
int decomp_svrmu_update_f64(const double* D, int64_t ldd, const double* P, int64_t ldp, const double* Q, int64_t ldq,
                            double alpha, int64_t rows, int64_t cols, double* out, int64_t ldo, void* stream);
/* Q = I - (*step) * G for a [k,k] (complex: interleaved) Gram matrix and out = (*scalar_dev) * A: the operands of
 * DECOMP_EPI_PROXQ, i.e. lasso.py:245-246  w + (yAh - w G)/L  written as  yAh/L + w (I - G/L). */
int decomp_lasso_q_f64(const double* G, int64_t ldg, int64_t k, int32_t is_complex, const double* step, double* Q,
                       int64_t ldq, void* stream);
int decomp_scale_scalar_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, const double* scalar_dev,
                            double* out, int64_t ldo, void* stream);
/* out = x * max(num, 0) / max(den, eps), elementwise (masked D update, grads.py:93 with :122-125) */
int decomp_mu_update_f64(const double* x, int64_t ldx, const double* num, int64_t ldn, const double* den, int64_t ldd,
                         int64_t rows, int64_t cols, double* out, int64_t ldo, const int32_t* skip_if, void* stream);
/* result[1] = max |A - B| (result[0] is a zero-initialised accumulator); sets *tol_latch = latch_value when the
 * maximum is < tol   (dictionary_learning.py:161,224) */
int decomp_max_abs_diff_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t rows, int64_t cols,
                            int32_t is_complex, double tol, int32_t* tol_latch, int32_t latch_value, double* result,
                            int32_t* scratch, const int32_t* skip_if, void* stream);

/* ---- TF32-split ("3xTF32") variant of the unmasked ISTA/FISTA iteration (tcgen05 tensor cores, FP32 accumulate) ----
 * Not bit-compatible with the FP64 path: products carry ~22 significant bits (see DESIGN.md for the tolerance).
 * hi/lo: FP32 arrays holding TF32-valued pieces, v ~= hi + lo; row pitch `ldh` in floats, multiple of 4. */
int decomp_split_tf32_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ldh,
                          void* stream);
/* P[M,N] (FP32, row-major) = A_lo.B_hi^T + A_hi.B_lo^T + A_hi.B_hi^T; A_* [M,K], B_* [N,K] K-contiguous FP32
 * (row pitches multiples of 4 floats, 16-byte aligned).  Any M, N, K: tiles of 128 x 256, edges zero-filled by TMA.
 * Replaces xp.tensordot(x0, AAt) of lasso.py:245 (with B = (I - G/L)^T) and x.dot(D D^T) of grads.py:110. */
int decomp_gemm_nt_tf32x3(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi, const float* B_lo,
                          int64_t ldb, int64_t M, int64_t N, int64_t K, float* P, int64_t ldp, const int32_t* skip_if,
                          void* stream);
/* out[M,N] (FP64) = A . B^T as above with the contraction cut into pieces of `k_per_split`: every piece is
 * accumulated in FP32 in tensor memory, written to an FP32 slab of `workspace`, and the slabs are summed in FP64 in a
 * fixed order (deterministic; the FP32 accumulation length is bounded by k_per_split whatever K is).
 * The sample-axis contractions x^T y, x^T x of grads.py:119-121 with A = x^T, B = y^T (both K-major).
 * k_blocked != 0: the operands are stored K-blocked, [ceil(K / k_per_split)][rows][k_per_split] FP32 with the tail of
 * the last block zero-filled (lda / ldb unused): piece z then reads block z, rows k_per_split * 4 bytes apart instead
 * of K * 4 -- the layout decomp_split_transpose_tf32_f64 and decomp_nmf_xupdate_tf32x3 produce for x^T and y^T. */
size_t decomp_gemm_nt_tf32x3_splitk_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t k_per_split);
int decomp_gemm_nt_tf32x3_splitk_f64(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi,
                                     const float* B_lo, int64_t ldb, int64_t M, int64_t N, int64_t K,
                                     int64_t k_per_split, int32_t k_blocked, double* out, int64_t ldo,
                                     void* workspace, size_t workspace_bytes, const int32_t* skip_if, void* stream);
/* NMF x update (grads.py:77-84, 108-111) with y D^T on the tcgen05 tensor cores:
 *   X <- X * max(Y D^T, 0) / max(NEG, 1e-15)    Y_* [n,f], D_* [k,f] TF32 pairs, NEG = X (D D^T) [n,k] FP32,
 * evaluated in FP64 from the FP32 accumulator; the new X is written as FP64 [n,k], as its TF32 pair row-major
 * (X_hi, X_lo: next sweep's x (D D^T)) and transposed (XT_hi, XT_lo: this sweep's x^T y, x^T x) -- [k][ldxt] when
 * xt_block == 0, else K-blocked [ceil(n / xt_block)][k][xt_block] (xt_block % 128 == 0; the caller zero-fills the
 * tail of the last block once).  k % 32 == 0, k <= 256. */
int decomp_nmf_xupdate_tf32x3(const float* Y_hi, const float* Y_lo, int64_t ldy, const float* D_hi, const float* D_lo,
                              int64_t ldd, int64_t n, int64_t k, int64_t f, double* X, int64_t ldx, const float* NEG,
                              int64_t ldneg, float* X_hi, float* X_lo, int64_t ldxh, float* XT_hi, float* XT_lo,
                              int64_t ldxt, int64_t xt_block, const int32_t* skip_if, void* stream);
/* The masked model's [rows, f] intermediate on the tcgen05 tensor cores (grads.py:112,114,122,124: f = x.dot(d) * mask;
 * lasso.py:262 the same with w, A):  F = (A . B^T) * mask, A_* [M,K], B_* [N,K] TF32 pairs, mask [M, N / cwidth] FP32 (null: no
 * mask; cwidth 2: interleaved complex columns share one entry), the product rounded once to FP32 and written as a TF32 pair -- row-major (F_hi, F_lo [M][ldf]: A operand of
 * F D^T) when F_hi != NULL, and / or transposed (FT_hi, FT_lo: [N][ldft] when ft_block == 0, else K-blocked
 * [ceil(M / ft_block)][N][ft_block], ft_block % 128 == 0, tail zero-filled by the caller once: B operand of x^T F)
 * when FT_hi != NULL.  Any M, N, K. */
int decomp_gemm_nt_mask_tf32x3(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi, const float* B_lo,
                               int64_t ldb, int64_t M, int64_t N, int64_t K, const float* mask, int64_t ldmask,
                               int32_t cwidth, float* F_hi, float* F_lo, int64_t ldf, float* FT_hi, float* FT_lo,
                               int64_t ldft, int64_t ft_block, const int32_t* skip_if, void* stream);
/* out (FP32) = A (FP64), round to nearest: the mask of the TF32-split masked path, once per solve */
int decomp_to_f32_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, float* out, int64_t ldo, void* stream);
/* TF32 pair of A^T from A [rows, cols] FP64 (y^T, once per solve): hiT/loT [cols][ldt] FP32 when block == 0, else
 * K-blocked [ceil(rows / block)][cols][block] with the tail of the last block zero-filled. */
int decomp_split_transpose_tf32_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, float* hiT, float* loT,
                                    int64_t ldt, int64_t block, void* stream);
/* The FP64 part of the iteration in one streaming pass (lasso.py:192-256, 405-414 with z = other + P):
 * uses epi->{out, other, prev, colvec (= step*alpha), colvec2, momentum, shrink, check, latch, scratch, latch_value};
 * w_next is written as the TF32 pair (w_hi, w_lo) the next decomp_gemm_nt_tf32x3 reads. */
int decomp_proxq_apply_f64(const float* P, int64_t ldp, const decomp_epilogue_t* epi, float* w_hi, float* w_lo,
                           int64_t ldw, int64_t M, int64_t N, const int32_t* skip_if, void* stream);
/* The same pass for the masked iteration (lasso.py:259-271; the DECOMP_EPI_PROX epilogue of the FP64 GEMMs):
 * z = w + step (other - P) with w = w_hi + w_lo and P = ((w A) * mask) A^H, threshold step * colvec * rowvec
 * (rowvec NULL: step * colvec, or colvec itself with DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD); epi->step is required. */
int decomp_prox_apply_f64(const float* P, int64_t ldp, const decomp_epilogue_t* epi, float* w_hi, float* w_lo,
                          int64_t ldw, int64_t M, int64_t N, const int32_t* skip_if, void* stream);

/* `iters` (1..DECOMP_LASSO_RESIDENT_MAX_ITERS) unmasked ISTA / FISTA iterations in ONE launch with the iterate
 * resident on chip (lasso.py:244-271, 405-414 with the gradient step folded into Q as for DECOMP_EPI_PROXQ):
 *     z = c + w Q^T...;  x = shrink(z, thr);  w = x + momentum[i] (x - x_prev)
 * Uses epi->{x (= w, in/out), ldx, other (= c = (y A^H)/L), ldother, out (= x, in: x_prev, out: x after the last
 * iteration), ldo, colvec (= step * alpha, DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD), colvec2, shrink, check (the test of
 * lasso.py:293/409 is evaluated in the LAST iteration of the launch), latch, scratch, latch_value}.
 * Q is the [N, N] NT operand of decomp_gemm_nt_f64 (B = (I - G/L)^T); `momentum` is a HOST array of `iters` values.
 * N must be 32, 64, 128 or 256 (decomp_lasso_resident_supported); results are bitwise those of `iters` launches of
 * decomp_gemm_nt_f64 with DECOMP_EPI_PROXQ. */
#define DECOMP_LASSO_RESIDENT_MAX_ITERS 32
int decomp_lasso_resident_supported(int64_t N);
int decomp_lasso_resident_f64(const double* Q, int64_t ldq, int64_t M, int64_t N, const decomp_epilogue_t* epi,
                              int32_t iters, const double* momentum, const int32_t* skip_if, void* stream);

/* ---- dictionary-learning basis update ------------------------------------------------- */
/* Gauss-Seidel atom sweep, dictionary_learning.py:154-159:
 *   for a in 0..k-1: u = (T[a] - S[a].D) / (S[a][a] + eps) + D[a];  D[a] = u / sqrt(max(|u|^2, 1))
 * in place on D (initialised by the caller with the old dictionary). One cooperative launch; `workspace` (block
 * partials of |u|^2 and the grid barrier counter) is sized by decomp_dl_sweep_workspace_bytes(). */
size_t decomp_dl_sweep_workspace_bytes(int64_t k, int64_t f, int32_t is_complex);
int decomp_dl_sweep_f64(const double* S, int64_t lds, const double* T, int64_t ldt, double* D, int64_t ldd,
                        int64_t k, int64_t f, int32_t is_complex, void* workspace, size_t workspace_bytes,
                        void* stream);
/* Masked statistics, dictionary_learning.py:210-213:  S[a][j][b] = beta*S[a][j][b] + sum_i conj(x_ia) x_ib m_ij.
 * Per atom a this is the TN product  Mask^T . W_a  with  W_a[i][b] = conj(x_ia) x_ib ; this entry point forms W_a
 * (interleaved complex when is_complex) and the caller feeds it to decomp_gemm_tn_f64(combine=1, beta) with
 * A = Mask, out = S[a].  Internal layout of S is [k][f][k*cw] doubles (never returned to the user). */
int decomp_dl_atom_weighted_f64(const double* X, int64_t ldx, int64_t rows, int64_t k, int32_t is_complex,
                                int64_t atom, double* W, int64_t ldw, void* stream);
/* Masked statistics of several atoms per GEMM (dictionary_learning.py:210-213): for `width` (atom a, b >= a) pairs
 * Wt[c*cw + part][i] = (conj(x[i][colA[c]]) x[i][colB[c]]).part, from the transposed real view Xt of the codes
 * (Xt[b*cw + part][i]); the caller runs the NT GEMM  mask^T[f, rows] . Wt^T -> P [f, width*cw]  and then
 * S[colA[c]][j][colB[c]] = beta * S[..] + P[j][c].  colA / colB: device int32 vectors. */
int decomp_dl_pair_products_t_f64(const double* Xt, int64_t ldx, int64_t rows, int32_t is_complex, const int32_t* colA,
                                  const int32_t* colB, int64_t width, double* Wt, int64_t ldw, void* stream);
/* slab_channels: S is stored as channel slabs [f / slab_channels][k][slab_channels][k*cw] (0 or f: the plain
 * [k][f][k*cw] tensor); with slab_channels = ceil(f / ranks) it is the send buffer of a reduce-scatter along f. */
int decomp_dl_scatter_stats_f64(const double* P, int64_t ldp, int64_t f, int64_t width, int32_t is_complex,
                                const int32_t* colA, const int32_t* colB, int64_t k, double beta, double* S,
                                int64_t slab_channels, void* stream);
/* S[b][j][a] = conj(S[a][j][b]) for b > a on the [k][f][k*cw] tensor: the statistics of dictionary_learning.py:210-213
 * are Hermitian in (a, b), so the GEMMs accumulate b >= a only (half the flops) and this fills in the rest. */
int decomp_dl_mirror_f64(double* S, int64_t k, int64_t f, int32_t is_complex, void* stream);
/* Masked Jacobi atom update, dictionary_learning.py:216-222:
 *   SaD[j] = sum_b S[a][j][b] D[b][j];  Saa = sum_j (S[a][j][a] + eps);  u = (T[a] - SaD)/Saa + D[a];  l2(u)
 * `workspace` holds the transposed dictionary: f * k * cw doubles. */
int decomp_dl_masked_update_f64(const double* S, const double* T, int64_t ldt, const double* D, int64_t ldd,
                                int64_t k, int64_t f, int32_t is_complex, double* D_out, int64_t ldo,
                                double* workspace, void* stream);

/* The same update on ONE channel slab (channels j0 .. j0 + slab_channels of S, [k][slab_channels][k*cw]), for statistics
 * reduce-scattered along f over several GPUs (the update is channel-local, dictionary_learning.py:218), in three phases
 * separated by the sums over all channels the ranks all-reduce: phase 1 writes T[a] - S_a D into D_slab_out
 * [k][slab_channels*cw] and the partial sum_j S[a][j][a] into stats[4a .. 4a+1]; phase 2 (after the all-reduce of
 * stats) forms u and its partial |u|^2 in stats[4a+2]; phase 3 (after the second all-reduce) divides by
 * sqrt(max(|u|^2, 1)).  `workspace`: slab_channels * k * cw doubles (phase 1). */
int decomp_dl_masked_update_phase_f64(int32_t phase, const double* S_slab, int64_t slab_channels, int64_t j0,
                                      const double* T, int64_t ldt, const double* D, int64_t ldd, int64_t k, int64_t f,
                                      int32_t is_complex, double* D_slab_out, double* stats, double* workspace,
                                      void* stream);

/* ---- whole NMF-MU runs of small problems in one cooperative launch ------------------------ */
/* batch_mu.solve (nmf_methods/batch_mu.py:8-26) for problems whose rows / dictionary fit in shared memory (k <= 32;
 * BASELINE configs[0] and the sizes of the reference's own tests): `sweeps` = maxiter - 1 multiplicative updates of
 * x [n,k] (in place) and D (D_in: rows already l2_strict-normalised, nmf.py:70; D_out: result), 'l2' likelihood,
 * optional mask [n,f].  *it_out = 0, or the sweep at which max |D - D_new| < tol (tol <= 0: never).  Same arithmetic
 * as the sweep-by-sweep kernels up to summation order.  workspace: decomp_nmf_mu_small_workspace_bytes(). */
int decomp_nmf_mu_small_supported(int64_t n, int64_t f, int64_t k, int32_t masked);
size_t decomp_nmf_mu_small_workspace_bytes(int64_t n, int64_t f, int64_t k, int32_t masked);
int decomp_nmf_mu_small_f64(const double* y, int64_t ldy, const double* mask, int64_t ldm, double* x, int64_t ldx,
                            const double* D_in, int64_t ldd, double* D_out, int64_t ldo, int64_t n, int64_t f,
                            int64_t k, int32_t sweeps, double tol, int32_t* it_out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- host-side staging of pageable inputs -------------------------------------------------- */
/* dst_device[0:bytes] = src_host[0:bytes] for an ordinary (pageable) host array, the kind of array the reference's
 * solve() functions receive (decomp/lasso.py:19, decomp/nmf.py:16, decomp/dictionary_learning.py:12): `threads` host
 * threads memcpy pieces of `slot_bytes` into the caller's page-locked ring (`slots` x `slot_bytes` bytes) while the
 * calling thread enqueues cudaMemcpyAsync of the filled slots on `stream`.  Returns when the last piece has left the
 * ring (the device copy itself is complete in stream order).  Touches no Python state. */
int decomp_staged_upload(void* dst_device, const void* src_host, size_t bytes, void* pinned_ring, size_t slot_bytes,
                         int32_t slots, int32_t threads, void* stream);

/* ---- collectives of the sharded solves (thin NCCL wrappers, bound to the process's libnccl.so.2 at run time) ----
 * One communicator per process (= per GPU).  decomp_comm_unique_id on one rank, the 128 bytes handed to every rank by
 * the host, decomp_comm_init on all ranks.  The calls enqueue on `stream`.
 *   all-reduce (sum, f64):    x^T y [k,f], x^T x [k,k] per NMF sweep / dictionary-learning minibatch (grads.py:117-125,
 *                             dictionary_learning.py:147-152), the two [k] sums of the sharded masked update
 *   all-reduce (min, i32):    the convergence latch of the batched Lasso (lasso.py:293/409: one max over ALL problems)
 *   reduce-scatter / all-gather (f64): the masked [k,f,k] statistic along f and the new dictionary
 *                             (dictionary_learning.py:206-222) */
#define DECOMP_COMM_ID_BYTES 128
int decomp_comm_unique_id(void* id_out);
int decomp_comm_init(const void* id, int32_t nranks, int32_t rank, void** comm_out);
int decomp_comm_destroy(void* comm);
int decomp_comm_allreduce_sum_f64(void* comm, double* buf, int64_t count, void* stream);
int decomp_comm_allreduce_min_i32(void* comm, int32_t* buf, int64_t count, void* stream);
int decomp_comm_reduce_scatter_sum_f64(void* comm, const double* send, double* recv, int64_t recv_count,
                                       void* stream);
int decomp_comm_allgather_f64(void* comm, const double* send, double* recv, int64_t send_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DECOMP_B200_H_ */

// Bandwidth-bound vector / reduction kernels of the hot path (see include/decomp_b200.h for the
// reference lines each one replaces).  All are coalesced along the contiguous dimension, use
// warp-shuffle reductions, and size their grids from the SM count.
#include "common.h"

namespace dcp {

constexpr double kEps = 1.0e-15;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reductions; `red` is a shared array of >= 32 doubles. Result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ double block_max_nan(double v, double* red) {
  // NaN-propagating max for non-negative inputs: compares IEEE bit patterns
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t > b ? t : b;
  }
  __syncthreads();
  if (lane == 0) red[w] = __longlong_as_double((long long)b);
  __syncthreads();
  unsigned long long r = (lane < nw) ? (unsigned long long)__double_as_longlong(red[lane]) : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, r, o);
    r = t > r ? t : r;
  }
  return __longlong_as_double((long long)r);
}

// ------------------------------------------------------------------------------------ make_rhs
__global__ void transpose_real_kernel(const double* __restrict__ S, long long lds, int p, int q,
                                      double* __restrict__ B, long long ldb, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  __shared__ double tile[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, j = j0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < p && j < q) ? S[(long long)i * lds + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r, i = i0 + threadIdx.x;
    if (j < q && i < p) B[(long long)j * ldb + i] = tile[threadIdx.x][r];
  }
}

__global__ void copy2d_kernel(const double* __restrict__ S, long long lds, long long rows, long long cols,
                              double* __restrict__ B, long long ldb, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    B[r * ldb + c] = S[r * lds + c];
  }
}

// complex S [p, q] interleaved -> real block embedding, see decomp_b200.h
__global__ void embed_complex_kernel(const double* __restrict__ S, long long lds, int p, int q, int conj_transpose,
                                     double* __restrict__ B, long long ldb, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  const long long total = (long long)p * q;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / q, j = idx % q;
    const double sr = S[i * lds + 2 * j], si = S[i * lds + 2 * j + 1];
    if (conj_transpose) {
      // X . S^H : B [2p, 2q]
      double* b0 = B + (2 * i) * ldb + 2 * j;
      double* b1 = b0 + ldb;
      b0[0] = sr;
      b0[1] = si;
      b1[0] = -si;
      b1[1] = sr;
    } else {
      // X . S : B [2q, 2p]
      double* b0 = B + (2 * j) * ldb + 2 * i;
      double* b1 = b0 + ldb;
      b0[0] = sr;
      b0[1] = -si;
      b1[0] = si;
      b1[1] = sr;
    }
  }
}

// ------------------------------------------------------------------------------------ row norms
__global__ void row_norms_kernel(const double* __restrict__ A, long long lda, long long cols_real,
                                 double* __restrict__ out) {
  __shared__ double red[32];
  const long long row = blockIdx.x;
  const double* a = A + row * lda;
  double s = 0.0;
  for (long long c = threadIdx.x; c < cols_real; c += blockDim.x) {
    const double v = a[c];
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[row] = sqrt(s);
}

// ------------------------------------------------------------------------------------ scaling / masking
__global__ void scale_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols, int cw,
                             const double* __restrict__ rowscale, int invert_row, const double* __restrict__ colscale,
                             int invert_col, double* __restrict__ out, long long ldo) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    double v = A[r * lda + c];
    if (rowscale != nullptr) v = invert_row ? v / rowscale[r] : v * rowscale[r];
    if (colscale != nullptr) {
      const double s = colscale[cw == 2 ? (c >> 1) : c];
      v = invert_col ? v / s : v * s;
    }
    out[r * ldo + c] = v;
  }
}

__global__ void mask_mul_kernel(const double* __restrict__ A, long long lda, const double* __restrict__ mask,
                                long long ldm, long long rows, long long cols, int cw, double* __restrict__ out,
                                long long ldo) {
  // two doubles per thread (cols is even for complex; the scalar tail handles odd real widths)
  const long long pairs = (cols + 1) / 2;
  const long long total = rows * pairs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / pairs, c = (idx % pairs) * 2;
    const double* a = A + r * lda + c;
    double* o = out + r * ldo + c;
    if (c + 1 < cols) {
      const double2 v = *reinterpret_cast<const double2*>(a);
      double m0, m1;
      if (cw == 2) {
        m0 = m1 = mask[r * ldm + (c >> 1)];
      } else {
        const double2 m = *reinterpret_cast<const double2*>(mask + r * ldm + c);
        m0 = m.x;
        m1 = m.y;
      }
      *reinterpret_cast<double2*>(o) = make_double2(v.x * m0, v.y * m1);
    } else {
      o[0] = a[0] * mask[r * ldm + c];
    }
  }
}

// out[j] = scale * sum_i A[i][j] ; one thread per column, coalesced across the warp, rows split over blockIdx.y
__global__ void col_sums_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols,
                                double* __restrict__ partial, int row_chunks) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= cols) return;
  const long long per = (rows + row_chunks - 1) / row_chunks;
  const long long r0 = blockIdx.y * per;
  long long r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  double s = 0.0;
  for (long long r = r0; r < r1; ++r) s += A[r * lda + j];
  partial[(long long)blockIdx.y * cols + j] = s;
}
__global__ void col_sums_finish_kernel(const double* __restrict__ partial, long long cols, int row_chunks,
                                       double scale, double* __restrict__ out) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double s = 0.0;
  for (int z = 0; z < row_chunks; ++z) s += partial[(long long)z * cols + j];
  out[j] = s * scale;
}

// out[i] = scale * sum_j A[i][j] ; one warp per row
__global__ void row_sums_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols,
                                double scale, double* __restrict__ out) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const double* a = A + r * lda;
    double s = 0.0;
    for (long long c = lane; c < cols; c += 32) s += a[c];
    s = warp_sum(s);
    if (lane == 0) out[r] = s * scale;
  }
}

// ------------------------------------------------------------------------------------ Gershgorin step
__global__ void gershgorin_kernel(const double* __restrict__ G, long long ldg, int k, int is_complex,
                                  const double* __restrict__ alpha, double* __restrict__ step_out,
                                  double* __restrict__ thr_out) {
  __shared__ double red[32];
  __shared__ double step_s;
  __shared__ double part[32][33];
  // column sums of |G|: 32 columns at a time (coalesced), the rows dealt out to the 32 warps, partial sums combined
  // in warp order
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double best = 0.0;
  for (int jb = 0; jb < k; jb += 32) {
    const int j = jb + tx;
    double s = 0.0;
    if (j < k) {
      if (is_complex) {
        for (int i = ty; i < k; i += 32) s += hypot(G[(long long)i * ldg + 2 * j], G[(long long)i * ldg + 2 * j + 1]);
      } else {
        for (int i = ty; i < k; i += 32) s += fabs(G[(long long)i * ldg + j]);
      }
    }
    part[ty][tx] = s;
    __syncthreads();
    if (ty == 0) {
      double c = 0.0;
#pragma unroll
      for (int t = 0; t < 32; ++t) c += part[t][tx];
      // sums of magnitudes are >= 0 or NaN: ordering by bit pattern keeps a NaN, like numpy's max
      if ((unsigned long long)__double_as_longlong(c) > (unsigned long long)__double_as_longlong(best)) best = c;
    }
    __syncthreads();
  }
  best = block_max_nan(best, red);
  if (threadIdx.x == 0) {
    step_s = 1.0 / best;
    *step_out = step_s;
  }
  __syncthreads();
  if (thr_out != nullptr && alpha != nullptr)
    for (int j = threadIdx.x; j < k; j += blockDim.x) thr_out[j] = step_s * alpha[j];
}

// ------------------------------------------------------------------------------------ row normalisation + latch
// maxdiff[0] is a running accumulator kept at zero between calls, maxdiff[1] receives the result.
__global__ void normalize_rows_kernel(const double* __restrict__ Din, long long ldi, long long cols_real, int strict,
                                      double* __restrict__ Dout, long long ldo, const double* __restrict__ Dref,
                                      long long ldr, double tol, int* __restrict__ tol_latch, int latch_value,
                                      double* __restrict__ maxdiff, int* __restrict__ scratch,
                                      const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  __shared__ double red[32];
  const long long row = blockIdx.x;
  const double* a = Din + row * ldi;
  double s = 0.0;
  for (long long c = threadIdx.x; c < cols_real; c += blockDim.x) {
    const double v = a[c];
    s += v * v;
  }
  s = block_sum(s, red);
  const double nrm = strict ? sqrt(s) : sqrt(fmax(s, 1.0));
  double md = 0.0;
  const double* ref = Dref != nullptr ? Dref + row * ldr : nullptr;
  for (long long c = threadIdx.x; c < cols_real; c += blockDim.x) {
    const double v = a[c] / nrm;
    if (ref != nullptr) {
      const double d = fabs(ref[c] - v);
      // NaN must win the max like numpy's max does
      md = (d != d) ? d : ((md != md) ? md : fmax(md, d));
    }
    Dout[row * ldo + c] = v;
  }
  if (maxdiff == nullptr) return;
  md = block_max_nan(md, red);
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(maxdiff), (unsigned long long)__double_as_longlong(md));
    __threadfence();
    const int ticket = atomicAdd(&scratch[0], 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const unsigned long long bits = atomicMax(reinterpret_cast<unsigned long long*>(maxdiff), 0ull);
      const double m = __longlong_as_double((long long)bits);
      maxdiff[1] = m;
      // latch_value < 0: the value is the number of this call, counted in scratch[1] (launches replayed from a CUDA
      // graph cannot carry a different value per replay)
      int value = latch_value;
      if (latch_value < 0) {
        value = scratch[1] + 1;
        scratch[1] = value;
      }
      if (tol_latch != nullptr && m < tol) *tol_latch = value;
      maxdiff[0] = 0.0;
      scratch[0] = 0;
    }
  }
}

// complex-aware variant is not needed: |re + i im|^2 summed over interleaved doubles equals the sum of
// squares over the real view, but max|D - D_new| for complex data is the complex modulus:
__global__ void normalize_rows_complex_kernel(const double* __restrict__ Din, long long ldi, long long cols_c,
                                              int strict, double* __restrict__ Dout, long long ldo,
                                              const double* __restrict__ Dref, long long ldr, double tol,
                                              int* __restrict__ tol_latch, int latch_value,
                                              double* __restrict__ maxdiff, int* __restrict__ scratch,
                                              const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  __shared__ double red[32];
  const long long row = blockIdx.x;
  const double* a = Din + row * ldi;
  double s = 0.0;
  for (long long c = threadIdx.x; c < cols_c; c += blockDim.x) {
    const double2 v = *reinterpret_cast<const double2*>(a + 2 * c);
    s += v.x * v.x + v.y * v.y;
  }
  s = block_sum(s, red);
  const double nrm = strict ? sqrt(s) : sqrt(fmax(s, 1.0));
  double md = 0.0;
  const double* ref = Dref != nullptr ? Dref + row * ldr : nullptr;
  for (long long c = threadIdx.x; c < cols_c; c += blockDim.x) {
    const double2 v = *reinterpret_cast<const double2*>(a + 2 * c);
    const double vr = v.x / nrm, vi = v.y / nrm;
    if (ref != nullptr) {
      const double d = hypot(ref[2 * c] - vr, ref[2 * c + 1] - vi);
      md = (d != d) ? d : ((md != md) ? md : fmax(md, d));
    }
    *reinterpret_cast<double2*>(Dout + row * ldo + 2 * c) = make_double2(vr, vi);
  }
  if (maxdiff == nullptr) return;
  md = block_max_nan(md, red);
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(maxdiff), (unsigned long long)__double_as_longlong(md));
    __threadfence();
    const int ticket = atomicAdd(&scratch[0], 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const unsigned long long bits = atomicMax(reinterpret_cast<unsigned long long*>(maxdiff), 0ull);
      const double m = __longlong_as_double((long long)bits);
      maxdiff[1] = m;
      // latch_value < 0: the value is the number of this call, counted in scratch[1] (launches replayed from a CUDA
      // graph cannot carry a different value per replay)
      int value = latch_value;
      if (latch_value < 0) {
        value = scratch[1] + 1;
        scratch[1] = value;
      }
      if (tol_latch != nullptr && m < tol) *tol_latch = value;
      maxdiff[0] = 0.0;
      scratch[0] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------ row gather
__global__ void gather_rows_kernel(const double* __restrict__ in, long long ldi, const long long* __restrict__ index,
                                   long long rows, long long cols, double* __restrict__ out, long long ldo) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const double* src = in + index[r] * ldi;
    double* dst = out + r * ldo;
    for (long long c = lane; c < cols; c += 32) dst[c] = src[c];
  }
}

// out[index[r]] = in[r]: the codes of a minibatch go back to their rows
__global__ void scatter_rows_kernel(const double* __restrict__ in, long long ldi, const long long* __restrict__ index,
                                    long long rows, long long cols, double* __restrict__ out, long long ldo) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const double* src = in + r * ldi;
    double* dst = out + index[r] * ldo;
    for (long long c = lane; c < cols; c += 32) dst[c] = src[c];
  }
}

// ------------------------------------------------------------------------------------ Lasso prologue vectors
// alpha_out[j] = (alpha / s[j]) * mult ; tol_out[j] = tol * s[j]      (lasso.py:129-130, 136-138)
__global__ void lasso_vectors_kernel(const double* __restrict__ s, int k, double alpha, double tol, double mult,
                                     const double* __restrict__ mult_dev, double* __restrict__ alpha_out,
                                     double* __restrict__ tol_out) {
  const double m = mult_dev != nullptr ? *mult_dev : mult;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    alpha_out[j] = (alpha / s[j]) * m;
    tol_out[j] = tol * s[j];
  }
}

// ------------------------------------------------------------------------------------ folded gradient step
// Q = I - step * G (complex: interleaved pairs), the right-hand operand that turns  w + step * (yAh - w G)  into
// step * yAh + w Q  (lasso.py:245-246 re-associated; exact in real arithmetic)
__global__ void lasso_q_kernel(const double* __restrict__ G, long long ldg, int k, int cw,
                               const double* __restrict__ step, double* __restrict__ Q, long long ldq) {
  const double st = *step;
  const long long total = (long long)k * k * cw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / (k * cw), c = idx % (k * cw);
    const bool diag_re = (c == i * cw);
    Q[i * ldq + c] = (diag_re ? 1.0 : 0.0) - st * G[i * ldg + c];
  }
}

__global__ void scale_scalar_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols,
                                    const double* __restrict__ scalar, double* __restrict__ out, long long ldo) {
  const double sc = *scalar;
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    out[r * ldo + c] = sc * A[r * lda + c];
  }
}

// ------------------------------------------------------------------------------------ out = a * X + b * Y
__global__ void axpby_kernel(double a, const double* __restrict__ X, long long ldx, double b,
                             const double* __restrict__ Y, long long ldy, long long rows, long long cols,
                             double* __restrict__ out, long long ldo) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    out[r * ldo + c] = a * X[r * ldx + c] + b * Y[r * ldy + c];
  }
}

// ------------------------------------------------------------------------------------ SVRMU basis step
// out = max(D * ((1 - alpha) + alpha * P / max(Q, eps)), 0)            (nmf_methods/kasai.py:75-77)
__global__ void svrmu_update_kernel(const double* __restrict__ D, long long ldd, const double* __restrict__ P,
                                    long long ldp, const double* __restrict__ Q, long long ldq, double alpha,
                                    long long rows, long long cols, double* __restrict__ out, long long ldo) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    const double ratio = __ddiv_rn(__dmul_rn(alpha, P[r * ldp + c]), fmax(Q[r * ldq + c], kEps));
    const double v = __dmul_rn(D[r * ldd + c], __dadd_rn(1.0 - alpha, ratio));
    out[r * ldo + c] = fmax(v, 0.0);
  }
}

// ------------------------------------------------------------------------------------ elementwise MU ratio
// out = x * max(num, 0) / max(den, eps)                                (grads.py:84,93)
__global__ void mu_update_kernel(const double* __restrict__ x, long long ldx, const double* __restrict__ num,
                                 long long ldn, const double* __restrict__ den, long long ldd, long long rows,
                                 long long cols, double* __restrict__ out, long long ldo,
                                 const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    out[r * ldo + c] = __ddiv_rn(__dmul_rn(x[r * ldx + c], fmax(num[r * ldn + c], 0.0)), fmax(den[r * ldd + c], kEps));
  }
}

// ------------------------------------------------------------------------------------ max |A - B| with latch
// result[0] accumulator (kept zero between calls), result[1] = max |A - B| (complex modulus when is_complex)
__global__ void max_abs_diff_kernel(const double* __restrict__ A, long long lda, const double* __restrict__ B,
                                    long long ldb, long long rows, long long cols, int is_complex, double tol,
                                    int* __restrict__ tol_latch, int latch_value, double* __restrict__ result,
                                    int* __restrict__ scratch, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  __shared__ double red[32];
  const long long total = rows * cols;
  double md = 0.0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    double d;
    if (is_complex)
      d = hypot(A[r * lda + 2 * c] - B[r * ldb + 2 * c], A[r * lda + 2 * c + 1] - B[r * ldb + 2 * c + 1]);
    else
      d = fabs(A[r * lda + c] - B[r * ldb + c]);
    md = (d != d) ? d : ((md != md) ? md : fmax(md, d));
  }
  md = block_max_nan(md, red);
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(result), (unsigned long long)__double_as_longlong(md));
    __threadfence();
    const int ticket = atomicAdd(&scratch[0], 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const unsigned long long bits = atomicMax(reinterpret_cast<unsigned long long*>(result), 0ull);
      const double m = __longlong_as_double((long long)bits);
      result[1] = m;
      if (tol_latch != nullptr && m < tol) *tol_latch = latch_value;
      result[0] = 0.0;
      scratch[0] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------ DMMA issue-rate probe
// Roofline denominator for the FP64 tensor path (MEASURED_PEAKS.json has no FP64 entry): every warp issues
// independent DMMA.8x8x4 chains from registers, no memory traffic.
__global__ void dmma_rate_kernel(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == 123.456) out[threadIdx.x] = s;
}

static int grid_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dcp

using namespace dcp;

extern "C" {

int decomp_probe_dmma_tflops(double* tflops_out) {
  if (tflops_out == nullptr) return DECOMP_ERR_INVALID;
  double* buf = nullptr;
  int rc = check_cuda(cudaMalloc(&buf, (64 + 1024) * sizeof(double)), "probe alloc");
  if (rc != DECOMP_OK) return rc;
  cudaMemset(buf, 0, (64 + 1024) * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 8192, warps = 16, sms = num_sms();
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    dmma_rate_kernel<<<sms, warps * 32>>>(buf + 64, buf, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  if (e != cudaSuccess) return check_cuda(e, "dmma probe");
  *tflops_out = 2.0 * 256.0 * 8.0 * iters * warps * sms / (best * 1e-3) / 1e12;
  return DECOMP_OK;
}

int decomp_make_rhs_f64(const double* S, int64_t lds, int64_t p, int64_t q, int32_t is_complex,
                        int32_t conj_transpose, double* B, int64_t ldb, const int32_t* skip_if, void* stream) {
  if (p <= 0 || q <= 0) return DECOMP_OK;
  cudaStream_t st = as_stream(stream);
  if (is_complex) {
    embed_complex_kernel<<<grid_for(p * q, 256), 256, 0, st>>>(S, lds, (int)p, (int)q, conj_transpose, B, ldb, skip_if);
  } else if (conj_transpose) {
    copy2d_kernel<<<grid_for(p * q, 256), 256, 0, st>>>(S, lds, p, q, B, ldb, skip_if);
  } else {
    dim3 grid((unsigned)((q + 31) / 32), (unsigned)((p + 31) / 32));
    transpose_real_kernel<<<grid, dim3(32, 8), 0, st>>>(S, lds, (int)p, (int)q, B, ldb, skip_if);
  }
  DCP_CHECK_LAUNCH("make_rhs");
  return DECOMP_OK;
}

int decomp_row_norms_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, int32_t is_complex, double* out,
                         void* stream) {
  if (rows <= 0) return DECOMP_OK;
  row_norms_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(A, lda, is_complex ? 2 * cols : cols, out);
  DCP_CHECK_LAUNCH("row_norms");
  return DECOMP_OK;
}

int decomp_scale_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, int32_t cwidth, const double* rowscale,
                     int32_t invert_row, const double* colscale, int32_t invert_col, double* out, int64_t ldo,
                     void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  scale_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(A, lda, rows, cols, cwidth, rowscale,
                                                                         invert_row, colscale, invert_col, out, ldo);
  DCP_CHECK_LAUNCH("scale");
  return DECOMP_OK;
}

int decomp_mask_mul_f64(const double* A, int64_t lda, const double* mask, int64_t ldm, int64_t rows, int64_t cols,
                        int32_t cwidth, double* out, int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  if ((lda & 1) || (ldo & 1) || (cwidth != 2 && (ldm & 1))) {
    set_error("decomp_mask_mul_f64: leading dimensions must be even");
    return DECOMP_ERR_INVALID;
  }
  mask_mul_kernel<<<grid_for(rows * ((cols + 1) / 2), 256), 256, 0, as_stream(stream)>>>(A, lda, mask, ldm, rows, cols,
                                                                                       cwidth, out, ldo);
  DCP_CHECK_LAUNCH("mask_mul");
  return DECOMP_OK;
}

static int col_sums_chunks(int64_t rows) {
  int chunks = (int)((rows + 4095) / 4096);
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  return chunks;
}

size_t decomp_col_sums_workspace_bytes(int64_t rows, int64_t cols) {
  if (cols <= 0) return 0;
  return (size_t)col_sums_chunks(rows) * (size_t)cols * sizeof(double);
}

// column sums in two deterministic stages: per row chunk into the caller's workspace, then over the chunks
int decomp_col_sums_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, double scale, double* out,
                        void* workspace, size_t workspace_bytes, void* stream) {
  if (cols <= 0) return DECOMP_OK;
  if (workspace == nullptr || workspace_bytes < decomp_col_sums_workspace_bytes(rows, cols)) {
    set_error("decomp_col_sums_f64: workspace too small (%zu < %zu)", workspace_bytes,
              decomp_col_sums_workspace_bytes(rows, cols));
    return DECOMP_ERR_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  const int chunks = col_sums_chunks(rows);
  double* partial = reinterpret_cast<double*>(workspace);
  dim3 grid((unsigned)((cols + 127) / 128), (unsigned)chunks);
  col_sums_kernel<<<grid, 128, 0, st>>>(A, lda, rows, cols, partial, chunks);
  col_sums_finish_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(partial, cols, chunks, scale, out);
  return check_cuda(cudaGetLastError(), "col_sums");
}

int decomp_row_sums_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, double scale, double* out,
                        void* stream) {
  if (rows <= 0) return DECOMP_OK;
  row_sums_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(A, lda, rows, cols, scale, out);
  DCP_CHECK_LAUNCH("row_sums");
  return DECOMP_OK;
}

int decomp_gershgorin_step_f64(const double* G, int64_t ldg, int64_t k, int32_t is_complex, const double* alpha_scaled,
                               double* step_out, double* thr_out, void* stream) {
  if (k <= 0) return DECOMP_OK;
  gershgorin_kernel<<<1, 1024, 0, as_stream(stream)>>>(G, ldg, (int)k, is_complex, alpha_scaled, step_out, thr_out);
  DCP_CHECK_LAUNCH("gershgorin");
  return DECOMP_OK;
}

int decomp_normalize_rows_f64(const double* D_in, int64_t ldi, int64_t rows, int64_t cols, int32_t is_complex,
                              int32_t strict, double* D_out, int64_t ldo, const double* D_ref, int64_t ldr, double tol,
                              int32_t* tol_latch, int32_t latch_value, double* maxdiff, int32_t* scratch,
                              const int32_t* skip_if, void* stream) {
  if (rows <= 0) return DECOMP_OK;
  if (maxdiff != nullptr && scratch == nullptr) {
    set_error("decomp_normalize_rows_f64: maxdiff needs scratch");
    return DECOMP_ERR_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  if (is_complex) {
    normalize_rows_complex_kernel<<<(unsigned)rows, 256, 0, st>>>(D_in, ldi, cols, strict, D_out, ldo, D_ref, ldr, tol,
                                                                  tol_latch, latch_value, maxdiff, scratch, skip_if);
  } else {
    normalize_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(D_in, ldi, cols, strict, D_out, ldo, D_ref, ldr, tol,
                                                          tol_latch, latch_value, maxdiff, scratch, skip_if);
  }
  DCP_CHECK_LAUNCH("normalize_rows");
  return DECOMP_OK;
}

int decomp_gather_rows_f64(const double* in, int64_t ldi, const int64_t* index, int64_t rows, int64_t cols, double* out,
                           int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  gather_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(
      in, ldi, reinterpret_cast<const long long*>(index), rows, cols, out, ldo);
  DCP_CHECK_LAUNCH("gather_rows");
  return DECOMP_OK;
}

int decomp_scatter_rows_f64(const double* in, int64_t ldi, const int64_t* index, int64_t rows, int64_t cols, double* out,
                            int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  scatter_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(
      in, ldi, reinterpret_cast<const long long*>(index), rows, cols, out, ldo);
  DCP_CHECK_LAUNCH("scatter_rows");
  return DECOMP_OK;
}

int decomp_lasso_vectors_f64(const double* s, int64_t k, double alpha, double tol, double mult, const double* mult_dev,
                             double* alpha_out, double* tol_out, void* stream) {
  if (k <= 0) return DECOMP_OK;
  lasso_vectors_kernel<<<grid_for(k, 256), 256, 0, as_stream(stream)>>>(s, (int)k, alpha, tol, mult, mult_dev,
                                                                       alpha_out, tol_out);
  DCP_CHECK_LAUNCH("lasso_vectors");
  return DECOMP_OK;
}

int decomp_svrmu_update_f64(const double* D, int64_t ldd, const double* P, int64_t ldp, const double* Q, int64_t ldq,
                            double alpha, int64_t rows, int64_t cols, double* out, int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  svrmu_update_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(D, ldd, P, ldp, Q, ldq, alpha, rows,
                                                                                cols, out, ldo);
  DCP_CHECK_LAUNCH("svrmu_update");
  return DECOMP_OK;
}

int decomp_axpby_f64(double a, const double* X, int64_t ldx, double b, const double* Y, int64_t ldy, int64_t rows,
                     int64_t cols, double* out, int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  axpby_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(a, X, ldx, b, Y, ldy, rows, cols, out, ldo);
  DCP_CHECK_LAUNCH("axpby");
  return DECOMP_OK;
}

int decomp_lasso_q_f64(const double* G, int64_t ldg, int64_t k, int32_t is_complex, const double* step, double* Q,
                       int64_t ldq, void* stream) {
  if (k <= 0) return DECOMP_OK;
  const int cw = is_complex ? 2 : 1;
  lasso_q_kernel<<<grid_for(k * k * cw, 256), 256, 0, as_stream(stream)>>>(G, ldg, (int)k, cw, step, Q, ldq);
  DCP_CHECK_LAUNCH("lasso_q");
  return DECOMP_OK;
}

int decomp_scale_scalar_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, const double* scalar_dev,
                            double* out, int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  scale_scalar_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(A, lda, rows, cols, scalar_dev, out,
                                                                                ldo);
  DCP_CHECK_LAUNCH("scale_scalar");
  return DECOMP_OK;
}

int decomp_mu_update_f64(const double* x, int64_t ldx, const double* num, int64_t ldn, const double* den, int64_t ldd,
                         int64_t rows, int64_t cols, double* out, int64_t ldo, const int32_t* skip_if, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  mu_update_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(x, ldx, num, ldn, den, ldd, rows, cols,
                                                                             out, ldo, skip_if);
  DCP_CHECK_LAUNCH("mu_update");
  return DECOMP_OK;
}

int decomp_max_abs_diff_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t rows, int64_t cols,
                            int32_t is_complex, double tol, int32_t* tol_latch, int32_t latch_value, double* result,
                            int32_t* scratch, const int32_t* skip_if, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  if (result == nullptr || scratch == nullptr) {
    set_error("decomp_max_abs_diff_f64: result and scratch are required");
    return DECOMP_ERR_INVALID;
  }
  int blocks = grid_for(rows * cols, 256);
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  max_abs_diff_kernel<<<blocks, 256, 0, as_stream(stream)>>>(A, lda, B, ldb, rows, cols, is_complex, tol, tol_latch,
                                                            latch_value, result, scratch, skip_if);
  DCP_CHECK_LAUNCH("max_abs_diff");
  return DECOMP_OK;
}

}  // extern "C"

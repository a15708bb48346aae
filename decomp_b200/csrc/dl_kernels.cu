// Dictionary (basis) update kernels: the sequential Gauss-Seidel atom sweep as ONE cooperative launch
// (k grid-wide barriers instead of k host round trips), and the masked Jacobi update, which streams the
// [k, f, k] statistics tensor once from HBM.  Reference: decomp/dictionary_learning.py:154-159, 206-222.
#include <cooperative_groups.h>
#include <type_traits>

#include "common.h"

namespace cg = cooperative_groups;

namespace dcp {

constexpr double kEpsDl = 1.0e-15;

__device__ __forceinline__ double warp_sum_dl(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (a + bi) / (c + di)
__device__ __forceinline__ double2 cdiv(double a, double b, double c, double d) {
  const double den = c * c + d * d;
  return make_double2((a * c + b * d) / den, (b * c - a * d) / den);
}

// ------------------------------------------------------------------------------------------------
// Gauss-Seidel sweep.  Block = 32 column lanes x 8 row groups; a block owns column groups
// cgp = blockIdx.x, blockIdx.x + gridDim.x, ...  For atom a:
//   u_j = (T[a][j] - sum_b S[a][b] D[b][j]) / (S[a][a] + eps) + D[a][j]      (D rows < a already updated)
//   D[a][j] = u_j / sqrt(max(sum_j |u_j|^2, 1))
// The only grid-wide dependency is the norm, exchanged through a double-buffered array of block partials.
// ------------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(256) dl_sweep_kernel(const double* __restrict__ S, long long lds,
                                                       const double* __restrict__ T, long long ldt, double* D,
                                                       long long ldd, int k, int f, double* partials) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red_r[8][33];
  __shared__ double red_i[8][33];
  __shared__ double blk[8];
  __shared__ double bcast;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int groups = (f + 31) / 32;
  constexpr int CW = CPLX ? 2 : 1;

  for (int a = 0; a < k; ++a) {
    double local = 0.0;
    const double* Sa = S + (long long)a * lds;
    for (int cgp = blockIdx.x; cgp < groups; cgp += gridDim.x) {
      const int j = cgp * 32 + tx;
      double sr = 0.0, si = 0.0;
      if (j < f) {
        for (int b = ty; b < k; b += 8) {
          if (CPLX) {
            const double2 s = *reinterpret_cast<const double2*>(Sa + 2 * b);
            const double2 d = *reinterpret_cast<const double2*>(D + (long long)b * ldd + 2 * j);
            sr += s.x * d.x - s.y * d.y;
            si += s.x * d.y + s.y * d.x;
          } else {
            sr += Sa[b] * D[(long long)b * ldd + j];
          }
        }
      }
      red_r[ty][tx] = sr;
      if (CPLX) red_i[ty][tx] = si;
      __syncthreads();
      if (ty == 0 && j < f) {
        double dr = 0.0, di = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          dr += red_r[t][tx];
          if (CPLX) di += red_i[t][tx];
        }
        double* dj = D + (long long)a * ldd + CW * j;
        if (CPLX) {
          const double2 t2 = *reinterpret_cast<const double2*>(T + (long long)a * ldt + 2 * j);
          const double2 saa = *reinterpret_cast<const double2*>(Sa + 2 * a);
          const double2 qv = cdiv(t2.x - dr, t2.y - di, saa.x + kEpsDl, saa.y);
          const double ur = qv.x + dj[0], ui = qv.y + dj[1];
          dj[0] = ur;
          dj[1] = ui;
          local += ur * ur + ui * ui;
        } else {
          const double u = (T[(long long)a * ldt + j] - dr) / (Sa[a] + kEpsDl) + dj[0];
          dj[0] = u;
          local += u * u;
        }
      }
      __syncthreads();
    }
    // block partial of |u|^2 (only ty == 0 lanes contribute)
    local = warp_sum_dl(local);
    if (tx == 0) blk[ty] = local;
    __syncthreads();
    if (threadIdx.x == 0) partials[(a & 1) * gridDim.x + blockIdx.x] = blk[0];
    __threadfence();
    grid.sync();
    if (ty == 0) {
      double tot = 0.0;
      for (int b = tx; b < (int)gridDim.x; b += 32) tot += partials[(a & 1) * gridDim.x + b];
      tot = warp_sum_dl(tot);
      if (tx == 0) bcast = sqrt(fmax(tot, 1.0));
    }
    __syncthreads();
    const double nrm = bcast;
    if (ty == 0) {
      for (int cgp = blockIdx.x; cgp < groups; cgp += gridDim.x) {
        const int j = cgp * 32 + tx;
        if (j < f) {
          double* dj = D + (long long)a * ldd + CW * j;
          dj[0] = dj[0] / nrm;
          if (CPLX) dj[1] = dj[1] / nrm;
        }
      }
    }
    __syncthreads();
  }
}

// W[i][b] = conj(x_ia) x_ib
template <bool CPLX>
__global__ void atom_weighted_kernel(const double* __restrict__ X, long long ldx, long long rows, int k, int atom,
                                     double* __restrict__ W, long long ldw) {
  const long long total = rows * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / k;
    const int b = (int)(idx % k);
    if (CPLX) {
      const double2 xa = *reinterpret_cast<const double2*>(X + i * ldx + 2 * atom);
      const double2 xb = *reinterpret_cast<const double2*>(X + i * ldx + 2 * b);
      *reinterpret_cast<double2*>(W + i * ldw + 2 * b) =
          make_double2(xa.x * xb.x + xa.y * xb.y, xa.x * xb.y - xa.y * xb.x);
    } else {
      W[i * ldw + b] = X[i * ldx + atom] * X[i * ldx + b];
    }
  }
}

// Pair products of several atoms side by side, so that one full-width GEMM produces the statistics of all of them.
// Transposed form for the NT GEMM (contraction index contiguous): Xt is the transposed real view of the codes,
// Xt[b*cw + part][i]; Wt[c*cw + part][i] = (conj(x_i,colA[c]) x_i,colB[c]).part -- coalesced along the rows i.
template <bool CPLX>
__global__ void pair_products_t_kernel(const double* __restrict__ Xt, long long ldx, long long rows,
                                       const int* __restrict__ colA, const int* __restrict__ colB, int width,
                                       double* __restrict__ Wt, long long ldw) {
  const int c = blockIdx.y;
  const int a = __ldg(colA + c), b = __ldg(colB + c);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows; i += (long long)gridDim.x * blockDim.x) {
    if (CPLX) {
      const double ar = Xt[(2LL * a) * ldx + i], ai = Xt[(2LL * a + 1) * ldx + i];
      const double br = Xt[(2LL * b) * ldx + i], bi = Xt[(2LL * b + 1) * ldx + i];
      Wt[(2LL * c) * ldw + i] = ar * br + ai * bi;
      Wt[(2LL * c + 1) * ldw + i] = ar * bi - ai * br;
    } else {
      Wt[(long long)c * ldw + i] = Xt[(long long)a * ldx + i] * Xt[(long long)b * ldx + i];
    }
  }
}

// S[colA[c]][j][colB[c]] = beta * S[...] + P[j][c].  S is cut into channel slabs of `fs` channels, each a contiguous
// [k][fs][k] tensor (fs = f: the plain [k][f][k] layout; fs = ceil(f / ranks): the input of a reduce-scatter along f)
template <bool CPLX>
__global__ void scatter_stats_kernel(const double* __restrict__ P, long long ldp, int f, int width,
                                     const int* __restrict__ colA, const int* __restrict__ colB, int k, double beta,
                                     double* __restrict__ S, int fs) {
  constexpr int CW = CPLX ? 2 : 1;
  const long long total = (long long)f * width;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long j = idx / width;
    const int c = (int)(idx % width);
    const int a = __ldg(colA + c), b = __ldg(colB + c);
    double* dst = S + ((((j / fs) * k + a) * fs + j % fs) * k + b) * CW;
    const double* src = P + j * ldp + (long long)c * CW;
    if (CPLX) {
      const double2 o = *reinterpret_cast<const double2*>(dst);
      const double2 v = *reinterpret_cast<const double2*>(src);
      *reinterpret_cast<double2*>(dst) = make_double2(beta * o.x + v.x, beta * o.y + v.y);
    } else {
      dst[0] = beta * dst[0] + src[0];
    }
  }
}

// S[b][j][a] = conj(S[a][j][b]) for b > a: the masked statistics are Hermitian in (a, b), so only b >= a is
// accumulated by the GEMMs and the rest is mirrored.  One block per feature j and 32x32 tile pair of the upper
// triangle, transposed through shared memory so that reads and writes are both 32 contiguous elements per warp.
template <bool CPLX>
__global__ void __launch_bounds__(256) dl_mirror_kernel(double* __restrict__ S, int k, int f) {
  using Elem = typename std::conditional<CPLX, double2, double>::type;
  __shared__ Elem tile[32][33];
  Elem* E = reinterpret_cast<Elem*>(S);
  const long long j = blockIdx.x;
  const int T = (k + 31) / 32;
  int p = blockIdx.y, ta = 0;
  while (p >= T - ta) {
    p -= T - ta;
    ++ta;
  }
  const int tb = ta + p;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = ty + 8 * r;
    const int a = ta * 32 + i, b = tb * 32 + tx;
    if (a < k && b < k) tile[i][tx] = E[((long long)a * f + j) * k + b];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = ty + 8 * r;
    const int b = tb * 32 + i, a = ta * 32 + tx;
    if (b < k && a < k && b > a) {
      Elem v = tile[tx][i];
      if constexpr (CPLX) v.y = -v.y;
      E[((long long)b * f + j) * k + a] = v;
    }
  }
}

// Dt[j][b] = D[b][j]  (elements are complex pairs when CPLX)
template <bool CPLX>
__global__ void transpose_elems_kernel(const double* __restrict__ D, long long ldd, int k, int f,
                                       double* __restrict__ Dt) {
  const long long total = (long long)k * f;
  constexpr int CW = CPLX ? 2 : 1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long j = idx / k;
    const int b = (int)(idx % k);
    Dt[(j * k + b) * CW] = D[(long long)b * ldd + CW * j];
    if (CPLX) Dt[(j * k + b) * CW + 1] = D[(long long)b * ldd + CW * j + 1];
  }
}

// Masked Jacobi update: one CTA per atom a, one warp per column j, lanes over b.
template <bool CPLX>
__global__ void __launch_bounds__(256) dl_masked_update_kernel(const double* __restrict__ S,
                                                               const double* __restrict__ T, long long ldt,
                                                               const double* __restrict__ D, long long ldd,
                                                               const double* __restrict__ Dt, int k, int f,
                                                               double* __restrict__ Dout, long long ldo) {
  constexpr int CW = CPLX ? 2 : 1;
  __shared__ double red[2][8];
  __shared__ double bc[2];
  const int a = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double* Sa = S + (long long)a * f * k * CW;
  double saa_r = 0.0, saa_i = 0.0;
  for (int j = w; j < f; j += 8) {
    const double* srow = Sa + (long long)j * k * CW;
    const double* drow = Dt + (long long)j * k * CW;
    double pr = 0.0, pi = 0.0;
    for (int b = lane; b < k; b += 32) {
      if (CPLX) {
        const double2 s = *reinterpret_cast<const double2*>(srow + 2 * b);
        const double2 d = *reinterpret_cast<const double2*>(drow + 2 * b);
        pr += s.x * d.x - s.y * d.y;
        pi += s.x * d.y + s.y * d.x;
      } else {
        pr += srow[b] * drow[b];
      }
    }
    pr = warp_sum_dl(pr);
    if (CPLX) pi = warp_sum_dl(pi);
    if (lane == 0) {
      // numerator T[a][j] - SaD[j], parked in the output row until Saa is known
      Dout[(long long)a * ldo + CW * j] = T[(long long)a * ldt + CW * j] - pr;
      if (CPLX) Dout[(long long)a * ldo + 2 * j + 1] = T[(long long)a * ldt + 2 * j + 1] - pi;
      saa_r += srow[CW * a] + kEpsDl;
      if (CPLX) saa_i += srow[2 * a + 1];
    }
  }
  if (lane == 0) {
    red[0][w] = saa_r;
    red[1][w] = saa_i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0, i = 0.0;
    for (int t = 0; t < 8; ++t) {
      r += red[0][t];
      i += red[1][t];
    }
    bc[0] = r;
    bc[1] = i;
  }
  __syncthreads();
  const double sr = bc[0], si = bc[1];
  double local = 0.0;
  for (int j = threadIdx.x; j < f; j += blockDim.x) {
    double* o = Dout + (long long)a * ldo + CW * j;
    const double* d = D + (long long)a * ldd + CW * j;
    if (CPLX) {
      const double2 qv = cdiv(o[0], o[1], sr, si);
      const double ur = qv.x + d[0], ui = qv.y + d[1];
      o[0] = ur;
      o[1] = ui;
      local += ur * ur + ui * ui;
    } else {
      const double u = o[0] / sr + d[0];
      o[0] = u;
      local += u * u;
    }
  }
  local = warp_sum_dl(local);
  __syncthreads();
  if (lane == 0) red[0][w] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int t = 0; t < 8; ++t) r += red[0][t];
    bc[0] = sqrt(fmax(r, 1.0));
  }
  __syncthreads();
  const double nrm = bc[0];
  for (int j = threadIdx.x; j < f; j += blockDim.x) {
    double* o = Dout + (long long)a * ldo + CW * j;
    o[0] = o[0] / nrm;
    if (CPLX) o[1] = o[1] / nrm;
  }
}

// The masked Jacobi update on a channel slab (the statistics reduce-scattered along f over the ranks): the same
// arithmetic as dl_masked_update_kernel in three phases, separated by the two [k] sums over all channels that the ranks
// exchange (sum_j S[a][j][a] and |u_a|^2).  S: [k][fs][k] slab holding channels j0 .. j0 + fs; fv valid channels.
// stats[a] = {Re saa, Im saa, |u|^2, -}.  Dt: the slab of D transposed, [fs][k].  One CTA per atom.
template <bool CPLX>
__global__ void __launch_bounds__(256) dl_masked_phase1_kernel(const double* __restrict__ S, int fs, int fv, int j0,
                                                               const double* __restrict__ T, long long ldt,
                                                               const double* __restrict__ Dt, int k,
                                                               double* __restrict__ out, double* __restrict__ stats) {
  constexpr int CW = CPLX ? 2 : 1;
  __shared__ double red[2][8];
  const int a = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double* Sa = S + (long long)a * fs * k * CW;
  double saa_r = 0.0, saa_i = 0.0;
  for (int jj = w; jj < fv; jj += 8) {
    const double* srow = Sa + (long long)jj * k * CW;
    const double* drow = Dt + (long long)jj * k * CW;
    double pr = 0.0, pi = 0.0;
    for (int b = lane; b < k; b += 32) {
      if (CPLX) {
        const double2 sv = *reinterpret_cast<const double2*>(srow + 2 * b);
        const double2 d = *reinterpret_cast<const double2*>(drow + 2 * b);
        pr += sv.x * d.x - sv.y * d.y;
        pi += sv.x * d.y + sv.y * d.x;
      } else {
        pr += srow[b] * drow[b];
      }
    }
    pr = warp_sum_dl(pr);
    if (CPLX) pi = warp_sum_dl(pi);
    if (lane == 0) {
      out[((long long)a * fs + jj) * CW] = T[(long long)a * ldt + CW * (j0 + jj)] - pr;
      if (CPLX) out[((long long)a * fs + jj) * 2 + 1] = T[(long long)a * ldt + 2 * (j0 + jj) + 1] - pi;
      saa_r += srow[CW * a] + kEpsDl;
      if (CPLX) saa_i += srow[2 * a + 1];
    }
  }
  if (lane == 0) {
    red[0][w] = saa_r;
    red[1][w] = saa_i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0, i = 0.0;
    for (int t = 0; t < 8; ++t) {
      r += red[0][t];
      i += red[1][t];
    }
    stats[4 * a] = r;
    stats[4 * a + 1] = i;
    stats[4 * a + 2] = 0.0;
    stats[4 * a + 3] = 0.0;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) dl_masked_phase2_kernel(int fs, int fv, int j0, const double* __restrict__ D,
                                                               long long ldd, double* __restrict__ out,
                                                               double* __restrict__ stats) {
  constexpr int CW = CPLX ? 2 : 1;
  __shared__ double red[8];
  const int a = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const double sr = stats[4 * a], si = stats[4 * a + 1];
  double local = 0.0;
  for (int jj = threadIdx.x; jj < fv; jj += blockDim.x) {
    double* o = out + ((long long)a * fs + jj) * CW;
    const double* d = D + (long long)a * ldd + CW * (j0 + jj);
    if (CPLX) {
      const double2 qv = cdiv(o[0], o[1], sr, si);
      const double ur = qv.x + d[0], ui = qv.y + d[1];
      o[0] = ur;
      o[1] = ui;
      local += ur * ur + ui * ui;
    } else {
      const double u = o[0] / sr + d[0];
      o[0] = u;
      local += u * u;
    }
  }
  local = warp_sum_dl(local);
  if (lane == 0) red[w] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
    for (int t = 0; t < 8; ++t) r += red[t];
    stats[4 * a + 2] = r;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) dl_masked_phase3_kernel(int fs, int fv, double* __restrict__ out,
                                                               const double* __restrict__ stats) {
  constexpr int CW = CPLX ? 2 : 1;
  const int a = blockIdx.x;
  const double nrm = sqrt(fmax(stats[4 * a + 2], 1.0));
  for (int jj = threadIdx.x; jj < fv; jj += blockDim.x) {
    double* o = out + ((long long)a * fs + jj) * CW;
    o[0] = o[0] / nrm;
    if (CPLX) o[1] = o[1] / nrm;
  }
}

// ------------------------------------------------------------------------------------------------
// Gauss-Seidel sweep, slice-resident version: a block owns w consecutive columns of D and keeps its [k, w] slice in
// shared memory for the whole sweep, so an atom costs one row of S (prefetched one atom ahead), k*w multiply-adds out
// of shared memory and ONE grid-wide exchange (the norm of u).  The exchange is a hand-rolled barrier on a
// monotonic counter plus block partials summed in block order by everybody (deterministic).  Thread = (column jj,
// row group bg): bg strides the atoms b = bg, bg + NBG, ...
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool CPLX>
__global__ void __launch_bounds__(256) dl_sweep_slice_kernel(const double* __restrict__ S, long long lds,
                                                             const double* __restrict__ T, long long ldt, double* D,
                                                             long long ldd, int k, int f, int w, int wpad,
                                                             double* partials, unsigned* counter) {
  using Elem = typename std::conditional<CPLX, double2, double>::type;
  constexpr int CW = CPLX ? 2 : 1;
  extern __shared__ __align__(16) unsigned char dl_smem[];
  Elem* Ds = reinterpret_cast<Elem*>(dl_smem);                       // [k][w]
  Elem* Sr = Ds + (size_t)k * w;                                     // [2][k]
  Elem* red = Sr + 2 * (size_t)k;                                    // [NBG][wpad]
  __shared__ double blk[8];
  __shared__ double bcast;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbg = 256 / wpad, jj = tid % wpad, bg = tid / wpad;
  const int j0 = blockIdx.x * w;
  const int wv = min(w, f - j0);                                     // valid columns of this slice (> 0)
  const unsigned nblk = gridDim.x;

  for (int idx = tid; idx < k * w; idx += 256) {
    const int b = idx / w, c = idx % w;
    Elem v;
    if constexpr (CPLX) v = make_double2(0.0, 0.0); else v = 0.0;
    if (c < wv) v = *reinterpret_cast<const Elem*>(D + (long long)b * ldd + (long long)CW * (j0 + c));
    Ds[idx] = v;
  }
  for (int b = tid; b < k; b += 256) Sr[b] = *reinterpret_cast<const Elem*>(S + CW * b);
  __syncthreads();

  for (int a = 0; a < k; ++a) {
    const Elem* sa = Sr + (size_t)(a & 1) * k;
    // next atom's row of S (does not depend on this atom's update)
    if (a + 1 < k) {
      Elem* sn = Sr + (size_t)((a + 1) & 1) * k;
      for (int b = tid; b < k; b += 256) sn[b] = *reinterpret_cast<const Elem*>(S + (long long)(a + 1) * lds + CW * b);
    }
    double sr = 0.0, si = 0.0;
    if (jj < wv) {
      for (int b = bg; b < k; b += nbg) {
        if constexpr (CPLX) {
          const double2 s = sa[b], d = Ds[(size_t)b * w + jj];
          sr += s.x * d.x - s.y * d.y;
          si += s.x * d.y + s.y * d.x;
        } else {
          sr += sa[b] * Ds[(size_t)b * w + jj];
        }
      }
    }
    if constexpr (CPLX) red[bg * wpad + jj] = make_double2(sr, si); else red[bg * wpad + jj] = sr;
    __syncthreads();
    double local = 0.0;
    if (bg == 0 && jj < wv) {
      double dr = 0.0, di = 0.0;
      for (int t = 0; t < nbg; ++t) {
        if constexpr (CPLX) {
          const double2 r = red[t * wpad + jj];
          dr += r.x;
          di += r.y;
        } else {
          dr += red[t * wpad + jj];
        }
      }
      const long long j = j0 + jj;
      if constexpr (CPLX) {
        const double2 t2 = *reinterpret_cast<const double2*>(T + (long long)a * ldt + 2 * j);
        const double2 saa = sa[a], dj = Ds[(size_t)a * w + jj];
        const double2 qv = cdiv(t2.x - dr, t2.y - di, saa.x + kEpsDl, saa.y);
        const double ur = qv.x + dj.x, ui = qv.y + dj.y;
        Ds[(size_t)a * w + jj] = make_double2(ur, ui);
        local = ur * ur + ui * ui;
      } else {
        const double u = (T[(long long)a * ldt + j] - dr) / (sa[a] + kEpsDl) + Ds[(size_t)a * w + jj];
        Ds[(size_t)a * w + jj] = u;
        local = u * u;
      }
    }
    // |u|^2 of this slice -> partials, grid-wide exchange
    local = warp_sum_dl(local);
    if (lane == 0) blk[warp] = local;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int t = 0; t < 8; ++t) tot += blk[t];
      partials[(size_t)(a & 1) * nblk + blockIdx.x] = tot;
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned target = (unsigned)(a + 1) * nblk;
      while (ld_acquire_u32(counter) < target) {
      }
    }
    __syncthreads();
    if (warp == 0) {
      double tot = 0.0;
      const double* pp = partials + (size_t)(a & 1) * nblk;
      for (unsigned b = lane; b < nblk; b += 32) tot += __ldcg(pp + b);
      tot = warp_sum_dl(tot);
      if (lane == 0) bcast = sqrt(fmax(tot, 1.0));
    }
    __syncthreads();
    const double nrm = bcast;
    if (bg == 0 && jj < wv) {
      if constexpr (CPLX) {
        double2 v = Ds[(size_t)a * w + jj];
        Ds[(size_t)a * w + jj] = make_double2(v.x / nrm, v.y / nrm);
      } else {
        Ds[(size_t)a * w + jj] = Ds[(size_t)a * w + jj] / nrm;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < k * w; idx += 256) {
    const int b = idx / w, c = idx % w;
    if (c < wv) *reinterpret_cast<Elem*>(D + (long long)b * ldd + (long long)CW * (j0 + c)) = Ds[idx];
  }
}

}  // namespace dcp

using namespace dcp;

extern "C" {

size_t decomp_dl_sweep_workspace_bytes(int64_t k, int64_t f, int32_t is_complex) {
  (void)k;
  (void)is_complex;
  if (f <= 0) return 0;
  // two partial sums per block (slice kernel: <= f blocks; L2-streaming kernel: <= 8 blocks per SM) + the counter
  const long long cap = 8LL * num_sms();
  const long long blocks = f > cap ? f : cap;
  return sizeof(double) * (size_t)(2 * blocks + 2);
}

int decomp_dl_sweep_f64(const double* S, int64_t lds, const double* T, int64_t ldt, double* D, int64_t ldd, int64_t k,
                        int64_t f, int32_t is_complex, void* workspace, size_t workspace_bytes, void* stream) {
  if (k <= 0 || f <= 0) return DECOMP_OK;
  if (workspace == nullptr || workspace_bytes < decomp_dl_sweep_workspace_bytes(k, f, is_complex)) {
    set_error("decomp_dl_sweep_f64: workspace too small (%zu < %zu)", workspace_bytes,
              decomp_dl_sweep_workspace_bytes(k, f, is_complex));
    return DECOMP_ERR_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  {
    // slice-resident sweep when a block's [k, w] slice of D (+ two rows of S + the reduction scratch) fits
    const int cw = is_complex ? 2 : 1;
    const int sms = num_sms();
    const long long w = (f + sms - 1) / sms;
    int wpad = 1;
    while (wpad < w) wpad <<= 1;
    const long long blocks = (f + w - 1) / w;
    const size_t smem = (size_t)cw * 8 * ((size_t)k * w + 2 * (size_t)k + 256);
    if (wpad <= 256 && smem <= 200 * 1024) {
      void* kern = is_complex ? (void*)dl_sweep_slice_kernel<true> : (void*)dl_sweep_slice_kernel<false>;
      int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "dl_sweep smem");
      if (rc != DECOMP_OK) return rc;
      double* scratch = reinterpret_cast<double*>(workspace);   // 2 * blocks partials, then the barrier counter
      unsigned* counter = reinterpret_cast<unsigned*>(scratch + 2 * blocks);
      cudaMemsetAsync(counter, 0, sizeof(double), st);
      long long lds_ = lds, ldt_ = ldt, ldd_ = ldd;
      int k_ = (int)k, f_ = (int)f, w_ = (int)w;
      void* args[] = {(void*)&S, (void*)&lds_, (void*)&T, (void*)&ldt_, (void*)&D, (void*)&ldd_, (void*)&k_, (void*)&f_,
                      (void*)&w_, (void*)&wpad, (void*)&scratch, (void*)&counter};
      cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(256), args, smem, st);
      return check_cuda(e, "dl_sweep slice launch");
    }
  }
  void* kern = is_complex ? (void*)dl_sweep_kernel<true> : (void*)dl_sweep_kernel<false>;
  int per_sm = 0;
  int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0), "dl_sweep occupancy");
  if (rc != DECOMP_OK) return rc;
  if (per_sm < 1) per_sm = 1;
  int groups = (int)((f + 31) / 32);
  int blocks = groups;
  const int cap = per_sm * num_sms();
  if (blocks > cap) blocks = cap;
  double* partials = reinterpret_cast<double*>(workspace);
  long long lds_ = lds, ldt_ = ldt, ldd_ = ldd;
  int k_ = (int)k, f_ = (int)f;
  void* args[] = {(void*)&S, (void*)&lds_, (void*)&T, (void*)&ldt_, (void*)&D, (void*)&ldd_, (void*)&k_, (void*)&f_,
                  (void*)&partials};
  cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(256), args, 0, st);
  return check_cuda(e, "dl_sweep launch");
}

int decomp_dl_atom_weighted_f64(const double* X, int64_t ldx, int64_t rows, int64_t k, int32_t is_complex, int64_t atom,
                                double* W, int64_t ldw, void* stream) {
  if (rows <= 0 || k <= 0) return DECOMP_OK;
  long long total = rows * k;
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (is_complex)
    atom_weighted_kernel<true><<<(unsigned)b, 256, 0, as_stream(stream)>>>(X, ldx, rows, (int)k, (int)atom, W, ldw);
  else
    atom_weighted_kernel<false><<<(unsigned)b, 256, 0, as_stream(stream)>>>(X, ldx, rows, (int)k, (int)atom, W, ldw);
  DCP_CHECK_LAUNCH("dl_atom_weighted");
  return DECOMP_OK;
}

int decomp_dl_pair_products_t_f64(const double* Xt, int64_t ldx, int64_t rows, int32_t is_complex, const int32_t* colA,
                                  const int32_t* colB, int64_t width, double* Wt, int64_t ldw, void* stream) {
  if (rows <= 0 || width <= 0) return DECOMP_OK;
  long long bx = (rows + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)width);
  if (is_complex)
    pair_products_t_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(Xt, ldx, rows, colA, colB, (int)width, Wt, ldw);
  else
    pair_products_t_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(Xt, ldx, rows, colA, colB, (int)width, Wt, ldw);
  DCP_CHECK_LAUNCH("dl_pair_products_t");
  return DECOMP_OK;
}

int decomp_dl_scatter_stats_f64(const double* P, int64_t ldp, int64_t f, int64_t width, int32_t is_complex,
                                const int32_t* colA, const int32_t* colB, int64_t k, double beta, double* S,
                                int64_t slab_channels, void* stream) {
  if (f <= 0 || width <= 0) return DECOMP_OK;
  const int fs = (int)(slab_channels > 0 ? slab_channels : f);
  long long b = (f * width + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (is_complex)
    scatter_stats_kernel<true><<<(unsigned)b, 256, 0, as_stream(stream)>>>(P, ldp, (int)f, (int)width, colA, colB, (int)k,
                                                                           beta, S, fs);
  else
    scatter_stats_kernel<false><<<(unsigned)b, 256, 0, as_stream(stream)>>>(P, ldp, (int)f, (int)width, colA, colB, (int)k,
                                                                            beta, S, fs);
  DCP_CHECK_LAUNCH("dl_scatter_stats");
  return DECOMP_OK;
}

int decomp_dl_mirror_f64(double* S, int64_t k, int64_t f, int32_t is_complex, void* stream) {
  if (k <= 1 || f <= 0) return DECOMP_OK;
  const long long T = (k + 31) / 32, pairs = T * (T + 1) / 2;
  if (f > 2147483647LL || pairs > 65535) {
    set_error("decomp_dl_mirror_f64: too many tiles (k <= 11552)");
    return DECOMP_ERR_INVALID;
  }
  const dim3 grid((unsigned)f, (unsigned)pairs), block(32, 8);
  if (is_complex)
    dl_mirror_kernel<true><<<grid, block, 0, as_stream(stream)>>>(S, (int)k, (int)f);
  else
    dl_mirror_kernel<false><<<grid, block, 0, as_stream(stream)>>>(S, (int)k, (int)f);
  DCP_CHECK_LAUNCH("dl_mirror");
  return DECOMP_OK;
}

int decomp_dl_masked_update_f64(const double* S, const double* T, int64_t ldt, const double* D, int64_t ldd, int64_t k,
                                int64_t f, int32_t is_complex, double* D_out, int64_t ldo, double* workspace,
                                void* stream) {
  if (k <= 0 || f <= 0) return DECOMP_OK;
  if (workspace == nullptr) {
    set_error("decomp_dl_masked_update_f64: workspace (f*k*cw doubles) required");
    return DECOMP_ERR_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  long long total = k * f;
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (is_complex) {
    transpose_elems_kernel<true><<<(unsigned)b, 256, 0, st>>>(D, ldd, (int)k, (int)f, workspace);
    dl_masked_update_kernel<true><<<(unsigned)k, 256, 0, st>>>(S, T, ldt, D, ldd, workspace, (int)k, (int)f, D_out, ldo);
  } else {
    transpose_elems_kernel<false><<<(unsigned)b, 256, 0, st>>>(D, ldd, (int)k, (int)f, workspace);
    dl_masked_update_kernel<false><<<(unsigned)k, 256, 0, st>>>(S, T, ldt, D, ldd, workspace, (int)k, (int)f, D_out, ldo);
  }
  DCP_CHECK_LAUNCH("dl_masked_update");
  return DECOMP_OK;
}

int decomp_dl_masked_update_phase_f64(int32_t phase, const double* S_slab, int64_t slab_channels, int64_t j0,
                                      const double* T, int64_t ldt, const double* D, int64_t ldd, int64_t k, int64_t f,
                                      int32_t is_complex, double* D_slab_out, double* stats, double* workspace,
                                      void* stream) {
  if (k <= 0 || f <= 0 || slab_channels <= 0) return DECOMP_OK;
  if (phase < 1 || phase > 3 || D_slab_out == nullptr || stats == nullptr || (phase == 1 && workspace == nullptr)) {
    set_error("decomp_dl_masked_update_phase_f64: invalid argument");
    return DECOMP_ERR_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  const int fs = (int)slab_channels;
  long long fvl = f - j0;
  if (fvl > fs) fvl = fs;
  if (fvl < 0) fvl = 0;
  const int fv = (int)fvl, cw = is_complex ? 2 : 1;
  if (phase == 1) {
    if (fv > 0) {
      long long b = ((long long)k * fv + 255) / 256;
      const long long cap = (long long)num_sms() * 16;
      if (b > cap) b = cap;
      if (is_complex)
        transpose_elems_kernel<true><<<(unsigned)b, 256, 0, st>>>(D + cw * j0, ldd, (int)k, fv, workspace);
      else
        transpose_elems_kernel<false><<<(unsigned)b, 256, 0, st>>>(D + cw * j0, ldd, (int)k, fv, workspace);
    }
    if (is_complex)
      dl_masked_phase1_kernel<true><<<(unsigned)k, 256, 0, st>>>(S_slab, fs, fv, (int)j0, T, ldt, workspace, (int)k,
                                                                  D_slab_out, stats);
    else
      dl_masked_phase1_kernel<false><<<(unsigned)k, 256, 0, st>>>(S_slab, fs, fv, (int)j0, T, ldt, workspace, (int)k,
                                                                   D_slab_out, stats);
  } else if (phase == 2) {
    if (is_complex)
      dl_masked_phase2_kernel<true><<<(unsigned)k, 256, 0, st>>>(fs, fv, (int)j0, D, ldd, D_slab_out, stats);
    else
      dl_masked_phase2_kernel<false><<<(unsigned)k, 256, 0, st>>>(fs, fv, (int)j0, D, ldd, D_slab_out, stats);
  } else {
    if (is_complex)
      dl_masked_phase3_kernel<true><<<(unsigned)k, 256, 0, st>>>(fs, fv, D_slab_out, stats);
    else
      dl_masked_phase3_kernel<false><<<(unsigned)k, 256, 0, st>>>(fs, fv, D_slab_out, stats);
  }
  DCP_CHECK_LAUNCH("dl_masked_update_phase");
  return DECOMP_OK;
}

}  // extern "C"

// FP64 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128-byte swizzle) feeds a
// multi-stage shared-memory ring guarded by mbarriers; one producer warp issues the copies,
// the consumer warps run DMMA.8x8x4 out of conflict-free swizzled LDS.64 reads and apply the
// fused epilogue (MU ratio / mask / ISTA-FISTA proximal step) straight from registers.
//
//   NT: acc[m][n] = sum_k A[m][k] B[n][k]   both operands K-contiguous   (y.dot(d.T), x.dot(G), ...)
//   TN: acc[m][n] = sum_k A[k][m] B[k][n]   both operands M/N-contiguous (x.T.dot(y), contraction over samples)
//
// Shared-memory layouts (BK = 16 doubles = one 128-byte swizzle row):
//   NT tile: [rows][16]      one TMA box {16, rows};            element (r, kk) at r*128 + (((kk>>1)^(r&7))<<4) + (kk&1)*8
//   TN tile: [rows/16][16][16] one TMA box {16, 16} per 16 rows; element (kk, m) at (m>>4)*2048 + kk*128 +
//                                                                 ((((m&15)>>1)^(kk&7))<<4) + (m&1)*8
// The k index each lane feeds to a given DMMA is permuted (identically for A and B, so the sum is
// unchanged) such that every half-warp LDS.64 touches 16 distinct 8-byte banks:
//   NT: step s, lane q -> kk = 2*(s + 4*(q>>1)) + (q&1)
//   TN: step s, lane q -> kk = 2*q + (s&1) + 8*(s>>1)
#pragma once
#include "../../include/decomp_b200.h"
#include "ptx.cuh"

namespace dcp {

constexpr int BK = 16;
constexpr double kEps = 1.0e-15;  // the reference's _JITTER

struct GemmGeom {
  long long M, N, K;
  int tiles_m, tiles_n, splits, kblocks_per_split, kblocks_total;
  long long ld_partial;  // TN: leading dimension of one partial slab (even)
};

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int NCONS = WARPS_M * WARPS_N;
  static constexpr int THREADS = (NCONS + 1) * 32;
  static constexpr int MI = WM / 8, NJ = WN / 8;
  static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8;
  static_assert(WM % 16 == 0 && WN % 16 == 0, "warp tile must be a multiple of 16 (TN sub-boxes)");
  static_assert(BM % 16 == 0 && BN % 16 == 0, "CTA tile must be a multiple of 16");
};

// --------------------------------------------------------------------------------------------------
// Epilogues.  Each handles the adjacent column pair (col, col+1) a lane owns in one 8x8 DMMA tile.
// --------------------------------------------------------------------------------------------------
struct Pair {
  double a, b;
};

__device__ __forceinline__ Pair ld_pair(const double* p, bool two) {
  Pair r;
  if (two) {
    double2 v = *reinterpret_cast<const double2*>(p);
    r.a = v.x;
    r.b = v.y;
  } else {
    r.a = p[0];
    r.b = 0.0;
  }
  return r;
}

__device__ __forceinline__ void st_pair(double* p, double a, double b, bool two) {
  if (two) {
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
  } else {
    p[0] = a;
  }
}

// Each epilogue processes one output row of a warp tile at a time: the NJ column pairs (col = cbase + 8 j,
// col + 1) a lane owns in that row.  All global loads of the row are issued before the first store so that
// they are in flight together (the operands may alias the outputs -- in-place updates -- so the compiler
// cannot hoist loads over stores by itself).
template <int KIND, int NJ>
struct Epilogue {
  // returns true if the convergence test is violated by any pair of this row (PROX with check only)
  static __device__ __forceinline__ bool row(const decomp_epilogue_t& ep, long long row, long long cbase, long long N,
                                             const double (&acc)[NJ][2]) {
    bool ok[NJ], two[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      ok[j] = cbase + 8 * j < N;
      two[j] = cbase + 8 * j + 1 < N;
    }
    if constexpr (KIND == DECOMP_EPI_STORE) {
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (ok[j]) st_pair(ep.out + row * ep.ldo + cbase + 8 * j, acc[j][0], acc[j][1], two[j]);
    } else if constexpr (KIND == DECOMP_EPI_STORE_MASK) {
      Pair m[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        const long long col = cbase + 8 * j;
        if (ep.cwidth == 2) {
          m[j].a = m[j].b = ep.mask[row * ep.ldmask + (col >> 1)];
        } else {
          m[j] = ld_pair(ep.mask + row * ep.ldmask + col, two[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (ok[j]) st_pair(ep.out + row * ep.ldo + cbase + 8 * j, acc[j][0] * m[j].a, acc[j][1] * m[j].b, two[j]);
    } else if constexpr (KIND == DECOMP_EPI_MU_NUM || KIND == DECOMP_EPI_MU_DEN) {
      Pair x[NJ], o[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        const long long col = cbase + 8 * j;
        x[j] = ld_pair(ep.x + row * ep.ldx + col, two[j]);
        o[j] = ld_pair(ep.other + row * ep.ldother + col, two[j]);
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        double n0, n1, d0, d1;
        if constexpr (KIND == DECOMP_EPI_MU_NUM) {
          n0 = acc[j][0]; n1 = acc[j][1]; d0 = o[j].a; d1 = o[j].b;
        } else {
          n0 = o[j].a; n1 = o[j].b; d0 = acc[j][0]; d1 = acc[j][1];
        }
        // x * max(pos, 0) / max(neg, eps), evaluated left to right like the reference (grads.py:84,93)
        const double r0 = __ddiv_rn(__dmul_rn(x[j].a, fmax(n0, 0.0)), fmax(d0, kEps));
        const double r1 = __ddiv_rn(__dmul_rn(x[j].b, fmax(n1, 0.0)), fmax(d1, kEps));
        st_pair(ep.out + row * ep.ldo + cbase + 8 * j, r0, r1, two[j]);
      }
    } else if constexpr (KIND == DECOMP_EPI_KL_RATIO) {
      Pair y[NJ], m[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        const long long col = cbase + 8 * j;
        y[j] = ld_pair(ep.other + row * ep.ldother + col, two[j]);
        if (ep.mask != nullptr) m[j] = ld_pair(ep.mask + row * ep.ldmask + col, two[j]);
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        if (ep.mask != nullptr) {
          y[j].a *= m[j].a;
          y[j].b *= m[j].b;
        }
        st_pair(ep.out + row * ep.ldo + cbase + 8 * j, y[j].a / (acc[j][0] + kEps), y[j].b / (acc[j][1] + kEps),
                two[j]);
      }
    } else if constexpr (KIND == DECOMP_EPI_PROX) {
      const double step = *ep.step;
      const double rowfac = ep.rowvec != nullptr ? ep.rowvec[row] : 1.0;
      Pair w[NJ], ya[NJ], xp[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        const long long col = cbase + 8 * j;
        w[j] = ld_pair(ep.x + row * ep.ldx + col, two[j]);
        ya[j] = ld_pair(ep.other + row * ep.ldother + col, two[j]);
        xp[j] = ld_pair(ep.prev + row * ep.ldprev + col, two[j]);
      }
      bool bad = false;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (!ok[j]) continue;
        const long long col = cbase + 8 * j;
        // z = w + step * (yAt - w.G)   (lasso.py:245-246)
        const double z0 = w[j].a + step * (ya[j].a - acc[j][0]);
        const double z1 = w[j].b + step * (ya[j].b - acc[j][1]);
        // threshold step * alpha  (lasso.py:287; full mask: alpha carries the per-problem mask count, :163);
        // the per-column vectors are a few kB and stay in L1
        Pair thr, tol;
        tol.a = tol.b = 0.0;
        if (ep.shrink == DECOMP_SHRINK_COMPLEX) {
          thr.a = thr.b = __ldg(ep.colvec + (col >> 1));
          if (ep.check) tol.a = __ldg(ep.colvec2 + (col >> 1));
        } else {
          thr.a = __ldg(ep.colvec + col);
          thr.b = two[j] ? __ldg(ep.colvec + col + 1) : 0.0;
          if (ep.check) {
            tol.a = __ldg(ep.colvec2 + col);
            tol.b = two[j] ? __ldg(ep.colvec2 + col + 1) : 0.0;
          }
        }
        const double t0 = ep.rowvec != nullptr ? step * (thr.a * rowfac) : step * thr.a;
        const double t1 = ep.rowvec != nullptr ? step * (thr.b * rowfac) : step * thr.b;
        double x0, x1;
        if (ep.shrink == DECOMP_SHRINK_COMPLEX) {
          const double r = hypot(z0, z1);
          const double den = r + kEps;
          const double mag = fmax(r - t0, 0.0);
          x0 = mag * (z0 / den);
          x1 = mag * (z1 / den);
          if (ep.check) bad = bad || !(hypot(x0 - xp[j].a, x1 - xp[j].b) - tol.a < 0.0);
        } else {
          if (ep.shrink == DECOMP_SHRINK_POSITIVE) {
            x0 = fmax(z0 - t0, 0.0);
            x1 = fmax(z1 - t1, 0.0);
          } else {
            // max(|z| - t, 0) * sign(z)   (lasso.py:206-207)
            const double s0 = (z0 > 0.0) ? 1.0 : ((z0 < 0.0) ? -1.0 : z0);
            const double s1 = (z1 > 0.0) ? 1.0 : ((z1 < 0.0) ? -1.0 : z1);
            x0 = fmax(fabs(z0) - t0, 0.0) * s0;
            x1 = fmax(fabs(z1) - t1, 0.0) * s1;
          }
          if (ep.check) {
            bad = bad || !(fabs(x0 - xp[j].a) - tol.a < 0.0);
            if (two[j]) bad = bad || !(fabs(x1 - xp[j].b) - tol.b < 0.0);
          }
        }
        st_pair(ep.out + row * ep.ldo + col, x0, x1, two[j]);
        if (ep.out2 != nullptr) {
          // w_next = x_new + momentum * (x_new - x_prev)   (lasso.py:412)
          st_pair(ep.out2 + row * ep.ldo2 + col, x0 + ep.momentum * (x0 - xp[j].a),
                  x1 + ep.momentum * (x1 - xp[j].b), two[j]);
        }
      }
      return bad;
    }
    return false;
  }
};

constexpr int EPI_PARTIAL = 100;  // TN split-K: raw partial tile into the workspace slab of this split

// --------------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------------
template <class C, bool TN, int EPI>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
gemm_f64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmGeom gs,
                const decomp_epilogue_t ep, double* __restrict__ partial, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;

  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int tile = blockIdx.x;
  const int tn = tile % gs.tiles_n;
  tile /= gs.tiles_n;
  const int tm = tile % gs.tiles_m;
  const int z = tile / gs.tiles_m;
  const int m0 = tm * C::BM, n0 = tn * C::BN;
  const int kb0 = z * gs.kblocks_per_split;
  int nkb = gs.kblocks_total - kb0;
  if (nkb > gs.kblocks_per_split) nkb = gs.kblocks_per_split;
  if (nkb < 0) nkb = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::NCONS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  bool violated = false;

  if (warp == C::NCONS) {
    // ---------------------------------------------------------------- TMA producer (one lane)
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
        unsigned char* sa = smem + s * C::STAGE_BYTES;
        unsigned char* sb = sa + C::A_BYTES;
        const int k0 = (kb0 + i) * BK;
        if constexpr (!TN) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int b = 0; b < C::BM / 16; ++b) tma_load_2d(sa + b * 2048, &tmA, &full_bar[s], m0 + 16 * b, k0);
#pragma unroll
          for (int b = 0; b < C::BN / 16; ++b) tma_load_2d(sb + b * 2048, &tmB, &full_bar[s], n0 + 16 * b, k0);
        }
        if (++s == C::STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- DMMA consumers
    const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
    const int g = lane >> 2, q = lane & 3;

    double acc[C::MI][C::NJ][2];
#pragma unroll
    for (int i = 0; i < C::MI; ++i)
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // per-lane byte offsets inside a stage for the four k-steps of one 16-wide k-block
    int offA[4], offB[4];
    if constexpr (!TN) {
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int o = (((s4 + 4 * (q >> 1)) ^ g) << 4) | ((q & 1) << 3);
        offA[s4] = (wm * C::WM + g) * 128 + o;
        offB[s4] = C::A_BYTES + (wn * C::WN + g) * 128 + o;
      }
    } else {
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int kk = 2 * q + (s4 & 1) + 8 * (s4 >> 1);
        // the (i & 1) dependent part of the swizzle is added in the loop (chunk = 4*(i&1) + (g>>1))
        offA[s4] = (wm * C::WM / 16) * 2048 + kk * 128 + ((g & 1) << 3);
        offB[s4] = C::A_BYTES + (wn * C::WN / 16) * 2048 + kk * 128 + ((g & 1) << 3);
      }
    }

    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(&full_bar[s], ph);
      const unsigned char* st = smem + s * C::STAGE_BYTES;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        double a[C::MI], b[C::NJ];
        if constexpr (!TN) {
#pragma unroll
          for (int i = 0; i < C::MI; ++i) a[i] = *reinterpret_cast<const double*>(st + offA[s4] + i * 1024);
#pragma unroll
          for (int j = 0; j < C::NJ; ++j) b[j] = *reinterpret_cast<const double*>(st + offB[s4] + j * 1024);
        } else {
          const int kx = 2 * q + (s4 & 1);  // kk & 7
#pragma unroll
          for (int i = 0; i < C::MI; ++i)
            a[i] = *reinterpret_cast<const double*>(st + offA[s4] + (i >> 1) * 2048 +
                                                    (((4 * (i & 1) + (g >> 1)) ^ kx) << 4));
#pragma unroll
          for (int j = 0; j < C::NJ; ++j)
            b[j] = *reinterpret_cast<const double*>(st + offB[s4] + (j >> 1) * 2048 +
                                                    (((4 * (j & 1) + (g >> 1)) ^ kx) << 4));
        }
#pragma unroll
        for (int i = 0; i < C::MI; ++i)
#pragma unroll
          for (int j = 0; j < C::NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == C::STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }

    // ---------------------------------------------------------------- epilogue from registers
    const long long rbase = (long long)m0 + wm * C::WM + g;
    const long long cbase = (long long)n0 + wn * C::WN + 2 * q;
#pragma unroll
    for (int i = 0; i < C::MI; ++i) {
      const long long row = rbase + 8 * i;
      if (row < gs.M) {
        if constexpr (EPI == EPI_PARTIAL) {
#pragma unroll
          for (int j = 0; j < C::NJ; ++j) {
            const long long col = cbase + 8 * j;
            if (col < gs.N)
              st_pair(partial + ((long long)z * gs.M + row) * gs.ld_partial + col, acc[i][j][0], acc[i][j][1],
                      col + 1 < gs.N);
          }
        } else {
          violated |= Epilogue<EPI, C::NJ>::row(ep, row, cbase, gs.N, acc[i]);
        }
      }
    }
  }

  if constexpr (EPI == DECOMP_EPI_PROX) {
    // convergence latch (lasso.py:293/409): the last CTA to finish decides for the whole batch
    if (ep.check) {
      const int any = __syncthreads_or(violated ? 1 : 0);
      if (threadIdx.x == 0) {
        if (any) atomicOr(&ep.scratch[0], 1);
        __threadfence();
        const int ticket = atomicAdd(&ep.scratch[1], 1);
        if (ticket == (int)gridDim.x - 1) {
          __threadfence();
          const int v = atomicOr(&ep.scratch[0], 0);
          if (v == 0) *ep.latch = ep.latch_value;
          ep.scratch[0] = 0;
          ep.scratch[1] = 0;
        }
      }
    }
  }
}

// out = [beta * out +] sum_z partial[z]   (fixed summation order -> bitwise reproducible)
template <int COMBINE>
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int splits, long long M, long long N,
                                       long long ldp, double* __restrict__ out, long long ldo, double beta,
                                       const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  if constexpr (COMBINE == 0 || COMBINE == 1) {
    const long long total = M * N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const long long m = idx / N, n = idx % N;
      double v = 0.0;
      for (int z = 0; z < splits; ++z) v += partial[((long long)z * M + m) * ldp + n];
      double* o = out + m * ldo + n;
      *o = (COMBINE == 1) ? beta * (*o) + v : v;
    }
  } else {
    // complex: out[i][j] = sum_k conj(a_ki) b_kj from the real [M, N] = [2*Mi, 2*Nj] product
    const long long Mi = M / 2, Nj = N / 2;
    const long long total = Mi * Nj;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const long long i = idx / Nj, j = idx % Nj;
      double rr = 0.0, ii = 0.0, ri = 0.0, ir = 0.0;
      for (int z = 0; z < splits; ++z) {
        const double* p0 = partial + ((long long)z * M + 2 * i) * ldp + 2 * j;
        const double* p1 = p0 + ldp;
        rr += p0[0];
        ri += p0[1];
        ir += p1[0];
        ii += p1[1];
      }
      const double re = rr + ii, im = ri - ir;
      double* o = out + i * ldo + 2 * j;
      if (COMBINE == 3) {
        o[0] = beta * o[0] + re;
        o[1] = beta * o[1] + im;
      } else {
        o[0] = re;
        o[1] = im;
      }
    }
  }
}

}  // namespace dcp

// FP64 tensor-core GEMM for sm_100a, persistent and warp-specialised:
//
//   * grid = (CTAs per SM) x (SM count); every CTA walks the tile list with stride gridDim.x
//   * operand k-blocks stream through a shared-memory ring filled by TMA (cp.async.bulk.tensor, 128-byte
//     swizzle) and guarded by full/empty mbarriers; one elected thread refills the stage that was consumed one
//     k-block earlier, and the ring keeps running across tile boundaries, so the next tile's operands arrive
//     while the current tile's epilogue is still executing.  There is deliberately no dedicated producer warp:
//     4 warps per CTA x 2 CTAs = 2 warps per SM sub-partition leaves 255 registers per thread (a fifth warp
//     would cap it at 168 and spill the 128 accumulator registers in the epilogue)
//   * four warps run DMMA.8x8x4 out of conflict-free swizzled LDS.64 reads
//   * epilogue: the accumulators are staged through a padded shared-memory buffer (half a tile at a time) and the
//     consumer warps then sweep it row-contiguously (512-byte rows, 16 bytes per lane), loading the epilogue
//     operands of four elements per thread before the first store; the fused update (MU ratio, mask, ISTA/FISTA
//     proximal step + momentum + convergence test) runs in that sweep
//   * two CTAs per SM, so one CTA's epilogue overlaps the other's mainloop on the DMMA pipe
//
//   NT: acc[m][n] = sum_k A[m][k] B[n][k]   both operands K-contiguous   (y.dot(d.T), x.dot(G), ...)
//   TN: acc[m][n] = sum_k A[k][m] B[k][n]   both operands M/N-contiguous (x.T.dot(y), contraction over samples)
//
// Shared-memory operand layouts (BK = 16 doubles = one 128-byte swizzle row):
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

// internal epilogue selectors (template arguments); the PROX kind of the ABI is split by shrink rule
constexpr int EPI_PROX_REAL = 40, EPI_PROX_COMPLEX = 41, EPI_PROX_POSITIVE = 42;
constexpr int EPI_PARTIAL = 100;  // TN split-K: raw partial tile into the workspace slab of this split

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
  static constexpr int CONS_THREADS = NCONS * 32;
  static constexpr int THREADS = NCONS * 32;
  static constexpr int MI = WM / 8, NJ = WN / 8;
  static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  // epilogue staging: WM rows of the tile at a time, row pitch BN + 8 doubles (pitch = 64 bytes mod 128, so that
  // the 8 rows x 64 bytes a warp stores per instruction fall into 4 conflict-free wavefronts)
  static constexpr int EPI_ROWS = WM, EPI_PITCH = BN + 8;
  static constexpr int EPI_BYTES = EPI_ROWS * EPI_PITCH * 8;
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + 2 * STAGES * 8;
  static constexpr int EPI_BATCH = 4;
  static constexpr int EPI_PAIRS = EPI_ROWS * (BN / 2);
  static_assert(WM % 16 == 0 && WN % 16 == 0, "warp tile must be a multiple of 16 (TN sub-boxes)");
  static_assert(BM % 16 == 0 && BN % 16 == 0, "CTA tile must be a multiple of 16");
  static_assert(EPI_PAIRS % (CONS_THREADS * EPI_BATCH) == 0, "epilogue sweep must divide evenly");
  static_assert(BN / 2 == 32, "the epilogue sweep maps one warp to one staged row");
  static_assert(RING_BYTES % 1024 == 0, "staging buffer must stay 16-byte aligned behind the ring");
};

// --------------------------------------------------------------------------------------------------
// Epilogues.  Each handles the adjacent column pair (col, col+1); `two` is false for the last odd column.
// load() only issues global loads, apply() computes and stores, so that a batch of loads is in flight
// before the first store (operands may alias outputs -- in-place updates -- which forbids the compiler
// from hoisting loads over stores by itself).
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

struct EpiIn {
  Pair p0, p1, p2;
  double rowfac;
};

template <int KIND>
struct Epilogue {
  static constexpr bool kProx = KIND == EPI_PROX_REAL || KIND == EPI_PROX_COMPLEX || KIND == EPI_PROX_POSITIVE;

  static __device__ __forceinline__ void load(const decomp_epilogue_t& ep, long long row, long long col, bool two,
                                              EpiIn& in) {
    if constexpr (KIND == DECOMP_EPI_STORE_MASK) {
      if (ep.cwidth == 2) {
        in.p0.a = in.p0.b = ep.mask[row * ep.ldmask + (col >> 1)];
      } else {
        in.p0 = ld_pair(ep.mask + row * ep.ldmask + col, two);
      }
    } else if constexpr (KIND == DECOMP_EPI_MU_NUM || KIND == DECOMP_EPI_MU_DEN) {
      in.p0 = ld_pair(ep.x + row * ep.ldx + col, two);
      in.p1 = ld_pair(ep.other + row * ep.ldother + col, two);
    } else if constexpr (KIND == DECOMP_EPI_KL_RATIO) {
      in.p0 = ld_pair(ep.other + row * ep.ldother + col, two);
      if (ep.mask != nullptr) in.p1 = ld_pair(ep.mask + row * ep.ldmask + col, two);
    } else if constexpr (kProx) {
      in.p0 = ld_pair(ep.x + row * ep.ldx + col, two);
      in.p1 = ld_pair(ep.other + row * ep.ldother + col, two);
      in.p2 = ld_pair(ep.prev + row * ep.ldprev + col, two);
      in.rowfac = ep.rowvec != nullptr ? ep.rowvec[row] : 1.0;
    }
  }

  // returns true if the convergence test is violated by this pair (PROX with check only)
  static __device__ __forceinline__ bool apply(const decomp_epilogue_t& ep, double* __restrict__ pbase,
                                               long long ldp, long long row, long long col, bool two, double v0,
                                               double v1, const EpiIn& in, double step) {
    if constexpr (KIND == EPI_PARTIAL) {
      st_pair(pbase + row * ldp + col, v0, v1, two);
    } else if constexpr (KIND == DECOMP_EPI_STORE) {
      st_pair(ep.out + row * ep.ldo + col, v0, v1, two);
    } else if constexpr (KIND == DECOMP_EPI_STORE_MASK) {
      st_pair(ep.out + row * ep.ldo + col, v0 * in.p0.a, v1 * in.p0.b, two);
    } else if constexpr (KIND == DECOMP_EPI_MU_NUM || KIND == DECOMP_EPI_MU_DEN) {
      double n0, n1, d0, d1;
      if constexpr (KIND == DECOMP_EPI_MU_NUM) {
        n0 = v0; n1 = v1; d0 = in.p1.a; d1 = in.p1.b;
      } else {
        n0 = in.p1.a; n1 = in.p1.b; d0 = v0; d1 = v1;
      }
      // x * max(pos, 0) / max(neg, eps), evaluated left to right like the reference (grads.py:84,93)
      const double r0 = __ddiv_rn(__dmul_rn(in.p0.a, fmax(n0, 0.0)), fmax(d0, kEps));
      const double r1 = __ddiv_rn(__dmul_rn(in.p0.b, fmax(n1, 0.0)), fmax(d1, kEps));
      st_pair(ep.out + row * ep.ldo + col, r0, r1, two);
    } else if constexpr (KIND == DECOMP_EPI_KL_RATIO) {
      double y0 = in.p0.a, y1 = in.p0.b;
      if (ep.mask != nullptr) {
        y0 *= in.p1.a;
        y1 *= in.p1.b;
      }
      st_pair(ep.out + row * ep.ldo + col, y0 / (v0 + kEps), y1 / (v1 + kEps), two);
    } else if constexpr (kProx) {
      // z = w + step * (yAt - w.G)   (lasso.py:245-246)
      const double z0 = in.p0.a + step * (in.p1.a - v0);
      const double z1 = in.p0.b + step * (in.p1.b - v1);
      // threshold step * alpha (lasso.py:287); under a full mask alpha carries the per-problem mask count (:163).
      // The per-column vectors are a few kB and stay in L1.
      double a0, a1, tol0 = 0.0, tol1 = 0.0;
      if constexpr (KIND == EPI_PROX_COMPLEX) {
        a0 = a1 = __ldg(ep.colvec + (col >> 1));
        if (ep.check) tol0 = __ldg(ep.colvec2 + (col >> 1));
      } else {
        a0 = __ldg(ep.colvec + col);
        a1 = two ? __ldg(ep.colvec + col + 1) : 0.0;
        if (ep.check) {
          tol0 = __ldg(ep.colvec2 + col);
          tol1 = two ? __ldg(ep.colvec2 + col + 1) : 0.0;
        }
      }
      const double t0 = ep.rowvec != nullptr ? step * (a0 * in.rowfac) : step * a0;
      const double t1 = ep.rowvec != nullptr ? step * (a1 * in.rowfac) : step * a1;
      double x0, x1;
      bool bad = false;
      if constexpr (KIND == EPI_PROX_COMPLEX) {
        // z / (|z| + eps) * max(|z| - t, 0)   (lasso.py:210-225)
        const double r = hypot(z0, z1);
        const double den = r + kEps;
        const double mag = fmax(r - t0, 0.0);
        x0 = mag * (z0 / den);
        x1 = mag * (z1 / den);
        if (ep.check) bad = !(hypot(x0 - in.p2.a, x1 - in.p2.b) - tol0 < 0.0);
      } else {
        if constexpr (KIND == EPI_PROX_POSITIVE) {
          x0 = fmax(z0 - t0, 0.0);   // lasso.py:228-241
          x1 = fmax(z1 - t1, 0.0);
        } else {
          // max(|z| - t, 0) * sign(z)   (lasso.py:206-207)
          const double s0 = (z0 > 0.0) ? 1.0 : ((z0 < 0.0) ? -1.0 : z0);
          const double s1 = (z1 > 0.0) ? 1.0 : ((z1 < 0.0) ? -1.0 : z1);
          x0 = fmax(fabs(z0) - t0, 0.0) * s0;
          x1 = fmax(fabs(z1) - t1, 0.0) * s1;
        }
        if (ep.check) {
          bad = !(fabs(x0 - in.p2.a) - tol0 < 0.0);
          if (two) bad = bad || !(fabs(x1 - in.p2.b) - tol1 < 0.0);
        }
      }
      st_pair(ep.out + row * ep.ldo + col, x0, x1, two);
      if (ep.out2 != nullptr) {
        // w_next = x_new + momentum * (x_new - x_prev)   (lasso.py:412)
        st_pair(ep.out2 + row * ep.ldo2 + col, x0 + ep.momentum * (x0 - in.p2.a), x1 + ep.momentum * (x1 - in.p2.b),
                two);
      }
      return bad;
    }
    return false;
  }
};

struct TileInfo {
  int m0, n0, kb0, nkb, z;
};

__device__ __forceinline__ TileInfo tile_info(const GemmGeom& gs, int tiles_mn, int tile, int BM, int BN) {
  TileInfo t;
  const int tn = tile % gs.tiles_n;
  const int tm = (tile / gs.tiles_n) % gs.tiles_m;
  t.z = tile / tiles_mn;
  t.m0 = tm * BM;
  t.n0 = tn * BN;
  t.kb0 = t.z * gs.kblocks_per_split;
  t.nkb = gs.kblocks_total - t.kb0;
  if (t.nkb > gs.kblocks_per_split) t.nkb = gs.kblocks_per_split;
  if (t.nkb < 0) t.nkb = 0;
  return t;
}

// --------------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------------
template <class C, bool TN, int EPI>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
gemm_f64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmGeom gs,
                const decomp_epilogue_t ep, double* __restrict__ partial, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;

  extern __shared__ __align__(1024) unsigned char smem[];
  double* epi_buf = reinterpret_cast<double*>(smem + C::RING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES + C::EPI_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = gs.tiles_m * gs.tiles_n;
  const int tiles_total = tiles_mn * gs.splits;
  const bool elected = threadIdx.x == 0;

  // ---- producer cursor (meaningful in the elected thread only): next k-block to request
  int p_tile = blockIdx.x, p_i = 0, p_stage = 0;
  uint32_t p_phase = 0;      // parity of the `empty` completion the next refill of p_stage has to wait for
  bool p_wait = false;       // false while the ring is being filled for the first time
  TileInfo p_t = tile_info(gs, tiles_mn, p_tile < tiles_total ? p_tile : 0, C::BM, C::BN);

  auto produce_one = [&]() {
    // requests the next k-block of this CTA's tile sequence into stage p_stage (elected thread only)
    while (p_tile < tiles_total && p_i >= p_t.nkb) {
      p_tile += gridDim.x;
      p_i = 0;
      if (p_tile < tiles_total) p_t = tile_info(gs, tiles_mn, p_tile, C::BM, C::BN);
    }
    if (p_tile >= tiles_total) return;
    if (p_wait) mbar_wait(&empty_bar[p_stage], p_phase);
    mbar_arrive_expect_tx(&full_bar[p_stage], C::STAGE_BYTES);
    unsigned char* sa = smem + p_stage * C::STAGE_BYTES;
    unsigned char* sb = sa + C::A_BYTES;
    const int k0 = (p_t.kb0 + p_i) * BK;
    if constexpr (!TN) {
      tma_load_2d(sa, &tmA, &full_bar[p_stage], k0, p_t.m0);
      tma_load_2d(sb, &tmB, &full_bar[p_stage], k0, p_t.n0);
    } else {
#pragma unroll
      for (int b = 0; b < C::BM / 16; ++b) tma_load_2d(sa + b * 2048, &tmA, &full_bar[p_stage], p_t.m0 + 16 * b, k0);
#pragma unroll
      for (int b = 0; b < C::BN / 16; ++b) tma_load_2d(sb + b * 2048, &tmB, &full_bar[p_stage], p_t.n0 + 16 * b, k0);
    }
    ++p_i;
    if (++p_stage == C::STAGES) {
      p_stage = 0;
      if (p_wait) p_phase ^= 1u;
      p_wait = true;
    }
  };

  if (elected) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::NCONS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  __syncthreads();
  if (elected) {
#pragma unroll 1
    for (int s = 0; s < C::STAGES; ++s) produce_one();   // fill the ring
  }

  bool violated = false;
  const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
  const int g = lane >> 2, q = lane & 3;

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
  double step = 0.0;
  if constexpr (Epilogue<EPI>::kProx) step = *ep.step;

  int s = 0;
  uint32_t ph = 0;
  bool first = true;   // no refill after the very first k-block: the ring was filled STAGES deep
#pragma unroll 1
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
    const TileInfo t = tile_info(gs, tiles_mn, tile, C::BM, C::BN);

    double acc[C::MI][C::NJ][2];
#pragma unroll
    for (int i = 0; i < C::MI; ++i)
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll 1
    for (int it = 0; it < t.nkb; ++it) {
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
      // refill the stage consumed one k-block ago (prefetch distance STAGES - 1)
      if (elected && !first) produce_one();
      first = false;
      if (++s == C::STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }

    // -------------------------------------------------------------- epilogue through shared memory
    double* pbase = nullptr;
    if constexpr (EPI == EPI_PARTIAL) pbase = partial + (long long)t.z * gs.M * gs.ld_partial;
#pragma unroll
    for (int h = 0; h < C::WARPS_M; ++h) {
      __syncthreads();  // the previous sweep has finished reading the staging buffer
      if (wm == h) {
#pragma unroll
        for (int i = 0; i < C::MI; ++i)
#pragma unroll
          for (int j = 0; j < C::NJ; ++j)
            *reinterpret_cast<double2*>(epi_buf + (g + 8 * i) * C::EPI_PITCH + wn * C::WN + 8 * j + 2 * q) =
                make_double2(acc[i][j][0], acc[i][j][1]);
      }
      __syncthreads();
      const long long row0 = (long long)t.m0 + h * C::EPI_ROWS;
      // thread -> (row r_in + 4 e, column pair c2) of the staged half tile: a warp covers one 512-byte row
      const int r_in = threadIdx.x >> 5, c2 = threadIdx.x & 31;
      const long long col = (long long)t.n0 + 2 * c2;
      constexpr int ROW_STEP = C::CONS_THREADS / (C::BN / 2);
      const bool full = row0 + C::EPI_ROWS <= gs.M && (long long)t.n0 + C::BN <= gs.N;
      if (full) {
        // interior tile: no predicates, every load of a batch is issued before the first store
#pragma unroll 1
        for (int r = r_in; r < C::EPI_ROWS; r += ROW_STEP * C::EPI_BATCH) {
          EpiIn in[C::EPI_BATCH];
          double2 v[C::EPI_BATCH];
#pragma unroll
          for (int b = 0; b < C::EPI_BATCH; ++b) {
            Epilogue<EPI>::load(ep, row0 + r + ROW_STEP * b, col, true, in[b]);
            v[b] = *reinterpret_cast<const double2*>(epi_buf + (r + ROW_STEP * b) * C::EPI_PITCH + 2 * c2);
          }
#pragma unroll
          for (int b = 0; b < C::EPI_BATCH; ++b)
            violated |= Epilogue<EPI>::apply(ep, pbase, gs.ld_partial, row0 + r + ROW_STEP * b, col, true, v[b].x,
                                             v[b].y, in[b], step);
        }
      } else {
        // edge tile: rows beyond M / columns beyond N are masked off, an odd last column is handled alone
        const bool col_ok = col < gs.N, two = col + 1 < gs.N;
#pragma unroll 1
        for (int r = r_in; r < C::EPI_ROWS; r += ROW_STEP) {
          const long long row = row0 + r;
          if (row < gs.M && col_ok) {
            EpiIn in;
            Epilogue<EPI>::load(ep, row, col, two, in);
            const double2 v = *reinterpret_cast<const double2*>(epi_buf + r * C::EPI_PITCH + 2 * c2);
            violated |= Epilogue<EPI>::apply(ep, pbase, gs.ld_partial, row, col, two, v.x, v.y, in, step);
          }
        }
      }
    }
  }

  if constexpr (Epilogue<EPI>::kProx) {
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

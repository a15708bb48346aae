// FP64 tensor-core GEMM for sm_100a, persistent and warp-specialised:
//
//   * grid = one CTA per SM; every CTA walks the tile list with stride gridDim.x
//   * 8 MMA warps (warp tile 32x32 of the 128x64 CTA tile) run DMMA.8x8x4 out of conflict-free swizzled
//     LDS.64 reads.  Two MMA warps per SM sub-partition are what saturates the FP64 tensor pipe (one warp per
//     sub-partition reaches ~71 % of it, measured)
//   * operand k-blocks stream through a shared-memory ring filled by TMA (cp.async.bulk.tensor, 128-byte
//     swizzle) and guarded by full/empty mbarriers; a dedicated producer thread (own warp group) keeps it full
//     across tile boundaries.  A stage is handed back one k-block late, when the DMMAs that consumed it have
//     issued (see MmaPipe::run for why the obvious place is a race)
//   * 4 epilogue warps own all global-memory traffic of the fused update: when a tile's mainloop ends the MMA
//     warps park the accumulators in one of two padded shared-memory staging buffers and immediately start the
//     next tile; the epilogue warps sweep the staged tile row-contiguously (one warp = one 512-byte row, 16
//     bytes per lane), issue all operand loads of a batch before its first store, and apply the fused update (MU
//     ratio, mask, ISTA/FISTA proximal step + momentum + convergence test).  The memory-bound epilogue of tile t
//     therefore runs under the DMMA mainloop of tile t+1 instead of stalling it
//   * 512 threads x 128 registers: enough for the 64 accumulator registers of a 32x32 warp tile and for the
//     epilogue batches, no spills
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
  int zero;              // always 0, opaque to the compiler
  int tn3d;              // TN: operands are described by 3-D tensor maps (one TMA instruction per operand)
  int m_fast;            // tile order: the m tiles of one (n tile, K split) run side by side.  TN x^T y: the big operand
                         // is B (y), read once per m tile -- with k = 256 = two m tiles 60 % of y came from DRAM twice
                         // when the two CTAs that share a y tile sat 64 CTAs apart (ncu: 51.6 GB for a 32.8 GB y)
};

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int NBUF_, int EPI_WARPS_, int EPI_BATCH_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, NBUF = NBUF_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int MMA_WARPS = WARPS_M * WARPS_N, EPI_WARPS = EPI_WARPS_;
  static constexpr int MMA_THREADS = MMA_WARPS * 32, EPI_THREADS = EPI_WARPS * 32;
  static constexpr int PROD_THREADS = 128;   // producer warp group: one active thread, registers handed to the others
  static constexpr int THREADS = MMA_THREADS + EPI_THREADS + PROD_THREADS;
  static constexpr int MI = WM / 8, NJ = WN / 8;
  static constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  // accumulator staging: whole tile, row pitch BN + 8 doubles (pitch = 64 bytes mod 128, so that the 8 rows x
  // 64 bytes a warp stores per instruction fall into 4 conflict-free wavefronts)
  static constexpr int EPI_PITCH = BN + 8;
  static constexpr int BUF_BYTES = BM * EPI_PITCH * 8;
  static constexpr int NUM_BARRIERS = 2 * STAGES + 2 * NBUF;
  static constexpr int SMEM_BYTES = RING_BYTES + NBUF * BUF_BYTES + NUM_BARRIERS * 8;
  static constexpr int EPI_BATCH = EPI_BATCH_;
  static constexpr int ROW_STEP = EPI_THREADS / (BN / 2);      // rows covered by one pass of the epilogue warps
  static_assert(WM % 16 == 0 && WN % 16 == 0, "warp tile must be a multiple of 16 (TN sub-boxes)");
  static_assert(BM % 16 == 0 && BN % 16 == 0, "CTA tile must be a multiple of 16");
  static_assert(BN / 2 == 32, "the epilogue sweep maps one warp to one staged row");
  static_assert(BM % (ROW_STEP * EPI_BATCH) == 0, "epilogue sweep must divide evenly");
  static_assert(RING_BYTES % 1024 == 0 && BUF_BYTES % 16 == 0, "staging buffers must stay 16-byte aligned");
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

// ---- bit-level helpers: the same results as the floating-point forms without touching the FP64 pipe
// max(v, 0) that keeps NaN (like numpy's maximum): negative and not NaN  <=>  0x80000000 <= hi <= 0xfff00000
__device__ __forceinline__ double max_zero(double v) {
  const unsigned hi = (unsigned)__double2hiint(v);
  return (hi - 0x80000000u) <= 0x7ff00000u ? 0.0 : v;
}
// mag * sign(s) for mag >= 0 (or NaN): sign(0) = 0 gives 0 either way because mag is then max(-t, 0) = 0
__device__ __forceinline__ double with_sign_of(double mag, double s) {
  return __hiloint2double(__double2hiint(mag) | (__double2hiint(s) & 0x80000000), __double2loint(mag));
}
__device__ __forceinline__ unsigned long long abs_bits(double v) {
  return (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
}

// v * m without touching the FP64 pipe when m is exactly 1.0 or +0.0 (what masks hold): the product is then v, or a
// zero with v's sign (NaN for a non-finite v), formed with integer instructions; any other weight multiplies.
// Scalar FP64 instructions of the epilogue warps compete with the DMMAs for the same pipe.
__device__ __forceinline__ double mask_mul(double v, double m) {
  const unsigned long long mb = (unsigned long long)__double_as_longlong(m);
  if (mb == 0x3ff0000000000000ull) return v;
  if (mb == 0ull) {
    const unsigned hi = (unsigned)__double2hiint(v);
    if ((hi & 0x7ff00000u) == 0x7ff00000u) return __longlong_as_double(0x7ff8000000000000ll);   // inf * 0, NaN * 0
    return __hiloint2double((int)(hi & 0x80000000u), 0);
  }
  return v * m;
}

struct EpiIn {
  Pair p0, p1, p2;
};

__device__ __forceinline__ Pair ld_pair16(const double* p) {
  const double2 v = *reinterpret_cast<const double2*>(p);
  Pair r;
  r.a = v.x;
  r.b = v.y;
  return r;
}

template <int KIND>
struct Epilogue {
  static constexpr bool kProx = KIND == EPI_PROX_REAL || KIND == EPI_PROX_COMPLEX || KIND == EPI_PROX_POSITIVE;
  static constexpr bool kLoads = KIND != DECOMP_EPI_STORE && KIND != EPI_PARTIAL;
  // column pairs per thread whose operand loads are in flight together (8 pairs x 2 streams for the MU ratio,
  // 4 pairs x up to 3 streams elsewhere: what fits in 128 registers without spills)
  static constexpr int kBatch = (KIND == DECOMP_EPI_MU_NUM || KIND == DECOMP_EPI_MU_DEN) ? 8 : 4;

  // Issues the global loads of one column pair.  (row, col) has been clamped into the matrix by the caller, so
  // the loads are unconditional 16-byte accesses (the even row pitch keeps an odd last column in bounds); what
  // is out of range is discarded in apply().
  static __device__ __forceinline__ void load(const decomp_epilogue_t& ep, long long row, long long col, EpiIn& in) {
    if constexpr (KIND == DECOMP_EPI_STORE_MASK) {
      if (ep.cwidth == 2) {
        in.p0.a = in.p0.b = ep.mask[row * ep.ldmask + (col >> 1)];
      } else {
        in.p0 = ld_pair16(ep.mask + row * ep.ldmask + col);
      }
    } else if constexpr (KIND == DECOMP_EPI_MU_NUM || KIND == DECOMP_EPI_MU_DEN) {
      in.p0 = ld_pair16(ep.x + row * ep.ldx + col);
      in.p1 = ld_pair16(ep.other + row * ep.ldother + col);
    } else if constexpr (KIND == DECOMP_EPI_KL_RATIO) {
      in.p0 = ld_pair16(ep.other + row * ep.ldother + col);
      if (ep.mask != nullptr) in.p1 = ld_pair16(ep.mask + row * ep.ldmask + col);
    } else if constexpr (kProx) {
      in.p0 = ld_pair16(ep.x + row * ep.ldx + col);
      in.p1 = ld_pair16(ep.other + row * ep.ldother + col);
      in.p2 = ld_pair16(ep.prev + row * ep.ldprev + col);
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
      st_pair(ep.out + row * ep.ldo + col, mask_mul(v0, in.p0.a), mask_mul(v1, in.p0.b), two);
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
      // Scalar FP64 instructions share the DMMA pipe and cost the MMA warps a tensor slot each (measured: ~16
      // pipe cycles per warp instruction while DMMAs are in flight), so this update is written with the
      // minimum of them -- 5 per element: abs / max(.,0) / sign / the convergence comparison are bit operations.
      // z = w + step * (yAt - w.G)   (lasso.py:245-246)
      const double z0 = in.p0.a + step * (in.p1.a - v0);
      const double z1 = in.p0.b + step * (in.p1.b - v1);
      // threshold step * alpha (lasso.py:287): colvec holds it ready-made (flag bit 0) or alpha, in which case
      // it is formed here; under a full mask alpha carries the per-problem mask count (lasso.py:163).
      double t0, t1;
      unsigned long long tolb0 = 0ull, tolb1 = 0ull;
      if constexpr (KIND == EPI_PROX_COMPLEX) {
        t0 = t1 = __ldg(ep.colvec + (col >> 1));
        if (ep.check) tolb0 = (unsigned long long)__double_as_longlong(__ldg(ep.colvec2 + (col >> 1)));
      } else {
        const double2 a = __ldg(reinterpret_cast<const double2*>(ep.colvec + col));   // colvec is padded to even
        t0 = a.x;
        t1 = a.y;
        if (ep.check) {
          const double2 tl = __ldg(reinterpret_cast<const double2*>(ep.colvec2 + col));
          tolb0 = (unsigned long long)__double_as_longlong(tl.x);
          tolb1 = (unsigned long long)__double_as_longlong(tl.y);
        }
      }
      if (ep.rowvec != nullptr) {
        const double rowfac = __ldg(ep.rowvec + row);
        t0 = step * (t0 * rowfac);
        t1 = step * (t1 * rowfac);
      } else if (!(ep.flags & DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD)) {
        t0 = step * t0;
        t1 = step * t1;
      }
      double x0, x1, d0, d1;
      bool bad = false;
      if constexpr (KIND == EPI_PROX_COMPLEX) {
        // z / (|z| + eps) * max(|z| - t, 0)   (lasso.py:210-225)
        const double r = hypot(z0, z1);
        const double den = r + kEps;
        const double mag = max_zero(r - t0);
        x0 = mag * (z0 / den);
        x1 = mag * (z1 / den);
        d0 = x0 - in.p2.a;
        d1 = x1 - in.p2.b;
        if (ep.check) bad = !(abs_bits(hypot(d0, d1)) < tolb0);
      } else {
        if constexpr (KIND == EPI_PROX_POSITIVE) {
          x0 = max_zero(z0 - t0);   // lasso.py:228-241
          x1 = max_zero(z1 - t1);
        } else {
          // max(|z| - t, 0) * sign(z)   (lasso.py:206-207); the product with +-1 / 0 is a sign transfer
          x0 = with_sign_of(max_zero(fabs(z0) - t0), z0);
          x1 = with_sign_of(max_zero(fabs(z1) - t1), z1);
        }
        d0 = x0 - in.p2.a;
        d1 = x1 - in.p2.b;
        // |d| - tol < 0  <=>  |d| < tol; for non-negative doubles that is the order of their bit patterns, and a
        // NaN compares "not smaller" exactly as in the reference's max(...) < 0
        if (ep.check) bad = !(abs_bits(d0) < tolb0) || (two && !(abs_bits(d1) < tolb1));
      }
      st_pair(ep.out + row * ep.ldo + col, x0, x1, two);
      if (ep.out2 != nullptr) {
        // w_next = x_new + momentum * (x_new - x_prev)   (lasso.py:412)
        st_pair(ep.out2 + row * ep.ldo2 + col, x0 + ep.momentum * d0, x1 + ep.momentum * d1, two);
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
  const int tn = gs.m_fast ? (tile / gs.tiles_m) % gs.tiles_n : tile % gs.tiles_n;
  const int tm = gs.m_fast ? tile % gs.tiles_m : (tile / gs.tiles_n) % gs.tiles_m;
  t.z = tile / tiles_mn;
  t.m0 = tm * BM;
  t.n0 = tn * BN;
  t.kb0 = t.z * gs.kblocks_per_split;
  t.nkb = gs.kblocks_total - t.kb0;
  if (t.nkb > gs.kblocks_per_split) t.nkb = gs.kblocks_per_split;
  if (t.nkb < 0) t.nkb = 0;
  return t;
}

// Zero that the instruction scheduler has to wait for: `zero` is a kernel argument that is always 0, so the result
// is 0, but it cannot be formed before the registers behind `dep` have been written.  Used as `lane == after(..)`
// in front of an mbarrier arrival that must not overtake the shared-memory reads feeding those registers.
__device__ __forceinline__ int after(int dep, int zero) { return dep & zero; }

// --------------------------------------------------------------------------------------------------
// Operand pipeline + DMMA mainloop of the 8 MMA warps, shared by both kernels below.
// --------------------------------------------------------------------------------------------------
template <class C, bool TN>
struct MmaPipe {
  const CUtensorMap* tmA;
  const CUtensorMap* tmB;
  const GemmGeom* gs;
  unsigned char* smem;
  uint64_t* full_bar;
  uint64_t* empty_bar;
  int tiles_mn, tiles_total;
  int lane, wm, wn, g, q;
  // producer cursor (meaningful in the producer thread only): next k-block to request
  int p_tile, p_i, p_stage;
  uint32_t p_phase;   // parity of the `empty` completion the next refill of p_stage has to wait for
  bool p_wait;        // false while the ring is being filled for the first time
  TileInfo p_t;
  // consumer cursor
  int s;
  uint32_t ph;
  int offA[4], offB[4];   // per-lane byte offsets inside a stage for the four k-steps of one 16-wide k-block

  __device__ __forceinline__ void init(const CUtensorMap* a, const CUtensorMap* b, const GemmGeom* geom,
                                       unsigned char* sm, uint64_t* full, uint64_t* empty) {
    tmA = a;
    tmB = b;
    gs = geom;
    smem = sm;
    full_bar = full;
    empty_bar = empty;
    tiles_mn = geom->tiles_m * geom->tiles_n;
    tiles_total = tiles_mn * geom->splits;
    const int warp = threadIdx.x >> 5;
    lane = threadIdx.x & 31;
    wm = warp / C::WARPS_N;
    wn = warp % C::WARPS_N;
    g = lane >> 2;
    q = lane & 3;
    p_tile = blockIdx.x;
    p_i = 0;
    p_stage = 0;
    p_phase = 0;
    p_wait = false;
    p_t = tile_info(*geom, tiles_mn, p_tile < tiles_total ? p_tile : 0, C::BM, C::BN);
    s = 0;
    ph = 0;
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
  }

  // moves the producer cursor to the next k-block that exists (skips past the end of a tile)
  __device__ __forceinline__ void advance_cursor() {
    while (p_tile < tiles_total && p_i >= p_t.nkb) {
      p_tile += gridDim.x;
      p_i = 0;
      if (p_tile < tiles_total) p_t = tile_info(*gs, tiles_mn, p_tile, C::BM, C::BN);
    }
  }

  // requests the next k-block of this CTA's tile sequence into stage p_stage (producer thread only)
  __device__ __forceinline__ void produce_one() {
    advance_cursor();
    if (p_tile >= tiles_total) return;
    if (p_wait) mbar_wait(&empty_bar[p_stage], p_phase);
    mbar_arrive_expect_tx(&full_bar[p_stage], C::STAGE_BYTES);
    unsigned char* sa = smem + p_stage * C::STAGE_BYTES;
    unsigned char* sb = sa + C::A_BYTES;
    const int k0 = (p_t.kb0 + p_i) * BK;
    if constexpr (!TN) {
      tma_load_2d(sa, tmA, &full_bar[p_stage], k0, p_t.m0);
      tma_load_2d(sb, tmB, &full_bar[p_stage], k0, p_t.n0);
    } else if (gs->tn3d) {
      tma_load_3d(sa, tmA, &full_bar[p_stage], 0, k0, p_t.m0 / 16);
      tma_load_3d(sb, tmB, &full_bar[p_stage], 0, k0, p_t.n0 / 16);
    } else {
#pragma unroll
      for (int b = 0; b < C::BM / 16; ++b) tma_load_2d(sa + b * 2048, tmA, &full_bar[p_stage], p_t.m0 + 16 * b, k0);
#pragma unroll
      for (int b = 0; b < C::BN / 16; ++b) tma_load_2d(sb + b * 2048, tmB, &full_bar[p_stage], p_t.n0 + 16 * b, k0);
    }
    ++p_i;
    if (++p_stage == C::STAGES) {
      p_stage = 0;
      if (p_wait) p_phase ^= 1u;
      p_wait = true;
    }
  }

  // hands stage `stage` back to the producer (one arrival per MMA warp).  No __syncwarp: the caller has passed
  // warp-synchronous DMMAs that consumed every lane's reads of that stage.
  __device__ __forceinline__ void release(int stage) {
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
  }

  // acc = sum over the tile's k-blocks
  __device__ __forceinline__ void run(const TileInfo& t, double (&acc)[C::MI][C::NJ][2]) {
#pragma unroll
    for (int i = 0; i < C::MI; ++i)
#pragma unroll
      for (int j = 0; j < C::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    int held = -1;   // stage whose release is still owed (see below)
#pragma unroll 1
    for (int it = 0; it < t.nkb; ++it) {
      mbar_wait(&full_bar[s], ph);
      // The stage read in the previous iteration is handed back only here.  Program order alone does not keep the
      // arrival behind the shared-memory reads: ptxas hoists it above the DMMAs that consume the loaded registers,
      // and a refill by TMA (async proxy) then raced with reads still in flight -- seen as wrong B fragments in one
      // warp behind K = 4096 mainloops.  At this point every DMMA of the previous k-block has been issued, hence
      // its operands had arrived in registers.
      if (held >= 0) release(held);
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
      held = s;
      if (++s == C::STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
    // last k-block of the tile: the arrival is made to depend on the finished accumulators instead
    if (held >= 0) {
      int dep = 0;
#pragma unroll
      for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) dep |= __double2hiint(acc[i][j][0]);
      __syncwarp();
      if (lane == after(dep, gs->zero)) mbar_arrive(&empty_bar[held]);
    }
  }
};

// last-CTA-done convergence latch shared by the Lasso kernels (lasso.py:293/409): every thread passes its flag
__device__ __forceinline__ void latch_vote(int* scratch, int* latch, int latch_value, bool violated) {
  const int any = __syncthreads_or(violated ? 1 : 0);
  if (threadIdx.x == 0) {
    if (any) atomicOr(&scratch[0], 1);
    __threadfence();
    const int ticket = atomicAdd(&scratch[1], 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const int v = atomicOr(&scratch[0], 0);
      if (v == 0) *latch = latch_value;
      scratch[0] = 0;
      scratch[1] = 0;
    }
  }
}
__device__ __forceinline__ void convergence_latch(const decomp_epilogue_t& ep, bool violated) {
  latch_vote(ep.scratch, ep.latch, ep.latch_value, violated);
}

// --------------------------------------------------------------------------------------------------
// Generic kernel: 8 MMA warps + 4 epilogue warps, accumulators handed over through shared memory
// --------------------------------------------------------------------------------------------------
template <class C, bool TN, int EPI>
__global__ void __launch_bounds__(C::THREADS, 1)
gemm_f64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmGeom gs,
                const decomp_epilogue_t ep, double* __restrict__ partial, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;

  extern __shared__ __align__(1024) unsigned char smem[];
  double* stage_buf = reinterpret_cast<double*>(smem + C::RING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES + C::NBUF * C::BUF_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* staged_bar = empty_bar + C::STAGES;    // MMA warps -> epilogue warps: accumulators parked
  uint64_t* drained_bar = staged_bar + C::NBUF;    // epilogue warps -> MMA warps: staging buffer free again

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_mn = gs.tiles_m * gs.tiles_n;
  const int tiles_total = tiles_mn * gs.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::MMA_WARPS);
    }
    for (int b = 0; b < C::NBUF; ++b) {
      mbar_init(&staged_bar[b], C::MMA_WARPS);
      mbar_init(&drained_bar[b], C::EPI_WARPS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  __syncthreads();

  bool violated = false;

  // 512 threads x 128 registers fill the register file exactly; the MMA warps (64 accumulator registers) and the
  // epilogue batches both fit in 128 without spills, so no register re-balancing (setmaxnreg) is needed here.
  if (threadIdx.x >= C::MMA_THREADS + C::EPI_THREADS) {
    // ================================================================ producer warp group (one active thread)
    if (threadIdx.x == C::MMA_THREADS + C::EPI_THREADS) {
      MmaPipe<C, TN> prod;
      prod.init(&tmA, &tmB, &gs, smem, full_bar, empty_bar);
      for (;;) {
        prod.advance_cursor();
        if (prod.p_tile >= tiles_total) break;
        prod.produce_one();
      }
    }
  } else if (warp < C::MMA_WARPS) {
    // ================================================================ MMA warps
    MmaPipe<C, TN> pipe;
    pipe.init(&tmA, &tmB, &gs, smem, full_bar, empty_bar);
    const int wm = pipe.wm, wn = pipe.wn, g = pipe.g, q = pipe.q;

    int buf = 0;
    uint32_t buf_phase = 0;   // parity of the `drained` completion to wait for before re-using `buf`
    bool buf_wait = false;    // false until every staging buffer has been used once
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
      const TileInfo t = tile_info(gs, tiles_mn, tile, C::BM, C::BN);
      double acc[C::MI][C::NJ][2];
      pipe.run(t, acc);

      // park the accumulators for the epilogue warps and move on
      if (buf_wait) mbar_wait(&drained_bar[buf], buf_phase);
      double* sb = stage_buf + buf * (C::BUF_BYTES / 8);
#pragma unroll
      for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NJ; ++j)
          *reinterpret_cast<double2*>(sb + (wm * C::WM + g + 8 * i) * C::EPI_PITCH + wn * C::WN + 8 * j + 2 * q) =
              make_double2(acc[i][j][0], acc[i][j][1]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&staged_bar[buf]);   // release: the stores above are visible to the waiters
      if (++buf == C::NBUF) {
        buf = 0;
        if (buf_wait) buf_phase ^= 1u;
        buf_wait = true;
      }
    }
  } else {
    // ================================================================ epilogue warps
    // A tile is swept in BM / (ROW_STEP * EPI_BATCH) batches of EPI_BATCH column pairs per thread; all operand
    // loads of a batch are issued before its first store.
    const int e = threadIdx.x - C::MMA_THREADS;
    const int r_in = e >> 5, c2 = e & 31;   // thread -> (row r_in + ROW_STEP j, column pair c2): a warp covers one row
    constexpr int BATCH = Epilogue<EPI>::kBatch;
    constexpr int NBATCH = C::BM / (C::ROW_STEP * BATCH);
    static_assert(C::BM % (C::ROW_STEP * BATCH) == 0, "epilogue sweep must divide evenly");
    double step = 0.0;
    if constexpr (Epilogue<EPI>::kProx) step = *ep.step;

    struct Where {
      long long row0, col, colc;   // first row of the tile, this thread's column, the column clamped into [0, N-2]
      bool col_ok, two;
    };
    auto locate = [&](const TileInfo& t) {
      Where w;
      w.row0 = t.m0;
      w.col = (long long)t.n0 + 2 * c2;
      w.col_ok = w.col < gs.N;
      w.two = w.col + 1 < gs.N;
      long long cmax = (gs.N - 1) & ~1LL;    // last even column: a 16-byte load there stays inside the even pitch
      w.colc = w.col < cmax ? w.col : cmax;
      return w;
    };
    auto issue = [&](const Where& w, int batch, EpiIn (&in)[BATCH]) {
      if constexpr (Epilogue<EPI>::kLoads) {
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          long long row = w.row0 + r_in + C::ROW_STEP * (batch * BATCH + b);
          if (row > gs.M - 1) row = gs.M - 1;
          Epilogue<EPI>::load(ep, row, w.colc, in[b]);
        }
      }
    };
    auto finish = [&](const Where& w, int batch, const EpiIn (&in)[BATCH], const double* sb, double* pbase) {
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int r = r_in + C::ROW_STEP * (batch * BATCH + b);
        const long long row = w.row0 + r;
        const double2 v = *reinterpret_cast<const double2*>(sb + r * C::EPI_PITCH + 2 * c2);
        if (row < gs.M && w.col_ok)
          violated |= Epilogue<EPI>::apply(ep, pbase, gs.ld_partial, row, w.col, w.two, v.x, v.y, in[b], step);
      }
    };

    int buf = 0;
    uint32_t buf_phase = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
      const TileInfo t = tile_info(gs, tiles_mn, tile, C::BM, C::BN);
      const Where w = locate(t);
      double* pbase = nullptr;
      if constexpr (EPI == EPI_PARTIAL) pbase = partial + (long long)t.z * gs.M * gs.ld_partial;
      const double* sb = stage_buf + buf * (C::BUF_BYTES / 8);
      EpiIn in[BATCH];
      issue(w, 0, in);                                     // operand loads do not depend on the accumulators
      mbar_wait(&staged_bar[buf], buf_phase);              // accumulators of this tile are parked
#pragma unroll 1
      for (int batch = 0; batch < NBATCH; ++batch) {
        if (batch > 0) issue(w, batch, in);
        finish(w, batch, in, sb, pbase);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&drained_bar[buf]);
      if (++buf == C::NBUF) {
        buf = 0;
        buf_phase ^= 1u;
      }
    }
  }

  if constexpr (Epilogue<EPI>::kProx) {
    if (ep.check) convergence_latch(ep, violated);
  }
}

// --------------------------------------------------------------------------------------------------
// Unmasked ISTA / FISTA iteration in ONE launch, accumulators never leave the registers:
//     z = c + w Q,   Q = I - G / L,  c = (y A^H) / L        (= w + (yAh - w G) / L, lasso.py:245-246)
//     x_new = shrink(z, alpha / L);   w_next = x_new + momentum (x_new - x_prev)
// 8 MMA warps, one CTA per SM.  The two epilogue operand tiles (c and x_prev, 64 KB each) are brought into
// shared memory by TMA while the tile's mainloop runs (128-byte swizzle, so that the fragment-layout reads are
// conflict free), and when the mainloop ends every warp updates its own 32x32 fragment and stores x_new and
// w_next straight from registers.  All scalar FP64 work of a tile therefore happens in one short burst in which
// no DMMA is in flight, instead of trickling through the shared FP64/DMMA pipe under the next tile's mainloop.
// --------------------------------------------------------------------------------------------------
template <class C>
struct ProxqSmem {
  static constexpr int OPND_BYTES = C::BM * C::BN * 8;
  static constexpr int BOX_BYTES = C::BM * 128;          // one TMA box: BM rows x 16 doubles, 128-byte swizzle
  static constexpr int SMEM_BYTES = C::RING_BYTES + 2 * OPND_BYTES + (2 * C::STAGES + 2) * 8;
};

template <class C, int EPI>
__global__ void __launch_bounds__(C::MMA_THREADS + 128, 1)
gemm_f64_proxq_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
                      const GemmGeom gs, const decomp_epilogue_t ep, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  using S = ProxqSmem<C>;

  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* opnd_c = smem + C::RING_BYTES;
  unsigned char* opnd_p = opnd_c + S::OPND_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(opnd_p + S::OPND_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* opnd_bar = empty_bar + C::STAGES;   // operand tiles of the current tile have landed
  uint64_t* opnd_free = opnd_bar + 1;            // every MMA warp is done with them

  const int tiles_mn = gs.tiles_m * gs.tiles_n;
  const int tiles_total = tiles_mn * gs.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C::MMA_WARPS);
    }
    mbar_init(opnd_bar, 1);
    mbar_init(opnd_free, C::MMA_WARPS);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmP);
  }
  __syncthreads();

  // The operand tiles of a tile are requested box by box (16 KB each) between the ring refills of that tile's own
  // mainloop, so that they never queue in front of the GEMM operands (producer thread only).
  constexpr int NBOX = C::BN / 16;
  auto request_box = [&](const TileInfo& t, int b) {
    if (b == 0) mbar_arrive_expect_tx(opnd_bar, 2 * S::OPND_BYTES);
    if (b < NBOX) {
      tma_load_2d(opnd_c + b * S::BOX_BYTES, &tmC, opnd_bar, t.n0 + 16 * b, t.m0);
    } else {
      tma_load_2d(opnd_p + (b - NBOX) * S::BOX_BYTES, &tmP, opnd_bar, t.n0 + 16 * (b - NBOX), t.m0);
    }
  };

  bool violated = false;
  if (threadIdx.x >= C::MMA_THREADS) {
    // ================================================================ producer warp group (one active thread)
    // A dedicated TMA thread: with the elected-MMA-thread scheme that thread's warp falls behind the other seven
    // by the time it spends waiting for `empty` slots and issuing copies, and everybody waits for it at the next
    // refill.  The group gives its registers to the MMA warps (setmaxnreg).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (threadIdx.x == C::MMA_THREADS) {
      MmaPipe<C, false> prod;
      prod.init(&tmA, &tmB, &gs, smem, full_bar, empty_bar);
      uint32_t free_phase = 0;
      bool free_wait = false;       // the operand tiles have not been used yet
      for (;;) {
        prod.advance_cursor();
        if (prod.p_tile >= tiles_total) break;
        const int i = prod.p_i;
        const TileInfo t = prod.p_t;
        prod.produce_one();
        {
          // k-block i of that tile is on its way; trickle that tile's operand boxes behind k-blocks STAGES, STAGES+1, ...
          // (by then the MMA warps have started this tile, i.e. finished the previous tile's update)
          int b0 = i - C::STAGES, b1 = b0 + 1;
          if (i == t.nkb - 1) b1 = 2 * NBOX;                    // last k-block: whatever is left
          if (b0 < 0) b0 = 0;
          for (int b = b0; b < b1 && b < 2 * NBOX; ++b) {
            if (b == 0) {
              if (free_wait) {
                mbar_wait(opnd_free, free_phase);
                free_phase ^= 1u;
              }
              free_wait = true;
            }
            request_box(t, b);
          }
        }
      }
    }
  } else {
    // ================================================================ MMA warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    MmaPipe<C, false> pipe;
    pipe.init(&tmA, &tmB, &gs, smem, full_bar, empty_bar);
    const int wm = pipe.wm, wn = pipe.wn, g = pipe.g, q = pipe.q;

    double step = 0.0;
    if (!(ep.flags & DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD)) step = *ep.step;
    uint32_t opnd_phase = 0;
  #pragma unroll 1
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
      const TileInfo t = tile_info(gs, tiles_mn, tile, C::BM, C::BN);
      double acc[C::MI][C::NJ][2];
        pipe.run(t, acc);

      mbar_wait(opnd_bar, opnd_phase);
      opnd_phase ^= 1u;

      // per-column vectors of this lane's NJ column pairs: threshold step * alpha (lasso.py:287) and tolerance
      // tol * s as bit patterns (lasso.py:130); columns beyond N get harmless zeros and are never stored
      const long long col_lane = (long long)t.n0 + wn * C::WN + 2 * q;
      double thr0[C::NJ], thr1[C::NJ];
      unsigned long long tolb0[C::NJ], tolb1[C::NJ];
  #pragma unroll
      for (int j = 0; j < C::NJ; ++j) {
        const long long col = col_lane + 8 * j;
        thr0[j] = thr1[j] = 0.0;
        tolb0[j] = tolb1[j] = 0ull;
        if (col < gs.N) {
          if constexpr (EPI == EPI_PROX_COMPLEX) {
            thr0[j] = thr1[j] = __ldg(ep.colvec + (col >> 1));
            if (ep.check) tolb0[j] = (unsigned long long)__double_as_longlong(__ldg(ep.colvec2 + (col >> 1)));
          } else {
            const double2 a = __ldg(reinterpret_cast<const double2*>(ep.colvec + col));   // padded to even
            thr0[j] = a.x;
            thr1[j] = a.y;
            if (ep.check) {
              const double2 tl = __ldg(reinterpret_cast<const double2*>(ep.colvec2 + col));
              tolb0[j] = (unsigned long long)__double_as_longlong(tl.x);
              tolb1[j] = (unsigned long long)__double_as_longlong(tl.y);
            }
          }
          if (!(ep.flags & DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD)) {
            thr0[j] = step * thr0[j];
            thr1[j] = step * thr1[j];
          }
        }
      }

      // fragment walk: row r = wm*WM + g + 8 i (so r & 7 == g), column c = wn*WN + 8 j + 2 q inside the tile
      const long long row_lane = (long long)t.m0 + wm * C::WM + g;
      const int c_lane = wn * C::WN + 2 * q;
      const unsigned char* sc = opnd_c + (wm * C::WM + g) * 128;
      const unsigned char* sp = opnd_p + (wm * C::WM + g) * 128;
      double* po = ep.out + row_lane * ep.ldo + col_lane;
      double* pw = ep.out2 != nullptr ? ep.out2 + row_lane * ep.ldo2 + col_lane : nullptr;
      const bool interior = (long long)t.m0 + C::BM <= gs.M && (long long)t.n0 + C::BN <= gs.N;
      const double mom = ep.momentum;

      int opnd_dep = 0;
      auto update = [&](int i, int j, bool store, bool two) {
        const int c = c_lane + 8 * j;
        const int off = (c >> 4) * S::BOX_BYTES + i * 1024 + ((((c & 15) >> 1) ^ g) << 4);
        const double2 cc = *reinterpret_cast<const double2*>(sc + off);
        const double2 pp = *reinterpret_cast<const double2*>(sp + off);
        const double z0 = acc[i][j][0] + cc.x;
        const double z1 = acc[i][j][1] + cc.y;
        double x0, x1, d0, d1;
        bool bad = false;
        if constexpr (EPI == EPI_PROX_COMPLEX) {
          // z / (|z| + eps) * max(|z| - t, 0)   (lasso.py:210-225)
          const double rr = hypot(z0, z1);
          const double den = rr + kEps;
          const double mag = max_zero(rr - thr0[j]);
          x0 = mag * (z0 / den);
          x1 = mag * (z1 / den);
          d0 = x0 - pp.x;
          d1 = x1 - pp.y;
          if (ep.check) bad = !(abs_bits(hypot(d0, d1)) < tolb0[j]);
        } else {
          if constexpr (EPI == EPI_PROX_POSITIVE) {
            x0 = max_zero(z0 - thr0[j]);   // lasso.py:228-241
            x1 = max_zero(z1 - thr1[j]);
          } else {
            // max(|z| - t, 0) * sign(z)   (lasso.py:206-207)
            x0 = with_sign_of(max_zero(fabs(z0) - thr0[j]), z0);
            x1 = with_sign_of(max_zero(fabs(z1) - thr1[j]), z1);
          }
          d0 = x0 - pp.x;
          d1 = x1 - pp.y;
          if (ep.check) bad = !(abs_bits(d0) < tolb0[j]) || (two && !(abs_bits(d1) < tolb1[j]));
        }
        opnd_dep |= __double2hiint(d0) | __double2hiint(d1);
        if (store) {
          violated |= bad;
          st_pair(po + 8 * j, x0, x1, two);
          // w_next = x_new + momentum * (x_new - x_prev)   (lasso.py:412)
          if (pw != nullptr) st_pair(pw + 8 * j, x0 + mom * d0, x1 + mom * d1, two);
        }
      };

      if (interior) {
  #pragma unroll
        for (int i = 0; i < C::MI; ++i) {
  #pragma unroll
          for (int j = 0; j < C::NJ; ++j) update(i, j, true, true);
          po += 8 * ep.ldo;
          if (pw != nullptr) pw += 8 * ep.ldo2;
        }
      } else {
  #pragma unroll
        for (int i = 0; i < C::MI; ++i) {
          const bool row_ok = row_lane + 8 * i < gs.M;
  #pragma unroll
          for (int j = 0; j < C::NJ; ++j) {
            const long long col = col_lane + 8 * j;
            update(i, j, row_ok && col < gs.N, col + 1 < gs.N);
          }
          po += 8 * ep.ldo;
          if (pw != nullptr) pw += 8 * ep.ldo2;
        }
      }
      // this warp is done with the operand tiles; TMA overwrites them next, so the arrival waits for the values
      // that were computed from them (same reasoning as in MmaPipe::run)
      __syncwarp();
      if (pipe.lane == after(opnd_dep, gs.zero)) mbar_arrive(opnd_free);
    }
  }
  if (ep.check) convergence_latch(ep, violated);
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

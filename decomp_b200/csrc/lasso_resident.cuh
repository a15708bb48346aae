// Unmasked ISTA / FISTA with the iterate resident on chip: one launch runs `iters` iterations
//     z = c + w Q;   x_new = shrink(z, thr);   w_next = x_new + momentum_i (x_new - x_prev)      (lasso.py:244-271, 405-414)
// A batch row only ever needs its own row of w, so a CTA keeps a block of BM rows for all iterations of the launch:
//   * w (the A operand of the next GEMM) lives in shared memory in the swizzled k-block layout the DMMA fragment
//     loads expect, and the update writes w_next straight back into it
//   * c = (y A^H)/L sits in tensor memory (thread-private columns, see RES_TMEM_COLS), x_prev in registers (every
//     thread owns the same accumulator fragment in every iteration)
//   * only Q = (I - G/L)^T streams, from L2, through a TMA ring
// so HBM sees one read of (w, c, x) and one write of (x, w) per launch instead of per iteration, and there is one
// launch per `iters` iterations.  BM * N = 8192 elements; 16 MMA warps, each owning a 16 x 32 piece of the block
// (N = 256 -> 32 rows, 2 x 8 warps; N = 128 -> 64 rows, 4 x 4 warps; ...): four warps per SM sub-partition take
// turns on the FP64 tensor pipe, so the latency of one warp's fragment loads, barrier waits and update is covered
// by the DMMAs of the other three, and a thread carries 32 accumulator + 32 x_prev registers instead of 64 + 64
// (the 8-warp version of round 1 sat at 168 registers with spills and could not overlap its fragment loads with
// its DMMAs: 84 % of the DMMA issue rate).  The k-loop is straight-line code per number of live 8-row groups.
// The 16 warps form two groups of 8 that own the two halves of the row block and share nothing but the Q ring
// (rows are independent): each group has its own w tile, TMA barrier and named barrier, and group 1 starts every
// iteration three k-blocks behind group 0 (of the five stages of the ring; a stage is refilled when both groups
// have released it, so the lead comes out of the prefetch distance -- with the three stages the ring had while c
// lived in shared memory a lead bought nothing).  When group 0 reaches its update --
// scalar FP64 work plus two barriers during which its warps issue no DMMA -- group 1 still has two k-blocks to go and
// has the tensor pipe to itself, and vice versa at the start of the next iteration: the update of one half runs
// under the DMMAs of the other instead of idling the pipe (it was 7 % of the iteration with all warps in lockstep).
// Scalar FP64 instructions run on the same pipe as the DMMAs (about one DMMA slot per warp instruction), so the
// update is kept to 3 of them per element: the accumulators start from c instead of zero (no z = acc + c), |z| - t,
// and the two of the extrapolation.  (gemm_f64_proxq_kernel adds c after the sum, so the two kernels agree to
// rounding, not to the bit.)
#pragma once
#include <type_traits>

#include "gemm.cuh"

namespace dcp {

constexpr int RES_MMA_WARPS = 16;
constexpr int RES_MMA_THREADS = RES_MMA_WARPS * 32;
constexpr int RES_THREADS = RES_MMA_THREADS + 128;   // + the producer warp group (one active thread)
constexpr int RES_MAX_ITERS = 32;
constexpr int RES_MAX_STAGES = 8;

// c (16 doubles per thread, 64 KB per CTA) lives in TENSOR MEMORY: the DMMA path does not use it otherwise, a
// tcgen05.ld/st.32x32b gives every thread 32 private 32-bit columns of its lane, and the 64 KB of shared memory this
// frees make the Q ring five stages deep instead of three -- which is what the lead of warp group 0 over group 1
// needs (a stage is refilled only when both groups have released it: lead + prefetch distance must fit the ring).
constexpr int RES_TMEM_COLS = 128;   // 4 warps per TMEM lane quarter x 32 columns

struct ResidentSmem {
  static constexpr int W_BYTES = 65536, RING_BYTES = 163840;
  static constexpr int RING_OFF = W_BYTES;
  static constexpr int BAR_OFF = RING_OFF + RING_BYTES;
  // full[8] + empty[8] + wbar[2] + the skew counter + the tensor-memory base address
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * RES_MAX_STAGES + 4) * 8;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_res(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct ResidentArgs {
  long long M;
  int N, iters, check, zero;
  int skew, prefetch;  // tuning knobs (DECOMP_RESIDENT_SKEW: k-blocks group 1 trails group 0; DECOMP_RESIDENT_PREFETCH)
  const double* c;     // [M, N]  (y A^H) / L
  long long ldc;
  double* x;           // [M, N]  in: x_prev, out: x after `iters` iterations
  long long ldx;
  double* w;           // [M, N]  in: the point the gradient is taken at (also behind tmW), out: the next one
  long long ldw;
  const double* thr;   // threshold step * alpha per column (complex: per column pair)
  const double* tol;   // tolerance per column (read when check)
  int* latch;
  int* scratch;
  int latch_value;
  double momentum[RES_MAX_ITERS];
};

// barrier of one group of 8 MMA warps (named barriers 1 and 2)
__device__ __forceinline__ void group_sync(int group) {
  if (group == 0)
    asm volatile("bar.sync 1, 256;" ::: "memory");
  else
    asm volatile("bar.sync 2, 256;" ::: "memory");
}

// One iteration's GEMM for a warp with MI live 8-row groups (its 16 rows x 32 columns): KB k-blocks of
// LDS.64 -> DMMA.8x8x4, straight-line so that the fragment loads of a k-step are scheduled under the DMMAs of the
// one before.  MI = 0: the warp only walks the ring (its arrivals are part of every stage's hand-over).
template <int MI>
__device__ __forceinline__ void resident_gemm(double (&acc)[2][4][2], const unsigned char* Wt,
                                              const unsigned char* ring, uint64_t* full_bar, uint64_t* empty_bar,
                                              int KB, int kb_bytes, int stage_bytes, int stages,
                                              const int (&offA)[4], const int (&offB)[4], int& s, uint32_t& ph,
                                              int lane, int zero, volatile int* signal, int signal_kb,
                                              int signal_value) {
  int held = -1;
#pragma unroll 1
  for (int kb = 0; kb < KB; ++kb) {
    // group 0 tells group 1 that it is `signal_kb` k-blocks into this iteration (one warp speaks for the group)
    if (signal != nullptr && kb == signal_kb && lane == 0) *signal = signal_value;
    mbar_wait(&full_bar[s], ph);
    // late release of the stage read one k-block ago, see MmaPipe::run
    if (held >= 0 && lane == 0) mbar_arrive(&empty_bar[held]);
    if constexpr (MI > 0) {
      const unsigned char* sa = Wt + kb * kb_bytes;
      const unsigned char* sb = ring + s * stage_bytes;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        double fa[MI], fb[4];
#pragma unroll
        for (int i = 0; i < MI; ++i) fa[i] = *reinterpret_cast<const double*>(sa + offA[s4] + i * 1024);
#pragma unroll
        for (int j = 0; j < 4; ++j) fb[j] = *reinterpret_cast<const double*>(sb + offB[s4] + j * 1024);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
      }
    }
    held = s;
    if (++s == stages) {
      s = 0;
      ph ^= 1u;
    }
  }
  // last k-block: the arrival waits for the accumulators computed from the stage
  int dep = 0;
  if constexpr (MI > 0) {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dep |= __double2hiint(acc[i][j][0]);
  }
  if (lane == after(dep, zero)) mbar_arrive(&empty_bar[held]);
}

template <int SHRINK>
__global__ void __launch_bounds__(RES_THREADS, 1)
lasso_resident_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmQ,
                      const ResidentArgs a, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  using S = ResidentSmem;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Wt = smem;
  unsigned char* ring = smem + S::RING_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + RES_MAX_STAGES;
  uint64_t* wbar = empty_bar + RES_MAX_STAGES;

  volatile int* skew = reinterpret_cast<volatile int*>(wbar + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 3);

  // two groups of 8 warps, each owning BMG rows of the block of BM = 2 BMG rows
  const int N = a.N, KB = N >> 4, WN = N >> 5, WMG = 8 / WN, BMG = 16 * WMG, BM = 2 * BMG;
  const int stage_bytes = N * 128;
  const int stages = S::RING_BYTES / stage_bytes < RES_MAX_STAGES ? S::RING_BYTES / stage_bytes : RES_MAX_STAGES;
  const int kb_bytes = BMG * 128;   // one k-block of a group's resident w tile
  // group 1 starts an iteration when group 0 is this many k-blocks into it (the ring lets group 0 lead by stages - 1)
  int skew_kb = a.skew;
  if (skew_kb > KB - 1) skew_kb = KB - 1;
  if (skew_kb > stages - 1) skew_kb = stages - 1;
  // every CTA owns one contiguous range of rows (a multiple of the DMMA row granularity 8) and walks it in blocks of
  // BM rows; the ragged last block only computes the 8-row groups it has, so the grid is loaded evenly to 8 rows
  const long long per_cta = (((a.M + gridDim.x - 1) / gridDim.x) + 7) & ~7LL;
  const long long row_begin = (long long)blockIdx.x * per_cta;
  const long long row_end = row_begin + per_cta < a.M ? row_begin + per_cta : a.M;
  const int tiles = row_begin < row_end ? (int)((row_end - row_begin + BM - 1) / BM) : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RES_MMA_WARPS);
    }
    mbar_init(&wbar[0], 1);
    mbar_init(&wbar[1], 1);
    *skew = 0;
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmQ);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(RES_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  bool violated = false;
  // Registers are allocated to warps in groups of four: 640 threads leave 96 each.  The producer group hands most of
  // its share to the four MMA groups (4 x 128 x 112 + 128 x 24 = 60416 of the 640 x 96 = 61440 the CTA was launched with).
  if (warp >= RES_MMA_WARPS) {
    // ================================================================ producer: Q k-blocks, round and round
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == RES_MMA_WARPS && lane == 0) {
      const long long total = (long long)tiles * a.iters * KB;
      int s = 0, kb = 0;
      uint32_t ph = 0;
      bool wait = false;
      const long long per_tile = (long long)a.iters * KB;
      long long next_tile_at = 0;
      int tile = 0;
      for (long long n = 0; n < total; ++n) {
        if (n == next_tile_at) {
          // a row block starts: pull the NEXT one's w, c and x towards L2, so that its prologue, which every CTA
          // reaches at about the same time, is not a burst of HBM reads
          next_tile_at += per_tile;
          ++tile;
          if (tile < tiles && a.prefetch) {
            const long long m1 = row_begin + (long long)tile * BM;
            for (int kb2 = 0; kb2 < KB; ++kb2) {
              tma_prefetch_l2_2d(&tmW, kb2 * BK, (int)m1);
              if (m1 + BMG < a.M) tma_prefetch_l2_2d(&tmW, kb2 * BK, (int)m1 + BMG);   // one box per warp group
            }
            const long long rows = row_end - m1 < BM ? row_end - m1 : BM;
            for (long long r = 0; r < rows; ++r) {
              bulk_prefetch_l2(a.c + (m1 + r) * a.ldc, (uint32_t)N * 8u);
              bulk_prefetch_l2(a.x + (m1 + r) * a.ldx, (uint32_t)N * 8u);
            }
          }
        }
        if (wait) mbar_wait(&empty_bar[s], ph);
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        tma_load_2d(ring + s * stage_bytes, &tmQ, &full_bar[s], kb * BK, 0);
        if (++kb == KB) kb = 0;
        if (++s == stages) {
          s = 0;
          if (wait) ph ^= 1u;
          wait = true;
        }
      }
    }
  } else {
    // ================================================================ MMA warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int group = warp >> 3, lw = warp & 7;
    const int wn = lw % WN, wm = lw / WN, g = lane >> 2, q = lane & 3;
    unsigned char* Wg = Wt + group * (S::W_BYTES / 2);
    int iter_no = 0;   // iterations this CTA has started, over all its row blocks
    int offA[4], offB[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int o = (((s4 + 4 * (q >> 1)) ^ g) << 4) | ((q & 1) << 3);
      offA[s4] = (wm * 16 + g) * 128 + o;
      offB[s4] = (wn * 32 + g) * 128 + o;
    }
    const int col_lane = wn * 32 + 2 * q;   // + 8 j
    // this thread's 32 columns of tensor memory: lane quarter warp % 4 (the only one the warp may touch), lane = lane
    const uint32_t c_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
    // w_next goes back into the swizzled tile as 16-byte stores; a quarter warp (rows g = 2a, 2a + 1) would hit every
    // bank twice if all its lanes stored the same column pair j, so odd rows store pair j ^ 1 first (chunks
    // (q + 4 (j & 1)) ^ g: the two rows then cover all eight 16-byte chunks of the 128-byte line)
    const int odd = g & 1;
    const int chunk_a = ((q + 4 * odd) ^ g) << 4, chunk_b = ((q + 4 * (1 - odd)) ^ g) << 4;

    int s = 0;
    uint32_t ph = 0, wph = 0;
#pragma unroll 1
    for (int tile = 0; tile < tiles; ++tile) {
      const long long m0 = row_begin + (long long)tile * BM + group * BMG;   // first row of this group's half
      // 8-row groups of this warp's 16 rows that lie inside the CTA's range
      const long long left = row_end - (m0 + wm * 16);
      const int mi = left >= 16 ? 2 : (left <= 0 ? 0 : (int)((left + 7) >> 3));
      if (lw == 0 && lane == 0) {
        // the group's previous w tile was read and written through the generic proxy (every warp of the group is
        // past its last read: barrier of the last iteration); TMA overwrites it now
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&wbar[group], S::W_BYTES / 2);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(Wg + kb * kb_bytes, &tmW, &wbar[group], kb * BK, (int)m0);
      }
      // this thread's fragment of x_prev (registers) and c (private shared-memory column): rows beyond M are clamped
      // into the matrix, computed on and never stored
      const long long row_lane = m0 + wm * 16 + g;
      double2 prev[2][4];
      {
        uint32_t cbits[32];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          long long r = row_lane + 8 * i;
          if (r > a.M - 1) r = a.M - 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            prev[i][j] = *reinterpret_cast<const double2*>(a.x + r * a.ldx + col_lane + 8 * j);
            const double2 cc = *reinterpret_cast<const double2*>(a.c + r * a.ldc + col_lane + 8 * j);
            cbits[(i * 4 + j) * 4 + 0] = (uint32_t)__double2loint(cc.x);
            cbits[(i * 4 + j) * 4 + 1] = (uint32_t)__double2hiint(cc.x);
            cbits[(i * 4 + j) * 4 + 2] = (uint32_t)__double2loint(cc.y);
            cbits[(i * 4 + j) * 4 + 3] = (uint32_t)__double2hiint(cc.y);
          }
        }
        tmem_st32(c_taddr, cbits);
      }
      mbar_wait(&wbar[group], wph);
      wph ^= 1u;

#pragma unroll 1
      for (int it = 0; it < a.iters; ++it) {
        // the accumulators start from c, so the mainloop ends with z = c + w Q
        double acc[2][4][2];
        {
          uint32_t cbits[32];
          tmem_ld32_res(c_taddr, cbits);
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][j][0] = __hiloint2double((int)cbits[(i * 4 + j) * 4 + 1], (int)cbits[(i * 4 + j) * 4 + 0]);
              acc[i][j][1] = __hiloint2double((int)cbits[(i * 4 + j) * 4 + 3], (int)cbits[(i * 4 + j) * 4 + 2]);
            }
        }
        ++iter_no;
        volatile int* signal = nullptr;
        if (group == 0) {
          if (lw == 0) signal = skew;
        } else {
          while (*skew < iter_no) __nanosleep(32);   // stay skew_kb k-blocks behind group 0
        }
        if (mi == 2)
          resident_gemm<2>(acc, Wg, ring, full_bar, empty_bar, KB, kb_bytes, stage_bytes, stages, offA, offB, s, ph, lane,
                           a.zero, signal, skew_kb, iter_no);
        else if (mi == 1)
          resident_gemm<1>(acc, Wg, ring, full_bar, empty_bar, KB, kb_bytes, stage_bytes, stages, offA, offB, s, ph, lane,
                           a.zero, signal, skew_kb, iter_no);
        else
          resident_gemm<0>(acc, Wg, ring, full_bar, empty_bar, KB, kb_bytes, stage_bytes, stages, offA, offB, s, ph, lane,
                           a.zero, signal, skew_kb, iter_no);
        group_sync(group);   // every warp of the group is done reading its w tile

        const bool last = it == a.iters - 1;
        const double mom = a.momentum[it];
        // Three straight-line variants (inner iteration / last / last with the convergence test): with the mode
        // tested per element the compiler kept one basic block per column pair and the dependency chains of a
        // thread ran one after the other.
        auto update = [&](auto LAST, auto CHECK) {
          constexpr bool kLast = decltype(LAST)::value, kCheck = decltype(CHECK)::value;
          double thr0[4], thr1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col_lane + 8 * j;
            if constexpr (SHRINK == DECOMP_SHRINK_COMPLEX) {
              thr0[j] = thr1[j] = __ldg(a.thr + (col >> 1));
            } else {
              const double2 t = __ldg(reinterpret_cast<const double2*>(a.thr + col));   // col is even
              thr0[j] = t.x;
              thr1[j] = t.y;
            }
          }
          unsigned long long tb0[4], tb1[4];
          if constexpr (kCheck) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = col_lane + 8 * j;
              if constexpr (SHRINK == DECOMP_SHRINK_COMPLEX) {
                tb0[j] = tb1[j] = (unsigned long long)__double_as_longlong(__ldg(a.tol + (col >> 1)));
              } else {
                tb0[j] = (unsigned long long)__double_as_longlong(__ldg(a.tol + col));
                tb1[j] = (unsigned long long)__double_as_longlong(__ldg(a.tol + col + 1));
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const long long row = row_lane + 8 * i;
            const bool row_ok = row < row_end;
            unsigned char* wrow = Wg + (wm * 16 + g + 8 * i) * 128;
            double* grow = a.w + row * a.ldw + col_lane;
            double2 wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double z0 = acc[i][j][0], z1 = acc[i][j][1];
              double x0, x1, d0, d1;
              if constexpr (SHRINK == DECOMP_SHRINK_COMPLEX) {
                // z / (|z| + eps) * max(|z| - t, 0)   (lasso.py:210-225)
                const double rr = hypot(z0, z1);
                const double den = rr + kEps;
                const double mag = max_zero(rr - thr0[j]);
                x0 = mag * (z0 / den);
                x1 = mag * (z1 / den);
              } else if constexpr (SHRINK == DECOMP_SHRINK_POSITIVE) {
                x0 = max_zero(z0 - thr0[j]);   // lasso.py:228-241
                x1 = max_zero(z1 - thr1[j]);
              } else {
                // max(|z| - t, 0) * sign(z)   (lasso.py:206-207)
                x0 = with_sign_of(max_zero(fabs(z0) - thr0[j]), z0);
                x1 = with_sign_of(max_zero(fabs(z1) - thr1[j]), z1);
              }
              d0 = x0 - prev[i][j].x;
              d1 = x1 - prev[i][j].y;
              if constexpr (kCheck) {
                bool bad;
                if constexpr (SHRINK == DECOMP_SHRINK_COMPLEX)
                  bad = !(abs_bits(hypot(d0, d1)) < tb0[j]);
                else
                  bad = !(abs_bits(d0) < tb0[j]) || !(abs_bits(d1) < tb1[j]);
                violated |= bad && row_ok;
              }
              prev[i][j] = make_double2(x0, x1);
              // w_next = x_new + momentum * (x_new - x_prev)   (lasso.py:412)
              wv[j] = make_double2(x0 + mom * d0, x1 + mom * d1);
            }
            if constexpr (!kLast) {
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                // columns 16 m .. 16 m + 15 of this warp's 32 = k-block 2 wn + m of the tile
                unsigned char* blk = wrow + (2 * wn + m) * kb_bytes;
                const double2 va = odd ? wv[2 * m + 1] : wv[2 * m];
                const double2 vb = odd ? wv[2 * m] : wv[2 * m + 1];
                *reinterpret_cast<double2*>(blk + chunk_a) = va;
                *reinterpret_cast<double2*>(blk + chunk_b) = vb;
              }
            } else {
              if (row_ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(grow + 8 * j) = wv[j];
              }
            }
          }
        };
        if (!last)
          update(std::false_type{}, std::false_type{});
        else if (a.check == 0)
          update(std::true_type{}, std::false_type{});
        else
          update(std::true_type{}, std::true_type{});
        if (!last) group_sync(group);   // the group's w_next is complete
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const long long row = row_lane + 8 * i;
        if (row < row_end) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(a.x + row * a.ldx + col_lane + 8 * j) = prev[i][j];
        }
      }
    }
  }
  if (a.check) latch_vote(a.scratch, a.latch, a.latch_value, violated);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();       // nobody uses the tensor memory any more
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(RES_TMEM_COLS) : "memory");
  }
}

}  // namespace dcp

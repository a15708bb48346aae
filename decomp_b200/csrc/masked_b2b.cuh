// Back-to-back FP64 GEMM for the masked updates:   acc[m][n] = sum_j ((sum_k W[m][k] R[j][k]) * mask[m][j]) R[j][n]
//     masked ISTA / FISTA    dx = ((w A) * M) A^H            lasso.py:259-271      (W = w, R = the NT operand of w A)
//     masked NMF             neg = ((x D) * M) D^T           grads.py:112-115      (W = x, R = D^T)
// without the [rows, f] intermediate ever leaving the SM (the two-GEMM path writes it to HBM and reads it back:
// 16 B per element and iteration, and its second GEMM cannot start before the first has finished a tile).
// One operand serves both products: the second GEMM needs R^T, and for complex data the real embedding of A^H is the
// transpose of the embedding of A, so the same shared-memory tile of R is read with the two index roles swapped.
//
//   * a CTA owns 128 rows; warp w of 8 owns rows 16 w .. 16 w + 15 and ALL N = K1 <= 128 output columns, so the
//     contraction over j is never split across warps: acc (16 x 128 per warp) = 128 accumulator registers
//   * W [128, K1] is loaded once per row block (TMA, swizzled k-block layout) and is the A operand of the first GEMM
//   * R streams through a TMA ring in chunks of 16 rows j (16 x K1 doubles); per chunk a warp computes
//     P [16 x 16] = W R_chunk^T (K1 / 4 k-steps), multiplies by its mask fragment, and feeds P straight back as the
//     A operand of the second GEMM: the DMMA accumulator fragment (row g; columns 2q, 2q+1) IS an A fragment for the
//     two k-steps kk = 2q + s, provided the B fragment is read with the same k permutation -- B2[kk][n] =
//     R[j0 + 8 jb + 2q + s][8 nb + g], a conflict-free LDS.64 pattern on the swizzled tile (row 2q+s, chunk
//     (4 (nb & 1) + (g >> 1)) ^ (2q + s), half g & 1: 16 distinct 8-byte slots per half warp)
//   * shared-memory traffic: 0.75 LDS.64 per DMMA, as in the resident Lasso kernel
//   * the epilogue is the generic one (PROX real / complex / positive with convergence vote, or plain STORE), applied
//     from the accumulator fragments
#pragma once
#include "gemm.cuh"

namespace dcp {

constexpr int B2B_BM = 128, B2B_FC = 16, B2B_STAGES = 4, B2B_MMA_WARPS = 8;
constexpr int B2B_THREADS = B2B_MMA_WARPS * 32 + 128;   // + producer warp group

template <int KB1>   // K1 / 16
struct B2bSmem {
  static constexpr int K1 = KB1 * 16;
  static constexpr int W_BYTES = B2B_BM * K1 * 8;
  static constexpr int STAGE_BYTES = B2B_FC * K1 * 8;
  static constexpr int RING_OFF = W_BYTES;
  static constexpr int BAR_OFF = RING_OFF + B2B_STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * B2B_STAGES + 2) * 8;
};

struct B2bArgs {
  long long M;     // rows
  int K1, F;       // width of W / of the result; rows of R (channels, real width)
  int zero;
};

template <int KB1, int EPI>
__global__ void __launch_bounds__(B2B_THREADS, 1)
masked_b2b_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmR, const B2bArgs a,
                  const decomp_epilogue_t ep, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  using S = B2bSmem<KB1>;
  constexpr int K1 = S::K1, NB = K1 / 8;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Wt = smem;
  unsigned char* ring = smem + S::RING_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + B2B_STAGES;
  uint64_t* wfull = empty_bar + B2B_STAGES;   // the row block's W tile has landed
  uint64_t* wfree = wfull + 1;                // every MMA warp is done reading it

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (int)((a.M + B2B_BM - 1) / B2B_BM);
  const int chunks = (a.F + B2B_FC - 1) / B2B_FC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < B2B_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], B2B_MMA_WARPS);
    }
    mbar_init(wfull, 1);
    mbar_init(wfree, B2B_MMA_WARPS);
    fence_barrier_init();
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmR);
  }
  __syncthreads();

  bool violated = false;
  if (warp >= B2B_MMA_WARPS) {
    // ================================================================ producer warp group (one active thread)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == B2B_MMA_WARPS && lane == 0) {
      int s = 0;
      uint32_t ph = 0, fph = 0;
      bool wait = false, fwait = false;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (fwait) {
          mbar_wait(wfree, fph);
          fph ^= 1u;
        }
        fwait = true;
        mbar_arrive_expect_tx(wfull, S::W_BYTES);
        for (int kb = 0; kb < KB1; ++kb) tma_load_2d(Wt + kb * (B2B_BM * 128), &tmW, wfull, kb * 16, tile * B2B_BM);
        for (int c = 0; c < chunks; ++c) {
          if (wait) mbar_wait(&empty_bar[s], ph);
          mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
          unsigned char* st = ring + s * S::STAGE_BYTES;
          for (int kb = 0; kb < KB1; ++kb) tma_load_2d(st + kb * (B2B_FC * 128), &tmR, &full_bar[s], kb * 16, c * B2B_FC);
          if (++s == B2B_STAGES) {
            s = 0;
            if (wait) ph ^= 1u;
            wait = true;
          }
        }
      }
    }
  } else {
    // ================================================================ MMA warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = lane >> 2, q = lane & 3;
    int offA[4], offB1[4], offB2[2][2];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int o = (((s4 + 4 * (q >> 1)) ^ g) << 4) | ((q & 1) << 3);
      offA[s4] = (warp * 16 + g) * 128 + o;    // + i * 1024: rows 16 warp + g + 8 i
      offB1[s4] = g * 128 + o;                 // + j * 1024: chunk rows 8 j + g
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int nbp = 0; nbp < 2; ++nbp)
        offB2[s][nbp] = (2 * q + s) * 128 + (((4 * nbp + (g >> 1)) ^ (2 * q + s)) << 4) + ((g & 1) << 3);

    double step = 0.0;
    if constexpr (Epilogue<EPI>::kProx) step = *ep.step;
    int s = 0;
    uint32_t ph = 0, wph = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long m0 = (long long)tile * B2B_BM;
      const long long row_lane = m0 + warp * 16 + g;   // + 8 i
      long long mrow[2];                                // rows clamped into the matrix for the mask loads
#pragma unroll
      for (int i = 0; i < 2; ++i) mrow[i] = row_lane + 8 * i < a.M ? row_lane + 8 * i : a.M - 1;

      double acc[2][NB][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[i][nb][0] = acc[i][nb][1] = 0.0;

      mbar_wait(wfull, wph);
      wph ^= 1u;
      int held = -1;
#pragma unroll 1
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&full_bar[s], ph);
        // late release of the chunk read one step ago, see MmaPipe::run
        if (held >= 0 && lane == 0) mbar_arrive(&empty_bar[held]);
        const unsigned char* st = ring + s * S::STAGE_BYTES;

        // ---- mask fragment of this chunk (issued first: consumed after the first GEMM)
        double2 mk[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int col = c * B2B_FC + 8 * j + 2 * q;          // real channel column of P[i][j][0]
            mk[i][j] = make_double2(0.0, 0.0);
            if (col < a.F) {
              if (ep.cwidth == 2) {
                const double v = __ldg(ep.mask + mrow[i] * ep.ldmask + (col >> 1));
                mk[i][j] = make_double2(v, v);
              } else if (col + 1 < a.F) {
                mk[i][j] = __ldg(reinterpret_cast<const double2*>(ep.mask + mrow[i] * ep.ldmask + col));
              } else {
                mk[i][j].x = __ldg(ep.mask + mrow[i] * ep.ldmask + col);
              }
            }
          }

        // ---- first GEMM: P [16 x 16] = W rows . R_chunk^T, contraction over K1
        double p[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) p[i][j][0] = p[i][j][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < KB1; ++kb) {
          const unsigned char* sa = Wt + kb * (B2B_BM * 128);
          const unsigned char* sb = st + kb * (B2B_FC * 128);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            double fa[2], fb[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) fa[i] = *reinterpret_cast<const double*>(sa + offA[s4] + i * 1024);
#pragma unroll
            for (int j = 0; j < 2; ++j) fb[j] = *reinterpret_cast<const double*>(sb + offB1[s4] + j * 1024);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 2; ++j) dmma884(p[i][j][0], p[i][j][1], fa[i], fb[j]);
          }
        }
        if (c == chunks - 1) {
          // last read of the W tile: the producer may overwrite it with the next row block's
          const int dep = __double2hiint(p[0][0][0]) | __double2hiint(p[1][1][1]);
          if (lane == after(dep, a.zero)) mbar_arrive(wfree);
        }
        // ---- mask (grads.py:113, lasso.py:261)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            p[i][j][0] = mask_mul(p[i][j][0], mk[i][j].x);
            p[i][j][1] = mask_mul(p[i][j][1], mk[i][j].y);
          }
        // ---- second GEMM: acc [16 x K1] += P . R_chunk, the accumulator fragments of P as A fragments
#pragma unroll
        for (int jb = 0; jb < 2; ++jb)
#pragma unroll
          for (int ss = 0; ss < 2; ++ss) {
            double fb2[NB];
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
              fb2[nb] = *reinterpret_cast<const double*>(st + (nb >> 1) * (B2B_FC * 128) + jb * 1024 + offB2[ss][nb & 1]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) dmma884(acc[i][nb][0], acc[i][nb][1], p[i][jb][ss], fb2[nb]);
          }
        held = s;
        if (++s == B2B_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      {
        // the last chunk's stage: released once the accumulators that consumed it exist
        int dep = 0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) dep |= __double2hiint(acc[0][nb][0]) | __double2hiint(acc[1][nb][1]);
        if (held >= 0 && lane == after(dep, a.zero)) mbar_arrive(&empty_bar[held]);
      }

      // ---- epilogue from the fragments: row = row_lane + 8 i, column = 8 nb + 2 q (+ 1)
      constexpr int BATCH = 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const long long row = row_lane + 8 * i;
        const long long rowc = row < a.M ? row : a.M - 1;
#pragma unroll
        for (int nb0 = 0; nb0 < NB; nb0 += BATCH) {
          EpiIn in[BATCH];
          if constexpr (Epilogue<EPI>::kLoads) {
#pragma unroll
            for (int b = 0; b < BATCH; ++b) Epilogue<EPI>::load(ep, rowc, 8 * (nb0 + b) + 2 * q, in[b]);
          }
          if (row < a.M) {
#pragma unroll
            for (int b = 0; b < BATCH; ++b)
              violated |= Epilogue<EPI>::apply(ep, nullptr, 0, row, 8 * (nb0 + b) + 2 * q, true, acc[i][nb0 + b][0],
                                               acc[i][nb0 + b][1], in[b], step);
          }
        }
      }
    }
  }
  if constexpr (Epilogue<EPI>::kProx) {
    if (ep.check) convergence_latch(ep, violated);
  }
}

}  // namespace dcp

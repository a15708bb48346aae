// TF32-split ("3xTF32") kernels on the 5th-generation tensor cores (precision='tf32x3'): the unmasked ISTA/FISTA
// iteration, the NMF multiplicative update (x update fused into y D^T, split-K statistics x^T y / x^T x) and the masked
// models' [rows, f] intermediate (x D) * mask / (w A) * mask.
//
//   A B^T ~= A_lo B_hi^T + A_hi B_lo^T + A_hi B_hi^T          A = A_hi + A_lo, B = B_hi + B_lo (TF32 pieces: 11 + 11
//                                                             significant bits; three tcgen05.mma kind::tf32 per
//                                                             k-step, FP32 accumulators in tensor memory)
//
// tf32x3_gemm_kernel<MODE> (one CTA per SM, M = 128) and tf32x3_gemm_pair_kernel<MODE> (cluster of two CTAs,
// tcgen05.mma.cta_group::2, M = 256): warp-specialised, persistent.
//   warp 0   TMA producer: per 32-wide k-block the [128, 32] tiles of A_hi / A_lo and the [N, 32] tiles of B_hi / B_lo
//            (FP32 in memory, already TF32-valued) land in shared memory in the 128-byte-swizzled K-major layout that
//            the UMMA shared-memory descriptors describe (ring of stages, mbarrier full/empty)
//   warp 1   MMA issuer: one elected thread issues 3 x 4 tcgen05.mma (N <= 256, K = 8) per k-block into one of two
//            TMEM accumulators (2 x N columns) and commits to the ring's `empty` barrier / the accumulator's `full`
//            barrier
//   warp 2   allocates / frees the tensor memory
//   warps 4-11 epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> memory by MODE (STORE / PARTIAL / XUPD /
//            FMASK, see below; two warps per TMEM lane quarter take alternate chunks), then release the accumulator
// proxq_apply_kernel: the FP64 part of a Lasso iteration as one coalesced streaming pass
//   unmasked: z = c + P;  masked: z = w + step (c - P);  x_new = shrink(z, thr);  w_next = x_new + momentum (x_new -
//   x_prev) -> written directly as the TF32 pair (w_hi, w_lo) the next GEMM reads; convergence test + last-CTA latch as
//   in the FP64 kernels.
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace dcp {

constexpr double kEpsT = 1.0e-15;

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(double v, float& hi, float& lo) {
  hi = to_tf32(__double2float_rn(v));
  lo = to_tf32(__double2float_rn(v - (double)hi));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\n.reg .b32 r;\nelect.sync r|p, 0xffffffff;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, TF32 inputs, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// ------------------------------------------------------------------------------------------------ split kernels
__global__ void split_tf32_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols,
                                  float* __restrict__ hi, float* __restrict__ lo, long long ldh) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    float h, l;
    split_tf32(A[r * lda + c], h, l);
    hi[r * ldh + c] = h;
    lo[r * ldh + c] = l;
  }
}

__global__ void to_f32_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols,
                              float* __restrict__ out, long long ldo) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx % cols;
    out[r * ldo + c] = __double2float_rn(A[r * lda + c]);
  }
}

// ------------------------------------------------------------------------------------------------ GEMM
constexpr int TBM = 128, TBK = 32, TSTAGES = 2, TNMAX = 256;
constexpr int TA_BYTES = TBM * TBK * 4;          // 16 KB
constexpr int TB_BYTES = TNMAX * TBK * 4;        // 32 KB
constexpr int TSTAGE_BYTES = 2 * TA_BYTES + 2 * TB_BYTES;   // 96 KB
constexpr int TSMEM_BYTES = TSTAGES * TSTAGE_BYTES + 16 * 8;
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue -- two warps per TMEM lane quarter
// (warp % 4), taking alternate 32-column chunks of the accumulator
// (the NMF x update keeps four: its epilogue holds a chunk of x and of the denominator in registers besides the
// accumulator and needs the 255-register budget of a 256-thread CTA)
__host__ __device__ constexpr int tf_epi_warps(int mode) { return mode == 2 ? 4 : 8; }
__host__ __device__ constexpr int tf_threads(int mode) { return (4 + tf_epi_warps(mode)) * 32; }
// FMASK: one padded 32 x 32 staging tile per epilogue warp behind the barriers
__host__ __device__ constexpr int tf_tile_bytes(int mode) { return mode == 3 ? tf_epi_warps(mode) * 32 * 32 * 4 : 0; }

// what the epilogue warps do with a finished 128 x mma_n accumulator tile
constexpr int TF_STORE = 0;     // P[row][n0 + c] = acc                              (FP32)
constexpr int TF_PARTIAL = 1;   // P[(z * M + row)][n0 + c] = acc of K-split z        (FP32 slabs, reduced in FP64 later)
constexpr int TF_XUPD = 2;      // NMF x update: x <- x * max(acc, 0) / max(neg, eps)  (grads.py:84) in FP64, written as
                                // FP64 x, its TF32 pair row-major and its TF32 pair transposed
constexpr int TF_FMASK = 3;     // masked model: F = acc * mask (grads.py:112,122; lasso.py:262) written as a TF32 pair,
                                // row-major (A operand of F D^T) and / or K-blocked transposed (B operand of x^T F)

struct Tf32Args {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split, kb_total;
  int mma_n;              // N of one MMA = rows of one B box (multiple of 32, <= 256); tiles_n boxes cover N
  int n_fast;             // tile order: the n tiles of one m tile run side by side (B small enough to stay in L2: A is
                          // then read from HBM once instead of once per n tile)
  int blocked;            // operands stored K-blocked, [K / kblock][rows][kblock] with kblock = 32 kb_per_split: split z
                          // reads block z (3-D tensor maps).  A sample-axis contraction over row-major x^T, y^T would
                          // touch one 128-byte piece per row 4 n bytes apart -- hundreds of pages per TMA box
  long long xt_block;     // XUPD: block length of the K-blocked transposed output (0: plain [N][ldxt])
  float* P;               // STORE / PARTIAL destination
  long long ldp;
  // XUPD
  double* X;              // [M, N] in / out
  long long ldx;
  const float* NEG;       // XUPD: [M, N] x (D D^T);  FMASK: the [M, N / cw] mask (FP32) or null
  long long ldneg;
  int cw;                 // FMASK: 1 real, 2 complex (interleaved re / im columns share one mask entry)
  float* Xh;              // [M, N] TF32 pair of the new x (FMASK: of F, or null), row-major
  float* Xl;
  long long ldxh;
  float* XTh;             // [N, M] TF32 pair of the new x (FMASK: of F, or null), transposed (K-major for x^T y)
  float* XTl;
  long long ldxt;
};

// one 32-column chunk of an accumulator row (already in registers) -> memory, by epilogue mode
template <int MODE>
__device__ __forceinline__ void tf32_epilogue_chunk(const Tf32Args& a, const uint32_t (&v)[32], long long row, int n0,
                                                    int c0, int z) {
  if constexpr (MODE == TF_STORE || MODE == TF_PARTIAL) {
    const long long prow = MODE == TF_PARTIAL ? (long long)z * a.M + row : row;
    float* dst = a.P + prow * a.ldp + n0 + c0;
    if (n0 + c0 + 32 <= a.N) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + c0 + j < a.N) dst[j] = __uint_as_float(v[j]);
    }
  } else {
    // x <- x * max(pos, 0) / max(neg, eps), left to right like the reference (grads.py:84); N % 32 == 0 here.
    // All loads of the chunk are issued before its first store: x is updated in place, so the compiler may not move
    // a load above an earlier store by itself, and a thread that alternates 16-byte loads and stores pays one DRAM
    // round trip per four columns (the epilogue, not the MMAs, was then the critical path of this mode).
    double* xr = a.X + row * a.ldx + c0;
    const float* ng = a.NEG + row * a.ldneg + c0;
    float* xh = a.Xh + row * a.ldxh + c0;
    float* xl = a.Xl + row * a.ldxh + c0;
    double2 xin[16];
    float4 nin[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) xin[j] = *reinterpret_cast<const double2*>(xr + 2 * j);
#pragma unroll
    for (int j = 0; j < 8; ++j) nin[j] = *reinterpret_cast<const float4*>(ng + 4 * j);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const double xv[4] = {xin[j / 2].x, xin[j / 2].y, xin[j / 2 + 1].x, xin[j / 2 + 1].y};
      const float nv[4] = {nin[j / 4].x, nin[j / 4].y, nin[j / 4].z, nin[j / 4].w};
      double r[4];
      float h[4], l[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double pos = (double)__uint_as_float(v[j + t]);
        r[t] = __ddiv_rn(__dmul_rn(xv[t], fmax(pos, 0.0)), fmax((double)nv[t], kEpsT));
        split_tf32(r[t], h[t], l[t]);
        // a warp writes 32 consecutive rows: 128 contiguous bytes in either layout
        const long long ti = a.xt_block > 0
                                 ? ((row / a.xt_block) * a.N + (c0 + j + t)) * a.xt_block + row % a.xt_block
                                 : (long long)(c0 + j + t) * a.ldxt + row;
        a.XTh[ti] = h[t];
        a.XTl[ti] = l[t];
      }
      *reinterpret_cast<double2*>(xr + j) = make_double2(r[0], r[1]);
      *reinterpret_cast<double2*>(xr + j + 2) = make_double2(r[2], r[3]);
      *reinterpret_cast<float4*>(xh + j) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(xl + j) = make_float4(l[0], l[1], l[2], l[3]);
    }
  }
}

// ---- TF_FMASK epilogue: F = acc * mask in FP32 (one rounding, 2^-24: below the 3 x TF32 products' own 2^-21), split
// exactly into hi + lo.  With K = k (a few k-blocks per tile) this mode is all epilogue: 12 bytes of HBM traffic per
// accumulator element.  tcgen05.ld hands every thread one ROW of the tile, and a warp that stores 16-byte pieces of 32
// different rows (4 KB apart) writes half sectors into 32 lines per instruction -- measured 2.5 TB/s.  So a 32 x 32
// chunk is turned through a padded shared-memory tile: lane = column for the mask loads and the row-major stores
// (128 contiguous bytes per instruction), lane = row again for the transposed stores.  The mask values of a chunk are
// fetched one chunk ahead of the accumulator they multiply.
// The interior of the matrix (whole 32 x 32 chunks) takes a path of 16-byte accesses with the lanes laid out as 8
// column quads x 4 rows -- the first version with 4-byte accesses and per-element predicates executed ~1600 warp
// instructions per chunk and was issue-bound (ncu: 42 % issue active with two busy warps per scheduler).
constexpr int TF_TILE_BYTES = 32 * 32 * 4;                  // per epilogue warp
// staging tile: 32 rows of 8 16-byte column quads, quad q of row r stored at position q ^ (r & 7) -- a quarter warp
// that writes one quad of 8 consecutive rows, or reads the 8 quads of one row, touches all 32 banks once
__device__ __forceinline__ int tidx(int row, int col) {
  return row * 32 + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3));
}

__device__ __forceinline__ bool fmask_interior(const Tf32Args& a, long long r0, int col0) {
  return r0 + 32 <= a.M && col0 + 32 <= a.N;
}

// interior chunk: this lane's mask values for rows r0 + (lane >> 3) + 4 t, t = 0..7, columns col0 + 4 (lane & 7) .. + 3
__device__ __forceinline__ void fmask_prefetch(const Tf32Args& a, long long r0, int col0, int lane, float4 (&m)[8]) {
  if (a.NEG == nullptr || !fmask_interior(a, r0, col0)) {
#pragma unroll
    for (int t = 0; t < 8; ++t) m[t] = make_float4(1.f, 1.f, 1.f, 1.f);
    return;
  }
  const int cq = (lane & 7) * 4;
  const long long rr = r0 + (lane >> 3);
  if (a.cw == 2) {
    const float* mk = a.NEG + rr * a.ldneg + ((col0 + cq) >> 1);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float2 q = __ldcs(reinterpret_cast<const float2*>(mk + 4 * t * a.ldneg));
      m[t] = make_float4(q.x, q.x, q.y, q.y);
    }
  } else {
    const float* mk = a.NEG + rr * a.ldneg + col0 + cq;
#pragma unroll
    for (int t = 0; t < 8; ++t) m[t] = __ldcs(reinterpret_cast<const float4*>(mk + 4 * t * a.ldneg));
  }
}

__device__ __forceinline__ void split4(const float4 p, float4& h, float4& l) {
  h = make_float4(to_tf32(p.x), to_tf32(p.y), to_tf32(p.z), to_tf32(p.w));
  l = make_float4(to_tf32(p.x - h.x), to_tf32(p.y - h.y), to_tf32(p.z - h.z), to_tf32(p.w - h.w));
}

// v: this thread's row of the chunk (lane = row); m: fmask_prefetch's values
__device__ __forceinline__ void fmask_chunk_interior(const Tf32Args& a, const uint32_t (&v)[32], const float4 (&m)[8],
                                                     long long r0, int col0, float* tile, int lane) {
  const int cq = (lane & 7) * 4, rq = lane >> 3;
  if (a.Xh != nullptr) {
    // accumulator rows -> tile -> (4 rows x 8 column quads) per instruction: 128 contiguous bytes per row
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(tile + tidx(lane, 4 * j)) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                      __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    float* fh = a.Xh + (r0 + rq) * a.ldxh + col0 + cq;
    float* fl = a.Xl + (r0 + rq) * a.ldxh + col0 + cq;
    const long long step4 = 4 * a.ldxh;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float4 acc = *reinterpret_cast<const float4*>(tile + tidx(rq + 4 * t, cq));
      float4 h, l;
      split4(make_float4(acc.x * m[t].x, acc.y * m[t].y, acc.z * m[t].z, acc.w * m[t].w), h, l);
      *reinterpret_cast<float4*>(fh) = h;
      *reinterpret_cast<float4*>(fl) = l;
      fh += step4;
      fl += step4;
    }
    __syncwarp();
  }
  if (a.XTh != nullptr) {
    // mask -> tile -> this thread's row; a warp then writes 32 consecutive rows of one column: 128 contiguous bytes
#pragma unroll
    for (int t = 0; t < 8; ++t) *reinterpret_cast<float4*>(tile + tidx(rq + 4 * t, cq)) = m[t];
    __syncwarp();
    const long long row = r0 + lane;
    const long long tstride = a.xt_block > 0 ? a.xt_block : a.ldxt;
    const long long tbase = (a.xt_block > 0 ? (row / a.xt_block) * (long long)a.N * a.xt_block + row % a.xt_block : row) +
                            (long long)col0 * tstride;
    float* th = a.XTh + tbase;
    float* tl = a.XTl + tbase;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 mm = *reinterpret_cast<const float4*>(tile + tidx(lane, 4 * j));
      float4 h, l;
      split4(make_float4(__uint_as_float(v[4 * j]) * mm.x, __uint_as_float(v[4 * j + 1]) * mm.y,
                         __uint_as_float(v[4 * j + 2]) * mm.z, __uint_as_float(v[4 * j + 3]) * mm.w), h, l);
      th[0] = h.x; tl[0] = l.x;
      th[tstride] = h.y; tl[tstride] = l.y;
      th[2 * tstride] = h.z; tl[2 * tstride] = l.z;
      th[3 * tstride] = h.w; tl[3 * tstride] = l.w;
      th += 4 * tstride;
      tl += 4 * tstride;
    }
    __syncwarp();
  }
}

// edge chunk (ragged rows / columns), accumulator rows already in the tile: per-element predicates, 4-byte accesses,
// its own mask loads
// (not inlined, and the accumulator chunk comes through the tile: its register needs stay out of the interior path)
__device__ __noinline__ void fmask_chunk_edge(const Tf32Args& a, long long r0, int col0, float* tile, int lane) {
  const int col = col0 + lane;
  const bool live = col < a.N;
  const float* mk = a.NEG != nullptr ? a.NEG + r0 * a.ldneg + (a.cw == 2 ? col >> 1 : col) : nullptr;
  float p[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float m = mk == nullptr ? 1.f : ((live && r0 + i < a.M) ? mk[i * a.ldneg] : 0.f);
    p[i] = tile[tidx(i, lane)] * m;
  }
  if (a.Xh != nullptr) {
    float* fh = a.Xh + r0 * a.ldxh + col;
    float* fl = a.Xl + r0 * a.ldxh + col;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (live && r0 + i < a.M) {
        const float h = to_tf32(p[i]);
        fh[i * a.ldxh] = h;
        fl[i * a.ldxh] = to_tf32(p[i] - h);
      }
  }
  if (a.XTh != nullptr) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) tile[tidx(i, lane)] = p[i];
    __syncwarp();
    const long long row = r0 + lane;
    const long long tbase = a.xt_block > 0 ? (row / a.xt_block) * (long long)a.N * a.xt_block + row % a.xt_block : row;
    const long long tstride = a.xt_block > 0 ? a.xt_block : a.ldxt;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (row < a.M && col0 + j < a.N) {
        const float q = tile[tidx(lane, j)];
        const float h = to_tf32(q);
        const long long ti = tbase + (long long)(col0 + j) * tstride;
        a.XTh[ti] = h;
        a.XTl[ti] = to_tf32(q - h);
      }
  }
  __syncwarp();      // the tile is rewritten by the next chunk
}

// the epilogue warps' pass over one finished accumulator: 32-column chunks c0 = first, first + step, ...
// row: this thread's row (lane = row), r0 = row - lane; tile: this warp's staging tile (FMASK only)
template <int MODE>
__device__ __forceinline__ void tf32_epilogue_tile(const Tf32Args& a, uint32_t taddr, long long row, int n0, int N,
                                                   int z, int first, int step, uint64_t* acc_full, uint32_t acc_ph,
                                                   float* tile, int lane) {
  if constexpr (MODE == TF_FMASK) {
    const long long r0 = row - lane;
    float4 mcur[8];
    fmask_prefetch(a, r0, n0 + first, lane, mcur);       // in flight while the MMAs of this tile are still running
    mbar_wait(acc_full, acc_ph);
    tc_fence_after();
    for (int c0 = first; c0 < N; c0 += step) {
      float4 mnext[8];
      if (c0 + step < N) fmask_prefetch(a, r0, n0 + c0 + step, lane, mnext);
      uint32_t v[32];
      tmem_ld32(taddr + (uint32_t)c0, v);
      if (r0 < a.M && n0 + c0 < a.N) {                   // warp-uniform conditions
        if (fmask_interior(a, r0, n0 + c0)) {
          fmask_chunk_interior(a, v, mcur, r0, n0 + c0, tile, lane);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(tile + tidx(lane, 4 * j)) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
          __syncwarp();
          fmask_chunk_edge(a, r0, n0 + c0, tile, lane);
        }
      }
      if (c0 + step < N) {
#pragma unroll
        for (int t = 0; t < 8; ++t) mcur[t] = mnext[t];
      }
    }
  } else {
    mbar_wait(acc_full, acc_ph);
    tc_fence_after();
    for (int c0 = first; c0 < N; c0 += step) {
      uint32_t v[32];
      tmem_ld32(taddr + (uint32_t)c0, v);
      if (row >= a.M || n0 + c0 >= a.N) continue;
      tf32_epilogue_chunk<MODE>(a, v, row, n0, c0, z);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(tf_threads(MODE), 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const Tf32Args a, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TSTAGES * TSTAGE_BYTES);
  uint64_t* empty_bar = full_bar + TSTAGES;
  uint64_t* acc_full = empty_bar + TSTAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.mma_n;
  const int tiles_mn = a.tiles_m * a.tiles_n;
  const int items = tiles_mn * a.splits;
  // two accumulators of N columns each; allocation is a power of two >= 32 columns
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * N)) cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], tf_epi_warps(MODE));   // one arrival per epilogue warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmAl);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmBl);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (m tile, n tile, K split): consecutive items share the split and the B tile, so that CTAs running side by
  // side hit the same operand lines in L2
  auto decode = [&](int item, int& m0, int& n0, int& kb0, int& nkb, int& z) {
    const int tm = a.n_fast ? (item / a.tiles_n) % a.tiles_m : item % a.tiles_m;
    const int tn = a.n_fast ? item % a.tiles_n : (item / a.tiles_m) % a.tiles_n;
    z = item / tiles_mn;
    m0 = tm * TBM;
    n0 = tn * N;
    kb0 = z * a.kb_per_split;
    nkb = a.kb_total - kb0 < a.kb_per_split ? a.kb_total - kb0 : a.kb_per_split;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = 2u * TA_BYTES + 2u * (uint32_t)N * TBK * 4u;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int m0, n0, kb0, nkb, z;
        decode(item, m0, n0, kb0, nkb, z);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          unsigned char* st = smem + s * TSTAGE_BYTES;
          if (a.blocked) {
            const int k0 = kb * TBK;   // inside block z
            tma_load_3d(st, &tmAh, &full_bar[s], k0, m0, z);
            tma_load_3d(st + TA_BYTES, &tmAl, &full_bar[s], k0, m0, z);
            tma_load_3d(st + 2 * TA_BYTES, &tmBh, &full_bar[s], k0, n0, z);
            tma_load_3d(st + 2 * TA_BYTES + TB_BYTES, &tmBl, &full_bar[s], k0, n0, z);
          } else {
            const int k0 = (kb0 + kb) * TBK;
            tma_load_2d(st, &tmAh, &full_bar[s], k0, m0);
            tma_load_2d(st + TA_BYTES, &tmAl, &full_bar[s], k0, m0);
            tma_load_2d(st + 2 * TA_BYTES, &tmBh, &full_bar[s], k0, n0);
            tma_load_2d(st + 2 * TA_BYTES + TB_BYTES, &tmBl, &full_bar[s], k0, n0);
          }
          if (++s == TSTAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int m0, n0, kb0, nkb, z;
      decode(item, m0, n0, kb0, nkb, z);
      mbar_wait(&acc_empty[acc], acc_ph ^ 1u);   // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * N);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = smem_u32(smem + s * TSTAGE_BYTES);
          const uint32_t a_lo = a_hi + TA_BYTES;
          const uint32_t b_hi = a_hi + 2 * TA_BYTES;
          const uint32_t b_lo = b_hi + TB_BYTES;
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {
            const uint32_t off = (uint32_t)k * 32u;   // 8 TF32 = 32 bytes along K inside the swizzle atom
            const uint64_t dah = umma_desc_k_sw128(a_hi + off), dal = umma_desc_k_sw128(a_lo + off);
            const uint64_t dbh = umma_desc_k_sw128(b_hi + off), dbl = umma_desc_k_sw128(b_lo + off);
            umma_tf32(tmem_d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);   // small terms first
            umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
          tc_commit(&empty_bar[s]);                      // smem stage is free once these MMAs have read it
          if (kb == nkb - 1) tc_commit(&acc_full[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++s == TSTAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> memory
    const int e = (warp - 4) & 3;           // TMEM lane quarter this warp may access (warp % 4)
    const int half = (warp - 4) >> 2;       // which of the alternate chunks
    float* tile = reinterpret_cast<float*>(smem + TSTAGES * TSTAGE_BYTES + 16 * 8) + (warp - 4) * (32 * 32);
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int m0, n0, kb0, nkb, z;
      decode(item, m0, n0, kb0, nkb, z);
      const long long row = (long long)m0 + e * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * N);
      tf32_epilogue_tile<MODE>(a, taddr, row, n0, N, z, half * 32, tf_epi_warps(MODE) * 8, &acc_full[acc], acc_ph,
                               tile, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ CTA-pair GEMM
// The same GEMM with tcgen05.mma.cta_group::2: two CTAs of a cluster (the two SMs of a TPC) compute one 256 x N tile.
// Each CTA stages ITS 128 rows of A and ITS half of B's rows; the pair's MMA (issued by the leader CTA only, M = 256)
// reads A from each CTA's own shared memory and the two halves of B from both.  A stage is 64 KB instead of 96 KB
// (three stages fit instead of two) and each SM ingests 64 KB per k-block instead of 96 KB -- the operand stream into
// the SM, not the tensor pipe, is what bounds the single-CTA kernel.
// Barrier protocol (as in the public sm_100 2-SM GEMMs): full[s] lives in the leader, count 2 (one arrival per CTA's
// producer), both CTAs' TMA loads complete_tx on it; empty[s] and acc_full[b] exist in both CTAs and are signalled
// by the leader's tcgen05.commit multicast; acc_empty[b] lives in the leader, count 16 (8 epilogue warps x 2 CTAs).
constexpr int T2_STAGES = 3;
constexpr int T2_HALF_N_BYTES = (TNMAX / 2) * TBK * 4;                   // 16 KB: this CTA's half of a B box
constexpr int T2_STAGE_BYTES = 2 * TA_BYTES + 2 * T2_HALF_N_BYTES;        // 64 KB
constexpr int T2_SMEM_BYTES = T2_STAGES * T2_STAGE_BYTES + 16 * 8;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory, completion signalled on a barrier given as a shared::cluster address
// (the leader's): the .cta_group::2 form allows the barrier to sit in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same offset in both CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tf_threads(MODE), 1)
tf32x3_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                        const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                        const Tf32Args a, const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;     // uniform over the grid: both CTAs of a pair leave together
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T2_STAGES * T2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + T2_STAGES;
  uint64_t* acc_full = empty_bar + T2_STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int N = a.mma_n, NH = N / 2;             // this CTA stages NH rows of each B box
  const int tiles_m2 = (a.M + 2 * TBM - 1) / (2 * TBM);
  const int tiles_mn = tiles_m2 * a.tiles_n;
  const int items = tiles_mn * a.splits;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * N)) cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_STAGES; ++s) {
      mbar_init(&full_bar[s], 2);      // used in the leader: one arrival per CTA's producer
      mbar_init(&empty_bar[s], 1);     // one multicast commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);      // one multicast commit
      mbar_init(&acc_empty[b], 2 * tf_epi_warps(MODE));     // used in the leader: the epilogue warps of both CTAs
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmAl);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmBl);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                  // the peer's barriers are initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& m0, int& n0, int& kb0, int& nkb, int& z) {
    const int tm = a.n_fast ? (item / a.tiles_n) % tiles_m2 : item % tiles_m2;
    const int tn = a.n_fast ? item % a.tiles_n : (item / tiles_m2) % a.tiles_n;
    z = item / tiles_mn;
    m0 = tm * 2 * TBM;
    n0 = tn * N;
    kb0 = z * a.kb_per_split;
    nkb = a.kb_total - kb0 < a.kb_per_split ? a.kb_total - kb0 : a.kb_per_split;
  };
  const int first = (int)cluster_id_x(), stride = (int)num_clusters_x();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t pair_bytes = 2u * (2u * TA_BYTES + 2u * (uint32_t)NH * TBK * 4u);
      for (int item = first; item < items; item += stride) {
        int m0, n0, kb0, nkb, z;
        decode(item, m0, n0, kb0, nkb, z);
        const int am = m0 + (int)rank * TBM, bn = n0 + (int)rank * NH;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const uint32_t lead_full = mapa_u32(&full_bar[s], 0);
          unsigned char* st = smem + s * T2_STAGE_BYTES;
          if (a.blocked) {
            const int k0 = kb * TBK;
            tma_load_3d_pair(st, &tmAh, lead_full, k0, am, z);
            tma_load_3d_pair(st + TA_BYTES, &tmAl, lead_full, k0, am, z);
            tma_load_3d_pair(st + 2 * TA_BYTES, &tmBh, lead_full, k0, bn, z);
            tma_load_3d_pair(st + 2 * TA_BYTES + T2_HALF_N_BYTES, &tmBl, lead_full, k0, bn, z);
          } else {
            const int k0 = (kb0 + kb) * TBK;
            tma_load_2d_pair(st, &tmAh, lead_full, k0, am);
            tma_load_2d_pair(st + TA_BYTES, &tmAl, lead_full, k0, am);
            tma_load_2d_pair(st + 2 * TA_BYTES, &tmBh, lead_full, k0, bn);
            tma_load_2d_pair(st + 2 * TA_BYTES + T2_HALF_N_BYTES, &tmBl, lead_full, k0, bn);
          }
          if (leader)
            mbar_arrive_expect_tx(&full_bar[s], pair_bytes);
          else
            mbar_arrive_remote(lead_full);
          if (++s == T2_STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M = 256 >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * TBM) >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = first; item < items; item += stride) {
      int m0, n0, kb0, nkb, z;
      decode(item, m0, n0, kb0, nkb, z);
      mbar_wait(&acc_empty[acc], acc_ph ^ 1u);   // both CTAs' epilogues have drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * N);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = smem_u32(smem + s * T2_STAGE_BYTES);
          const uint32_t a_lo = a_hi + TA_BYTES;
          const uint32_t b_hi = a_hi + 2 * TA_BYTES;
          const uint32_t b_lo = b_hi + T2_HALF_N_BYTES;
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {
            const uint32_t off = (uint32_t)k * 32u;
            const uint64_t dah = umma_desc_k_sw128(a_hi + off), dal = umma_desc_k_sw128(a_lo + off);
            const uint64_t dbh = umma_desc_k_sw128(b_hi + off), dbl = umma_desc_k_sw128(b_lo + off);
            umma_tf32_pair(tmem_d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);   // small terms first
            umma_tf32_pair(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32_pair(tmem_d, dah, dbh, idesc, 1u);
          }
          tc_commit_pair(&empty_bar[s]);                      // both CTAs may refill the stage
          if (kb == nkb - 1) tc_commit_pair(&acc_full[acc]);  // both CTAs' epilogues may read their half
        }
        __syncwarp();
        if (++s == T2_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int e = (warp - 4) & 3, half = (warp - 4) >> 2;
    float* tile = reinterpret_cast<float*>(smem + T2_STAGES * T2_STAGE_BYTES + 16 * 8) + (warp - 4) * (32 * 32);
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = first; item < items; item += stride) {
      int m0, n0, kb0, nkb, z;
      decode(item, m0, n0, kb0, nkb, z);
      const long long row = (long long)m0 + (long long)rank * TBM + e * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * N);
      tf32_epilogue_tile<MODE>(a, taddr, row, n0, N, z, half * 32, tf_epi_warps(MODE) * 8, &acc_full[acc], acc_ph,
                               tile, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_u32(&acc_empty[acc], 0));
      if (++acc == 2) {
        acc = 0;
        acc_ph ^= 1u;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                  // nobody frees tensor memory or leaves while the peer still uses the pair
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
}

// out[m][n] = sum_z P[(z * M + m)][n] in FP64, fixed order (deterministic): the K-split slabs of TF_PARTIAL
__global__ void reduce_partials_f32_kernel(const float* __restrict__ P, long long ldp, int splits, long long M,
                                           long long N, double* __restrict__ out, long long ldo,
                                           const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  const long long total = M * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / N, n = idx % N;
    double acc = 0.0;
    for (int z = 0; z < splits; ++z) acc += (double)P[((long long)z * M + m) * ldp + n];
    out[m * ldo + n] = acc;
  }
}

// A (float64 [rows, cols]) -> TF32 pair of A^T (float32 [cols, rows]): 32 x 32 tiles through shared memory
// block > 0: K-blocked output [ceil(rows / block)][cols][block], the rows beyond `rows` of the last block zero-filled
__global__ void split_transpose_tf32_kernel(const double* __restrict__ A, long long lda, long long rows, long long cols,
                                            float* __restrict__ hiT, float* __restrict__ loT, long long ldt,
                                            long long block) {
  __shared__ float th[32][33], tl[32][33];
  const long long rows_out = block > 0 ? (rows + block - 1) / block * block : rows;
  const long long tiles_c = (cols + 31) / 32, tiles_r = (rows_out + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
  for (long long t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const long long r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = r0 + ty + 8 * i, c = c0 + tx;
      float h = 0.f, l = 0.f;
      if (r < rows && c < cols) split_tf32(A[r * lda + c], h, l);
      th[ty + 8 * i][tx] = h;
      tl[ty + 8 * i][tx] = l;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long c = c0 + ty + 8 * i, r = r0 + tx;
      if (r < rows_out && c < cols) {
        const long long o = block > 0 ? ((r / block) * cols + c) * block + r % block : c * ldt + r;
        hiT[o] = th[tx][ty + 8 * i];
        loT[o] = tl[tx][ty + 8 * i];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ FP64 pass
__device__ __forceinline__ double max_zero_t(double v) {
  const unsigned hi = (unsigned)__double2hiint(v);
  return (hi - 0x80000000u) <= 0x7ff00000u ? 0.0 : v;
}
__device__ __forceinline__ double with_sign_of_t(double mag, double s) {
  return __hiloint2double(__double2hiint(mag) | (__double2hiint(s) & 0x80000000), __double2loint(mag));
}

// one thread = one column pair; grid-stride over rows * N/2 pairs, row-contiguous
// PROX == false: z = other + P (the gradient step folded into Q and other, unmasked iteration)
// PROX == true:  z = w + step (other - P) with w = w_hi + w_lo, P = ((w A) * M) A^H, threshold step * alpha_col * rowvec
//                (masked iteration, lasso.py:259-271; the EPI_PROX epilogue of the FP64 GEMMs)
template <int SHRINK, bool PROX>
__global__ void __launch_bounds__(256)
proxq_apply_kernel(const float* __restrict__ P, long long ldp, const decomp_epilogue_t ep, float* __restrict__ w_hi,
                   float* __restrict__ w_lo, long long ldw, long long M, long long N,
                   const int* __restrict__ skip_if) {
  if (skip_if != nullptr && *skip_if != 0) return;
  const long long pairs = N / 2;
  const long long total = M * pairs;
  bool violated = false;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / pairs, col = (idx % pairs) * 2;
    const float2 p = *reinterpret_cast<const float2*>(P + row * ldp + col);
    const double2 c = *reinterpret_cast<const double2*>(ep.other + row * ep.ldother + col);
    const double2 xp = *reinterpret_cast<const double2*>(ep.prev + row * ep.ldprev + col);
    double z0, z1;
    if constexpr (PROX) {
      const float2 wh = *reinterpret_cast<const float2*>(w_hi + row * ldw + col);
      const float2 wl = *reinterpret_cast<const float2*>(w_lo + row * ldw + col);
      const double stepv = __ldg(ep.step);
      z0 = ((double)wh.x + (double)wl.x) + stepv * (c.x - (double)p.x);
      z1 = ((double)wh.y + (double)wl.y) + stepv * (c.y - (double)p.y);
    } else {
      z0 = c.x + (double)p.x;
      z1 = c.y + (double)p.y;
    }
    double t0, t1, tol0 = 0.0, tol1 = 0.0;
    if (SHRINK == DECOMP_SHRINK_COMPLEX) {
      t0 = t1 = __ldg(ep.colvec + (col >> 1));
      if (ep.check) tol0 = __ldg(ep.colvec2 + (col >> 1));
    } else {
      t0 = __ldg(ep.colvec + col);
      t1 = __ldg(ep.colvec + col + 1);
      if (ep.check) {
        tol0 = __ldg(ep.colvec2 + col);
        tol1 = __ldg(ep.colvec2 + col + 1);
      }
    }
    if constexpr (PROX) {
      const double stepv = __ldg(ep.step);
      if (ep.rowvec != nullptr) {
        const double rowfac = __ldg(ep.rowvec + row);
        t0 = stepv * (t0 * rowfac);
        t1 = stepv * (t1 * rowfac);
      } else if (!(ep.flags & DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD)) {
        t0 = stepv * t0;
        t1 = stepv * t1;
      }
    }
    double x0, x1;
    if (SHRINK == DECOMP_SHRINK_COMPLEX) {
      const double r = hypot(z0, z1);
      const double den = r + kEpsT;
      const double mag = max_zero_t(r - t0);
      x0 = mag * (z0 / den);
      x1 = mag * (z1 / den);
      if (ep.check) violated |= !(hypot(x0 - xp.x, x1 - xp.y) - tol0 < 0.0);
    } else {
      if (SHRINK == DECOMP_SHRINK_POSITIVE) {
        x0 = max_zero_t(z0 - t0);
        x1 = max_zero_t(z1 - t1);
      } else {
        x0 = with_sign_of_t(max_zero_t(fabs(z0) - t0), z0);
        x1 = with_sign_of_t(max_zero_t(fabs(z1) - t1), z1);
      }
      if (ep.check) violated |= !(fabs(x0 - xp.x) - tol0 < 0.0) || !(fabs(x1 - xp.y) - tol1 < 0.0);
    }
    *reinterpret_cast<double2*>(ep.out + row * ep.ldo + col) = make_double2(x0, x1);
    const double w0 = x0 + ep.momentum * (x0 - xp.x), w1 = x1 + ep.momentum * (x1 - xp.y);
    float h0, l0, h1, l1;
    split_tf32(w0, h0, l0);
    split_tf32(w1, h1, l1);
    *reinterpret_cast<float2*>(w_hi + row * ldw + col) = make_float2(h0, h1);
    *reinterpret_cast<float2*>(w_lo + row * ldw + col) = make_float2(l0, l1);
  }
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

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// FP32 [outer, inner] row-major tensor map, box {32 floats = 128 bytes, box_outer rows}, 128-byte swizzle
static int make_map_f32(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld,
                        uint32_t box_outer) {
  auto enc = encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DECOMP_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld & 3u) != 0) {
    set_error("TF32 GEMM operand must be 16-byte aligned with a row pitch that is a multiple of 4 floats");
    return DECOMP_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f32) failed with CUresult %d", (int)r);
    return DECOMP_ERR_CUDA;
  }
  return DECOMP_OK;
}

// K-blocked FP32 operand [blocks][rows][block]: 3-D map {block, rows, blocks}, box {32, box_rows, 1}
static int make_map_f32_blocked(CUtensorMap* map, const float* base, uint64_t block, uint64_t rows, uint64_t blocks,
                                uint32_t box_rows) {
  auto enc = encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DECOMP_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (block & 31u) != 0) {
    set_error("K-blocked TF32 operand must be 16-byte aligned with a block length that is a multiple of 32");
    return DECOMP_ERR_INVALID;
  }
  cuuint64_t gdim[3] = {block, rows, blocks};
  cuuint64_t gstride[2] = {block * sizeof(float), rows * block * sizeof(float)};
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f32, blocked) failed with CUresult %d", (int)r);
    return DECOMP_ERR_CUDA;
  }
  return DECOMP_OK;
}

}  // namespace dcp

using namespace dcp;

// shared launcher of the three epilogue modes
template <int MODE>
static int launch_tf32(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi, const float* B_lo,
                       int64_t ldb, Tf32Args& a, int64_t k_per_split, const int32_t* skip_if, void* stream) {
  if (a.M <= 0 || a.N <= 0) return DECOMP_OK;
  if (a.K <= 0) {
    set_error("tf32x3 GEMM: K must be positive");
    return DECOMP_ERR_INVALID;
  }
  a.mma_n = a.N >= TNMAX ? TNMAX : (int)((a.N + 31) / 32 * 32);
  a.tiles_m = (a.M + TBM - 1) / TBM;
  a.tiles_n = (a.N + a.mma_n - 1) / a.mma_n;
  a.kb_total = (a.K + TBK - 1) / TBK;
  a.kb_per_split = k_per_split > 0 ? (int)((k_per_split + TBK - 1) / TBK) : a.kb_total;
  // (a K-blocked operand keeps its block length whatever K is: the tensor map is built from kb_per_split)
  if (!a.blocked && a.kb_per_split > a.kb_total) a.kb_per_split = a.kb_total;
  a.splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  a.n_fast = a.tiles_n > 1 && (long long)a.N * a.K * 8 <= (16ll << 20) ? 1 : 0;
  // CTA pairs (tcgen05.mma.cta_group::2) when there are at least two 128-row tiles: each CTA stages half of B
  static const bool pair_enabled = [] {
    const char* e = getenv("DECOMP_TF32_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  const bool pair = pair_enabled && a.M > TBM && num_sms() >= 2;
  const uint32_t b_box = pair ? (uint32_t)a.mma_n / 2 : (uint32_t)a.mma_n;
  CUtensorMap tah, tal, tbh, tbl;
  int rc;
  if (a.blocked) {
    const uint64_t block = (uint64_t)a.kb_per_split * TBK, blocks = (uint64_t)a.splits;
    rc = make_map_f32_blocked(&tah, A_hi, block, (uint64_t)a.M, blocks, TBM);
    if (rc == DECOMP_OK) rc = make_map_f32_blocked(&tal, A_lo, block, (uint64_t)a.M, blocks, TBM);
    if (rc == DECOMP_OK) rc = make_map_f32_blocked(&tbh, B_hi, block, (uint64_t)a.N, blocks, b_box);
    if (rc == DECOMP_OK) rc = make_map_f32_blocked(&tbl, B_lo, block, (uint64_t)a.N, blocks, b_box);
  } else {
    rc = make_map_f32(&tah, A_hi, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)lda, TBM);
    if (rc == DECOMP_OK) rc = make_map_f32(&tal, A_lo, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)lda, TBM);
    if (rc == DECOMP_OK) rc = make_map_f32(&tbh, B_hi, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)ldb, b_box);
    if (rc == DECOMP_OK) rc = make_map_f32(&tbl, B_lo, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)ldb, b_box);
  }
  if (rc != DECOMP_OK) return rc;
  if (pair) {
    auto kern2 = tf32x3_gemm_pair_kernel<MODE>;
    static bool configured2 = false;   // per instantiation
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES + tf_tile_bytes(MODE));
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tf32x3 pair smem)");
      configured2 = true;
    }
    const long long tiles_m2 = (a.M + 2 * TBM - 1) / (2 * TBM);
    const long long items2 = tiles_m2 * a.tiles_n * a.splits;
    if (items2 > 2147483647LL) {
      set_error("tf32x3 GEMM: too many tiles");
      return DECOMP_ERR_INVALID;
    }
    long long clusters = num_sms() / 2;
    if (clusters > items2) clusters = items2;
    kern2<<<(unsigned)(2 * clusters), tf_threads(MODE), T2_SMEM_BYTES + tf_tile_bytes(MODE), as_stream(stream)>>>(tah, tal, tbh, tbl, a, skip_if);
    return check_cuda(cudaGetLastError(), "tf32x3 pair gemm launch");
  }
  auto kern = tf32x3_gemm_kernel<MODE>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TSMEM_BYTES + tf_tile_bytes(MODE));
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(tf32x3 smem)");
    configured = true;
  }
  const long long items = (long long)a.tiles_m * a.tiles_n * a.splits;
  if (items > 2147483647LL) {
    set_error("tf32x3 GEMM: too many tiles");
    return DECOMP_ERR_INVALID;
  }
  long long ctas = num_sms();
  if (ctas > items) ctas = items;
  kern<<<(unsigned)ctas, tf_threads(MODE), TSMEM_BYTES + tf_tile_bytes(MODE), as_stream(stream)>>>(tah, tal, tbh, tbl, a, skip_if);
  return check_cuda(cudaGetLastError(), "tf32x3 gemm launch");
}

template <bool PROX>
static int launch_prox_apply(const float* P, int64_t ldp, const decomp_epilogue_t* epi, float* w_hi, float* w_lo,
                             int64_t ldw, int64_t M, int64_t N, const int32_t* skip_if, void* stream) {
  if (M <= 0 || N <= 0) return DECOMP_OK;
  if (epi == nullptr || (N & 1) || epi->other == nullptr || epi->prev == nullptr || epi->out == nullptr ||
      epi->colvec == nullptr || (ldp & 1) || (ldw & 1) || (PROX && epi->step == nullptr) ||
      (epi->check && (epi->latch == nullptr || epi->scratch == nullptr))) {
    set_error("decomp_prox%s_apply_f64: invalid argument", PROX ? "" : "q");
    return DECOMP_ERR_INVALID;
  }
  long long b = (M * (N / 2) + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  cudaStream_t st = as_stream(stream);
  switch (epi->shrink) {
    case DECOMP_SHRINK_REAL:
      proxq_apply_kernel<DECOMP_SHRINK_REAL, PROX><<<(unsigned)b, 256, 0, st>>>(P, ldp, *epi, w_hi, w_lo, ldw, M, N,
                                                                                  skip_if);
      break;
    case DECOMP_SHRINK_COMPLEX:
      proxq_apply_kernel<DECOMP_SHRINK_COMPLEX, PROX><<<(unsigned)b, 256, 0, st>>>(P, ldp, *epi, w_hi, w_lo, ldw, M, N,
                                                                                     skip_if);
      break;
    case DECOMP_SHRINK_POSITIVE:
      proxq_apply_kernel<DECOMP_SHRINK_POSITIVE, PROX><<<(unsigned)b, 256, 0, st>>>(P, ldp, *epi, w_hi, w_lo, ldw, M, N,
                                                                                      skip_if);
      break;
    default:
      set_error("decomp_prox_apply: unknown shrink kind %d", epi->shrink);
      return DECOMP_ERR_INVALID;
  }
  return check_cuda(cudaGetLastError(), "prox_apply launch");
}

extern "C" {

int decomp_split_tf32_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, float* hi, float* lo, int64_t ldh,
                          void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  long long b = (rows * cols + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  split_tf32_kernel<<<(unsigned)b, 256, 0, as_stream(stream)>>>(A, lda, rows, cols, hi, lo, ldh);
  DCP_CHECK_LAUNCH("split_tf32");
  return DECOMP_OK;
}

int decomp_gemm_nt_tf32x3(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi, const float* B_lo,
                          int64_t ldb, int64_t M, int64_t N, int64_t K, float* P, int64_t ldp, const int32_t* skip_if,
                          void* stream) {
  if (M <= 0) return DECOMP_OK;
  if (N <= 0 || K <= 0 || (ldp & 3) != 0 || M > 2147483647LL || N > 2147483647LL || K > 2147483647LL) {
    set_error("decomp_gemm_nt_tf32x3: needs N > 0, K > 0, ldp %% 4 == 0 (N=%lld K=%lld)", (long long)N, (long long)K);
    return DECOMP_ERR_UNSUPPORTED;
  }
  Tf32Args a;
  memset(&a, 0, sizeof(a));
  a.M = (int)M;
  a.N = (int)N;
  a.K = (int)K;
  a.P = P;
  a.ldp = ldp;
  return launch_tf32<TF_STORE>(A_hi, A_lo, lda, B_hi, B_lo, ldb, a, 0, skip_if, stream);
}

size_t decomp_gemm_nt_tf32x3_splitk_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t k_per_split) {
  if (M <= 0 || N <= 0 || K <= 0 || k_per_split <= 0) return 0;
  const int64_t kb_total = (K + TBK - 1) / TBK;
  int64_t per = (k_per_split + TBK - 1) / TBK;
  if (per > kb_total) per = kb_total;
  const int64_t splits = (kb_total + per - 1) / per;
  const int64_t ldp = (N + 3) / 4 * 4;
  return (size_t)splits * (size_t)M * (size_t)ldp * sizeof(float);
}

int decomp_gemm_nt_tf32x3_splitk_f64(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi,
                                     const float* B_lo, int64_t ldb, int64_t M, int64_t N, int64_t K,
                                     int64_t k_per_split, int32_t k_blocked, double* out, int64_t ldo, void* workspace,
                                     size_t workspace_bytes, const int32_t* skip_if, void* stream) {
  if (M <= 0 || N <= 0) return DECOMP_OK;
  if (K <= 0 || k_per_split <= 0 || out == nullptr || M > 2147483647LL || N > 2147483647LL || K > 2147483647LL ||
      (k_blocked && (k_per_split % TBK) != 0)) {
    set_error("decomp_gemm_nt_tf32x3_splitk_f64: invalid argument (a K-blocked layout needs k_per_split %% 32 == 0)");
    return DECOMP_ERR_INVALID;
  }
  const size_t need = decomp_gemm_nt_tf32x3_splitk_workspace_bytes(M, N, K, k_per_split);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("decomp_gemm_nt_tf32x3_splitk_f64: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DECOMP_ERR_INVALID;
  }
  Tf32Args a;
  memset(&a, 0, sizeof(a));
  a.M = (int)M;
  a.N = (int)N;
  a.K = (int)K;
  a.P = reinterpret_cast<float*>(workspace);
  a.ldp = (N + 3) / 4 * 4;
  a.blocked = k_blocked ? 1 : 0;
  int rc = launch_tf32<TF_PARTIAL>(A_hi, A_lo, lda, B_hi, B_lo, ldb, a, k_per_split, skip_if, stream);
  if (rc != DECOMP_OK) return rc;
  long long b = (M * N + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  reduce_partials_f32_kernel<<<(unsigned)b, 256, 0, as_stream(stream)>>>(a.P, a.ldp, a.splits, M, N, out, ldo, skip_if);
  return check_cuda(cudaGetLastError(), "reduce_partials_f32 launch");
}

int decomp_nmf_xupdate_tf32x3(const float* Y_hi, const float* Y_lo, int64_t ldy, const float* D_hi, const float* D_lo,
                              int64_t ldd, int64_t n, int64_t k, int64_t f, double* X, int64_t ldx, const float* NEG,
                              int64_t ldneg, float* X_hi, float* X_lo, int64_t ldxh, float* XT_hi, float* XT_lo,
                              int64_t ldxt, int64_t xt_block, const int32_t* skip_if, void* stream) {
  if (n <= 0) return DECOMP_OK;
  if (k <= 0 || k > TNMAX || (k % 32) != 0 || f <= 0 || X == nullptr || NEG == nullptr || X_hi == nullptr ||
      X_lo == nullptr || XT_hi == nullptr || XT_lo == nullptr || (ldx & 1) || (ldneg & 3) || (ldxh & 3) ||
      n > 2147483647LL || f > 2147483647LL) {
    set_error("decomp_nmf_xupdate_tf32x3: needs k %% 32 == 0, k <= 256, even ldx, ldneg and ldxh multiples of 4");
    return DECOMP_ERR_UNSUPPORTED;
  }
  if (xt_block > 0 && (xt_block % 128) != 0) {
    set_error("decomp_nmf_xupdate_tf32x3: the block length of the transposed output must be a multiple of 128");
    return DECOMP_ERR_INVALID;
  }
  Tf32Args a;
  memset(&a, 0, sizeof(a));
  a.M = (int)n;
  a.N = (int)k;
  a.K = (int)f;
  a.X = X;
  a.ldx = ldx;
  a.NEG = NEG;
  a.ldneg = ldneg;
  a.Xh = X_hi;
  a.Xl = X_lo;
  a.ldxh = ldxh;
  a.XTh = XT_hi;
  a.XTl = XT_lo;
  a.ldxt = ldxt;
  a.xt_block = xt_block > 0 ? xt_block : 0;
  return launch_tf32<TF_XUPD>(Y_hi, Y_lo, ldy, D_hi, D_lo, ldd, a, 0, skip_if, stream);
}

int decomp_gemm_nt_mask_tf32x3(const float* A_hi, const float* A_lo, int64_t lda, const float* B_hi, const float* B_lo,
                               int64_t ldb, int64_t M, int64_t N, int64_t K, const float* mask, int64_t ldmask,
                               int32_t cwidth, float* F_hi, float* F_lo, int64_t ldf, float* FT_hi, float* FT_lo,
                               int64_t ldft, int64_t ft_block, const int32_t* skip_if, void* stream) {
  if (M <= 0 || N <= 0) return DECOMP_OK;
  if ((cwidth != 1 && cwidth != 2) || (cwidth == 2 && (N & 1))) {
    set_error("decomp_gemm_nt_mask_tf32x3: cwidth must be 1 or 2 (and N even when 2)");
    return DECOMP_ERR_INVALID;
  }
  const bool rowmajor = F_hi != nullptr, transposed = FT_hi != nullptr;
  if (K <= 0 || (!rowmajor && !transposed) || (rowmajor && (F_lo == nullptr || (ldf & 3))) ||
      (transposed && FT_lo == nullptr) || (mask != nullptr && (ldmask & 3)) || M > 2147483647LL || N > 2147483647LL ||
      K > 2147483647LL) {
    set_error("decomp_gemm_nt_mask_tf32x3: needs K > 0, an output, ldf and ldmask multiples of 4");
    return DECOMP_ERR_INVALID;
  }
  if (mask != nullptr && (reinterpret_cast<uintptr_t>(mask) & 15u) != 0) {
    set_error("decomp_gemm_nt_mask_tf32x3: the mask must be 16-byte aligned");
    return DECOMP_ERR_INVALID;
  }
  if (transposed && ft_block > 0 && (ft_block % 128) != 0) {
    set_error("decomp_gemm_nt_mask_tf32x3: the block length of the transposed output must be a multiple of 128");
    return DECOMP_ERR_INVALID;
  }
  Tf32Args a;
  memset(&a, 0, sizeof(a));
  a.M = (int)M;
  a.N = (int)N;
  a.K = (int)K;
  a.NEG = mask;
  a.ldneg = ldmask;
  a.cw = cwidth;
  a.Xh = F_hi;
  a.Xl = F_lo;
  a.ldxh = ldf;
  a.XTh = FT_hi;
  a.XTl = FT_lo;
  a.ldxt = ldft;
  a.xt_block = ft_block > 0 ? ft_block : 0;
  return launch_tf32<TF_FMASK>(A_hi, A_lo, lda, B_hi, B_lo, ldb, a, 0, skip_if, stream);
}

int decomp_to_f32_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, float* out, int64_t ldo, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  long long b = (rows * cols + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  to_f32_kernel<<<(unsigned)b, 256, 0, as_stream(stream)>>>(A, lda, rows, cols, out, ldo);
  DCP_CHECK_LAUNCH("to_f32");
  return DECOMP_OK;
}

int decomp_split_transpose_tf32_f64(const double* A, int64_t lda, int64_t rows, int64_t cols, float* hiT, float* loT,
                                    int64_t ldt, int64_t block, void* stream) {
  if (rows <= 0 || cols <= 0) return DECOMP_OK;
  const long long rows_out = block > 0 ? (rows + block - 1) / block * block : rows;
  long long tiles = ((rows_out + 31) / 32) * ((cols + 31) / 32);
  const long long cap = (long long)num_sms() * 32;
  if (tiles > cap) tiles = cap;
  split_transpose_tf32_kernel<<<(unsigned)tiles, 256, 0, as_stream(stream)>>>(A, lda, rows, cols, hiT, loT, ldt,
                                                                              block > 0 ? block : 0);
  DCP_CHECK_LAUNCH("split_transpose_tf32");
  return DECOMP_OK;
}

int decomp_proxq_apply_f64(const float* P, int64_t ldp, const decomp_epilogue_t* epi, float* w_hi, float* w_lo,
                           int64_t ldw, int64_t M, int64_t N, const int32_t* skip_if, void* stream) {
  return launch_prox_apply<false>(P, ldp, epi, w_hi, w_lo, ldw, M, N, skip_if, stream);
}

int decomp_prox_apply_f64(const float* P, int64_t ldp, const decomp_epilogue_t* epi, float* w_hi, float* w_lo,
                          int64_t ldw, int64_t M, int64_t N, const int32_t* skip_if, void* stream) {
  return launch_prox_apply<true>(P, ldp, epi, w_hi, w_lo, ldw, M, N, skip_if, stream);
}

}  // extern "C"

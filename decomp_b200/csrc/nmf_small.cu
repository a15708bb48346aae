// Whole NMF multiplicative-update runs of SMALL problems in one cooperative launch (BASELINE configs[0]: y 1000 x 200,
// k = 20, 100 sweeps; the sizes of the reference's own tests).  At these sizes a sweep is a few microseconds of
// arithmetic; launched kernel by kernel (12 launches per sweep) it costs ~100 us.  Here every CTA keeps its rows of
// y, mask and x and the whole dictionary in shared memory for all sweeps, and a sweep is
//     x update of the CTA's rows            (grads.py:77-84, 108-115)
//     partial statistics of the CTA's rows  -> global, one slab per CTA
//     grid barrier; fixed-order reduction of the slabs (deterministic); grid barrier
//     D update, l2_strict, max |D - D_new|  (grads.py:86-93, 117-125; normalize.py:13-21; batch_mu.py:21-23),
//     computed redundantly and identically by every CTA, so all of them take the same exit without exchanging it
// Plain FP64 FMAs (the problems are far too small for the tensor pipe to matter).  k <= 32.
#include <cuda_runtime.h>

#include <stdlib.h>

#include "common.h"

namespace dcp {

constexpr double kEpsS = 1.0e-15;
constexpr int SMALL_THREADS = 256, SMALL_KMAX = 32;

struct SmallNmfArgs {
  const double* y;
  long long ldy;
  const double* mask;   // nullptr: unmasked
  long long ldm;
  double* x;            // [n, k] in / out
  long long ldx;
  const double* D_in;   // [k, f], rows already normalised (nmf.py:70)
  long long ldd;
  double* D_out;
  long long ldo;
  int n, f, k, sweeps, rows_per_cta;
  double tol;
  int* it_out;          // 0, or the sweep at which max |D - D_new| < tol
  double* partials;     // [grid][slab] then [slab] reduced
  unsigned* counter;    // grid barrier, zeroed by the host
};

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum_s(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums of 16 values per lane over the warp in 16 shuffles instead of 80: every level of the xor butterfly also halves
// the number of values a lane carries (the lane keeps the half its bit selects and receives the partner's partial sums
// of that half).  Same pairs added at the same levels as warp_sum_s, so the sums are bitwise those.  On return lane l
// holds the total of value (l >> 1) & 15.
__device__ __forceinline__ double warp_sum16(double (&v)[16], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
  double w8[8], w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double recv = __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[i + 8], 16);
    w8[i] = (b4 ? v[i + 8] : v[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double recv = __shfl_xor_sync(0xffffffffu, b3 ? w8[i] : w8[i + 4], 8);
    w4[i] = (b3 ? w8[i + 4] : w8[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double recv = __shfl_xor_sync(0xffffffffu, b2 ? w4[i] : w4[i + 2], 4);
    w2[i] = (b2 ? w4[i + 2] : w4[i]) + recv;
  }
  const double recv = __shfl_xor_sync(0xffffffffu, b1 ? w2[0] : w2[1], 2);
  double r = (b1 ? w2[1] : w2[0]) + recv;
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

template <bool MASKED>
__global__ void __launch_bounds__(SMALL_THREADS, 1) nmf_mu_small_kernel(const SmallNmfArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, f = a.f, k = a.k, R = a.rows_per_cta;
  const int slab_len = MASKED ? 2 * k * f : k * f + k * k;   // per-CTA statistics: (T, NEGD) or (T, S)
  const int slab = (slab_len + 3) & ~3;                      // slab pitch: 32-byte aligned for the reduction's loads
  double* D_s = sm;                     // [k][f]
  double* Dn_s = D_s + k * f;           // [k][f]
  double* G_s = Dn_s + k * f;           // [k][k]  the reduced S = x^T x (unmasked)
  double* y_s = G_s + k * k;            // [R][f]  y (masked: y * mask)
  double* x_s = y_s + R * f;            // [R][k]
  double* m_s = x_s + R * k;            // [R][f]  mask                     (masked only)
  double* F_s = m_s + (MASKED ? R * f : 0);   // [R][f]  x D (masked: (x D) * mask)
  double* red = F_s + R * f;                  // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * R;
  const int rows = row0 < n ? (n - row0 < R ? n - row0 : R) : 0;
  unsigned target = 0;

  for (int e = tid; e < k * f; e += SMALL_THREADS) D_s[e] = a.D_in[(long long)(e / f) * a.ldd + e % f];
  for (int e = tid; e < rows * f; e += SMALL_THREADS) {
    const int r = e / f, j = e % f;
    double v = a.y[(long long)(row0 + r) * a.ldy + j];
    if (MASKED) {
      const double m = a.mask[(long long)(row0 + r) * a.ldm + j];
      m_s[e] = m;
      v *= m;                                            // y * mask, once (grads.py:113,123)
    }
    y_s[e] = v;
  }
  for (int e = tid; e < rows * k; e += SMALL_THREADS) x_s[e] = a.x[(long long)(row0 + e / k) * a.ldx + e % k];
  __syncthreads();

  for (int e = slab_len + tid; e < slab; e += SMALL_THREADS) a.partials[(long long)blockIdx.x * slab + e] = 0.0;
  int converged_at = 0;
  for (int it = 1; it <= a.sweeps; ++it) {
    // ---------------------------------------------------------------- x update of this CTA's rows
    for (int r = warp; r < rows; r += SMALL_THREADS / 32) {
      const double* yr = y_s + r * f;
      double* xr = x_s + r * k;
      // F = x D for the row, times the mask if there is one (grads.py:110, 112-115): the denominator is F D^T in
      // the reference's own association (no D D^T)
      for (int j = lane; j < f; j += 32) {
        double s = 0.0;
        for (int b = 0; b < k; ++b) s += xr[b] * D_s[b * f + j];
        F_s[r * f + j] = MASKED ? s * m_s[r * f + j] : s;
      }
      __syncwarp();
      double mine_pos = 0.0, mine_neg = 0.0;   // lane c ends up with the sums of atom c
      // eight atoms at a time: one pass over the row feeds eight independent accumulator chains
      for (int c0 = 0; c0 < k; c0 += 8) {
        double p[8], q[8];
        const double* dc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          p[u] = q[u] = 0.0;
          dc[u] = D_s + (c0 + u < k ? c0 + u : k - 1) * f;
        }
        for (int j = lane; j < f; j += 32) {
          const double yv = yr[j];
          const double fv = F_s[r * f + j];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double d = dc[u][j];
            p[u] += yv * d;
            q[u] += fv * d;
          }
        }
        double pq[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          pq[u] = p[u];
          pq[8 + u] = q[u];
        }
        const double tot = warp_sum16(pq, lane);         // lane l: total of value (l >> 1) & 15
        const int u = lane - c0;                         // lane c0 + u takes the sums of atom c0 + u
        const double gp = __shfl_sync(0xffffffffu, tot, (2 * u) & 31);
        const double gq = __shfl_sync(0xffffffffu, tot, (2 * u + 16) & 31);
        if (u >= 0 && u < 8) {
          mine_pos = gp;
          mine_neg = gq;
        }
      }
      __syncwarp();
      if (lane < k) xr[lane] = xr[lane] * fmax(mine_pos, 0.0) / fmax(mine_neg, kEpsS);
      __syncwarp();
      if (MASKED) {
        // F with the NEW x: the D update uses it (grads.py:122-125)
        for (int j = lane; j < f; j += 32) {
          double s = 0.0;
          for (int b = 0; b < k; ++b) s += xr[b] * D_s[b * f + j];
          F_s[r * f + j] = s * m_s[r * f + j];
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- statistics of this CTA's rows
    double* mine = a.partials + (long long)blockIdx.x * slab;
    for (int e = tid; e < k * f; e += SMALL_THREADS) {
      const int c = e / f, j = e % f;
      double t = 0.0, g = 0.0;
      for (int r = 0; r < rows; ++r) {
        t += x_s[r * k + c] * y_s[r * f + j];
        if (MASKED) g += x_s[r * k + c] * F_s[r * f + j];
      }
      mine[e] = t;
      if (MASKED) mine[k * f + e] = g;
    }
    if (!MASKED) {
      for (int e = tid; e < k * k; e += SMALL_THREADS) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += x_s[r * k + e / k] * x_s[r * k + e % k];
        mine[k * f + e] = s;
      }
    }
    grid_barrier(a.counter, target);
    // ---------------------------------------------------------------- fixed-order reduction over the CTAs
    // (one warp per entry: lane l adds the slabs l, l + 32, ... in order, then a fixed shuffle tree -- the same
    // association every time, and the ~gridDim.x loads of an entry are in flight together instead of one by one)
    double* total = a.partials + (long long)gridDim.x * slab;
    for (int e4 = blockIdx.x * (SMALL_THREADS / 32) + warp; e4 < slab / 4; e4 += gridDim.x * (SMALL_THREADS / 32)) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (unsigned c = lane; c < gridDim.x; c += 32) {
        const double2* src = reinterpret_cast<const double2*>(a.partials + (long long)c * slab) + 2 * e4;
        const double2 lo = __ldcg(src), hi = __ldcg(src + 1);
        s0 += lo.x;
        s1 += lo.y;
        s2 += hi.x;
        s3 += hi.y;
      }
      s0 = warp_sum_s(s0);
      s1 = warp_sum_s(s1);
      s2 = warp_sum_s(s2);
      s3 = warp_sum_s(s3);
      if (lane == 0) {
        reinterpret_cast<double2*>(total)[2 * e4] = make_double2(s0, s1);
        reinterpret_cast<double2*>(total)[2 * e4 + 1] = make_double2(s2, s3);
      }
    }
    grid_barrier(a.counter, target);
    // ---------------------------------------------------------------- D update (every CTA, identically)
    if (!MASKED) {
      for (int e = tid; e < k * k; e += SMALL_THREADS) G_s[e] = __ldcg(total + k * f + e);   // S = x^T x
      __syncthreads();
    }
    if (!MASKED && (k & 1) == 0) {
      // den[c][j] = sum_b S[c][b] D[b][j]: a thread owns column j and all k denominators; per b one load of D[b][j] and
      // S[b][c..c+1] (= S[c..c+1][b]: x^T x is bitwise symmetric) as broadcast 16-byte reads -- a third of the
      // shared-memory loads of the entry-per-thread form below, same summation order
      for (int j = tid; j < f; j += SMALL_THREADS) {
        double den[SMALL_KMAX], num[SMALL_KMAX];
#pragma unroll
        for (int c = 0; c < SMALL_KMAX; ++c) {
          den[c] = 0.0;
          num[c] = c < k ? __ldcg(total + c * f + j) : 0.0;
        }
        for (int b = 0; b < k; ++b) {
          const double d = D_s[b * f + j];
          const double2* g = reinterpret_cast<const double2*>(G_s + b * k);
#pragma unroll
          for (int c = 0; c < SMALL_KMAX; c += 2)
            if (c < k) {
              const double2 g2 = g[c >> 1];
              den[c] += g2.x * d;
              den[c + 1] += g2.y * d;
            }
        }
#pragma unroll
        for (int c = 0; c < SMALL_KMAX; ++c)
          if (c < k) Dn_s[c * f + j] = D_s[c * f + j] * fmax(num[c], 0.0) / fmax(den[c], kEpsS);
      }
    } else
    for (int e0 = tid; e0 < k * f; e0 += 4 * SMALL_THREADS) {      // four entries per thread in flight
      double den[4], num[4];
      int ee[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * SMALL_THREADS;
        ee[u] = e < k * f ? e : k * f - 1;
        num[u] = __ldcg(total + ee[u]);
        den[u] = MASKED ? __ldcg(total + k * f + ee[u]) : 0.0;
      }
      if (!MASKED) {
        for (int b = 0; b < k; ++b) {
#pragma unroll
          for (int u = 0; u < 4; ++u) den[u] += G_s[(ee[u] / f) * k + b] * D_s[b * f + ee[u] % f];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u * SMALL_THREADS < k * f) Dn_s[ee[u]] = D_s[ee[u]] * fmax(num[u], 0.0) / fmax(den[u], kEpsS);
    }
    __syncthreads();
    double md = 0.0;
    for (int c = warp; c < k; c += SMALL_THREADS / 32) {
      double s = 0.0;
      for (int j = lane; j < f; j += 32) s += Dn_s[c * f + j] * Dn_s[c * f + j];
      const double nrm = sqrt(warp_sum_s(s));           // l2_strict: no floor (normalize.py:13-21)
      for (int j = lane; j < f; j += 32) {
        const double v = Dn_s[c * f + j] / nrm;
        const double d = fabs(D_s[c * f + j] - v);
        md = (d != d) ? d : ((md != md) ? md : fmax(md, d));   // NaN wins like numpy's max
        Dn_s[c * f + j] = v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, md, o);
      md = (other != other) ? other : ((md != md) ? md : fmax(md, other));
    }
    if (lane == 0) red[warp] = md;
    __syncthreads();
    for (int e = tid; e < k * f; e += SMALL_THREADS) D_s[e] = Dn_s[e];
    double all = red[0];
    for (int w = 1; w < SMALL_THREADS / 32; ++w) {
      const double other = red[w];
      all = (other != other) ? other : ((all != all) ? all : fmax(all, other));
    }
    __syncthreads();
    if (a.tol > 0.0 && all < a.tol) {      // batch_mu.py:22-23: identical on every CTA
      converged_at = it;
      break;
    }
    // the partial slabs are rewritten next sweep: everybody must be past the reduction (second barrier above) --
    // and the reduced slab `total` is only rewritten after the next first barrier, when all CTAs have read it
  }
  for (int e = tid; e < rows * k; e += SMALL_THREADS) a.x[(long long)(row0 + e / k) * a.ldx + e % k] = x_s[e];
  if (blockIdx.x == 0) {
    for (int e = tid; e < k * f; e += SMALL_THREADS) a.D_out[(long long)(e / f) * a.ldo + e % f] = D_s[e];
    if (tid == 0) *a.it_out = converged_at;
  }
}

static size_t small_smem_bytes(int f, int k, int R, bool masked) {
  return sizeof(double) * ((size_t)2 * k * f + (size_t)k * k + (size_t)R * f * (masked ? 3 : 2) + (size_t)R * k + 32);
}

static void small_plan(int64_t n, int* grid, int* R) {
  const int sms = num_sms();
  long long r = (n + sms - 1) / sms;
  // DECOMP_SMALL_ROWS: lower bound on the rows per CTA (fewer CTAs: cheaper grid barriers and slab sums, longer row pass)
  static const long long min_rows = [] {
    const char* e = getenv("DECOMP_SMALL_ROWS");
    return e != nullptr ? atoll(e) : 0ll;
  }();
  if (r < min_rows) r = min_rows;
  if (r < 1) r = 1;
  *R = (int)r;
  *grid = (int)((n + r - 1) / r);
}

}  // namespace dcp

using namespace dcp;

extern "C" {

int decomp_nmf_mu_small_supported(int64_t n, int64_t f, int64_t k, int32_t masked) {
  if (n <= 0 || f <= 0 || k <= 0 || k > SMALL_KMAX || n > (1 << 20) || f > (1 << 16)) return 0;
  int grid, R;
  small_plan(n, &grid, &R);
  return small_smem_bytes((int)f, (int)k, R, masked != 0) <= (size_t)200 * 1024;
}

size_t decomp_nmf_mu_small_workspace_bytes(int64_t n, int64_t f, int64_t k, int32_t masked) {
  if (!decomp_nmf_mu_small_supported(n, f, k, masked)) return 0;
  int grid, R;
  small_plan(n, &grid, &R);
  size_t slab = masked ? (size_t)2 * k * f : (size_t)k * f + (size_t)k * k;
  slab = (slab + 3) & ~(size_t)3;
  return sizeof(double) * slab * ((size_t)grid + 1) + 16;   // slabs + the reduced slab + the barrier counter
}

int decomp_nmf_mu_small_f64(const double* y, int64_t ldy, const double* mask, int64_t ldm, double* x, int64_t ldx,
                            const double* D_in, int64_t ldd, double* D_out, int64_t ldo, int64_t n, int64_t f,
                            int64_t k, int32_t sweeps, double tol, int32_t* it_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (!decomp_nmf_mu_small_supported(n, f, k, mask != nullptr) || y == nullptr || x == nullptr || D_in == nullptr ||
      D_out == nullptr || it_out == nullptr || sweeps < 0) {
    set_error("decomp_nmf_mu_small_f64: unsupported size or invalid argument");
    return DECOMP_ERR_INVALID;
  }
  const size_t need = decomp_nmf_mu_small_workspace_bytes(n, f, k, mask != nullptr);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("decomp_nmf_mu_small_f64: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DECOMP_ERR_INVALID;
  }
  int grid, R;
  small_plan(n, &grid, &R);
  const bool masked = mask != nullptr;
  size_t slab = masked ? (size_t)2 * k * f : (size_t)k * f + (size_t)k * k;
  slab = (slab + 3) & ~(size_t)3;
  SmallNmfArgs a;
  a.y = y;
  a.ldy = ldy;
  a.mask = mask;
  a.ldm = ldm;
  a.x = x;
  a.ldx = ldx;
  a.D_in = D_in;
  a.ldd = ldd;
  a.D_out = D_out;
  a.ldo = ldo;
  a.n = (int)n;
  a.f = (int)f;
  a.k = (int)k;
  a.sweeps = sweeps;
  a.rows_per_cta = R;
  a.tol = tol;
  a.it_out = it_out;
  a.partials = reinterpret_cast<double*>(workspace);
  a.counter = reinterpret_cast<unsigned*>(a.partials + slab * ((size_t)grid + 1));
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(a.counter, 0, 16, st);
  if (e != cudaSuccess) return check_cuda(e, "nmf_small counter");
  const size_t smem = small_smem_bytes((int)f, (int)k, R, masked);
  void* kern = masked ? (void*)nmf_mu_small_kernel<true> : (void*)nmf_mu_small_kernel<false>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return check_cuda(e, "nmf_small smem");
  void* args[] = {(void*)&a};
  e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(SMALL_THREADS), args, smem, st);
  return check_cuda(e, "nmf_small launch");
}

}  // extern "C"

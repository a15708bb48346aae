// C-ABI entry points of the FP64 GEMM family (see include/decomp_b200.h).
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.h"
#include "gemm.cuh"
#include "lasso_resident.cuh"
#include "masked_b2b.cuh"

namespace dcp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return DECOMP_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DECOMP_ERR_CUDA;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* map, const double* base, uint64_t inner, uint64_t outer, uint64_t ld,
                    uint32_t box_inner, uint32_t box_outer) {
  auto enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DECOMP_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld & 1u) != 0) {
    set_error("GEMM operand must be 16-byte aligned with an even leading dimension (ptr=%p ld=%llu)", (const void*)base,
              (unsigned long long)ld);
    return DECOMP_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {ld * sizeof(double)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return DECOMP_ERR_CUDA;
  }
  return DECOMP_OK;
}

int make_tensor_map_tn3d(CUtensorMap* map, const double* base, uint64_t width, uint64_t rows, uint64_t ld,
                         uint32_t blocks) {
  auto enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return DECOMP_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld & 1u) != 0 || (width & 15u) != 0) return DECOMP_ERR_INVALID;
  cuuint64_t gdim[3] = {16, rows, width / 16};
  cuuint64_t gstride[2] = {ld * sizeof(double), 16 * sizeof(double)};
  cuuint32_t box[3] = {16, BK, blocks};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DECOMP_OK : DECOMP_ERR_UNSUPPORTED;
}

// CTA tile 128x64: 8 MMA warps of 32x32, 4 epilogue warps (batches of 4 column pairs per thread), 3-stage operand
// ring (72 KB) + two 72 KB accumulator staging buffers = 216 KB -> one persistent CTA per SM.
using CfgMain = GemmCfg<128, 64, 32, 32, 3, 2, 4, 4>;
// the fused ISTA/FISTA kernel has no accumulator staging: 4-stage ring (96 KB) + two 64 KB operand tiles
using CfgProxq = GemmCfg<128, 64, 32, 32, 4, 2, 4, 4>;

template <class C, bool TN, int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmGeom& gs, const decomp_epilogue_t& ep,
                  double* partial, const int32_t* skip_if, cudaStream_t stream) {
  auto kern = gemm_f64_kernel<C, TN, EPI>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(gemm smem)");
    configured = true;
  }
  const long long tiles = (long long)gs.tiles_m * gs.tiles_n * gs.splits;
  if (tiles <= 0) return DECOMP_OK;
  if (tiles > 2147483647LL) {
    set_error("GEMM tile count too large");
    return DECOMP_ERR_INVALID;
  }
  long long ctas = num_sms();   // persistent grid: one CTA per SM walks the tile list
  if (ctas > tiles) ctas = tiles;
  kern<<<(unsigned)ctas, C::THREADS, C::SMEM_BYTES, stream>>>(ta, tb, gs, ep, partial, skip_if);
  return check_cuda(cudaGetLastError(), "gemm launch");
}

template <class C, int EPI>
static int launch_proxq(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp,
                        const GemmGeom& gs, const decomp_epilogue_t& ep, const int32_t* skip_if, cudaStream_t stream) {
  auto kern = gemm_f64_proxq_kernel<C, EPI>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ProxqSmem<C>::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(proxq smem)");
    configured = true;
  }
  const long long tiles = (long long)gs.tiles_m * gs.tiles_n;
  if (tiles <= 0) return DECOMP_OK;
  long long ctas = num_sms();
  if (ctas > tiles) ctas = tiles;
  kern<<<(unsigned)ctas, C::MMA_THREADS + 128, ProxqSmem<C>::SMEM_BYTES, stream>>>(ta, tb, tc, tp, gs, ep, skip_if);
  return check_cuda(cudaGetLastError(), "proxq launch");
}

template <int SHRINK>
static int launch_resident(const CUtensorMap& tw, const CUtensorMap& tq, const ResidentArgs& a, const int32_t* skip_if,
                           cudaStream_t stream) {
  auto kern = lasso_resident_kernel<SHRINK>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ResidentSmem::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(resident smem)");
    configured = true;
  }
  long long groups = (a.M + 7) / 8, ctas = num_sms();   // rows are dealt out in groups of 8
  if (ctas > groups) ctas = groups;
  kern<<<(unsigned)ctas, RES_THREADS, ResidentSmem::SMEM_BYTES, stream>>>(tw, tq, a, skip_if);
  return check_cuda(cudaGetLastError(), "resident lasso launch");
}

template <int KB1, int EPI>
static int launch_b2b(const CUtensorMap& tw, const CUtensorMap& tr, const B2bArgs& a, const decomp_epilogue_t& ep,
                      const int32_t* skip_if, cudaStream_t stream) {
  auto kern = masked_b2b_kernel<KB1, EPI>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, B2bSmem<KB1>::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(b2b smem)");
    configured = true;
  }
  long long tiles = (a.M + B2B_BM - 1) / B2B_BM, ctas = num_sms();
  if (ctas > tiles) ctas = tiles;
  kern<<<(unsigned)ctas, B2B_THREADS, B2bSmem<KB1>::SMEM_BYTES, stream>>>(tw, tr, a, ep, skip_if);
  return check_cuda(cudaGetLastError(), "b2b launch");
}

template <int KB1>
static int dispatch_b2b(const CUtensorMap& tw, const CUtensorMap& tr, const B2bArgs& a, const decomp_epilogue_t& ep,
                        const int32_t* skip_if, cudaStream_t st) {
  switch (ep.kind) {
    case DECOMP_EPI_STORE:
      return launch_b2b<KB1, DECOMP_EPI_STORE>(tw, tr, a, ep, skip_if, st);
    case DECOMP_EPI_PROX:
      switch (ep.shrink) {
        case DECOMP_SHRINK_REAL:
          return launch_b2b<KB1, EPI_PROX_REAL>(tw, tr, a, ep, skip_if, st);
        case DECOMP_SHRINK_COMPLEX:
          return launch_b2b<KB1, EPI_PROX_COMPLEX>(tw, tr, a, ep, skip_if, st);
        case DECOMP_SHRINK_POSITIVE:
          return launch_b2b<KB1, EPI_PROX_POSITIVE>(tw, tr, a, ep, skip_if, st);
        default:
          break;
      }
    default:
      set_error("decomp_gemm_b2b_masked_f64: epilogue must be STORE or PROX");
      return DECOMP_ERR_INVALID;
  }
}

static void tn_plan(long long M, long long N, long long K, GemmGeom* gs) {
  using C = CfgMain;
  gs->M = M;
  gs->N = N;
  gs->K = K;
  gs->tiles_m = (int)((M + C::BM - 1) / C::BM);
  gs->tiles_n = (int)((N + C::BN - 1) / C::BN);
  gs->kblocks_total = (int)((K + BK - 1) / BK);
  const long long tiles = (long long)gs->tiles_m * gs->tiles_n;
  // Split the contraction so that the persistent grid (one CTA per SM) is evenly loaded: minimise
  //   rounds(tiles * splits) * (k-blocks per split + fixed cost of parking / writing one partial tile)
  // over the split count; >= 128 rows (8 k-blocks) per split.
  const long long sms = num_sms();
  const long long max_splits = gs->kblocks_total / 8 > 0 ? gs->kblocks_total / 8 : 1;
  long long splits = 1;
  double best = 0.0;
  for (long long cand = 1; cand <= max_splits && cand <= 4096; ++cand) {
    const long long per = (gs->kblocks_total + cand - 1) / cand;
    const long long real = (gs->kblocks_total + per - 1) / per;   // splits that actually get work
    const long long rounds = (tiles * real + sms - 1) / sms;
    const double cost = (double)rounds * ((double)per + 6.0);
    if (cand == 1 || cost < best * 0.999) {
      best = cost;
      splits = cand;
    }
  }
  gs->kblocks_per_split = (int)((gs->kblocks_total + splits - 1) / splits);
  if (gs->kblocks_per_split < 1) gs->kblocks_per_split = 1;
  gs->splits = (gs->kblocks_total + gs->kblocks_per_split - 1) / gs->kblocks_per_split;
  if (gs->splits < 1) gs->splits = 1;
  gs->ld_partial = (N + 1) & ~1LL;
  gs->m_fast = 1;
}

}  // namespace dcp

using namespace dcp;

extern "C" {

const char* decomp_last_error(void) { return g_err; }
int decomp_abi_version(void) { return DECOMP_ABI_VERSION; }

int decomp_gemm_nt_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                       const decomp_epilogue_t* epi, const int32_t* skip_if, void* stream) {
  using C = CfgMain;
  if (epi == nullptr || M < 0 || N < 0 || K < 0) {
    set_error("decomp_gemm_nt_f64: invalid argument");
    return DECOMP_ERR_INVALID;
  }
  if (M == 0 || N == 0) return DECOMP_OK;
  if (K == 0) {
    set_error("decomp_gemm_nt_f64: K must be positive");
    return DECOMP_ERR_INVALID;
  }
  CUtensorMap ta, tb;
  int rc = make_tensor_map(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, C::BM);
  if (rc != DECOMP_OK) return rc;
  rc = make_tensor_map(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, C::BN);
  if (rc != DECOMP_OK) return rc;
  GemmGeom gs;
  gs.M = M;
  gs.N = N;
  gs.K = K;
  gs.tiles_m = (int)((M + C::BM - 1) / C::BM);
  gs.tiles_n = (int)((N + C::BN - 1) / C::BN);
  gs.splits = 1;
  gs.kblocks_total = (int)((K + BK - 1) / BK);
  gs.kblocks_per_split = gs.kblocks_total;
  gs.ld_partial = 0;
  gs.tn3d = 0;
  gs.zero = 0;
  gs.m_fast = 0;
  cudaStream_t st = as_stream(stream);
  switch (epi->kind) {
    case DECOMP_EPI_STORE:
      return launch<C, false, DECOMP_EPI_STORE>(ta, tb, gs, *epi, nullptr, skip_if, st);
    case DECOMP_EPI_STORE_MASK:
      return launch<C, false, DECOMP_EPI_STORE_MASK>(ta, tb, gs, *epi, nullptr, skip_if, st);
    case DECOMP_EPI_MU_NUM:
      return launch<C, false, DECOMP_EPI_MU_NUM>(ta, tb, gs, *epi, nullptr, skip_if, st);
    case DECOMP_EPI_MU_DEN:
      return launch<C, false, DECOMP_EPI_MU_DEN>(ta, tb, gs, *epi, nullptr, skip_if, st);
    case DECOMP_EPI_PROX:
      if (epi->check && (epi->latch == nullptr || epi->scratch == nullptr)) {
        set_error("PROX epilogue with check needs latch and scratch");
        return DECOMP_ERR_INVALID;
      }
      switch (epi->shrink) {
        case DECOMP_SHRINK_REAL:
          return launch<C, false, EPI_PROX_REAL>(ta, tb, gs, *epi, nullptr, skip_if, st);
        case DECOMP_SHRINK_COMPLEX:
          return launch<C, false, EPI_PROX_COMPLEX>(ta, tb, gs, *epi, nullptr, skip_if, st);
        case DECOMP_SHRINK_POSITIVE:
          return launch<C, false, EPI_PROX_POSITIVE>(ta, tb, gs, *epi, nullptr, skip_if, st);
        default:
          set_error("decomp_gemm_nt_f64: unknown shrink kind %d", epi->shrink);
          return DECOMP_ERR_INVALID;
      }
    case DECOMP_EPI_KL_RATIO:
      return launch<C, false, DECOMP_EPI_KL_RATIO>(ta, tb, gs, *epi, nullptr, skip_if, st);
    case DECOMP_EPI_PROXQ: {
      if (epi->check && (epi->latch == nullptr || epi->scratch == nullptr)) {
        set_error("PROXQ epilogue with check needs latch and scratch");
        return DECOMP_ERR_INVALID;
      }
      if (epi->other == nullptr || epi->prev == nullptr || epi->out == nullptr || epi->colvec == nullptr ||
          (!(epi->flags & DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD) && epi->step == nullptr)) {
        set_error("PROXQ epilogue needs out, other, prev, colvec (and step unless colvec is the threshold)");
        return DECOMP_ERR_INVALID;
      }
      CUtensorMap tc, tp;
      rc = make_tensor_map(&tc, epi->other, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldother, 16, C::BM);
      if (rc != DECOMP_OK) return rc;
      rc = make_tensor_map(&tp, epi->prev, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldprev, 16, C::BM);
      if (rc != DECOMP_OK) return rc;
      switch (epi->shrink) {
        case DECOMP_SHRINK_REAL:
          return launch_proxq<CfgProxq, EPI_PROX_REAL>(ta, tb, tc, tp, gs, *epi, skip_if, st);
        case DECOMP_SHRINK_COMPLEX:
          return launch_proxq<CfgProxq, EPI_PROX_COMPLEX>(ta, tb, tc, tp, gs, *epi, skip_if, st);
        case DECOMP_SHRINK_POSITIVE:
          return launch_proxq<CfgProxq, EPI_PROX_POSITIVE>(ta, tb, tc, tp, gs, *epi, skip_if, st);
        default:
          set_error("decomp_gemm_nt_f64: unknown shrink kind %d", epi->shrink);
          return DECOMP_ERR_INVALID;
      }
    }
    default:
      set_error("decomp_gemm_nt_f64: unknown epilogue kind %d", epi->kind);
      return DECOMP_ERR_INVALID;
  }
}

int decomp_lasso_resident_supported(int64_t N) { return N == 32 || N == 64 || N == 128 || N == 256; }

int decomp_lasso_resident_f64(const double* Q, int64_t ldq, int64_t M, int64_t N, const decomp_epilogue_t* epi,
                              int32_t iters, const double* momentum, const int32_t* skip_if, void* stream) {
  if (epi == nullptr || Q == nullptr || momentum == nullptr || M < 0 || iters < 1 ||
      iters > DECOMP_LASSO_RESIDENT_MAX_ITERS || !decomp_lasso_resident_supported(N)) {
    set_error("decomp_lasso_resident_f64: invalid argument (N must be 32/64/128/256, 1 <= iters <= %d)",
              DECOMP_LASSO_RESIDENT_MAX_ITERS);
    return DECOMP_ERR_INVALID;
  }
  if (M == 0) return DECOMP_OK;
  if (epi->x == nullptr || epi->other == nullptr || epi->out == nullptr || epi->colvec == nullptr ||
      !(epi->flags & DECOMP_EPI_FLAG_COLVEC_IS_THRESHOLD) ||
      (epi->check && (epi->latch == nullptr || epi->scratch == nullptr || epi->colvec2 == nullptr))) {
    set_error("decomp_lasso_resident_f64: needs x (w), other (c), out (x), colvec as threshold; with check also "
              "colvec2, latch and scratch");
    return DECOMP_ERR_INVALID;
  }
  if (((epi->ldo | epi->ldother) & 1) != 0 || ((reinterpret_cast<uintptr_t>(epi->out) |
                                                reinterpret_cast<uintptr_t>(epi->other)) & 15u) != 0) {
    set_error("decomp_lasso_resident_f64: x and c must be 16-byte aligned with even leading dimensions");
    return DECOMP_ERR_INVALID;
  }
  const int bm = (int)(4096 / N);   // rows of one warp group's half of the row block (one TMA box)
  CUtensorMap tw, tq;
  int rc = make_tensor_map(&tw, epi->x, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldx, BK, (uint32_t)bm);
  if (rc != DECOMP_OK) return rc;
  rc = make_tensor_map(&tq, Q, (uint64_t)N, (uint64_t)N, (uint64_t)ldq, BK, (uint32_t)N);
  if (rc != DECOMP_OK) return rc;
  ResidentArgs a;
  a.M = M;
  a.N = (int)N;
  a.iters = iters;
  a.check = epi->check;
  a.zero = 0;
  static const int knob_skew = [] {
    // k-blocks group 1 trails group 0; measured with the 5-stage ring (0 .. 4): 0.832 / 0.840 / 0.840 / 0.849 / 0.838
    const char* e = getenv("DECOMP_RESIDENT_SKEW");
    return e != nullptr ? atoi(e) : 3;
  }();
  static const int knob_prefetch = [] {
    // off by default: the L2 prefetch of the next row block made the launch read 842 MB instead of 640 MB from DRAM
    // (ncu; 614 MB algorithmic) for no measurable gain in time (8.46 vs 8.45 ms)
    const char* e = getenv("DECOMP_RESIDENT_PREFETCH");
    return e != nullptr ? atoi(e) : 0;
  }();
  a.skew = knob_skew < 0 ? 0 : knob_skew;
  a.prefetch = knob_prefetch;
  a.c = epi->other;
  a.ldc = epi->ldother;
  a.x = epi->out;
  a.ldx = epi->ldo;
  a.w = const_cast<double*>(epi->x);
  a.ldw = epi->ldx;
  a.thr = epi->colvec;
  a.tol = epi->colvec2;
  a.latch = epi->latch;
  a.scratch = epi->scratch;
  a.latch_value = epi->latch_value;
  for (int i = 0; i < RES_MAX_ITERS; ++i) a.momentum[i] = i < iters ? momentum[i] : 0.0;
  cudaStream_t st = as_stream(stream);
  switch (epi->shrink) {
    case DECOMP_SHRINK_REAL:
      return launch_resident<DECOMP_SHRINK_REAL>(tw, tq, a, skip_if, st);
    case DECOMP_SHRINK_COMPLEX:
      return launch_resident<DECOMP_SHRINK_COMPLEX>(tw, tq, a, skip_if, st);
    case DECOMP_SHRINK_POSITIVE:
      return launch_resident<DECOMP_SHRINK_POSITIVE>(tw, tq, a, skip_if, st);
    default:
      set_error("decomp_lasso_resident_f64: unknown shrink kind %d", epi->shrink);
      return DECOMP_ERR_INVALID;
  }
}

int decomp_gemm_b2b_masked_supported(int64_t K1) { return K1 == 32 || K1 == 64 || K1 == 128; }

int decomp_gemm_b2b_masked_f64(const double* W, int64_t ldw, const double* R, int64_t ldr, int64_t M, int64_t K1,
                               int64_t F, const decomp_epilogue_t* epi, const int32_t* skip_if, void* stream) {
  if (epi == nullptr || W == nullptr || R == nullptr || M < 0 || F <= 0 || !decomp_gemm_b2b_masked_supported(K1) ||
      F > 2147483647LL) {
    set_error("decomp_gemm_b2b_masked_f64: invalid argument (K1 must be 32, 64 or 128)");
    return DECOMP_ERR_INVALID;
  }
  if (M == 0) return DECOMP_OK;
  if (epi->mask == nullptr || epi->out == nullptr || (epi->cwidth != 1 && epi->cwidth != 2) ||
      (epi->cwidth == 1 && ((epi->ldmask & 1) != 0 || (reinterpret_cast<uintptr_t>(epi->mask) & 15u) != 0))) {
    set_error("decomp_gemm_b2b_masked_f64: needs out and a mask (real data: 16-byte aligned, even ldmask)");
    return DECOMP_ERR_INVALID;
  }
  if (epi->kind == DECOMP_EPI_PROX &&
      (epi->x == nullptr || epi->other == nullptr || epi->prev == nullptr || epi->colvec == nullptr ||
       epi->step == nullptr || (epi->check && (epi->latch == nullptr || epi->scratch == nullptr)))) {
    set_error("decomp_gemm_b2b_masked_f64: PROX epilogue needs x, other, prev, colvec, step (and latch, scratch with check)");
    return DECOMP_ERR_INVALID;
  }
  CUtensorMap tw, tr;
  int rc = make_tensor_map(&tw, W, (uint64_t)K1, (uint64_t)M, (uint64_t)ldw, BK, B2B_BM);
  if (rc != DECOMP_OK) return rc;
  rc = make_tensor_map(&tr, R, (uint64_t)K1, (uint64_t)F, (uint64_t)ldr, BK, B2B_FC);
  if (rc != DECOMP_OK) return rc;
  B2bArgs a;
  a.M = M;
  a.K1 = (int)K1;
  a.F = (int)F;
  a.zero = 0;
  cudaStream_t st = as_stream(stream);
  switch (K1) {
    case 32:
      return dispatch_b2b<2>(tw, tr, a, *epi, skip_if, st);
    case 64:
      return dispatch_b2b<4>(tw, tr, a, *epi, skip_if, st);
    default:
      return dispatch_b2b<8>(tw, tr, a, *epi, skip_if, st);
  }
}

size_t decomp_gemm_tn_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  GemmGeom gs;
  tn_plan(M, N, K, &gs);
  return (size_t)gs.splits * (size_t)M * (size_t)gs.ld_partial * sizeof(double);
}

int decomp_gemm_tn_f64(const double* A, int64_t lda, const double* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                       double* out, int64_t ldo, int32_t combine, double beta, void* workspace,
                       size_t workspace_bytes, const int32_t* skip_if, void* stream) {
  using C = CfgMain;
  if (M < 0 || N < 0 || K <= 0 || combine < 0 || combine > 3) {
    set_error("decomp_gemm_tn_f64: invalid argument");
    return DECOMP_ERR_INVALID;
  }
  if (M == 0 || N == 0) return DECOMP_OK;
  if (combine >= 2 && ((M & 1) || (N & 1))) {
    set_error("decomp_gemm_tn_f64: complex combine needs even M and N");
    return DECOMP_ERR_INVALID;
  }
  GemmGeom gs;
  tn_plan(M, N, K, &gs);
  const size_t need = (size_t)gs.splits * (size_t)M * (size_t)gs.ld_partial * sizeof(double);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("decomp_gemm_tn_f64: workspace too small (%zu < %zu)", workspace_bytes, need);
    return DECOMP_ERR_INVALID;
  }
  CUtensorMap ta, tb;
  int rc;
  // one 3-D TMA instruction per operand and k-block when both widths are multiples of 16, else 16x16 boxes
  gs.tn3d = 0;
  gs.zero = 0;
  if ((M % 16) == 0 && (N % 16) == 0 &&
      make_tensor_map_tn3d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, C::BM / 16) == DECOMP_OK &&
      make_tensor_map_tn3d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, C::BN / 16) == DECOMP_OK) {
    gs.tn3d = 1;
  } else {
    rc = make_tensor_map(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 16, BK);
    if (rc != DECOMP_OK) return rc;
    rc = make_tensor_map(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 16, BK);
    if (rc != DECOMP_OK) return rc;
  }
  cudaStream_t st = as_stream(stream);
  decomp_epilogue_t ep;
  memset(&ep, 0, sizeof(ep));
  double* partial = reinterpret_cast<double*>(workspace);
  rc = launch<C, true, EPI_PARTIAL>(ta, tb, gs, ep, partial, skip_if, st);
  if (rc != DECOMP_OK) return rc;
  const long long total = (combine >= 2) ? (M / 2) * (N / 2) : M * N;
  int blocks = (int)((total + 255) / 256);
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  switch (combine) {
    case 0:
      reduce_partials_kernel<0><<<blocks, 256, 0, st>>>(partial, gs.splits, M, N, gs.ld_partial, out, ldo, beta, skip_if);
      break;
    case 1:
      reduce_partials_kernel<1><<<blocks, 256, 0, st>>>(partial, gs.splits, M, N, gs.ld_partial, out, ldo, beta, skip_if);
      break;
    case 2:
      reduce_partials_kernel<2><<<blocks, 256, 0, st>>>(partial, gs.splits, M, N, gs.ld_partial, out, ldo, beta, skip_if);
      break;
    default:
      reduce_partials_kernel<3><<<blocks, 256, 0, st>>>(partial, gs.splits, M, N, gs.ld_partial, out, ldo, beta, skip_if);
      break;
  }
  return check_cuda(cudaGetLastError(), "reduce_partials launch");
}

}  // extern "C"

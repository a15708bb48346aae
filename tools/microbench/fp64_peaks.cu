// FP64 peak probes for B200 (sm_100a): DMMA.8x8x4 issue rate, plain DFMA rate, and a
// cuBLAS DGEMM used ONLY to set the roofline denominator (never on the product path).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu -lcublas
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_rate(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC][2];
#pragma unroll
  for (int j = 0; j < NACC; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_rate(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) c[j] = (double)j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) c[j] = fma(c[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < NACC; ++j) s += c[j];
  if (s == 123.456) out[threadIdx.x] = s;
}

static float time_ms(void (*launch)(void*), void* ctx, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(ctx); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(ctx); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

struct RateCtx { double* out; double* in; int iters; int blocks; int threads; int which; };
static void launch_rate(void* p) {
  RateCtx* c = (RateCtx*)p;
  if (c->which == 0) dmma_rate<8><<<c->blocks, c->threads>>>(c->out, c->in, c->iters);
  else if (c->which == 1) dmma_rate<2><<<c->blocks, c->threads>>>(c->out, c->in, c->iters);
  else dfma_rate<16><<<c->blocks, c->threads>>>(c->out, c->in, c->iters);
}

struct GemmCtx { cublasHandle_t h; double *A, *B, *C; int m, n, k; cublasOperation_t ta, tb; int lda, ldb, ldc; };
static void launch_gemm(void* p) {
  GemmCtx* g = (GemmCtx*)p; double one = 1.0, zero = 0.0;
  cublasDgemm(g->h, g->ta, g->tb, g->m, g->n, g->k, &one, g->A, g->lda, g->B, g->ldb, &zero, g->C, g->ldc);
}

int main(int argc, char** argv) {
  const char* outpath = argc > 1 ? argv[1] : "fp64_peaks.json";
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double *in, *out; CK(cudaMalloc(&in, 64 * 8)); CK(cudaMalloc(&out, 1024 * 8));
  std::vector<double> h(64, 1.0e-3); CK(cudaMemcpy(in, h.data(), 64 * 8, cudaMemcpyHostToDevice));
  FILE* f = fopen(outpath, "w");
  fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
  printf("gpu %s sms %d\n", prop.name, sms);

  // DMMA rate: warps/SM sweep
  int wps[] = {4, 8, 16, 32};
  for (int which = 0; which < 3; ++which) {
    for (int wi = 0; wi < 4; ++wi) {
      int warps = wps[wi];
      RateCtx c{out, in, 4096, sms, warps * 32, which};
      float ms = time_ms(launch_rate, &c, 5);
      double flops;
      const char* name;
      if (which == 0) { flops = 2.0 * 256 * 8 * (double)c.iters * warps * sms; name = "dmma884_acc8"; }
      else if (which == 1) { flops = 2.0 * 256 * 2 * (double)c.iters * warps * sms; name = "dmma884_acc2"; }
      else { flops = 2.0 * 32 * 16 * (double)c.iters * warps * sms; name = "dfma_acc16"; }
      double tf = flops / (ms * 1e-3) / 1e12;
      printf("%s warps/SM=%d: %.3f ms  %.2f TFLOP/s\n", name, warps, ms, tf);
      fprintf(f, ", \"%s_w%d_tflops\": %.3f", name, warps, tf);
    }
  }

  // cuBLAS DGEMM: peak denominator only
  cublasHandle_t hnd; cublasCreate(&hnd);
  struct Shape { const char* name; int m, n, k; cublasOperation_t ta, tb; } shapes[] = {
    {"dgemm_8192_nn", 8192, 8192, 8192, CUBLAS_OP_N, CUBLAS_OP_N},
    {"dgemm_8192_tn", 8192, 8192, 8192, CUBLAS_OP_T, CUBLAS_OP_N},
    // row-major Y[n,f]·Phi[k,f]^T -> [n,k]  == col-major (k x n) = Phi_cm^T(k x f) * Y_cm(f x n): TN, m=k, n=rows, k=f
    {"dgemm_nmf_ydt_256x131072x4096_tn", 256, 131072, 4096, CUBLAS_OP_T, CUBLAS_OP_N},
    // row-major W[B,k]·G[k,k] -> [B,k] == col-major (k x B) = G_cm(k x k) * W_cm(k x B): NN
    {"dgemm_fista_256x100000x256_nn", 256, 100000, 256, CUBLAS_OP_N, CUBLAS_OP_N},
    // row-major C^T[k,n]·Y[n,f] -> [k,f] == col-major (f x k) = Y_cm(f x n) * C_cm^T(n x k): NT, m=f, n=k, k=rows
    {"dgemm_nmf_cty_4096x256x131072_nt", 4096, 256, 131072, CUBLAS_OP_N, CUBLAS_OP_T},
  };
  for (auto& s : shapes) {
    size_t ea = (size_t)s.m * s.k, eb = (size_t)s.k * s.n, ec = (size_t)s.m * s.n;
    double *A, *B, *C; CK(cudaMalloc(&A, ea * 8)); CK(cudaMalloc(&B, eb * 8)); CK(cudaMalloc(&C, ec * 8));
    CK(cudaMemset(A, 0, ea * 8)); CK(cudaMemset(B, 0, eb * 8));
    GemmCtx g{hnd, A, B, C, s.m, s.n, s.k, s.ta, s.tb,
              s.ta == CUBLAS_OP_N ? s.m : s.k, s.tb == CUBLAS_OP_N ? s.k : s.n, s.m};
    float ms = time_ms(launch_gemm, &g, 5);
    double tf = 2.0 * s.m * (double)s.n * s.k / (ms * 1e-3) / 1e12;
    printf("%s: %.3f ms %.2f TFLOP/s\n", s.name, ms, tf);
    fprintf(f, ", \"%s_tflops\": %.3f", s.name, tf);
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  // sustained: 8192^3 back-to-back ~3 s
  {
    int n = 8192; double *A, *B, *C; CK(cudaMalloc(&A, (size_t)n * n * 8)); CK(cudaMalloc(&B, (size_t)n * n * 8)); CK(cudaMalloc(&C, (size_t)n * n * 8));
    CK(cudaMemset(A, 0, (size_t)n * n * 8)); CK(cudaMemset(B, 0, (size_t)n * n * 8));
    GemmCtx g{hnd, A, B, C, n, n, n, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int reps = 100; CK(cudaEventRecord(e0)); for (int i = 0; i < reps; ++i) launch_gemm(&g); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * n * (double)n * n * reps / (ms * 1e-3) / 1e12;
    printf("dgemm_8192 sustained x%d: %.1f ms total %.2f TFLOP/s\n", reps, ms, tf);
    fprintf(f, ", \"dgemm_8192_sustained_tflops\": %.3f", tf);
  }
  fprintf(f, "}\n"); fclose(f);
  return 0;
}

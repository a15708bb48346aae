// How many warps per SM sub-partition does the FP64 tensor pipe need?  DMMA.8x8x4 rate with 1..4 warps per scheduler,
// 8 independent accumulators per warp, operands in registers (no memory traffic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_warps dmma_warps.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int NACC>
__global__ void dmma_rate(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[NACC][2];
#pragma unroll
  for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < NACC; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
  if (s == 123.456) out[threadIdx.x] = s;
}

template <int NACC>
static double run(int sms, int warps, double* out, double* in) {
  const int iters = 8192;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    dmma_rate<NACC><<<sms, warps * 32>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return 2.0 * 256.0 * NACC * iters * warps * sms / (best * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *in, *out;
  cudaMalloc(&in, 64 * 8);
  cudaMalloc(&out, 1024 * 8);
  cudaMemset(in, 0, 64 * 8);
  printf("{\"gpu\": \"%s\", \"rows\": [\n", prop.name);
  const int warps[] = {4, 8, 12, 16, 32};
  for (int w = 0; w < 5; ++w) {
    printf("  {\"warps_per_sm\": %d, \"tflops_8acc\": %.2f, \"tflops_4acc\": %.2f, \"tflops_2acc\": %.2f}%s\n", warps[w],
           run<8>(sms, warps[w], out, in), run<4>(sms, warps[w], out, in), run<2>(sms, warps[w], out, in),
           w == 4 ? "" : ",");
  }
  printf("]}\n");
  return 0;
}

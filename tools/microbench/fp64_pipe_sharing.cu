// How expensive are scalar FP64 instructions while DMMA.8x8x4 is in flight on B200 (sm_100a)?
// Per SM: 8 "MMA" warps issue DMMA chains (2 per sub-partition = what saturates the pipe).
//   mode 0: DMMA only
//   mode 1: 4 extra warps issue independent DFMAs concurrently, `nfma` per thread per `period` DMMA rounds
//   mode 2: the MMA warps themselves run the same number of DFMAs in lockstep (bar.sync, then all do FP64 math)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_sharing fp64_pipe_sharing.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// rounds: outer iterations; per round every MMA warp issues 64*8 DMMAs (= one 32x32x(16*8) warp tile step)
__global__ void __launch_bounds__(384) kern(double* out, const double* in, int rounds, int mode, int nfma) {
  const int warp = threadIdx.x >> 5;
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  if (warp < 8) {
    double c[16][2];
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j][0] = c[j][1] = 0.0;
    double f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = a + j;
    for (int r = 0; r < rounds; ++r) {
      for (int k = 0; k < 32; ++k) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dmma(c[j][0], c[j][1], a, b);
      }
      if (mode == 2) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // 4 warps x nfma spread over 8 warps
        for (int i = 0; i < nfma / 16; ++i) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b);
        }
      }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j][0] + c[j][1];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
    if (s == 123.456) out[threadIdx.x] = s;
  } else if (mode >= 3) {
    // 3: sleep only; 4: integer work; 5: DFMA without pacing; 6: one DFMA burst of nfma then a long sleep
    unsigned x = threadIdx.x;
    double f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = a + j;
    for (int r = 0; r < rounds; ++r) {
      if (mode == 3) __nanosleep(2000);
      if (mode == 4) { for (int i = 0; i < nfma; ++i) x = x * 1664525u + 1013904223u; __nanosleep(2000); }
      if (mode == 5) { for (int i = 0; i < nfma / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b); } }
      if (mode == 6) { for (int i = 0; i < nfma / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b); } __nanosleep(8000); }
    }
    double s = x;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
    if (s == 123.456) out[threadIdx.x] = s;
  } else if (mode == 1) {
    double f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = a + j;
    for (int r = 0; r < rounds; ++r) {
      for (int i = 0; i < nfma / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b);
      }
      // pace the epilogue warps roughly with the MMA warps: ~512 DMMA * 16 cycles / 2 warps per SMSP
      __nanosleep(2000);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[j];
    if (s == 123.456) out[threadIdx.x] = s;
  }
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  double *in, *out; cudaMalloc(&in, 64 * 8); cudaMalloc(&out, 1024 * 8);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1e-3; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int rounds = 400;
  printf("per round and SM: 8 warps x 512 DMMA; DMMA-only ideal = 512*8/4*16 = 16384 cycles per round\n");
  for (int mode = 0; mode < 7; ++mode) {
    int nf[] = {0, 64, 128, 256, 512, 1024};
    for (int ni = 0; ni < (mode == 0 ? 1 : 6); ++ni) {
      int nfma = nf[ni];
      if (mode != 0 && nfma == 0) continue;
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<<<sms, 384>>>(out, in, rounds, mode, nfma);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double tf = 2.0 * 256 * 512 * 8 * (double)rounds * sms / (best * 1e-3) / 1e12;
      // scalar FP64 warp-instructions per round per SMSP: mode 1: 1 warp/SMSP x nfma; mode 2: 2 warps x nfma/2... same total
      printf("mode %d nfma/thread/round %4d: %.3f ms  DMMA %.2f TFLOP/s  (%.1f cycles per round)\n", mode, nfma, best, tf,
             best * 1e-3 * prop.clockRate * 1e3 / rounds);
    }
  }
  return 0;
}

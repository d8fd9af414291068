// Box probe: FP64 issue-rate microbenchmarks on sm_100a.
//   dmma : mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), register-resident operands
//   dfma : plain DFMA chains
//   mixed: both interleaved (do the pipes overlap?)
//   popc : LOP3+POPC integer rate (mask-count kernel ceiling)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

// per iteration: NACC DMMA + NF DFMA
template <int NACC, int NF>
__global__ void k_mixed(double *out, int iters, double a0, double b0) {
  double c[NACC][2], f[NF];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = i;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      dmma(c[i][0], c[i][1], a, b);
      if (i < NF) f[i] = fma(f[i], a0, b0);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void k_popc(unsigned *out, int iters, unsigned a0, unsigned b0) {
  unsigned acc[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = 0; x[i] = a0 * (threadIdx.x + i + 1); }
  unsigned y = b0 + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] += __popc(x[i] & y); }
    y = y * 1664525u + 1013904223u;
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
  if (s == 0x12345678u) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, clk);
  double *d; CK(cudaMalloc(&d, 1024));
  int nsm = p.multiProcessorCount;
  const int iters = 4096;
  for (int wps = 4; wps <= 16; wps *= 2) {      // warps per SM
    int threads = wps * 32;
    {
      float ms = time_ms([&] { k_dmma<8><<<nsm, threads>>>(d, iters, 1.0, 2.0); });
      double flops = (double)nsm * wps * iters * 8 * 512.0;
      printf("dmma  NACC=8  warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM @%d MHz nominal)\n", wps, ms,
             flops / ms * 1e-9, flops / 2 / (ms * 1e-3) / nsm / (clk * 1e3), clk / 1000);
    }
    {
      float ms = time_ms([&] { k_dmma<16><<<nsm, threads>>>(d, iters, 1.0, 2.0); });
      double flops = (double)nsm * wps * iters * 16 * 512.0;
      printf("dmma  NACC=16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { k_dmma<1><<<nsm, threads>>>(d, iters, 1.0, 2.0); });
      double flops = (double)nsm * wps * iters * 1 * 512.0;
      printf("dmma  NACC=1  warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s  (latency chain: %.1f ns per DMMA)\n", wps, ms, flops / ms * 1e-9,
             ms * 1e6 / iters);
    }
    {
      float ms = time_ms([&] { k_dfma<16><<<nsm, threads>>>(d, iters, 1.0000001, 1e-9); });
      double flops = (double)nsm * threads * iters * 16 * 2.0;
      printf("dfma  NACC=16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { k_mixed<8, 8><<<nsm, threads>>>(d, iters, 1.0000001, 1e-9); });
      double fl_m = (double)nsm * wps * iters * 8 * 512.0, fl_f = (double)nsm * threads * iters * 8 * 2.0;
      printf("mixed 8dmma+8dfma warps/SM=%2d : %8.3f ms  dmma %7.2f + dfma %7.2f TFLOP/s\n", wps, ms, fl_m / ms * 1e-9, fl_f / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { k_mixed<8, 2><<<nsm, threads>>>(d, iters, 1.0000001, 1e-9); });
      double fl_m = (double)nsm * wps * iters * 8 * 512.0, fl_f = (double)nsm * threads * iters * 2 * 2.0;
      printf("mixed 8dmma+2dfma warps/SM=%2d : %8.3f ms  dmma %7.2f + dfma %7.2f TFLOP/s\n", wps, ms, fl_m / ms * 1e-9, fl_f / ms * 1e-9);
    }
  }
  {
    int threads = 1024;
    float ms = time_ms([&] { k_popc<<<nsm * 2, threads>>>((unsigned *)d, iters, 3u, 5u); });
    double ops = (double)nsm * 2 * threads * iters * 8;
    printf("popc(and) : %8.3f ms  %7.2f Tpopc32/s (%.1f /clk/SM nominal)\n", ms, ops / ms * 1e-9, ops / (ms * 1e-3) / nsm / (clk * 1e3));
  }
  // sustained DMMA for ~2 s (power-capped clock)
  {
    int threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int n = 0;
    for (; n < 40; n++) k_dmma<8><<<nsm, threads>>>(d, iters * 16, 1.0, 2.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)n * nsm * 8 * (iters * 16.0) * 8 * 512.0;
    printf("dmma sustained: %.1f ms total, %7.2f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  return 0;
}

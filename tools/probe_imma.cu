// IMMA issue-rate probe on sm_100a: legacy mma.sync int8 (IMMA.16832.S8.S8) -- candidate for the called-genotype integer path.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC>
__global__ void k_imma(int *out, int iters) {
  int c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x + 1u, 5u};
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) imma(c[i], a, b);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 0x12345678) out[0] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int *d; cudaMalloc(&d, 64);
  int nsm = p.multiProcessorCount; const int iters = 8192;
  for (int wps = 4; wps <= 32; wps *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
      cudaEventRecord(e0); k_imma<8><<<nsm, wps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double macs = (double) nsm * wps * iters * 8 * (16.0 * 8 * 32);
    printf("imma.m16n8k32.s8 warps/SM=%2d: %.3f ms  %.1f TMAC/s  (%.0f MAC/clk/SM at %d MHz)\n", wps, best, macs / best * 1e-9,
           macs / (best * 1e-3) / nsm / (p.clockRate * 1e3), p.clockRate / 1000);
  }
  return 0;
}

#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(int* out, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 7, b1 = a0 * 13;
  int c[8][4] = {};
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  int s = 0;
  for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000, threads = 32 * warps;
    k<<<148, threads>>>(d, 100);
    cudaEventRecord(e0); k<<<148, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mma = 148.0 * warps * iters * 8;
    printf("warps/SM %d: %.1f T int8 MAC/s, %.3f mma/clk/SM (at 1.965 GHz)\n", warps, mma * 4096 / ms / 1e9, mma / 148 / (ms * 1e-3 * 1.965e9));
  }
  return 0;
}

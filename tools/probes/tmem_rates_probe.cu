// tmem_rates_probe.cu -- per-instruction cost of tcgen05.ld / tcgen05.st by shape and width on one B200 SM (12 warps, all SMs),
// to cost tensor-memory data movement against the shared-memory pipe (see tmem_xchg_probe.cu).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o variants/tmem_rates_probe tools/probes/tmem_rates_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define S8(r, o) "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
__device__ __forceinline__ void ld32x16(uint32_t* r, uint32_t ta) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : R8(r, 0), R8(r, 8) : "r"(ta));
}
__device__ __forceinline__ void st32x16(uint32_t ta, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(ta), S8(r, 0), S8(r, 8) : "memory");
}
__device__ __forceinline__ void ld256x2(uint32_t* r, uint32_t ta) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : R8(r, 0) : "r"(ta));
}
__device__ __forceinline__ void st256x2(uint32_t ta, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), S8(r, 0) : "memory");
}
__device__ __forceinline__ void ld256x4(uint32_t* r, uint32_t ta) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : R8(r, 0), R8(r, 8) : "r"(ta));
}
__device__ __forceinline__ void st256x4(uint32_t ta, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(ta), S8(r, 0), S8(r, 8) : "memory");
}
__device__ __forceinline__ void ld128x8(uint32_t* r, uint32_t ta) {  // 16x128b.x8: 16 registers
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : R8(r, 0), R8(r, 8) : "r"(ta));
}
__device__ __forceinline__ void ld64x16(uint32_t* r, uint32_t ta) {  // 16x64b.x16: 16 registers
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : R8(r, 0), R8(r, 8) : "r"(ta));
}
// MODE: 0 ld32x16 x4 | 1 st32x16 x4 | 2 ld256x2 x8 | 3 st256x2 x8 | 4 ld256x4 x4 | 5 st256x4 x4 | 6 ld128x8 x4 | 7 ld64x16 x4
//       8 st32x16 x4, wait, ld256x4 x4, wait  | 9 st256x4 x4, wait, ld32x16 x4, wait | 10: 8 with one warp only (latency)
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(double* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t ta = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; i++) r[i] = threadIdx.x * 64 + i;
  st32x16(ta, r); st32x16(ta + 16, r + 16); st32x16(ta + 32, r + 32); st32x16(ta + 48, r + 48);
  WAIT_ST();
  __syncthreads();
  const bool active = MODE != 10 || warp == 0;
  long long t0 = clock64();
  if (active)
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) { ld32x16(r, ta); ld32x16(r + 16, ta + 16); ld32x16(r + 32, ta + 32); ld32x16(r + 48, ta + 48); WAIT_LD(); }
    if (MODE == 1) { st32x16(ta, r); st32x16(ta + 16, r + 16); st32x16(ta + 32, r + 32); st32x16(ta + 48, r + 48); WAIT_ST(); r[it & 63] ^= it; }
    if (MODE == 2) {
#pragma unroll
      for (int g = 0; g < 4; g++) { ld256x2(r + 16 * g, ta + 16 * g); ld256x2(r + 16 * g + 8, ta + 16 * g + (16u << 16)); }
      WAIT_LD();
    }
    if (MODE == 3) {
#pragma unroll
      for (int g = 0; g < 4; g++) { st256x2(ta + 16 * g, r + 16 * g); st256x2(ta + 16 * g + (16u << 16), r + 16 * g + 8); }
      WAIT_ST(); r[it & 63] ^= it;
    }
    if (MODE == 4) { ld256x4(r, ta); ld256x4(r + 16, ta + (16u << 16)); ld256x4(r + 32, ta + 32); ld256x4(r + 48, ta + 32 + (16u << 16)); WAIT_LD(); }
    if (MODE == 5) { st256x4(ta, r); st256x4(ta + (16u << 16), r + 16); st256x4(ta + 32, r + 32); st256x4(ta + 32 + (16u << 16), r + 48); WAIT_ST(); r[it & 63] ^= it; }
    if (MODE == 6) { ld128x8(r, ta); ld128x8(r + 16, ta + (16u << 16)); ld128x8(r + 32, ta + 32); ld128x8(r + 48, ta + 32 + (16u << 16)); WAIT_LD(); }
    if (MODE == 7) { ld64x16(r, ta); ld64x16(r + 16, ta + (16u << 16)); ld64x16(r + 32, ta + 32); ld64x16(r + 48, ta + 32 + (16u << 16)); WAIT_LD(); }
    if (MODE == 8 || MODE == 10) {
      st32x16(ta, r); st32x16(ta + 16, r + 16); st32x16(ta + 32, r + 32); st32x16(ta + 48, r + 48); WAIT_ST();
      ld256x4(r, ta); ld256x4(r + 16, ta + (16u << 16)); ld256x4(r + 32, ta + 32); ld256x4(r + 48, ta + 32 + (16u << 16)); WAIT_LD();
    }
    if (MODE == 9) {
      st256x4(ta, r); st256x4(ta + (16u << 16), r + 16); st256x4(ta + 32, r + 32); st256x4(ta + 32 + (16u << 16), r + 48); WAIT_ST();
      ld32x16(r, ta); ld32x16(r + 16, ta + 16); ld32x16(r + 32, ta + 32); ld32x16(r + 48, ta + 48); WAIT_LD();
    }
  }
  long long t1 = clock64();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 64; i++) x ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
template <int MODE>
void run(const char* name, int iters, double kb_per_warp_iter) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, 384>>>(out, cyc, iters);
  k<MODE><<<148, 384>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  const int warps = MODE == 10 ? 1 : 12;
  printf("%-52s cycles/iter %8.1f  -> %7.1f B/clk/SM  (%s)\n", name, avg / iters, warps * kb_per_warp_iter * 1024 / (avg / iters), cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  const int it = 20000;
  run<0>("ld 32x32b.x16 x4 (8 KB/warp)", it, 8);
  run<1>("st 32x32b.x16 x4", it, 8);
  run<2>("ld 16x256b.x2 x8", it, 8);
  run<3>("st 16x256b.x2 x8", it, 8);
  run<4>("ld 16x256b.x4 x4", it, 8);
  run<5>("st 16x256b.x4 x4", it, 8);
  run<6>("ld 16x128b.x8 x4", it, 8);
  run<7>("ld 16x64b.x16 x4", it, 8);
  run<8>("round: st 32x32b, wait, ld 16x256b.x4, wait", it, 16);
  run<9>("round: st 16x256b.x4, wait, ld 32x32b, wait", it, 16);
  run<10>("same as the first round, ONE warp (latency)", it, 16);
  return 0;
}

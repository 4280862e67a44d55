// tmem_xchg_probe.cu -- can tensor memory serve as the exchange medium of the register FFT?
// A thread reaches only its own TMEM lane with the 32x32b shape, but the 16x256b shape maps thread t of a warp to lanes
// t/4 and 8 + t/4 (64-bit units (t%4) + 4j of each 256-bit group), so "store 32x32b, load 16x256b" moves two lane-index
// bits into the register index and two column bits into the lane index: two such rounds transpose a 16 x 16 tile between
// the threads of a warp without touching the shared-memory pipe.  This probe (1) checks that mapping, in both directions
// (store 32x32b / load 16x256b and store 16x256b / load 32x32b), (2) measures what a warp-level 16 x 16 complex-f64
// transpose costs that way -- full size (64 columns per warp) and chunked (16 columns per warp) -- next to the
// shared-memory exchange of fft16.cuh, and (3) whether the two paths and FP64 issue overlap.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tmem_xchg_probe tools/probes/tmem_xchg_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD256X2(r, o, ta)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                           \
               : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),         \
                 "=r"(r[o + 6]), "=r"(r[o + 7])                                                                        \
               : "r"(ta))
#define ST256X2(ta, r, o)                                                                                              \
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(ta), "r"(r[o + 0]), \
               "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])  \
               : "memory")
#define LD32X16(r, o, ta)                                                                                              \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
      : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]),  \
        "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]), "=r"(r[o + 12]),               \
        "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])                                                              \
      : "r"(ta))
#define ST32X16(ta, r, o)                                                                                              \
  asm volatile(                                                                                                        \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" \
      ::"r"(ta), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), \
      "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]), "r"(r[o + 12]), "r"(r[o + 13]),       \
      "r"(r[o + 14]), "r"(r[o + 15])                                                                                   \
      : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

__device__ __forceinline__ uint32_t tmem_alloc512(uint32_t* slot) {
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
__device__ __forceinline__ void tmem_free512(uint32_t base) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

// ---- (1) mapping ----------------------------------------------------------------------------------------------------
// out[dir][warp][t][16]: dir 0 = st 32x32b (value = lane << 8 | col) then ld 16x256b.x2 at lane bases 0 and 16;
// dir 1 = st 16x256b.x2 at lane bases 0 / 16 (value = 0x10000 | base16 << 12 | t << 4 | reg) then ld 32x32b.x16
__global__ void __launch_bounds__(128, 1) map_kernel(uint32_t* out) {
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc512(&slot);
  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  const uint32_t ta = base + ((uint32_t)(warp * 32) << 16);
  uint32_t r[16], s[16];
#pragma unroll
  for (int c = 0; c < 16; c++) r[c] = ((uint32_t)t << 8) | c;
  ST32X16(ta, r, 0);
  WAIT_ST();
  LD256X2(s, 0, ta);
  LD256X2(s, 8, ta + (16u << 16));
  WAIT_LD();
#pragma unroll
  for (int c = 0; c < 16; c++) out[((0 * 4 + warp) * 32 + t) * 16 + c] = s[c];
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 16; c++) r[c] = 0x10000u | ((uint32_t)(c >> 3) << 12) | ((uint32_t)t << 4) | (c & 7);
  ST256X2(ta + 32, r, 0);
  ST256X2(ta + 32 + (16u << 16), r, 8);
  WAIT_ST();
  LD32X16(s, 0, ta + 32);
  WAIT_LD();
#pragma unroll
  for (int c = 0; c < 16; c++) out[((1 * 4 + warp) * 32 + t) * 16 + c] = s[c];
  tmem_free512(base);
}

// ---- (2)/(3) cost ---------------------------------------------------------------------------------------------------
struct alignas(16) C2 { double x, y; };
constexpr int kPad = 65;

// MODE 0: two full rounds (64 columns per warp): 4 st.x16, wait, 8 ld.16x256b.x2, wait -- twice
// MODE 1: two rounds in four 16-column chunks each: (st.x16, wait, 2 ld.x2, wait) x 4 -- twice
// MODE 2: the shared-memory exchange: 16 STS.128, bar, 16 LDS.128, bar
// MODE 3: MODE 0 and MODE 2 in the same iteration (do the pipes overlap?)
// MODE 4: MODE 0 + 64 DFMA
// MODE 5: MODE 1 + 64 DFMA
// MODE 6: MODE 1 + MODE 2
// MODE 7: 64 DFMA only
template <int MODE>
__global__ void __launch_bounds__(384, 1) cost_kernel(double* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc512(&slot);
  const int warp = threadIdx.x >> 5;
  const int team = threadIdx.x / 64, u = threadIdx.x % 64;
  const int k1 = u & 15, q = u >> 4;
  C2* buf = reinterpret_cast<C2*>(smem) + team * 16 * kPad;
  constexpr bool kChunk = MODE == 1 || MODE == 5 || MODE == 6;
  const uint32_t ta = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * (kChunk ? 16 : 64);
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 64; i++) r[i] = threadIdx.x * 64 + i;
  C2 v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = C2{1.0 + threadIdx.x * 1e-3 + i, 0.5 + i};
  double a0 = 1.0, a1 = 1.1, a2 = 1.2, a3 = 1.3, a4 = 1.4, a5 = 1.5, a6 = 1.6, a7 = 1.7;
  const double m = 1.0000000001, c = 1e-12;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0 || MODE == 3 || MODE == 4) {
#pragma unroll
      for (int round = 0; round < 2; round++) {
        ST32X16(ta, r, 0); ST32X16(ta + 16, r, 16); ST32X16(ta + 32, r, 32); ST32X16(ta + 48, r, 48);
        WAIT_ST();
#pragma unroll
        for (int g = 0; g < 4; g++) {
          LD256X2(r, 16 * g, ta + 16 * g);
          LD256X2(r, 16 * g + 8, ta + 16 * g + (16u << 16));
        }
        WAIT_LD();
      }
    }
    if (kChunk) {
#pragma unroll
      for (int round = 0; round < 2; round++) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
          ST32X16(ta, r, 16 * g);
          WAIT_ST();
          LD256X2(r, 16 * g, ta);
          LD256X2(r, 16 * g + 8, ta + (16u << 16));
          WAIT_LD();
        }
      }
    }
    if (MODE == 2 || MODE == 3 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < 16; i++) buf[i * kPad + u] = v[i];
      asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory");
#pragma unroll
      for (int i = 0; i < 16; i++) { C2 t = buf[k1 * kPad + q + 4 * i]; v[i].x += t.x; v[i].y += t.y; }
      asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory");
    }
    if (MODE == 4 || MODE == 5 || MODE == 7) {
#pragma unroll
      for (int rr = 0; rr < 8; rr++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
      }
    }
  }
  long long t1 = clock64();
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
#pragma unroll
  for (int i = 0; i < 16; i++) s += v[i].x + v[i].y;
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 64; i++) x ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  tmem_free512(base);
}

template <int MODE>
void run(const char* name, int iters) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&cyc, 148 * 8);
  const int smem = 6 * 16 * kPad * 16;
  cudaFuncSetAttribute(cost_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cost_kernel<MODE><<<148, 384, smem>>>(out, cyc, iters);
  cost_kernel<MODE><<<148, 384, smem>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  printf("%-58s cycles/iter (12 warps) %8.1f  (%s)\n", name, avg / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  {
    uint32_t* d; cudaMalloc(&d, 2 * 4 * 32 * 16 * 4);
    map_kernel<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    static uint32_t h[2][4][32][16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("map kernel: %s\n", cudaGetErrorString(e));
    int bad0 = 0, bad1 = 0;
    for (int w = 0; w < 4; w++)
      for (int t = 0; t < 32; t++) {
        for (int c = 0; c < 16; c++) {
          // dir 0 prediction: reg c of the load at lane base 16*(c>>3): within the x2 load, reg j = c & 7:
          // lane = base + t/4 + 8*((j>>1)&1), col = 2*(t%4) + (j&1) + 8*(j>>2)
          const int j = c & 7, lane = 16 * (c >> 3) + t / 4 + 8 * ((j >> 1) & 1), col = 2 * (t % 4) + (j & 1) + 8 * (j >> 2);
          if (h[0][w][t][c] != (uint32_t)((lane << 8) | col)) bad0++;
        }
        // dir 1 prediction: lane t, column cc holds what thread ts stored from register rr of the store at base b16:
        // lane = 16 b16 + ts/4 + 8*((rr>>1)&1), col = 2*(ts%4) + (rr&1) + 8*(rr>>2)
        for (int cc = 0; cc < 16; cc++) {
          const int b16 = t >> 4, l = t & 15, hi8 = l >> 3, ts4 = l & 7;
          const int cgrp = cc >> 3, cin = cc & 7, ts = ts4 * 4 + (cin >> 1), rr = (cin & 1) | (hi8 << 1) | (cgrp << 2);
          const uint32_t want = 0x10000u | ((uint32_t)b16 << 12) | ((uint32_t)ts << 4) | rr;
          if (h[1][w][t][cc] != want) bad1++;
        }
      }
    printf("st 32x32b -> ld 16x256b.x2 mapping: %d mismatches of 2048; st 16x256b.x2 -> ld 32x32b: %d mismatches of 2048\n", bad0, bad1);
    printf("warp 0, dir 0 (lane<<8|col), threads 0..7:\n");
    for (int t = 0; t < 8; t++) { for (int c = 0; c < 16; c++) printf(" %04x", h[0][0][t][c]); printf("\n"); }
    printf("warp 1, dir 1, threads 0..3 and 16..17:\n");
    for (int t : {0, 1, 2, 3, 16, 17}) { for (int c = 0; c < 16; c++) printf(" %05x", h[1][1][t][c]); printf("\n"); }
    cudaFree(d);
  }
  const int it = 20000;
  run<0>("TMEM transpose, 2 full rounds (64 col/warp)", it);
  run<1>("TMEM transpose, 2 rounds x 4 chunks (16 col/warp)", it);
  run<2>("smem exchange: 16 STS.128, bar, 16 LDS.128, bar", it);
  run<3>("full TMEM rounds + smem exchange", it);
  run<6>("chunked TMEM rounds + smem exchange", it);
  run<7>("64 DFMA", it);
  run<4>("full TMEM rounds + 64 DFMA", it);
  run<5>("chunked TMEM rounds + 64 DFMA", it);
  return 0;
}

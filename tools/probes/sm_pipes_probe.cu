// sm_pipes_probe.cu -- what one B200 SM sustains on the pipes pbs_kernel leans on: shared-memory wavefronts for 128-bit
// loads / stores / their mix (with the exchange-buffer access pattern of fft16.cuh), FP64 issue next to that traffic, and
// tcgen05.ld (tensor-memory reads).  One CTA of 384 threads per SM like pbs_kernel.  Prints cycles and per-clock rates.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/sm_pipes_probe tools/probes/sm_pipes_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct alignas(16) C2 { double x, y; };
constexpr int kPad = 65;

template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(double* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int team = threadIdx.x / 64, u = threadIdx.x % 64;
  C2* buf = reinterpret_cast<C2*>(smem) + team * 16 * kPad;
  C2 v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = C2{1.0 + threadIdx.x * 1e-3 + i, 0.5 + i};
#pragma unroll
  for (int i = 0; i < 16; i++) buf[i * kPad + u] = v[i];
  __syncthreads();
  const int k1 = u & 15, q = u >> 4;
  double a0 = v[0].x, a1 = v[1].x, a2 = v[2].x, a3 = v[3].x, a4 = v[4].x, a5 = v[5].x, a6 = v[6].x, a7 = v[7].x;
  const double m = 1.0000000001, c = 1e-12;
  const bool fp_warp = (threadIdx.x >> 5) & 1;  // MODE 6: odd warps do FP64, even warps do smem
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0 || MODE == 2 || MODE == 3 || (MODE == 6 && !fp_warp)) {  // column-major stores (fwd_x1_write)
#pragma unroll
      for (int i = 0; i < 16; i++) buf[i * kPad + u] = v[i];
    }
    if (MODE == 3) asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory");
    if (MODE == 1 || MODE == 2 || MODE == 3 || (MODE == 6 && !fp_warp)) {  // row gathers (fwd_x1_read)
#pragma unroll
      for (int i = 0; i < 16; i++) { C2 t = buf[k1 * kPad + q + 4 * i]; v[i].x += t.x; v[i].y += t.y; }
    }
    if (MODE == 3) asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory");
    if (MODE == 4 || MODE == 5 || (MODE == 6 && fp_warp)) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
      }
    }
    if (MODE == 5) {  // FP64 and smem traffic from the same warp, back to back
#pragma unroll
      for (int i = 0; i < 16; i++) buf[i * kPad + u] = v[i];
#pragma unroll
      for (int i = 0; i < 16; i++) { C2 t = buf[k1 * kPad + q + 4 * i]; v[i].x += t.x; v[i].y += t.y; }
    }
  }
  long long t1 = clock64();
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
#pragma unroll
  for (int i = 0; i < 16; i++) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(384, 1) tmem_probe(double* out, long long* cyc, int iters) {
  __shared__ uint32_t base;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  const uint32_t ta = base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[16];
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(ta + 16 * c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[15];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int MODE>
void run(const char* name, double wf_per_iter_thread_warp, double fp_per_iter, int iters) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&cyc, 148 * 8);
  const int smem = 6 * 16 * kPad * 16;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<MODE><<<148, 384, smem>>>(out, cyc, iters);
  probe<MODE><<<148, 384, smem>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  // wavefronts per iteration per SM = 12 warps * wf_per_iter_thread_warp
  printf("%-34s cycles/iter %8.1f  smem wavefronts/clk %.3f  fp64 warp-instr/clk/SMSP %.3f  (%s)\n", name, avg / iters,
         12 * wf_per_iter_thread_warp / (avg / iters), 3 * fp_per_iter / (avg / iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int it = 20000;
  run<0>("STS.128 only (16 per iter)", 64, 0, it);
  run<1>("LDS.128 only (16 per iter)", 64, 0, it);
  run<2>("16 STS.128 + 16 LDS.128", 128, 0, it);
  run<3>("same + 2 named barriers", 128, 0, it);
  run<4>("DFMA only (64 per iter)", 0, 64, it);
  run<5>("64 DFMA + 16 STS + 16 LDS per warp", 128, 64, it);
  run<6>("odd warps DFMA, even warps smem", 64, 32, it);
  {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 384 * 8); cudaMalloc(&cyc, 148 * 8);
    tmem_probe<<<148, 384>>>(out, cyc, it);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("tcgen05.ld 32x32b.x16 (4 per iter per warp, waited): cycles/iter %.1f -> %.1f B/clk/SM (%s)\n", avg / it,
           12 * 4 * 2048.0 / (avg / it), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

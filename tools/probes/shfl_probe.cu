// shfl_probe.cu -- does SHFL share the shared-memory data pipe?  384 threads per SM; modes: SHFL only, LDS/STS only, both.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct alignas(16) C2 { double x, y; };
constexpr int kPad = 65;
template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(unsigned* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int team = threadIdx.x / 64, u = threadIdx.x % 64;
  C2* buf = reinterpret_cast<C2*>(smem) + team * 16 * kPad;
  C2 v[16];
  unsigned r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { v[i] = C2{1.0 + threadIdx.x + i, 0.5 + i}; r[i] = threadIdx.x * 17 + i; buf[i * kPad + u] = v[i]; }
  __syncthreads();
  const int k1 = u & 15, q = u >> 4;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE & 1) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = __shfl_xor_sync(0xFFFFFFFFu, r[i], 1 + (i & 3)) + 1;
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 16; i++) buf[i * kPad + u] = v[i];
#pragma unroll
      for (int i = 0; i < 16; i++) { C2 t = buf[k1 * kPad + q + 4 * i]; v[i].x += t.x; v[i].y += t.y; }
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i] + (unsigned)v[i].x;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
double run(int iters) {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&cyc, 148 * 8);
  const int smem = 6 * 16 * kPad * 16;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<MODE><<<148, 384, smem>>>(out, cyc, iters);
  probe<MODE><<<148, 384, smem>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
  cudaFree(out); cudaFree(cyc);
  return avg / iters;
}
int main() {
  const int it = 20000;
  const double a = run<1>(it), b = run<2>(it), c = run<3>(it);
  printf("64 SHFL per thread-iter (12 warps): %.1f cycles/iter -> %.3f warp-SHFL/clk/SM\n", a, 12 * 64 / a);
  printf("16 STS.128 + 16 LDS.128:            %.1f cycles/iter -> %.3f wavefronts/clk/SM\n", b, 12 * 128 / b);
  printf("both:                               %.1f cycles/iter (sum %.1f, max %.1f) %s\n", c, a + b, a > b ? a : b, cudaGetErrorString(cudaGetLastError()));
  return 0;
}

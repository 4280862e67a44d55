// Probe: one tcgen05.mma.kind::i8 tile (M=128, N=64, K=32*KS, u8 x u8 -> s32, K-major operands in
// shared memory without swizzle) against a CPU product.  Validates the instruction / shared-memory
// descriptor encodings used for a tcgen05 version of the keyswitch GEMM.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, KS = 4, KB = 32;  // KB bytes of K per MMA

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);            // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset (K direction), bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset (M/N direction), bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell)
  return d;                                          // base offset 0, layout type 0 (no swizzle)
}

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* A, const uint8_t* B, int* D, int* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint8_t* sa = smem;                          // KS tiles of [2 kc][16 row-groups][8 rows][16 B] = 4096 B each
  uint8_t* sb = smem + KS * M * KB;            // KS tiles of [2 kc][8 row-groups][8 rows][16 B] = 2048 B each
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fill operands: row r of A (all threads), row r of B (threads < 64), K-major core-matrix layout
  for (int ks = 0; ks < KS; ks++)
    for (int kc = 0; kc < 2; kc++) {
      const uint4 va = *reinterpret_cast<const uint4*>(A + (size_t)tid * KS * KB + ks * KB + kc * 16);
      *reinterpret_cast<uint4*>(sa + ks * M * KB + kc * (M / 8) * 128 + (tid / 8) * 128 + (tid % 8) * 16) = va;
      if (tid < N) {
        const uint4 vb = *reinterpret_cast<const uint4*>(B + (size_t)tid * KS * KB + ks * KB + kc * 16);
        *reinterpret_cast<uint4*>(sb + ks * N * KB + kc * (N / 8) * 128 + (tid / 8) * 128 + (tid % 8) * 16) = vb;
      }
    }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // S32 accum, u8 x u8, K-major
    for (int ks = 0; ks < KS; ks++) {
      const uint64_t da = smem_desc(s32(sa + ks * M * KB), (M / 8) * 128, 128);
      const uint64_t db = smem_desc(s32(sb + ks * N * KB), (N / 8) * 128, 128);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tacc), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
  }
  // bounded wait for the MMAs
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22) && !done; spin++)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(s32(&mbar)) : "memory");
  if (!done) { if (tid == 0) *status = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (done) {
    // thread = TMEM lane = output row; 64 columns in 4 loads of 16
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tacc + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; i++) D[tid * N + c + i] = (int)r[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
}

int main() {
  std::vector<uint8_t> hA(M * KS * KB), hB(N * KS * KB);
  srand(1);
  for (auto& x : hA) x = rand() & 3;      // digits 0..3
  for (auto& x : hB) x = rand() & 0xFF;   // key bytes
  uint8_t *dA, *dB; int *dD, *dS;
  cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size()); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, M * N * 4); cudaMemset(dS, 0, 4);
  const int smem = KS * (M + N) * KB;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<int> hD(M * N); int st = 0;
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      int ref = 0;
      for (int k = 0; k < KS * KB; k++) ref += (int)hA[m * KS * KB + k] * (int)hB[n * KS * KB + k];
      bad += ref != hD[m * N + n];
    }
  printf("umma_i8 probe: cuda=%s status=%d mismatches=%ld / %d  (D[0][0]=%d D[5][7]=%d)\n", cudaGetErrorString(e), st, bad, M * N,
         hD[0], hD[5 * N + 7]);
  return 0;
}

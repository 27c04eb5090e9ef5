// tc_rate — cycles per tcgen05.mma (cta_group::1, M=128) as a function of kind / N, measured with clock64
// around a burst of back-to-back MMAs on one SM.  Operands: zeros in shared memory (timing only).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) rate(int kind, uint32_t idesc, int reps, int distinct_k, int two_acc, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x * 16; i < 65536; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(sm);
    const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    // descriptors are precomputed and the burst is unrolled: the loop must not be issue-bound
    uint64_t ad[4], bd[4];
    for (int q = 0; q < 4; ++q) {
      const uint32_t ko = (uint32_t)(q % distinct_k) * 32;
      ad[q] = hi | (uint64_t)(((base + ko) >> 4) & 0x3fff);
      bd[q] = hi | (uint64_t)(((base + 32768 + ko) >> 4) & 0x3fff);
    }
    const uint32_t d1 = tmem_base + (two_acc ? 256u : 0u);
    t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t dt = (q & 1) ? d1 : tmem_base;
        if (kind == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dt), "l"(ad[q & 3]), "l"(bd[q & 3]), "r"(idesc), "r"(1) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dt), "l"(ad[q & 3]), "l"(bd[q & 3]), "r"(idesc), "r"(1) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    t1 = clock64();
    out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}
static uint32_t idesc(int N, int fmt) { return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24); }
int main() {
  long long* d; CK(cudaMalloc(&d, 8));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000));
  const int reps = 512;
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {32, 64, 128, 224, 256})
      for (int dk : {1, 4}) for (int ta = 0; ta < 2; ++ta) {
        if (ta && N > 256) continue;
        rate<<<1, 128, 70000>>>(kind, idesc(N, kind == 0 ? 2 : 1), reps, dk, ta, d);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
        printf("%s M=128 N=%3d distinct_k=%d two_acc=%d: %.1f cycles/MMA  (%.0f MAC/cycle)\n", kind == 0 ? "tf32 K=8 " : "bf16 K=16", N, dk, ta,
               (double)c / reps, 128.0 * N * (kind == 0 ? 8 : 16) * reps / (double)c);
      }
  // all SMs at once: does the rate hold chip-wide?
  rate<<<148, 128, 70000>>>(0, idesc(256, 2), reps, 4, 0, d);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
  printf("tf32 N=256 on 148 CTAs: %.1f cycles/MMA\n", (double)c / reps);
  return 0;
}

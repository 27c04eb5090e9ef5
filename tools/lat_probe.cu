// lat_probe — dependent-issue latencies of the instructions on the critical path of the in-register
// Cholesky sweep (one warp, one SM): SHFL.IDX, MUFU.RSQ, FFMA, FFMA2, STS->LDS round trip.
#include <cuda_runtime.h>
#include <cstdio>
#define N 1024
__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void probe(float seed, long long* out, float* sink) {
  __shared__ float sh[64];
  const int lane = threadIdx.x;
  float x = seed + lane * 1e-3f;
  long long t0, t1;
  // SHFL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1 + (i & 1)) & 31) + 1e-6f;  // rotation: result stays lane-dependent
  t1 = clock64(); if (lane == 0) out[0] = t1 - t0;
  // MUFU.RSQ chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsq(x);
  t1 = clock64(); if (lane == 0) out[1] = t1 - t0;
  // FFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fmaf(x, 0.999f, 1e-3f);
  t1 = clock64(); if (lane == 0) out[2] = t1 - t0;
  // SHFL + RSQ + FMUL (the pivot chain)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { float p = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); x = x * rsq(p) + 1.f + lane * 1e-3f; }
  t1 = clock64(); if (lane == 0) out[3] = t1 - t0;
  // STS -> syncwarp -> LDS round trip
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { sh[lane] = x; __syncwarp(); x = sh[(lane + 1) & 31] + 1e-3f; __syncwarp(); }
  t1 = clock64(); if (lane == 0) out[4] = t1 - t0;
  // independent SHFL throughput (32 independent values)
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = x + j;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __shfl_sync(0xffffffffu, v[j], (lane + j + 1) & 31);
  }
  t1 = clock64(); if (lane == 0) out[5] = t1 - t0;
#pragma unroll
  for (int j = 0; j < 16; ++j) x += v[j];
  // independent FFMA (3-reg) and FFMA2 throughput
  float w[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = x + j;
  float m = x * 0.5f;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = fmaf(w[j], m, v[j]);
  }
  t1 = clock64(); if (lane == 0) out[6] = t1 - t0;
  unsigned long long m2; asm("mov.b64 %0, {%1, %1};" : "=l"(m2) : "f"(m));
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; j += 2)
      asm volatile("{\n\t.reg .b64 rb, rc;\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\tfma.rn.f32x2 rc, %2, rc, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}" : "+f"(w[j]), "+f"(w[j + 1]) : "l"(m2), "f"(v[j]), "f"(v[j + 1]));
  }
  t1 = clock64(); if (lane == 0) out[7] = t1 - t0;
#pragma unroll
  for (int j = 0; j < 16; ++j) x += w[j];
  sink[lane] = x;
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 64); cudaMalloc(&s, 128);
  probe<<<1, 32>>>(1.5f, d, s); probe<<<1, 32>>>(1.5f, d, s);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const char* names[8] = {"SHFL.IDX + FADD dependent", "MUFU.RSQ dependent", "FFMA dependent", "SHFL+RSQ+FFMA chain", "STS->LDS round trip (2 syncwarp)",
                          "SHFL independent (per instr)", "FFMA 3-reg independent (per instr)", "FFMA2 independent (per instr)"};
  const double div[8] = {N, N, N, N, N, N, N, N / 2.0};
  for (int i = 0; i < 8; ++i) printf("%-40s %.1f cycles\n", names[i], h[i] / div[i]);
  return 0;
}

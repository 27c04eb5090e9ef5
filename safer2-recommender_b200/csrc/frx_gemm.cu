// C[out(i)][:] = A[i][:] * B for a tall A [M x d] and a square B [d x d] (d a multiple of 128), fp32 SIMT.
// Used for the basis changes of the dual-form row path: Et = E * Q (rotated fixed-side factors) and
// X[row] = Xt * Q^T (solutions back to the original basis, scattered to their row ids).
#include "frx_kernels.cuh"

namespace frx {

namespace {

constexpr int GB_M = 128, GB_N = 128, GB_K = 16, GB_THREADS = 256;

__global__ void __launch_bounds__(GB_THREADS) rows_gemm_kernel(const float* __restrict__ A, int M, int d,
                                                               const float* __restrict__ B, float* __restrict__ C,
                                                               const int* __restrict__ c_rows,
                                                               const int* __restrict__ c_map) {
  __shared__ __align__(16) float As[2][GB_K][GB_M];
  __shared__ __align__(16) float Bs[2][GB_K][GB_N];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * GB_M, n0 = blockIdx.y * GB_N;
  const int a_row = t >> 1, a_k4 = (t & 1) * 2;       // A tile: row a_row, float4 columns a_k4, a_k4+1
  const int b_k = t >> 4, b_n4 = (t & 15) * 2;        // B tile: row b_k, float4 columns b_n4, b_n4+1
  const int ty = t >> 4, tx = t & 15;
  const bool a_ok = m0 + a_row < M;
  const float4* a_src = reinterpret_cast<const float4*>(A + (size_t)(a_ok ? m0 + a_row : 0) * d);
  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      ra[j] = a_ok ? __ldg(a_src + (k0 >> 2) + a_k4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[j] = __ldg(reinterpret_cast<const float4*>(B + (size_t)(k0 + b_k) * d + n0) + b_n4 + j);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kk = 4 * (a_k4 + j);
      As[buf][kk + 0][a_row] = ra[j].x;
      As[buf][kk + 1][a_row] = ra[j].y;
      As[buf][kk + 2][a_row] = ra[j].z;
      As[buf][kk + 3][a_row] = ra[j].w;
      *reinterpret_cast<float4*>(&Bs[buf][b_k][4 * (b_n4 + j)]) = rb[j];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < d; k0 += GB_K) {
    const bool more = k0 + GB_K < d;
    if (more) fetch(k0 + GB_K);
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= M) continue;
    int orow = row;
    if (c_rows) orow = c_rows[row];
    if (c_map) orow = c_map[orow];
    float4* dst = reinterpret_cast<float4*>(C + (size_t)orow * d + n0 + tx * 8);
    dst[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    dst[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

}  // namespace

void launch_rows_gemm(const float* A, int M, int d, const float* B, float* C, const int* c_rows, const int* c_map,
                      cudaStream_t s, long long* launches) {
  if (M <= 0) return;
  dim3 grid((M + GB_M - 1) / GB_M, d / GB_N);
  rows_gemm_kernel<<<grid, GB_THREADS, 0, s>>>(A, M, d, B, C, c_rows, c_map);
  if (launches) ++*launches;
}

}  // namespace frx

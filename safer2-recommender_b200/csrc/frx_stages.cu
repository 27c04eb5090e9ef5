// Side stages of the epoch: Gramians, per-user loss, dual weights, the
// smoothed-quantile Newton iteration, exact quantile, prediction cache and the
// one-off Initialize() reductions.  Reference citations are per kernel.
#include "frx_kernels.cuh"
#include <cooperative_groups.h>
#include <cfloat>

namespace cg = cooperative_groups;

namespace frx {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of a double; result valid in every thread.  blockDim <= 1024.
__device__ double block_sum_d(double v, double* sh /*[33]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? sh[lane] : 0.0;
    t = warp_sum_d(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// ---------------------------------------------------------------------------
// Gramian, SIMT two-stage version: partial tiles per row slab, then a fixed-order
// reduction (deterministic).  Restates `X.transpose() * Y` at ials.h:321,
// safer2.h:55,294-295,504-509 and the block forms ialspp.h:356-365,
// safer2pp.h:534-544.
// ---------------------------------------------------------------------------
constexpr int GT = 64;   // output tile edge
constexpr int GK = 32;   // rows per staged chunk

__global__ void __launch_bounds__(256) gramian_partial_kernel(
    const float* __restrict__ E, int n, int d, int cs, int bd, int fs, int fd,
    const float* __restrict__ w, float* __restrict__ ws, int rows_per_slab, int tiles_j) {
  __shared__ __align__(16) float As[GK][GT];
  __shared__ __align__(16) float Bs[GK][GT];
  const int tile = blockIdx.y;
  const int i0 = (tile / tiles_j) * GT, j0 = (tile % tiles_j) * GT;
  const int slab = blockIdx.x;
  const int r_begin = slab * rows_per_slab;
  const int r_end = min(n, r_begin + rows_per_slab);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int r0 = r_begin; r0 < r_end; r0 += GK) {
    for (int idx = threadIdx.x; idx < GK * GT; idx += 256) {
      const int rr = idx / GT, c = idx % GT;
      const int row = r0 + rr;
      float a = 0.f, b = 0.f;
      if (row < r_end) {
        const float* er = E + (size_t)row * d;
        const float wr = w ? w[row] : 1.f;
        if (i0 + c < bd) a = er[cs + i0 + c] * wr;
        if (j0 + c < fd) b = er[fs + j0 + c];
      }
      As[rr][c] = a;
      Bs[rr][c] = b;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < GK; ++rr) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[rr][4 * ty]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[rr][4 * tx]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* out = ws + ((size_t)slab * gridDim.y + tile) * (GT * GT);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) out[(4 * ty + a) * GT + 4 * tx + b] = acc[a][b];
}

__global__ void gramian_reduce_kernel(const float* __restrict__ ws, int slabs, int tiles, int tiles_j,
                                      int bd, int fd, float* __restrict__ out, int ld_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= bd * fd) return;
  const int i = idx / fd, j = idx % fd;
  const int tile = (i / GT) * tiles_j + (j / GT);
  const int off = (i % GT) * GT + (j % GT);
  float s = 0.f;
  for (int sl = 0; sl < slabs; ++sl) s += ws[((size_t)sl * tiles + tile) * (GT * GT) + off];
  out[(size_t)i * ld_out + j] = s;
}

int gramian_slabs(int n, int tiles, int num_sms) {
  int want = (4 * num_sms + tiles - 1) / tiles;  // ~4 waves of CTAs
  int max_slabs = (n + 255) / 256;               // at least 256 rows per slab
  if (max_slabs < 1) max_slabs = 1;
  if (want > max_slabs) want = max_slabs;
  if (want < 1) want = 1;
  return want;
}

// ---------------------------------------------------------------------------
// Kernel functions of the smoothed quantile, safer2.h:599-647.  float in/out,
// double bodies, exactly the promotions of the reference expressions
// (SURVEY.md D.4); B-14: float fabs.  pow(x, 2.0) / pow(x, 3.0) / pow(x, -3.0) are
// written as products: glibc's pow is exact for these (what the reference gets),
// CUDA's generic pow() is not and costs several hundred FP64 instructions per call,
// which made the xi kernel compute-bound.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float gaussian_kernel(const float u, const float h) {
  const double t = (double)(u / h) * 0.70710678118654752440;
  return (float)(0.39894228040143267794 /* pow(2 pi, -0.5) */ * exp(-(t * t)) / (double)h);
}
__device__ __forceinline__ float gaussian_kernel_cdf(const float u, const float h) {
  return (float)(0.5 * erfc((double)(-(u / h)) * 0.70710678118654752440));
}
__device__ __forceinline__ float gaussian_loss(const float u, const float h, const float alpha) {
  const float ell = h * gaussian_kernel(u, h) + (u / h) * (1 - 2 * gaussian_kernel_cdf(-u, h));
  return (float)((double)((h / 2) * ell) + ((double)(1 - alpha) - 0.5) * (double)u);
}
__device__ __forceinline__ float epanechnikov_kernel(const float u, const float h) {
  const float uh = u / h;
  const double uhd = (double)uh;
  return (float)((3.0 / 4.0) * (1 - uhd * uhd) * (int)(fabsf(uh) < 1) / (double)h);
}
__device__ __forceinline__ float epanechnikov_kernel_cdf(const float u, const float h) {
  const float uh = u / h;
  const int in_supp = (int)(fabsf(uh) <= 1);
  const int pos = (int)(uh > 1);
  const double hd = h, ud = u;
  const double h2 = hd * hd, h3 = h2 * hd;
  return (float)((((1.0 / h3) / 4.0) *
                  (((double)(3 * u) * h2 - ud * ud * ud) + 2 * h3) * in_supp) +
                 (1 - in_supp) * pos);
}
__device__ __forceinline__ float epanechnikov_loss(const float u, const float h, const float alpha) {
  const float uh = u / h;
  const int in_supp = (int)(fabsf(uh) <= 1);
  const int pos = (int)(uh > 1);
  const double uh2 = (double)uh * (double)uh;
  const float ell = (float)(((3.0 / 4.0) * uh2 - (1.0 / 8.0) * (uh2 * uh2) +
                             (3.0 / 8.0)) * in_supp + (double)(fabsf(uh) * pos));
  return (float)((1.0 / 2.0) * (double)h * (double)ell + ((double)(1 - alpha) - 0.5) * (double)u);
}

// ---------------------------------------------------------------------------
// u^T G u for a tile of users (the `ireg` of ComputeLoss, safer2.h:97).
// ---------------------------------------------------------------------------
constexpr int QU = 16;  // users per CTA
__global__ void __launch_bounds__(256) quadform_kernel(const float* __restrict__ U, int num_users, int d,
                                                       const float* __restrict__ G, float* __restrict__ quad) {
  extern __shared__ __align__(16) float sh[];
  float* Us = sh;                 // [QU][dp]
  const int dp = (d + 3) & ~3;
  float* red = sh + QU * dp;      // [QU][8 warps]
  const int u0 = blockIdx.x * QU;
  for (int idx = threadIdx.x; idx < QU * dp; idx += 256) {
    const int uu = idx / dp, k = idx % dp;
    Us[idx] = (u0 + uu < num_users && k < d) ? U[(size_t)(u0 + uu) * d + k] : 0.f;
  }
  __syncthreads();
  float tot[QU];
#pragma unroll
  for (int uu = 0; uu < QU; ++uu) tot[uu] = 0.f;
  for (int j = threadIdx.x; j < d; j += 256) {
    float acc[QU];
#pragma unroll
    for (int uu = 0; uu < QU; ++uu) acc[uu] = 0.f;
    int i = 0;
    for (; i + 4 <= d; i += 4) {
      const float g0 = __ldg(G + (size_t)(i + 0) * d + j), g1 = __ldg(G + (size_t)(i + 1) * d + j);
      const float g2 = __ldg(G + (size_t)(i + 2) * d + j), g3 = __ldg(G + (size_t)(i + 3) * d + j);
#pragma unroll
      for (int uu = 0; uu < QU; ++uu) {
        const float4 u4 = *reinterpret_cast<const float4*>(&Us[uu * dp + i]);
        acc[uu] = fmaf(u4.x, g0, acc[uu]);
        acc[uu] = fmaf(u4.y, g1, acc[uu]);
        acc[uu] = fmaf(u4.z, g2, acc[uu]);
        acc[uu] = fmaf(u4.w, g3, acc[uu]);
      }
    }
    for (; i < d; ++i) {
      const float g0 = __ldg(G + (size_t)i * d + j);
#pragma unroll
      for (int uu = 0; uu < QU; ++uu) acc[uu] = fmaf(Us[uu * dp + i], g0, acc[uu]);
    }
#pragma unroll
    for (int uu = 0; uu < QU; ++uu) tot[uu] = fmaf(acc[uu], Us[uu * dp + j], tot[uu]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int uu = 0; uu < QU; ++uu) {
    const float v = warp_sum(tot[uu]);
    if (lane == 0) red[uu * 8 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < QU && u0 + threadIdx.x < num_users) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x * 8 + w];
    quad[u0 + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------
// Per-user loss, ComputeLoss safer2.h:85-101 / ials.h:70-86 / safer2pp.h:80-95.
// One warp per user; the squared residuals are added in history (file) order
// into a float through a double, as `loss += pow(x - 1, 2.0)` does.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) user_loss_kernel(LossParams p) {
  extern __shared__ __align__(16) float sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = p.d;
  float* us = sh + warp * ((d + 3) & ~3);
  const int gw = blockIdx.x * 8 + warp;
  const int stride = gridDim.x * 8;
  for (int ri = gw; ri < p.num_rows; ri += stride) {
    const int u = p.order[ri];
    const int beg = p.ptr[u], n = p.ptr[u + 1] - beg;
    float loss = 0.f;
    double obs = 0.0;
    if (p.pred) {
      // safer2pp.h:86-89: residuals from the cached predictions
      double part = 0.0;
      for (int e = lane; e < n; e += 32) {
        const float x = p.pred[p.tup[beg + e]] - 1.f;
        part += (double)x * (double)x;
      }
      obs = warp_sum_d(part);
      loss = (float)obs;
    } else {
      __syncwarp();
      for (int k = lane; k < d; k += 32) us[k] = p.U[(size_t)u * d + k];
      __syncwarp();
      int e = 0;
      if ((d & 127) == 0) {
        // vector path: each lane owns float4 slices (coalesced 512 B per warp load), 8 history entries in flight
        for (; e + 8 <= n; e += 8) {
          const float* vp[8];
          float acc[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            vp[q] = p.V + (size_t)__ldg(p.col + beg + e + q) * d;
            acc[q] = 0.f;
          }
          for (int k = lane * 4; k < d; k += 128) {
            const float4 u4 = *reinterpret_cast<const float4*>(us + k);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 v4 = __ldg(reinterpret_cast<const float4*>(vp[q] + k));
              acc[q] = fmaf(v4.x, u4.x, acc[q]);
              acc[q] = fmaf(v4.y, u4.y, acc[q]);
              acc[q] = fmaf(v4.z, u4.z, acc[q]);
              acc[q] = fmaf(v4.w, u4.w, acc[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const double sq = (double)(warp_sum(acc[q]) - 1.f);
            loss = (float)((double)loss + sq * sq);
            obs += sq * sq;
          }
        }
      }
      for (; e + 4 <= n; e += 4) {
        const float* v0 = p.V + (size_t)p.col[beg + e] * d;
        const float* v1 = p.V + (size_t)p.col[beg + e + 1] * d;
        const float* v2 = p.V + (size_t)p.col[beg + e + 2] * d;
        const float* v3 = p.V + (size_t)p.col[beg + e + 3] * d;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int k = lane; k < d; k += 32) {
          const float uk = us[k];
          a0 = fmaf(__ldg(v0 + k), uk, a0);
          a1 = fmaf(__ldg(v1 + k), uk, a1);
          a2 = fmaf(__ldg(v2 + k), uk, a2);
          a3 = fmaf(__ldg(v3 + k), uk, a3);
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
        const double s0 = (double)(a0 - 1.f), s1 = (double)(a1 - 1.f);
        const double s2 = (double)(a2 - 1.f), s3 = (double)(a3 - 1.f);
        loss = (float)((double)loss + s0 * s0);
        loss = (float)((double)loss + s1 * s1);
        loss = (float)((double)loss + s2 * s2);
        loss = (float)((double)loss + s3 * s3);
        obs += s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3;
      }
      for (; e < n; ++e) {
        const float* v0 = p.V + (size_t)p.col[beg + e] * d;
        float a0 = 0.f;
        for (int k = lane; k < d; k += 32) a0 = fmaf(__ldg(v0 + k), us[k], a0);
        a0 = warp_sum(a0);
        const double s0 = (double)(a0 - 1.f);
        loss = (float)((double)loss + s0 * s0);
        obs += s0 * s0;
      }
    }
    if (lane == 0) {
      loss /= (float)n;
      loss += p.beta * p.quad[u];
      if (p.halve) loss *= 0.5f;
      p.loss[u] = loss;
      if (p.obs_sq) p.obs_sq[u] = obs;
    }
  }
}

// Pass 1 of the two-pass user loss: resid[beg + e] = u . v_{col[beg+e]} - 1 for the entries of one chunk.
__global__ void __launch_bounds__(256) user_resid_kernel(LossParams p) {
  extern __shared__ __align__(16) float sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = p.d;
  float* us = sh + warp * ((d + 3) & ~3);
  for (int it = blockIdx.x * 8 + warp; it < p.num_chunks; it += gridDim.x * 8) {
    const int u = p.chunk_row[it], off = p.chunk_off[it];
    const int beg = p.ptr[u] + off;
    const int n = min(FRX_LOSS_CHUNK, p.ptr[u + 1] - beg);
    __syncwarp();
    for (int k = lane; k < d; k += 32) us[k] = p.U[(size_t)u * d + k];
    __syncwarp();
    int e = 0;
    if ((d & 127) == 0) {
      for (; e + 8 <= n; e += 8) {
        const float* vp[8];
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          vp[q] = p.V + (size_t)__ldg(p.col + beg + e + q) * d;
          acc[q] = 0.f;
        }
        for (int k = lane * 4; k < d; k += 128) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + k);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(vp[q] + k));
            acc[q] = fmaf(v4.x, u4.x, acc[q]);
            acc[q] = fmaf(v4.y, u4.y, acc[q]);
            acc[q] = fmaf(v4.z, u4.z, acc[q]);
            acc[q] = fmaf(v4.w, u4.w, acc[q]);
          }
        }
        float mine = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float r = warp_sum(acc[q]) - 1.f;
          if (lane == q) mine = r;
        }
        if (lane < 8) p.resid[beg + e + lane] = mine;
      }
    }
    for (; e < n; ++e) {
      const float* v0 = p.V + (size_t)p.col[beg + e] * d;
      float a0 = 0.f;
      for (int k = lane; k < d; k += 32) a0 = fmaf(__ldg(v0 + k), us[k], a0);
      a0 = warp_sum(a0);
      if (lane == 0) p.resid[beg + e] = a0 - 1.f;
    }
  }
}

// Pass 2: per user, the running float sum of the squared residuals in history order (safer2.h:86-99).
__global__ void __launch_bounds__(128) user_loss_finish_kernel(LossParams p) {
  const int ri = blockIdx.x * blockDim.x + threadIdx.x;
  if (ri >= p.num_rows) return;
  const int u = p.order[ri];
  const int beg = p.ptr[u], n = p.ptr[u + 1] - beg;
  float loss = 0.f;
  double obs = 0.0;
  const float* r = p.resid + beg;
  for (int e = 0; e < n; ++e) {
    const double s0 = (double)r[e];
    loss = (float)((double)loss + s0 * s0);
    obs += s0 * s0;
  }
  loss /= (float)n;
  loss += p.beta * p.quad[u];
  if (p.halve) loss *= 0.5f;
  p.loss[u] = loss;
  if (p.obs_sq) p.obs_sq[u] = obs;
}

// PredictDataset (ialspp.h:469-517) over FRX_LOSS_CHUNK-entry chunks of the rows (a popular item has tens of
// thousands of entries: one warp per ROW leaves the launch waiting for a single warp): eight entries per pass,
// pred[tup] = x_row . e_col.
__global__ void __launch_bounds__(256) predict_chunks_kernel(const int* __restrict__ ptr, const int* __restrict__ col,
                                                             const int* __restrict__ tup,
                                                             const int* __restrict__ chunk_row,
                                                             const int* __restrict__ chunk_off, int num_chunks,
                                                             const float* __restrict__ U, const int* __restrict__ xmap,
                                                             const float* __restrict__ V, int d,
                                                             float* __restrict__ pred) {
  extern __shared__ __align__(16) float sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* us = sh + warp * ((d + 3) & ~3);
  for (int it = blockIdx.x * 8 + warp; it < num_chunks; it += gridDim.x * 8) {
    const int u = chunk_row[it], off = chunk_off[it];
    const int xr = xmap ? xmap[u] : u;
    const int beg = ptr[u] + off;
    const int n = min(FRX_LOSS_CHUNK, ptr[u + 1] - beg);
    __syncwarp();
    for (int k = lane; k < d; k += 32) us[k] = U[(size_t)xr * d + k];
    __syncwarp();
    int e = 0;
    if ((d & 127) == 0) {
      for (; e + 8 <= n; e += 8) {
        const float* vp[8];
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          vp[q] = V + (size_t)__ldg(col + beg + e + q) * d;
          acc[q] = 0.f;
        }
        for (int k = lane * 4; k < d; k += 128) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + k);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(vp[q] + k));
            acc[q] = fmaf(v4.x, u4.x, acc[q]);
            acc[q] = fmaf(v4.y, u4.y, acc[q]);
            acc[q] = fmaf(v4.z, u4.z, acc[q]);
            acc[q] = fmaf(v4.w, u4.w, acc[q]);
          }
        }
        float mine = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float r = warp_sum(acc[q]);
          if (lane == q) mine = r;
        }
        if (lane < 8) pred[__ldg(tup + beg + e + lane)] = mine;
      }
    }
    for (; e < n; ++e) {
      const float* v0 = V + (size_t)col[beg + e] * d;
      float a0 = 0.f;
      for (int k = lane; k < d; k += 32) a0 = fmaf(__ldg(v0 + k), us[k], a0);
      a0 = warp_sum(a0);
      if (lane == 0) pred[tup[beg + e]] = a0;
    }
  }
}

__global__ void __launch_bounds__(256) predict_kernel(const int* __restrict__ ptr, const int* __restrict__ col,
                                                      const int* __restrict__ tup, const int* __restrict__ order,
                                                      int num_rows, const float* __restrict__ U,
                                                      const int* __restrict__ xmap, const float* __restrict__ V,
                                                      int d, float* __restrict__ pred) {
  extern __shared__ __align__(16) float sh[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* us = sh + warp * ((d + 3) & ~3);
  for (int ri = blockIdx.x * 8 + warp; ri < num_rows; ri += gridDim.x * 8) {
    const int u = order[ri];
    const int xr = xmap ? xmap[u] : u;
    const int beg = ptr[u], n = ptr[u + 1] - beg;
    __syncwarp();
    for (int k = lane; k < d; k += 32) us[k] = U[(size_t)xr * d + k];
    __syncwarp();
    for (int e = 0; e < n; ++e) {
      const float* v = V + (size_t)col[beg + e] * d;
      float a = 0.f;
      for (int k = lane; k < d; k += 32) a = fmaf(__ldg(v + k), us[k], a);
      a = warp_sum(a);
      if (lane == 0) pred[tup[beg + e]] = a;
    }
  }
}

__global__ void user_weights_kernel(const float* __restrict__ loss, const float* __restrict__ hist_size,
                                    int n, float* __restrict__ z, const float* __restrict__ xi_dev,
                                    float h, int kind, int only_with_history) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  if (only_with_history && !(hist_size[u] > 0.f)) return;
  const float xi = *xi_dev;
  const float r = loss[u] - xi;
  float nw;
  if (kind == 2) nw = (r >= 0) ? 1.f : 0.f;                          // cvar_mf.h:623
  else if (kind == 1) nw = 1 - epanechnikov_kernel_cdf(-r, h);       // safer2.h:773
  else nw = 1 - gaussian_kernel_cdf(-r, h);                          // safer2.h:775
  z[u] = nw;
}

// ---------------------------------------------------------------------------
// xi: Newton-Raphson with Armijo backtracking on the smoothed quantile
// objective, safer2.h:652-742, all iterations in ONE cooperative kernel.  Every
// EvaluateQuantile is a grid-wide 3-way reduction: per-CTA partial sums in
// double -> grid barrier -> every CTA adds the partials in the same fixed order,
// so all CTAs take identical branches.
// ---------------------------------------------------------------------------
constexpr int XI_MAXK = 8;  // Armijo trial points evaluated per pass
struct QEval { float value, grad, H; };

// K evaluation points at once: one pass over the (sub-sampled) losses, one grid barrier.  For every point the
// per-thread / per-CTA / cross-CTA summation order is the one a single evaluation uses, so the values (and with
// them the Newton / Armijo branches) do not depend on K.
template <int K>
__device__ void evaluate_quantile(const XiParams& p, const int* idx, int n, const float (&xi)[K], int parity,
                                  double* sh, cg::grid_group& grid, QEval (&out)[K]) {
  double s_cdf[K], s_pdf[K], s_loss[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s_cdf[k] = s_pdf[k] = s_loss[k] = 0.0;
  const float h = p.bandwidth, alpha = p.alpha;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float l = idx ? p.loss[idx[i]] : p.loss[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float u = l - xi[k];
      if (p.epanechnikov) {
        s_cdf[k] += epanechnikov_kernel_cdf(-u, h);
        s_pdf[k] += epanechnikov_kernel(-u, h);
        s_loss[k] += epanechnikov_loss(u, h, alpha);
      } else {
        s_cdf[k] += gaussian_kernel_cdf(-u, h);
        s_pdf[k] += gaussian_kernel(-u, h);
        s_loss[k] += gaussian_loss(u, h, alpha);
      }
    }
  }
  double* part = p.partials + (size_t)parity * gridDim.x * (3 * XI_MAXK);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double a = block_sum_d(s_cdf[k], sh);
    const double b = block_sum_d(s_pdf[k], sh);
    const double c = block_sum_d(s_loss[k], sh);
    if (threadIdx.x == 0) {
      part[(size_t)blockIdx.x * (3 * XI_MAXK) + 3 * k + 0] = a;
      part[(size_t)blockIdx.x * (3 * XI_MAXK) + 3 * k + 1] = b;
      part[(size_t)blockIdx.x * (3 * XI_MAXK) + 3 * k + 2] = c;
    }
  }
  grid.sync();
  // thread q < 3K adds quantity q over the CTAs in CTA order (the fixed order every CTA repeats)
  double* tot = sh + 40;
  if (threadIdx.x < 3 * K) {
    double t = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) t += part[(size_t)b * (3 * XI_MAXK) + threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float mean_cdf = (float)(tot[3 * k] / n), mean_pdf = (float)(tot[3 * k + 1] / n),
                mean_loss = (float)(tot[3 * k + 2] / n);
    out[k].grad = (-(1 - alpha) + mean_cdf) / alpha;  // safer2.h:659-686
    out[k].H = mean_pdf / alpha;
    out[k].value = mean_loss / alpha;
  }
  __syncthreads();  // tot is reused by the next evaluation
}

__global__ void __launch_bounds__(256) xi_newton_kernel(XiParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[40 + 3 * XI_MAXK];
  int parity = 0;
  float xi;
  if (p.start_from_mean) {  // Initialize: prev_xi = user_loss_.mean() (safer2.h:822)
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.num_users; i += gridDim.x * blockDim.x)
      s += (double)p.loss[i];
    s = block_sum_d(s, sh);
    double* part = p.partials + (size_t)parity * gridDim.x * (3 * XI_MAXK);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * (3 * XI_MAXK)] = s;
    grid.sync();
    double t = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) t += part[(size_t)b * (3 * XI_MAXK)];
    xi = (float)(t / p.num_users);
    parity ^= 1;
  } else {
    xi = *p.xi_io;
  }
  for (int t = 0; t < p.iters; ++t) {
    const int* idx = p.snr_idx ? p.snr_idx + (size_t)t * p.n_samples : nullptr;
    const int n = p.snr_idx ? p.n_samples : p.num_users;
    // ComputeXiDirection, safer2.h:692-712.  The reference halves gamma until the Armijo test passes (at most 32
    // times), one EvaluateQuantile per trial; here the trial points are evaluated XI_MAXK at a time and the first
    // one that passes is taken: same values, same choice, two grid barriers per Newton step in the usual case.
    QEval e0[1];
    {
      const float x0[1] = {xi};
      evaluate_quantile<1>(p, idx, n, x0, parity, sh, grid, e0);
      parity ^= 1;
    }
    const float d = e0[0].grad / e0[0].H;
    const float c = 1e-4f;
    float gamma = 1.0f;
    bool accepted = false;
    for (int kb = 0; kb < 32 && !accepted; kb += XI_MAXK) {
      float xs[XI_MAXK], gs[XI_MAXK];
      float g = gamma;
#pragma unroll
      for (int j = 0; j < XI_MAXK; ++j) {
        gs[j] = g;
        xs[j] = xi + g * (-d);
        g *= 0.5f;
      }
      QEval ex[XI_MAXK];
      evaluate_quantile<XI_MAXK>(p, idx, n, xs, parity, sh, grid, ex);
      parity ^= 1;
#pragma unroll
      for (int j = 0; j < XI_MAXK; ++j) {
        if (!accepted) {
          if (ex[j].value > e0[0].value + c * gs[j] * ex[j].grad * (-d)) {  // B-6: trial point's gradient
            gamma = gs[j] * 0.5f;
          } else {
            gamma = gs[j];
            accepted = true;
          }
        }
      }
    }
    xi = xi + (-gamma * d);
  }
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.xi_io = xi;
}

// ---------------------------------------------------------------------------
// Exact quantile, cvar_mf.h:582-595: nth_element of -loss at Q=(size_t)(n*alpha),
// returned negated.  Single-CTA 4-pass radix select on order-preserving keys.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

__global__ void __launch_bounds__(1024) exact_quantile_kernel(const float* __restrict__ loss, int n, float alpha,
                                                              float* __restrict__ xi_out) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_rank;
  const float Qf = (float)n * alpha;  // vals.size() * alpha_  (float product)
  unsigned rank = (unsigned)(size_t)Qf;  // 0-based rank in ascending order of -loss
  if (rank >= (unsigned)n) rank = n - 1;
  unsigned prefix = 0, mask = 0;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned k = f2key(-loss[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned acc = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] > rank) break;
        acc += hist[b];
      }
      if (b > 255) b = 255;
      s_prefix = prefix | ((unsigned)b << (8 * pass));
      s_rank = rank - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    rank = s_rank;
    mask |= 0xffu << (8 * pass);
    __syncthreads();
  }
  if (threadIdx.x == 0) *xi_out = -key2f(prefix);
}

__global__ void __launch_bounds__(256) weight_means_partial(const float* __restrict__ z, const float* __restrict__ loss,
                                                            int n, double* __restrict__ ws) {
  __shared__ double sh[33];
  double a = 0.0, b = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    a += (double)z[i];
    b += (double)(z[i] * loss[i]);  // (dual_weight_.array() * user_loss_.array()).mean(), safer2.h:300
  }
  a = block_sum_d(a, sh);
  b = block_sum_d(b, sh);
  if (threadIdx.x == 0) { ws[blockIdx.x * 2] = a; ws[blockIdx.x * 2 + 1] = b; }
}
__global__ void weight_means_final(const double* __restrict__ ws, int nb, int n, float* __restrict__ out2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nb; ++i) { a += ws[i * 2]; b += ws[i * 2 + 1]; }
    out2[0] = (float)(a / n);
    out2[1] = (float)(b / n);
  }
}

__global__ void hist_size_kernel(const int* __restrict__ uptr, int rows, float* __restrict__ hist_size) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < rows) {
    const int n = uptr[u + 1] - uptr[u];
    if (n > 0) hist_size[u] = (float)n;  // safer2.h:826-829
  }
}
// item_reg_(v) += 1.0 / user_history_size_(u) in item-history order, float += double (safer2.h:830-837)
__global__ void item_reg_kernel(const int* __restrict__ iptr, const int* __restrict__ icol, int rows,
                                const float* __restrict__ hist_size, float* __restrict__ item_reg) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= rows) return;
  float acc = item_reg[v];
  for (int t = iptr[v]; t < iptr[v + 1]; ++t) acc = (float)((double)acc + 1.0 / (double)hist_size[icol[t]]);
  item_reg[v] = acc;
}
__global__ void norm_weights_kernel(const float* __restrict__ z, const float* __restrict__ hs, int n,
                                    float* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) out[u] = z[u] / hs[u];
}
// PrintLosses / ComputeLosses regulariser sums (safer2.h:355-380, ials.h:246-262) on the device: one warp per row
// with a non-empty history: out[0] += |x_r|^2 * reg_r, out[1] += |x_r|^2 (double; atomics: a printed statistic).
// kind 0: iALS reg * (n + uw * num_other)^nu (ials.h:310-315); 1: SAFER2 user reg (1 + uw * num_other)
// (safer2.h:418-421); 2: SAFER2 item reg (item_reg_v + alpha * uw * num_other) (safer2.h:426-432).
__global__ void __launch_bounds__(256) reg_sums_kernel(const float* __restrict__ X, const int* __restrict__ ptr, int rows,
                                                       int d, int kind, float reg, float reg_exp, float uw, float alpha,
                                                       int num_other, const float* __restrict__ item_reg,
                                                       double* __restrict__ out) {
  __shared__ double sh[33];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  double a = 0.0, b = 0.0;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const int n = ptr[r + 1] - ptr[r];
    if (n == 0) continue;
    double n2 = 0.0;
    for (int k = lane; k < d; k += 32) {
      const double v = (double)X[(size_t)r * d + k];
      n2 += v * v;
    }
    n2 = warp_sum_d(n2);
    float rf;
    if (kind == 0) rf = (float)((double)reg * pow((double)((float)n + uw * (float)num_other), (double)reg_exp));
    else if (kind == 1) rf = reg * (1 + uw * num_other);
    else rf = reg * (item_reg[r] + alpha * uw * num_other);
    if (lane == 0) { a += n2 * (double)rf; b += n2; }
  }
  a = block_sum_d(a, sh);
  b = block_sum_d(b, sh);
  if (threadIdx.x == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
}
// *out += sum a[i] * b[i] (b may be null: sum a[i]); double
__global__ void __launch_bounds__(256) dot_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                      double* __restrict__ out) {
  __shared__ double sh[33];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s += b ? (double)a[i] * (double)b[i] : (double)a[i];
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__global__ void __launch_bounds__(256) sum_d_kernel(const double* __restrict__ a, size_t n, double* __restrict__ out) {
  __shared__ double sh[33];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += a[i];
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void __launch_bounds__(256) entry_weights_kernel(const int* __restrict__ col, const float* __restrict__ w,
                                                            size_t e_begin, size_t e_end, float* __restrict__ ew) {
  const size_t e = e_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < e_end) ew[e] = __ldg(w + __ldg(col + e));
}

// out[0] = max(out[0], bits(max |a[i]|)), out[1] likewise for w (non-negative floats compare like their bit patterns)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ a, size_t n, const float* __restrict__ w,
                                                     size_t nw, unsigned* __restrict__ out) {
  float m = 0.f, mw = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = n / 4;
  for (size_t i = i0; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  for (size_t i = 4 * n4 + i0; i < n; i += stride) m = fmaxf(m, fabsf(a[i]));
  for (size_t i = i0; i < nw; i += stride) mw = fmaxf(mw, fabsf(w[i]));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, __float_as_uint(m));
    if (w) atomicMax(out + 1, __float_as_uint(mw));
  }
}

// sum (a - b)^2 in double -> *out (atomic: a diagnostic, the order of the partial sums is not fixed)
__global__ void __launch_bounds__(256) sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                     double* __restrict__ out) {
  __shared__ double sh[33];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double dlt = (double)a[i] - (double)b[i];
    s += dlt * dlt;
  }
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__global__ void fill_kernel(float* p, size_t n, float v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

size_t gramian_workspace_floats(int n, int bd, int fd, int num_sms) {
  const int tiles = ((bd + GT - 1) / GT) * ((fd + GT - 1) / GT);
  return (size_t)gramian_slabs(n, tiles, num_sms) * tiles * GT * GT;
}

void launch_gramian(const float* E, int n, int d, int cs, int bd, int fs, int fd, const float* w,
                    float* out, int ld_out, float* workspace, size_t workspace_floats,
                    cudaStream_t s, int num_sms, long long* launches) {
  const int tiles_i = (bd + GT - 1) / GT, tiles_j = (fd + GT - 1) / GT;
  const int tiles = tiles_i * tiles_j;
  const int slabs = gramian_slabs(n, tiles, num_sms);
  (void)workspace_floats;
  int rows_per_slab = (n + slabs - 1) / slabs;
  rows_per_slab = ((rows_per_slab + GK - 1) / GK) * GK;
  dim3 grid(slabs, tiles);
  gramian_partial_kernel<<<grid, 256, 0, s>>>(E, n, d, cs, bd, fs, fd, w, workspace, rows_per_slab, tiles_j);
  gramian_reduce_kernel<<<(bd * fd + 255) / 256, 256, 0, s>>>(workspace, slabs, tiles, tiles_j, bd, fd, out, ld_out);
  if (launches) *launches += 2;
}

void launch_quadform(const LossParams& p, int user_begin, int user_end, cudaStream_t s, long long* launches) {
  const int d = p.d, dp = (d + 3) & ~3;
  if (user_end > user_begin) {  // u^T G u for the users [user_begin, user_end) the loss kernel will visit
    const int nu = user_end - user_begin;
    const size_t smem = sizeof(float) * (size_t)(QU * dp + QU * 8);
    cudaFuncSetAttribute(quadform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    quadform_kernel<<<(nu + QU - 1) / QU, 256, smem, s>>>(p.U + (size_t)user_begin * d, nu, d, p.G, p.quad + user_begin);
    if (launches) ++*launches;
  }
}

void launch_user_loss(const LossParams& p, int user_begin, int user_end, cudaStream_t s, int num_sms, long long* launches) {
  launch_quadform(p, user_begin, user_end, s, launches);
  launch_user_loss_rows(p, s, num_sms, launches);
}

void launch_user_loss_rows(const LossParams& p, cudaStream_t s, int num_sms, long long* launches) {
  const int d = p.d, dp = (d + 3) & ~3;
  if (p.num_rows > 0 && p.resid && !p.pred && p.num_chunks > 0) {
    const size_t smem = sizeof(float) * (size_t)(8 * dp);
    int grid = (p.num_chunks + 7) / 8;
    const int cap = num_sms * 8;
    if (grid > cap) grid = cap;
    cudaFuncSetAttribute(user_resid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    user_resid_kernel<<<grid, 256, smem, s>>>(p);
    user_loss_finish_kernel<<<(p.num_rows + 127) / 128, 128, 0, s>>>(p);
    if (launches) *launches += 2;
    return;
  }
  if (p.num_rows > 0) {
    const size_t smem = sizeof(float) * (size_t)(8 * dp);
    int grid = (p.num_rows + 7) / 8;
    const int cap = num_sms * 8;
    if (grid > cap) grid = cap;
    cudaFuncSetAttribute(user_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    user_loss_kernel<<<grid, 256, smem, s>>>(p);
    if (launches) ++*launches;
  }
}

void launch_predict(const int* ptr, const int* col, const int* tup, const int* order, int num_rows,
                    const float* U, const int* xmap, const float* V, int d, float* pred,
                    cudaStream_t s, long long* launches) {
  if (num_rows <= 0) return;
  const size_t smem = sizeof(float) * (size_t)(8 * ((d + 3) & ~3));
  int grid = (num_rows + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  predict_kernel<<<grid, 256, smem, s>>>(ptr, col, tup, order, num_rows, U, xmap, V, d, pred);
  if (launches) ++*launches;
}

void launch_predict_chunks(const int* ptr, const int* col, const int* tup, const int* chunk_row, const int* chunk_off,
                           int num_chunks, const float* U, const int* xmap, const float* V, int d, float* pred,
                           cudaStream_t s, int num_sms, long long* launches) {
  if (num_chunks <= 0) return;
  const size_t smem = sizeof(float) * (size_t)(8 * ((d + 3) & ~3));
  int grid = (num_chunks + 7) / 8;
  if (grid > num_sms * 8) grid = num_sms * 8;
  cudaFuncSetAttribute(predict_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  predict_chunks_kernel<<<grid, 256, smem, s>>>(ptr, col, tup, chunk_row, chunk_off, num_chunks, U, xmap, V, d, pred);
  if (launches) ++*launches;
}

void launch_user_weights(const float* loss, const float* hist_size, int num_users, float* z,
                         const float* xi_dev, float bandwidth, int kind, int only_with_history,
                         cudaStream_t s, long long* launches) {
  user_weights_kernel<<<(num_users + 255) / 256, 256, 0, s>>>(loss, hist_size, num_users, z, xi_dev,
                                                             bandwidth, kind, only_with_history);
  if (launches) ++*launches;
}

size_t xi_partials_doubles(int num_sms) { return (size_t)2 * num_sms * 3 * XI_MAXK; }

int launch_xi_newton(const XiParams& p_in, cudaStream_t s, int num_sms, long long* launches) {
  XiParams p = p_in;
  const int n = p.snr_idx ? p.n_samples : p.num_users;
  int work = n > p.num_users ? n : (p.start_from_mean ? p.num_users : n);
  int grid = (work + 255) / 256;
  if (grid > num_sms) grid = num_sms;
  if (grid < 1) grid = 1;
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)xi_newton_kernel, dim3(grid), dim3(256), args, 0, s);
  if (launches) ++*launches;
  return e == cudaSuccess ? 0 : -1;
}

void launch_exact_quantile(const float* loss, int n, float alpha, float* xi_out, unsigned* hist_ws,
                           cudaStream_t s, long long* launches) {
  (void)hist_ws;
  exact_quantile_kernel<<<1, 1024, 0, s>>>(loss, n, alpha, xi_out);
  if (launches) ++*launches;
}

void launch_weight_means(const float* z, const float* loss, int n, float* out2, double* ws,
                         cudaStream_t s, long long* launches) {
  int nb = (n + 255) / 256;
  if (nb > 128) nb = 128;
  weight_means_partial<<<nb, 256, 0, s>>>(z, loss, n, ws);
  weight_means_final<<<1, 32, 0, s>>>(ws, nb, n, out2);
  if (launches) *launches += 2;
}

void launch_hist_and_item_reg(const int* uptr, int num_users_ds, float* hist_size, const int* iptr,
                              const int* icol, int num_items_ds, float* item_reg, cudaStream_t s,
                              long long* launches) {
  if (num_users_ds > 0) hist_size_kernel<<<(num_users_ds + 255) / 256, 256, 0, s>>>(uptr, num_users_ds, hist_size);
  if (num_items_ds > 0) item_reg_kernel<<<(num_items_ds + 127) / 128, 128, 0, s>>>(iptr, icol, num_items_ds, hist_size, item_reg);
  if (launches) *launches += 2;
}

void launch_norm_weights(const float* z, const float* hist_size, int n, float* out, cudaStream_t s,
                         long long* launches) {
  norm_weights_kernel<<<(n + 255) / 256, 256, 0, s>>>(z, hist_size, n, out);
  if (launches) ++*launches;
}

void launch_reg_sums(const float* X, const int* ptr, int rows, int d, int kind, float reg, float reg_exp, float uw,
                     float alpha, int num_other, const float* item_reg, double* out2, cudaStream_t s, int num_sms,
                     long long* launches) {
  if (rows <= 0) return;
  int g = (rows + 7) / 8;
  if (g > num_sms * 8) g = num_sms * 8;
  reg_sums_kernel<<<g, 256, 0, s>>>(X, ptr, rows, d, kind, reg, reg_exp, uw, alpha, num_other, item_reg, out2);
  if (launches) ++*launches;
}
void launch_dot_sum(const float* a, const float* b, size_t n, double* out, cudaStream_t s, int num_sms, long long* launches) {
  if (n == 0) return;
  size_t g = (n + 255) / 256;
  if (g > (size_t)num_sms * 8) g = (size_t)num_sms * 8;
  dot_sum_kernel<<<(unsigned)g, 256, 0, s>>>(a, b, n, out);
  if (launches) ++*launches;
}
void launch_sum_d(const double* a, size_t n, double* out, cudaStream_t s, int num_sms, long long* launches) {
  if (n == 0) return;
  size_t g = (n + 255) / 256;
  if (g > (size_t)num_sms * 8) g = (size_t)num_sms * 8;
  sum_d_kernel<<<(unsigned)g, 256, 0, s>>>(a, n, out);
  if (launches) ++*launches;
}

void launch_entry_weights(const int* col, const float* w, size_t e_begin, size_t e_end, float* ew, cudaStream_t s,
                          long long* launches) {
  if (e_end <= e_begin) return;
  entry_weights_kernel<<<(unsigned)((e_end - e_begin + 255) / 256), 256, 0, s>>>(col, w, e_begin, e_end, ew);
  if (launches) ++*launches;
}
void launch_absmax(const float* E, size_t n, const float* w, size_t nw, unsigned* out, cudaStream_t s, int num_sms,
                   long long* launches) {
  cudaMemsetAsync(out, 0, 2 * sizeof(unsigned), s);
  size_t g = (n / 4 + 255) / 256;
  if (g > (size_t)num_sms * 8) g = (size_t)num_sms * 8;
  if (g == 0) g = 1;
  absmax_kernel<<<(unsigned)g, 256, 0, s>>>(E, n, w, w ? nw : 0, out);
  if (launches) ++*launches;
}
void launch_sqdiff(const float* a, const float* b, size_t n, double* out, cudaStream_t s, int num_sms, long long* launches) {
  if (n == 0) return;
  size_t g = (n + 255) / 256;
  if (g > (size_t)num_sms * 8) g = (size_t)num_sms * 8;
  sqdiff_kernel<<<(unsigned)g, 256, 0, s>>>(a, b, n, out);
  if (launches) ++*launches;
}

void launch_fill(float* p, size_t n, float v, cudaStream_t s, long long* launches) {
  if (n == 0) return;
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, v);
  if (launches) ++*launches;
}

}  // namespace frx

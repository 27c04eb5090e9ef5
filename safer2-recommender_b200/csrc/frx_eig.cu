// Orthogonal reduction of the d x d Gramian to tridiagonal form, G = H T H^T (d = 128 / 256), by Householder
// reflections in fp64 on one thread-block cluster of 8 CTAs.
//
// Why it exists: the dual-form (Woodbury) row kernel (frx_row_wb.cu) solves a row with n <= 128 history
// entries in the n x n space.  It needs (alpha*G + beta*I)^-1 applied to the row's gathered factor rows for
// per-row alpha, beta.  In the basis H that operator is the inverse of the SPD tridiagonal alpha*T + beta*I,
// whose L D L^T factors cost O(d) per row and apply in O(d) per vector.  The reference solves the same
// d x d system with Eigen::LLT (safer2.h:159-161, ials.h:140-142); the dual form is the same linear system
// by the push-through identity.  (A first version used a Jacobi eigen-solver: its ~4000 cluster-wide rounds
// each move the whole matrix through distributed shared memory, 17 B/clk per SM -> 21 ms; the Householder
// reduction exchanges only two d-vectors per step.)
//
// Layout: column j of the (symmetric, fully stored) working matrix lives in the shared memory of CTA
// j / (d/8); row r of the accumulated H in CTA r / (d/8).  Step k (LAPACK dsytd2, lower):
//   owner of column k   v, tau from A[k+1:, k]; publishes v                          -- cluster barrier --
//   every CTA           reads v (DSMEM), p_j = tau * A[:, j].v for its columns, partial v.p  -- cluster barrier --
//   every CTA           gathers p, w = p - (tau/2)(v.p) v, rank-2 update of its columns A[:, j] -= v w_j + w v_j,
//                       and H[r, :] -= tau (H[r, :].v) v^T for its rows.
// No atomics: every reduction has a fixed order, so all ranks of a multi-GPU job compute identical bits.
// Outputs (fp32): tdiag[d] = diag(T), tsub[d] with tsub[j] = T[j][j-1] (tsub[0] = 0), H and H^T row-major.
#include "frx_kernels.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace frx {

namespace {

constexpr int TRI_CLUSTER = 8;
constexpr int TRI_THREADS = 512;
constexpr int TRI_WARPS = TRI_THREADS / 32;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

template <int D>
__global__ void __cluster_dims__(TRI_CLUSTER, 1, 1) __launch_bounds__(TRI_THREADS, 1)
    sym_tridiag_kernel(const float* __restrict__ G, float* __restrict__ H, float* __restrict__ HT,
                       float* __restrict__ tdiag, float* __restrict__ tsub) {
  constexpr int CPC = D / TRI_CLUSTER;  // columns of A / rows of H per CTA
  constexpr int E = D / 32;             // elements per lane of a d-vector
  extern __shared__ double smem_tri[];
  double* Acol = smem_tri;              // [CPC][D], column lc contiguous
  double* Hrow = Acol + CPC * D;        // [CPC][D], row lr contiguous
  double* vbuf = Hrow + CPC * D;        // [D]   published by the owner of column k
  double* ploc = vbuf + D;              // [CPC] p of this CTA's columns
  double* vloc = ploc + CPC;            // [D]
  double* wloc = vloc + D;              // [D]
  double* red = wloc + D;               // [TRI_WARPS] per-warp partials
  double* scal = red + TRI_WARPS;       // [0] tau (published), [1] partial v.p (published), [2] K, [3] tau local
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int idx = tid; idx < CPC * D; idx += TRI_THREADS) {
    const int lc = idx / D, i = idx % D, c = rank * CPC + lc;
    Acol[idx] = 0.5 * ((double)G[(size_t)i * D + c] + (double)G[(size_t)c * D + i]);
    Hrow[idx] = (i == c) ? 1.0 : 0.0;
  }
  __syncthreads();

  for (int k = 0; k < D - 2; ++k) {
    const int owner = k / CPC, lk = k % CPC;
    if (rank == owner) {
      // Householder vector of A[k+1:, k]
      const double* x = Acol + lk * D;
      double part = 0.0;
      for (int i = k + 2 + tid; i < D; i += TRI_THREADS) part = fma(x[i], x[i], part);
      part = warp_sum_d(part);
      if (lane == 0) red[warp] = part;
      __syncthreads();
      if (tid == 0) {
        double sigma = 0.0;
        for (int wv = 0; wv < TRI_WARPS; ++wv) sigma += red[wv];
        const double alpha = x[k + 1];
        double tau = 0.0, beta = alpha, scale = 0.0;
        if (sigma > 0.0) {
          beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
          tau = (beta - alpha) / beta;
          scale = 1.0 / (alpha - beta);
        }
        scal[0] = tau;
        scal[2] = scale;
        tdiag[k] = (float)x[k];
        tsub[k + 1] = (float)beta;
      }
      __syncthreads();
      const double scale = scal[2];
      for (int i = tid; i < D; i += TRI_THREADS) vbuf[i] = i <= k ? 0.0 : (i == k + 1 ? 1.0 : x[i] * scale);
    }
    cluster.sync();  // A: v and tau are published
    {
      const double* vsrc = cluster.map_shared_rank(vbuf, owner);
      for (int i = tid; i < D; i += TRI_THREADS) vloc[i] = vsrc[i];
      if (tid == 0) scal[3] = *cluster.map_shared_rank(&scal[0], owner);
    }
    __syncthreads();
    const double tau = scal[3];
    // p_j = tau * A[:, j] . v for this CTA's columns (zero for j <= k), partial v.p in a fixed order
    double kp = 0.0;
    for (int lc = warp; lc < CPC; lc += TRI_WARPS) {
      const int j = rank * CPC + lc;
      double s = 0.0;
      if (j > k && tau != 0.0) {
        const double* col = Acol + lc * D;
#pragma unroll
        for (int e = 0; e < E; ++e)
          if (32 * e + 31 > k) s = fma(col[lane + 32 * e], vloc[lane + 32 * e], s);  // v is zero up to k
        s = warp_sum_d(s) * tau;
      }
      if (lane == 0) ploc[lc] = s;
      kp = fma(s, vloc[j], kp);
    }
    if (lane == 0) red[warp] = kp;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int wv = 0; wv < TRI_WARPS; ++wv) s += red[wv];
      scal[1] = s;
    }
    cluster.sync();  // B: every CTA's p and partial v.p are published
    if (tau != 0.0) {  // uniform over the cluster
      for (int j = tid; j < D; j += TRI_THREADS) wloc[j] = *cluster.map_shared_rank(&ploc[j % CPC], j / CPC);
      if (tid == 0) {
        double K = 0.0;
        for (int cta = 0; cta < TRI_CLUSTER; ++cta) K += *cluster.map_shared_rank(&scal[1], cta);
        scal[2] = K;
      }
      __syncthreads();
      const double hk = 0.5 * tau * scal[2];
      for (int j = tid; j < D; j += TRI_THREADS) wloc[j] = fma(-hk, vloc[j], wloc[j]);  // w = p - (tau K / 2) v
      __syncthreads();
      for (int lc = warp; lc < CPC; lc += TRI_WARPS) {
        const int j = rank * CPC + lc;
        if (j > k) {  // A[:, j] -= v w_j + w v_j
          double* col = Acol + lc * D;
          const double wj = wloc[j], vj = vloc[j];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int i = lane + 32 * e;
            if (32 * e + 31 > k) col[i] = fma(-vloc[i], wj, fma(-wloc[i], vj, col[i]));  // v, w are zero up to k
          }
        }
        {  // H[r, :] -= tau (H[r, :] . v) v^T
          double* row = Hrow + lc * D;
          double t = 0.0;
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (32 * e + 31 > k) t = fma(row[lane + 32 * e], vloc[lane + 32 * e], t);
          t = warp_sum_d(t) * tau;
#pragma unroll
          for (int e = 0; e < E; ++e)
            if (32 * e + 31 > k) row[lane + 32 * e] = fma(-t, vloc[lane + 32 * e], row[lane + 32 * e]);
        }
      }
    }
    __syncthreads();
  }
  cluster.sync();  // nobody reads another CTA's shared memory past this point

  // the last 2 x 2 block of T, and H
  for (int lc = tid; lc < CPC; lc += TRI_THREADS) {
    const int j = rank * CPC + lc;
    if (j >= D - 2) tdiag[j] = (float)Acol[lc * D + j];
    if (j == D - 2) tsub[D - 1] = (float)Acol[lc * D + D - 1];
    if (j == 0) tsub[0] = 0.f;
  }
  for (int idx = tid; idx < CPC * D; idx += TRI_THREADS) {
    const int lr = idx / D, j = idx % D, r = rank * CPC + lr;
    const float v = (float)Hrow[idx];
    H[(size_t)r * D + j] = v;
    HT[(size_t)j * D + r] = v;
  }
}

template <int D>
int launch_tridiag_instance(const float* G, float* H, float* HT, float* tdiag, float* tsub, cudaStream_t s) {
  constexpr int CPC = D / TRI_CLUSTER;
  const int smem = (2 * CPC * D + 3 * D + CPC + TRI_WARPS + 8) * (int)sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(sym_tridiag_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return 1;
  sym_tridiag_kernel<D><<<TRI_CLUSTER, TRI_THREADS, smem, s>>>(G, H, HT, tdiag, tsub);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace

bool sym_tridiag_supported(int d) { return d == 128 || d == 256; }

int launch_sym_tridiag(const float* G, int d, float* H, float* HT, float* tdiag, float* tsub, cudaStream_t s,
                       long long* launches) {
  int rc = 1;
  if (d == 256) rc = launch_tridiag_instance<256>(G, H, HT, tdiag, tsub, s);
  else if (d == 128) rc = launch_tridiag_instance<128>(G, H, HT, tdiag, tsub, s);
  if (rc == 0 && launches) ++*launches;
  return rc;
}

}  // namespace frx

// Symmetric eigendecomposition of the d x d Gramian (d = 128 / 256) by cyclic Jacobi rotations, fp64, one
// thread-block cluster of 8 CTAs with the matrices distributed over the cluster's shared memory (DSMEM).
//
// Why it exists: the dual-form (Woodbury) row kernel (frx_row_wb.cu) solves a row with n <= 128 history
// entries in the n x n space.  It needs (alpha*G + beta*I)^-1 for per-row alpha, beta, i.e. G = Q diag(lam)
// Q^T once per half-step; the per-row inverse is then the diagonal 1 / (alpha*lam + beta) in the rotated
// basis.  The reference solves the same d x d system with Eigen::LLT (safer2.h:159-161, ials.h:140-142);
// the dual form is the same linear system by the push-through identity, evaluated in fp32 to ~1e-7.
//
// Method: implicit two-sided Jacobi.  Keep A = G*J and J (J = product of the rotations, starts at I).  For a
// pair (p, q):  g_pp = j_p.a_p, g_qq = j_q.a_q, g_pq = j_p.a_q are the entries of J^T G J; the rotation that
// zeroes g_pq is applied to columns p, q of BOTH A and J (a right multiplication keeps A = G*J).  A round of
// the round-robin tournament has d/2 disjoint pairs: one warp per pair, one cluster barrier per round.
// Column c of A and J lives in the shared memory of CTA c / (d/8); a warp reads / writes its two columns
// through DSMEM.  Converged when a whole sweep applies no rotation (|g_pq| <= tol * max_i G_ii).
// Outputs: lam (Rayleigh quotients j_i.a_i), Q (row-major, column i = eigenvector i) and Q^T, as fp32.
#include "frx_kernels.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace frx {

namespace {

constexpr int EIG_CLUSTER = 8;
constexpr double EIG_TOL = 1e-9;   // off-diagonal threshold relative to max diag(G); fp32 consumers need ~6e-8
constexpr int EIG_MAX_SWEEPS = 40;

template <int D>
__global__ void __cluster_dims__(EIG_CLUSTER, 1, 1) __launch_bounds__((D / 2 / EIG_CLUSTER) * 32, 1)
    jacobi_eig_kernel(const float* __restrict__ G, float* __restrict__ Q, float* __restrict__ QT,
                      float* __restrict__ lam, int* __restrict__ info) {
  constexpr int CPC = D / EIG_CLUSTER;        // columns per CTA
  constexpr int WARPS = D / 2 / EIG_CLUSTER;  // pairs (= warps) per CTA
  constexpr int E = D / 64;                   // double2 elements per lane per column
  extern __shared__ double smem_eig[];
  double* Acol = smem_eig;             // [CPC][D]
  double* Jcol = smem_eig + CPC * D;   // [CPC][D]
  __shared__ int rotated_flag;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // A = (G + G^T) / 2 in fp64, J = I; this CTA's columns
  double gmax = 0.0;
  for (int i = lane; i < D; i += 32) gmax = fmax(gmax, fabs((double)G[(size_t)i * D + i]));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
  for (int idx = tid; idx < CPC * D; idx += WARPS * 32) {
    const int lc = idx / D, i = idx % D, c = rank * CPC + lc;
    Acol[idx] = 0.5 * ((double)G[(size_t)i * D + c] + (double)G[(size_t)c * D + i]);
    Jcol[idx] = (i == c) ? 1.0 : 0.0;
  }
  if (tid == 0) rotated_flag = 0;
  cluster.sync();

  const double thr = EIG_TOL * gmax;
  const int k = rank * WARPS + warp;  // this warp's pair index within a round
  int sweeps = 0;
  bool converged = (gmax == 0.0);
  while (!converged && sweeps < EIG_MAX_SWEEPS) {
    bool rotated = false;
    for (int r = 0; r < D - 1; ++r) {
      int p, q;
      if (k == 0) { p = D - 1; q = r; }
      else { p = (r + k) % (D - 1); q = (r - k + (D - 1)) % (D - 1); }
      double* ap_ptr = cluster.map_shared_rank(Acol + (p % CPC) * D, p / CPC);
      double* aq_ptr = cluster.map_shared_rank(Acol + (q % CPC) * D, q / CPC);
      double* jp_ptr = cluster.map_shared_rank(Jcol + (p % CPC) * D, p / CPC);
      double* jq_ptr = cluster.map_shared_rank(Jcol + (q % CPC) * D, q / CPC);
      double2 ap[E], aq[E], jp[E], jq[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int off = 2 * lane + 64 * e;
        ap[e] = *reinterpret_cast<const double2*>(ap_ptr + off);
        aq[e] = *reinterpret_cast<const double2*>(aq_ptr + off);
        jp[e] = *reinterpret_cast<const double2*>(jp_ptr + off);
        jq[e] = *reinterpret_cast<const double2*>(jq_ptr + off);
      }
      double gpp = 0.0, gqq = 0.0, gpq = 0.0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        gpp = fma(jp[e].x, ap[e].x, fma(jp[e].y, ap[e].y, gpp));
        gqq = fma(jq[e].x, aq[e].x, fma(jq[e].y, aq[e].y, gqq));
        gpq = fma(jp[e].x, aq[e].x, fma(jp[e].y, aq[e].y, gpq));
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        gpp += __shfl_xor_sync(0xffffffffu, gpp, off);
        gqq += __shfl_xor_sync(0xffffffffu, gqq, off);
        gpq += __shfl_xor_sync(0xffffffffu, gpq, off);
      }
      if (fabs(gpq) > thr) {  // warp-uniform: every lane holds the same sums
        rotated = true;
        const double zeta = (gqq - gpp) / (2.0 * gpq);
        const double t = (zeta == 0.0) ? 1.0 : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int off = 2 * lane + 64 * e;
          double2 np, nq;
          np.x = c * ap[e].x - s * aq[e].x; np.y = c * ap[e].y - s * aq[e].y;
          nq.x = s * ap[e].x + c * aq[e].x; nq.y = s * ap[e].y + c * aq[e].y;
          *reinterpret_cast<double2*>(ap_ptr + off) = np;
          *reinterpret_cast<double2*>(aq_ptr + off) = nq;
          np.x = c * jp[e].x - s * jq[e].x; np.y = c * jp[e].y - s * jq[e].y;
          nq.x = s * jp[e].x + c * jq[e].x; nq.y = s * jp[e].y + c * jq[e].y;
          *reinterpret_cast<double2*>(jp_ptr + off) = np;
          *reinterpret_cast<double2*>(jq_ptr + off) = nq;
        }
      }
      cluster.sync();  // the next round pairs the columns differently
    }
    ++sweeps;
    if (rotated && lane == 0) atomicOr(&rotated_flag, 1);
    cluster.sync();
    int any = 0;
    for (int cta = 0; cta < EIG_CLUSTER; ++cta) any |= *cluster.map_shared_rank(&rotated_flag, cta);
    cluster.sync();  // everyone has read the flags
    if (tid == 0) rotated_flag = 0;  // ordered before the next sweep's atomicOr by its D-1 round barriers
    converged = (any == 0);
  }

  // eigenvalues and vectors of this CTA's columns
  for (int lc = warp; lc < CPC; lc += WARPS) {
    const int c = rank * CPC + lc;
    double ray = 0.0;
    for (int i = lane; i < D; i += 32) ray = fma(Jcol[lc * D + i], Acol[lc * D + i], ray);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ray += __shfl_xor_sync(0xffffffffu, ray, off);
    if (lane == 0) lam[c] = (float)ray;
    for (int i = lane; i < D; i += 32) {
      const float v = (float)Jcol[lc * D + i];
      QT[(size_t)c * D + i] = v;  // row c of Q^T = eigenvector c
      Q[(size_t)i * D + c] = v;
    }
  }
  if (info && rank == 0 && tid == 0) info[0] = converged ? sweeps : -sweeps;
}

template <int D>
int launch_eig_instance(const float* G, float* Q, float* QT, float* lam, int* info, cudaStream_t s) {
  constexpr int CPC = D / EIG_CLUSTER;
  const int smem = 2 * CPC * D * (int)sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(jacobi_eig_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return 1;
  jacobi_eig_kernel<D><<<EIG_CLUSTER, (D / 2 / EIG_CLUSTER) * 32, smem, s>>>(G, Q, QT, lam, info);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace

bool sym_eig_supported(int d) { return d == 128 || d == 256; }

int launch_sym_eig(const float* G, int d, float* Q, float* QT, float* lam, int* info, cudaStream_t s,
                   long long* launches) {
  int rc = 1;
  if (d == 256) rc = launch_eig_instance<256>(G, Q, QT, lam, info, s);
  else if (d == 128) rc = launch_eig_instance<128>(G, Q, QT, lam, info, s);
  if (rc == 0 && launches) ++*launches;
  return rc;
}

}  // namespace frx

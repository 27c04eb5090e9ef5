// Generic fused row kernel: CSR gather -> weighted outer-product accumulation
// (SYRK, lower triangle) -> regularised d x d system -> in-shared-memory
// Cholesky / gradient step / block-subspace update, one CTA per row.
//
// This is the any-dimension SIMT path (tests use d=8 and block_size=4; the
// ML-1M configs d=32).  It restates, per RowMode, the reference projections
// cited in frx_kernels.cuh.  Data layout in shared memory: the lower triangle
// of the (dp+1) x (dp+1) augmented system [M; rhs^T] packed row by row
// (row i starts at i(i+1)/2); the rhs lives in row dp so that the Cholesky
// sweep performs the forward substitution L y = rhs for free.
#include "frx_kernels.cuh"
#include <cfloat>

namespace frx {

namespace {
constexpr int NT_MAX = 1024;  // max threads per CTA

constexpr int CH = 32;    // history entries staged per chunk

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

struct SmemLayout {
  int dp, chunk_off, qr_off, colk_off, ldiag_off, xv_off, sol_off, cg_off, mtx_off, total_no_mtx, ntri;
  __host__ __device__ SmemLayout(int bd, int d, int solver = 0) {
    dp = (bd + 3) & ~3;
    int dx = (d + 3) & ~3;
    chunk_off = 0;
    // the gather chunk [CH][dp]; the blocked Cholesky of the global-scratch case reuses it as its column-major
    // panel [32][dp + 8] (CH == 32)
    qr_off = chunk_off + CH * (dp + 8);
    colk_off = qr_off + CH;
    ldiag_off = colk_off + dp + 4;
    xv_off = ldiag_off + dp;
    sol_off = xv_off + dx;
    cg_off = sol_off + dp;                          // iterative solvers: 8 work vectors + reduction scratch
    mtx_off = cg_off + (solver ? 8 * dp + 32 : 0);
    total_no_mtx = mtx_off;
    ntri = ((dp + 1) * (dp + 2)) >> 1;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool VEC4>
__global__ void __launch_bounds__(NT_MAX) row_solve_generic_kernel(RowParams p) {
  const int NT = blockDim.x, NW = NT >> 5;
  extern __shared__ __align__(16) float smem[];
  const int bd = p.bd, d = p.d, cs = p.cs, mode = p.mode;
  const SmemLayout L(bd, d, p.solver);
  const int dp = L.dp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* chunk = smem + L.chunk_off;
  float* qr = smem + L.qr_off;
  float* colk = smem + L.colk_off;
  float* ldiag = smem + L.ldiag_off;
  float* xv = smem + L.xv_off;
  float* sol = smem + L.sol_off;
  float* Mtx = p.use_smem_matrix ? smem + L.mtx_off : p.scratch + (size_t)blockIdx.x * p.scratch_stride;
  float* rhs = Mtx + tri(dp);  // row dp of the packed augmented system
  const int T = dp >> 2;
  const int ntiles = (T * (T + 1)) >> 1;
  const bool item_side = (mode == RM_SAFER_V || mode == RM_CVAR_V || mode == RM_PP_SAFER_V);
  const bool stale_tail = (mode == RM_SAFER_V || mode == RM_CVAR_V);  // B-1 (not in safer2pp.h:203-207)
  const bool pp = (mode >= RM_PP_IALS);
  const bool need_x = (mode >= RM_CVAR_U);

  for (int ri = blockIdx.x; ri < p.num_rows; ri += gridDim.x) {
    const int r = p.order[ri];
    const int beg = p.ptr[r];
    const int n = p.ptr[r + 1] - beg;
    const int xr = p.xmap ? p.xmap[r] : r;
    __syncthreads();  // previous row fully consumed
    for (int i = tid; i < L.ntri; i += NT) Mtx[i] = 0.f;
    if (need_x)
      for (int k = tid; k < d; k += NT) xv[k] = p.Xread[(size_t)xr * d + k];
    int dup_lo = 0, dup_hi = 0;
    if (stale_tail && n > 128 && (n & 127) != 0) {
      const int kf = n >> 7;
      dup_lo = 128 * (kf - 1) + (n & 127);
      dup_hi = 128 * kf;
    }
    __syncthreads();

    // ---- gather + accumulate -------------------------------------------------
    for (int c0 = 0; c0 < n; c0 += CH) {
      const int ne = min(CH, n - c0);
      for (int e = warp; e < ne; e += NW) {
        const int t = beg + c0 + e;
        const int c = p.col[t];
        float s = 1.f, q = 1.f;
        if (item_side) { const float w = p.entry_w[c]; s = w; q = w; }
        if (pp) {
          const float residual = p.pred[p.tup[t]] - 1.0f;
          q = item_side ? residual * s : residual;
        }
        const int j = c0 + e;
        if (j >= dup_lo && j < dup_hi) s *= 2.f;  // stale columns added twice (B-1)
        const float sq = sqrtf(s);
        if (lane == 0) qr[e] = s > 0.f ? q / sq : 0.f;
        const float* src = p.E + (size_t)c * d + cs;
        float* dst = chunk + e * dp;
        if (VEC4) {
          for (int k = lane * 4; k < dp; k += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < bd) v = __ldg(reinterpret_cast<const float4*>(src + k));
            v.x *= sq; v.y *= sq; v.z *= sq; v.w *= sq;
            *reinterpret_cast<float4*>(dst + k) = v;
          }
        } else {
          for (int k = lane; k < dp; k += 32) dst[k] = (k < bd) ? __ldg(src + k) * sq : 0.f;
        }
      }
      __syncthreads();
      for (int k = tid; k < dp; k += NT) {
        float a = 0.f;
        for (int e = 0; e < ne; ++e) a += qr[e] * chunk[e * dp + k];
        rhs[k] += a;
      }
      for (int t = tid; t < ntiles; t += NT) {
        int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while (tri(ti + 1) <= t) ++ti;
        while (tri(ti) > t) --ti;
        const int tj = t - tri(ti);
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        const float* ca = chunk + 4 * ti;
        const float* cb = chunk + 4 * tj;
        for (int e = 0; e < ne; ++e) {
          const float4 a4 = *reinterpret_cast<const float4*>(ca + e * dp);
          const float4 b4 = *reinterpret_cast<const float4*>(cb + e * dp);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
          const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = 4 * ti + a;
          float* mrow = Mtx + tri(i) + 4 * tj;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (4 * tj + b <= i) mrow[b] += acc[a][b];
        }
      }
      __syncthreads();
    }

    // ---- per-row scalars --------------------------------------------------------
    float reg, weight = 1.f;
    if (mode == RM_IALS || mode == RM_PP_IALS) {
      // ials.h:310-315: reg * pow(n + uw * num_choices, reg_exp) (double pow, float result)
      reg = (float)((double)p.reg * pow((double)((float)n + p.uw * (float)p.num_other), (double)p.reg_exp));
    } else if (item_side) {
      // safer2.h:426-432
      reg = p.reg * (p.item_reg[r] + p.alpha * p.uw * (float)p.num_users_total);
    } else {
      reg = p.reg * (1.f + p.uw * (float)p.num_other);  // safer2.h:418-421
      if (mode == RM_CVAR_U) weight = p.stepsize;       // B-4: weight := stepsize_
      else weight = p.row_w ? p.row_w[r] : 1.f;
    }
    const bool user_form = (mode == RM_SAFER_U || mode == RM_PP_SAFER_U || mode == RM_CVAR_U);
    const float nf = (float)n;

    // ---- assemble the lower triangle of the system in place -----------------------
    for (int i = warp; i < dp; i += NW) {
      float* mrow = Mtx + tri(i);
      for (int j = lane; j <= i; j += 32) {
        float m;
        if (i >= bd) {
          m = (i == j) ? 1.f : 0.f;
        } else {
          const float g = __ldg(p.G + (size_t)(cs + i) * d + cs + j);
          const float sij = mrow[j];
          if (user_form) {  // safer2.h:143-150: /= n; += uw*G; *= weight; diag += reg
            m = sij / nf;
            m += p.uw * g;
            m *= weight;
            if (i == j) m += reg;
          } else if (mode == RM_IALS || mode == RM_PP_IALS || mode == RM_PP_SAFER_V) {
            m = p.uw * g;  // ials.h:101-105: uw*G, diag += reg, then rank updates
            if (i == j) m += reg;
            m += sij;
          } else {  // safer2.h:176,206-208: uw*G + rank updates, then diag += reg
            m = p.uw * g + sij;
            if (i == j) m += reg;
          }
        }
        mrow[j] = m;
      }
    }
    if (user_form) {
      const float sc = weight / nf;  // rhs *= weight / history_size
      for (int k = tid; k < dp; k += NT) rhs[k] *= sc;
    }
    __syncthreads();
    if (pp) {
      // rhs += uw * G_{B,:} x (* weight) ; rhs += reg * x_B   (ialspp.h:135-136, safer2pp.h:147-149,209-210)
      for (int i = warp; i < bd; i += NW) {
        const float* grow = p.G + (size_t)(cs + i) * d;
        float a = 0.f;
        for (int j = lane; j < d; j += 32) a = fmaf(__ldg(grow + j), xv[j], a);
        a = warp_sum(a);
        if (lane == 0) {
          float add = p.uw * a;
          if (mode == RM_PP_SAFER_U) add *= weight;
          rhs[i] += add;
          rhs[i] += reg * xv[cs + i];
        }
      }
      __syncthreads();
    }

    if (mode == RM_CVAR_U || mode == RM_CVAR_V) {
      // x - step * (matrix * x - rhs) with the half-updated full matrix (B-3).
      const float step = (mode == RM_CVAR_U) ? (p.row_w ? p.row_w[r] : 1.f) : p.stepsize;  // B-4
      const float gscale = (mode == RM_CVAR_U) ? weight * p.uw : p.uw;
      for (int i = warp; i < bd; i += NW) {
        const float* mrow = Mtx + tri(i);
        const float* grow = p.G + (size_t)i * d;
        float a = 0.f;
        for (int j = lane; j < bd; j += 32) {
          const float mij = (j <= i) ? mrow[j] : gscale * __ldg(grow + j);
          a = fmaf(mij, xv[j], a);
        }
        a = warp_sum(a);
        if (lane == 0) p.X[(size_t)xr * d + i] = xv[i] - step * (a - rhs[i]);
      }
      continue;
    }

    if (p.solver != 0) {
      // ---- --use_cg: the reference's iterative solvers instead of LLT (safer2.h:152-157, ials.h:133-138:
      // Eigen::ConjugateGradient<MatrixXf, Lower>; erm_mf.h:139-145, 198-204: Eigen::BiCGSTAB on the FULL matrix,
      // whose strict upper triangle holds only the Gramian term -- SURVEY.md B-5).  Both with the diagonal
      // preconditioner, x0 = 0 and Eigen's stopping rule |r|^2 < tol^2 |rhs|^2 or max_iterations. ----
      float* cgv = smem + L.cg_off;
      float* x_ = sol;
      float* r_ = cgv;            float* p_ = cgv + dp;      float* z_ = cgv + 2 * dp;  float* t_ = cgv + 3 * dp;
      float* r0_ = cgv + 4 * dp;  float* v_ = cgv + 5 * dp;  float* y_ = cgv + 6 * dp;  float* s_ = cgv + 7 * dp;
      float* red = cgv + 8 * dp;
      const bool full = (p.solver == 2);
      // strict upper triangle of the reference's full matrix: (weight *) uw * G only (B-2)
      const float gscale = user_form ? weight * p.uw : p.uw;
      auto bdot = [&](const float* a, const float* b) -> float {
        float acc = 0.f;
        for (int k = tid; k < bd; k += NT) acc = fmaf(a[k], b[k], acc);
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        float tot = 0.f;
        for (int wv = 0; wv < NW; ++wv) tot += red[wv];
        return tot;
      };
      auto matvec = [&](const float* in, float* out) {  // callers synchronise before and after
        for (int i = warp; i < bd; i += NW) {
          const float* mrow = Mtx + tri(i);
          float a = 0.f;
          for (int j = lane; j <= i; j += 32) a = fmaf(mrow[j], in[j], a);
          if (full) {
            const float* grow = p.G + (size_t)(cs + i) * d + cs;
            for (int j = i + 1 + lane; j < bd; j += 32) a = fmaf(gscale * __ldg(grow + j), in[j], a);
          } else {  // selfadjointView<Lower>
            for (int j = i + 1 + lane; j < bd; j += 32) a = fmaf(Mtx[tri(j) + i], in[j], a);
          }
          a = warp_sum(a);
          if (lane == 0) out[i] = a;
        }
      };
      auto dinv = [&](int k) -> float { const float dg = Mtx[tri(k) + k]; return dg != 0.f ? 1.f / dg : 1.f; };
      for (int k = tid; k < bd; k += NT) { x_[k] = 0.f; r_[k] = rhs[k]; r0_[k] = rhs[k]; v_[k] = 0.f; p_[k] = 0.f; }
      __syncthreads();
      const float rhs2 = bdot(rhs, rhs);
      const float tol2 = p.cg_tol * p.cg_tol * rhs2;
      if (rhs2 != 0.f && !full) {
        const float threshold = fmaxf(tol2, FLT_MIN);
        if (!(rhs2 < threshold)) {
          for (int k = tid; k < bd; k += NT) p_[k] = dinv(k) * r_[k];
          __syncthreads();
          float abs_new = bdot(r_, p_);
          for (int it = 0; it < p.cg_max_it; ++it) {
            __syncthreads();
            matvec(p_, t_);
            __syncthreads();
            const float alpha = abs_new / bdot(p_, t_);
            for (int k = tid; k < bd; k += NT) { x_[k] = fmaf(alpha, p_[k], x_[k]); r_[k] = fmaf(-alpha, t_[k], r_[k]); }
            __syncthreads();
            const float res2 = bdot(r_, r_);
            if (res2 < threshold) break;
            for (int k = tid; k < bd; k += NT) z_[k] = dinv(k) * r_[k];
            __syncthreads();
            const float abs_old = abs_new;
            abs_new = bdot(r_, z_);
            const float beta = abs_new / abs_old;
            for (int k = tid; k < bd; k += NT) p_[k] = fmaf(beta, p_[k], z_[k]);
          }
        }
      } else if (rhs2 != 0.f) {
        float r0_sq = rhs2, rho = 1.f, alpha = 1.f, w = 1.f;
        const float eps2 = FLT_EPSILON * FLT_EPSILON;
        int it = 0, restarts = 0;
        while (it < p.cg_max_it) {
          if (!(bdot(r_, r_) > tol2)) break;
          const float rho_old = rho;
          rho = bdot(r0_, r_);
          if (fabsf(rho) < eps2 * r0_sq) {  // r became orthogonal to r0: restart from the true residual
            __syncthreads();
            matvec(x_, t_);
            __syncthreads();
            for (int k = tid; k < bd; k += NT) { const float rr = rhs[k] - t_[k]; r_[k] = rr; r0_[k] = rr; }
            __syncthreads();
            rho = r0_sq = bdot(r_, r_);
            if (restarts++ == 0) it = 0;
          }
          const float beta = (rho / rho_old) * (alpha / w);
          for (int k = tid; k < bd; k += NT) {
            const float pk = r_[k] + beta * (p_[k] - w * v_[k]);
            p_[k] = pk;
            y_[k] = dinv(k) * pk;
          }
          __syncthreads();
          matvec(y_, v_);
          __syncthreads();
          alpha = rho / bdot(r0_, v_);
          for (int k = tid; k < bd; k += NT) {
            const float sk = r_[k] - alpha * v_[k];
            s_[k] = sk;
            z_[k] = dinv(k) * sk;
          }
          __syncthreads();
          matvec(z_, t_);
          __syncthreads();
          const float tt = bdot(t_, t_);
          w = tt > 0.f ? bdot(t_, s_) / tt : 0.f;
          for (int k = tid; k < bd; k += NT) {
            x_[k] += alpha * y_[k] + w * z_[k];
            r_[k] = s_[k] - w * t_[k];
          }
          __syncthreads();
          ++it;
        }
      }
      __syncthreads();
      for (int k = tid; k < bd; k += NT) p.X[(size_t)xr * d + k] = x_[k];
      continue;
    }

    // ---- Cholesky of the augmented system (rows 0..dp; row dp carries rhs -> y) -----
    if (!p.use_smem_matrix) {
      // The system lives in global scratch (d = 512: 525 KB): blocked right-looking factorisation, 32-wide panels
      // staged in shared memory (column-major: Pt[c * PS + (i - p0)]), so that every trailing element is read and
      // written once per PANEL with 32 FMAs in between instead of once per column (measured at the MSD shape,
      // d = 512: the unblocked sweep below moved ~360 MB through L2 per row and took 9.6 ms per row).
      float* Pt = chunk;
      const int PS = dp + 8;
      for (int p0 = 0; p0 < dp; p0 += 32) {
        const int nb = min(32, dp - p0);
        const int nrows = dp + 1 - p0;  // panel rows p0 .. dp (row dp: the rhs)
        for (int idx = tid; idx < nrows * nb; idx += NT) {
          const int ii = idx / nb, c = idx - ii * nb;
          const int i = p0 + ii;
          Pt[c * PS + ii] = (p0 + c <= i) ? Mtx[tri(i) + p0 + c] : 0.f;
        }
        __syncthreads();
        if (warp == 0) {  // diagonal block in registers: lane = row, column k broadcast by shuffles
          float a[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) a[c] = (c < nb && lane < nb) ? Pt[c * PS + lane] : (c == lane ? 1.f : 0.f);
          bool bad = false;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            float pivot = __shfl_sync(0xffffffffu, a[k], k);
            if (!(pivot > 0.f)) { bad = bad || k < nb; pivot = 1.f; }
            const float l = sqrtf(pivot);
            const float lk = lane == k ? l : a[k] / l;  // L[lane][k] (lanes >= k)
            a[k] = lk;
#pragma unroll
            for (int j = k + 1; j < 32; ++j) {
              const float ljk = __shfl_sync(0xffffffffu, lk, j);
              if (lane >= j) a[j] = fmaf(-lk, ljk, a[j]);
            }
            if (lane == k && k < nb) ldiag[p0 + k] = l;
          }
          if (bad && lane == 0) atomicExch(p.status, 1);
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < nb && lane < nb && c <= lane) Pt[c * PS + lane] = a[c];
        }
        __syncthreads();
        for (int ii = nb + tid; ii < nrows; ii += NT) {  // rows below: L21 row = A21 row * L11^-T (thread = row)
          float x[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) x[c] = c < nb ? Pt[c * PS + ii] : 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if (c < nb) {
              x[c] = x[c] / Pt[c * PS + c];
#pragma unroll
              for (int m2 = c + 1; m2 < 32; ++m2)
                if (m2 < nb) x[m2] = fmaf(-x[c], Pt[c * PS + m2], x[m2]);
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < nb) Pt[c * PS + ii] = x[c];
        }
        __syncthreads();
        for (int idx = tid; idx < nrows * nb; idx += NT) {  // the factor goes back to the packed matrix
          const int ii = idx / nb, c = idx - ii * nb;
          const int i = p0 + ii;
          if (p0 + c <= i && !(ii < nb && c == ii)) Mtx[tri(i) + p0 + c] = Pt[c * PS + ii];
        }
        const int p1 = p0 + nb, rem = dp - p1;  // trailing rows / columns p1 .. dp-1 (a multiple of 4), then the rhs row
        const int Tn = rem >> 2, ntl = (Tn * (Tn + 1)) >> 1;
        for (int t = tid; t < ntl; t += NT) {
          int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
          while (tri(ti + 1) <= t) ++ti;
          while (tri(ti) > t) --ti;
          const int tj = t - tri(ti);
          float acc[4][4], old[4][4];
          // the old values are requested first: their L2 latency passes under the 32 panel columns below
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const float* mrow = Mtx + tri(p1 + 4 * ti + a) + p1 + 4 * tj;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              acc[a][b] = 0.f;
              old[a][b] = (4 * tj + b <= 4 * ti + a) ? mrow[b] : 0.f;
            }
          }
          const float* pa = Pt + nb + 4 * ti;
          const float* pb = Pt + nb + 4 * tj;
          for (int c = 0; c < nb; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(pa + c * PS);
            const float4 b4 = *reinterpret_cast<const float4*>(pb + c * PS);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
          }
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int i = p1 + 4 * ti + a;
            float* mrow = Mtx + tri(i) + p1 + 4 * tj;
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (4 * tj + b <= 4 * ti + a) mrow[b] = old[a][b] - acc[a][b];
          }
        }
        for (int j = tid; j < rem; j += NT) {  // rhs row (row dp of the augmented system)
          float a = 0.f;
          for (int c = 0; c < nb; ++c) a = fmaf(Pt[c * PS + (dp - p0)], Pt[c * PS + nb + j], a);
          rhs[p1 + j] -= a;
        }
        __syncthreads();
      }
    } else
    for (int k = 0; k < dp; ++k) {
      float pivot = Mtx[tri(k) + k];
      if (!(pivot > 0.f)) {
        if (tid == 0) atomicExch(p.status, 1);
        pivot = 1.f;
      }
      const float l = sqrtf(pivot);
      for (int i = k + 1 + tid; i <= dp; i += NT) {
        const float v = Mtx[tri(i) + k] / l;
        Mtx[tri(i) + k] = v;
        colk[i] = v;
      }
      if (tid == 0) ldiag[k] = l;
      __syncthreads();
      for (int i = k + 1 + warp; i <= dp; i += NW) {
        const float lik = colk[i];
        float* mrow = Mtx + tri(i);
        const int jmax = min(i, dp - 1);
        for (int j = k + 1 + lane; j <= jmax; j += 32) mrow[j] = fmaf(-lik, colk[j], mrow[j]);
      }
      __syncthreads();
    }
    // ---- back substitution L^T x = y ---------------------------------------
    if (!p.use_smem_matrix) {
      // blocked like the factorisation: the 32 rows of a panel are staged in shared memory by all threads (one
      // global latency per panel instead of one per row), warp 0 solves the 32 x 32 triangle, then every thread
      // applies the panel's 32 solved unknowns to one earlier component
      float* S = chunk;  // [32][PS] row-major: S[(k - p0) * PS + j] = L[k][j], j <= k
      const int PS = dp + 8;
      for (int j = tid; j < dp; j += NT) sol[j] = rhs[j];
      __syncthreads();
      for (int p0 = ((dp - 1) / 32) * 32; p0 >= 0; p0 -= 32) {
        const int nb = min(32, dp - p0);
        for (int kk = warp; kk < nb; kk += NW) {
          const float* mrow = Mtx + tri(p0 + kk);
          for (int j = lane; j < p0 + kk; j += 32) S[kk * PS + j] = mrow[j];
        }
        __syncthreads();
        if (warp == 0) {
          for (int kk = nb - 1; kk >= 0; --kk) {
            const float xk = sol[p0 + kk] / ldiag[p0 + kk];
            __syncwarp();
            if (lane == kk) sol[p0 + kk] = xk;
            else if (lane < kk) sol[p0 + lane] = fmaf(-S[kk * PS + p0 + lane], xk, sol[p0 + lane]);
            __syncwarp();
          }
        }
        __syncthreads();
        for (int j = tid; j < p0; j += NT) {
          float a = 0.f;
          for (int kk = 0; kk < nb; ++kk) a = fmaf(S[kk * PS + j], sol[p0 + kk], a);
          sol[j] -= a;
        }
        __syncthreads();
      }
    } else if (warp == 0) {
      for (int j = lane; j < dp; j += 32) sol[j] = rhs[j];
      __syncwarp();
      for (int k = dp - 1; k >= 0; --k) {
        const float xk = sol[k] / ldiag[k];
        __syncwarp();
        if (lane == 0) sol[k] = xk;
        const float* mrow = Mtx + tri(k);
        for (int j = lane; j < k; j += 32) sol[j] = fmaf(-mrow[j], xk, sol[j]);
        __syncwarp();
      }
    }
    __syncthreads();

    if (!pp) {
      for (int k = tid; k < bd; k += NT) p.X[(size_t)xr * d + k] = sol[k];
    } else {
      // new = x_B - solve; delta = new - x_B; pred[t] += delta . e_B  (ialspp.h:392-406)
      for (int k = tid; k < bd; k += NT) {
        const float old = xv[cs + k];
        const float nv = old - sol[k];
        p.X[(size_t)xr * d + cs + k] = nv;
        colk[k] = nv - old;
      }
      __syncthreads();
      for (int e = warp; e < n; e += NW) {
        const int t = beg + e;
        const float* src = p.E + (size_t)p.col[t] * d + cs;
        float a = 0.f;
        for (int k = lane; k < bd; k += 32) a = fmaf(colk[k], __ldg(src + k), a);
        a = warp_sum(a);
        if (lane == 0) p.pred[p.tup[t]] += a;
      }
    }
  }
}

size_t smem_bytes_for(int bd, int d, bool with_matrix, int solver = 0) {
  SmemLayout L(bd, d, solver);
  return sizeof(float) * (size_t)(L.total_no_mtx + (with_matrix ? L.ntri : 0));
}
constexpr size_t kMaxSmem = 220 * 1024;
}  // namespace

size_t row_solve_generic_scratch_floats(int bd, int solver) {
  SmemLayout L(bd, bd, solver);
  if (smem_bytes_for(bd, bd, true, solver) <= kMaxSmem) return 0;
  return (size_t)((L.ntri + 31) & ~31);
}

int row_solve_generic_grid(int num_rows, int num_sms) {
  // Callers size scratch with this; the launch uses at most this many CTAs.
  int g = num_sms * 4;
  return num_rows < g ? (num_rows > 0 ? num_rows : 1) : g;
}

void launch_row_solve_generic(const RowParams& p_in, cudaStream_t s, int num_sms, long long* launches) {
  if (p_in.num_rows <= 0) return;
  RowParams p = p_in;
  const bool fits = smem_bytes_for(p.bd, p.d, true, p.solver) <= kMaxSmem;
  p.use_smem_matrix = fits ? 1 : 0;
  const size_t smem = smem_bytes_for(p.bd, p.d, fits, p.solver);
  int per_sm = (int)((size_t)(224 * 1024) / (smem + 1024));
  if (p.bd >= 96) per_sm = per_sm > 2 ? 2 : per_sm;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = num_sms * per_sm;
  if (grid > p.num_rows) grid = p.num_rows;
  const bool vec4 = (p.d % 4 == 0) && (p.cs % 4 == 0) && (p.bd % 4 == 0);
  const int NT = p.bd >= 96 ? 1024 : 256;  // large systems are latency-bound: more warps per SM
  if (vec4) {
    cudaFuncSetAttribute(row_solve_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    row_solve_generic_kernel<true><<<grid, NT, smem, s>>>(p);
  } else {
    cudaFuncSetAttribute(row_solve_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    row_solve_generic_kernel<false><<<grid, NT, smem, s>>>(p);
  }
  if (launches) ++*launches;
}

}  // namespace frx

// Fused row kernel on 5th-generation tensor cores (tcgen05 + TMEM), D = 128 / 256.
//
// One CTA per row r of the side being solved:
//   phase A  A_r = sum_c s_c e_c e_c^T  (+ rhs = sum_c q_c e_c)
//            16 loader warps gather the history rows (lane = history entry, 128 B per lane
//            per load), split each fp32 into tf32 hi + lo and store them TRANSPOSED into
//            K-major, 128B-swizzled operand tiles [feature][32 entries]; one thread issues
//            tcgen05.mma kind::tf32 for hi*hi + hi*lo + lo*hi (error-compensated 3xTF32:
//            fp32-level accuracy) into TMEM accumulators; lower-triangle tiles only.
//            Two operand stages, mbarrier full/empty pipeline, tcgen05.commit frees a stage.
//   phase B  TMEM -> packed lower triangle in shared memory -> M = a*A + b*G + reg*I ->
//            Cholesky of the augmented system [M; rhs^T] -> back substitution -> row of X.
//
// Hardware conventions were established with tools/tc_probe.cu on a B200:
//  * kind::tf32 works with K-major operands (SWIZZLE_128B, SBO = 1024 B, K step = +32 B on
//    the descriptor start address); MN-major tf32 operands produce zeros, hence the transpose.
//  * the MMA ignores the low 13 mantissa bits of fp32 inputs (truncation).
//  * tcgen05.ld 32x32b: warp w reads TMEM lanes 32*(w%4)..+31, thread = accumulator row.
// Restates ials.h:88-144, safer2.h:104-163 and safer2.h:166-221 (incl. the stale-tail quirk).
#include "frx_kernels.cuh"
#include <cstdint>

namespace frx {

namespace {

constexpr int TC_LOADER_WARPS = 16;                 // 2 groups of 8
constexpr int TC_THREADS = (TC_LOADER_WARPS + 1) * 32;  // + 1 MMA-issuing warp
constexpr int KT = 32;                              // history entries per operand tile (128 B rows)

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B smem descriptor: LBO unused (encoded 1), SBO = 1024 B, version 1.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, K-major A and B, M = 128.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums over the 32 lanes of C per-lane values; on return v[0] of lane l holds the sum
// of column (C == 32 ? l : l >> 1) (for C == 16 both lanes of a pair hold it).
template <int C>
__device__ __forceinline__ void transpose_reduce(float (&v)[C], int lane) {
  int cnt = C / 2;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
    if (cnt >= 1) {
#pragma unroll
      for (int i = 0; i < C / 2; ++i) {
        if (i < cnt) {
          const float send = upper ? v[i] : v[i + cnt];
          const float recv = __shfl_xor_sync(0xffffffffu, send, off);
          v[i] = (upper ? v[i + cnt] : v[i]) + recv;
        }
      }
      cnt >>= 1;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}

template <int D>
struct TcLayout {
  static constexpr int kTileBytes = D * 128;             // one operand tile: D rows x 128 B
  static constexpr int kStageBytes = 2 * kTileBytes;     // hi + lo
  static constexpr int kStagingBytes = 2 * kStageBytes;  // two stages
  static constexpr int kTriFloats = ((D + 1) * (D + 2)) / 2;
  static constexpr int kMtxBytes = ((kTriFloats * 4 + 1023) / 1024) * 1024;
  static constexpr int kBigBytes = kStagingBytes > kMtxBytes ? kStagingBytes : kMtxBytes;  // aliased region
  // after the aliased region: rhs partials [2][D], colk [D+4], ldiag [D], sol [D], barriers
  static constexpr int kRhsOff = kBigBytes;
  static constexpr int kColkOff = kRhsOff + 2 * D * 4;
  static constexpr int kLdiagOff = kColkOff + (D + 4) * 4;
  static constexpr int kSolOff = kLdiagOff + D * 4;
  static constexpr int kBarOff = kSolOff + D * 4;
  static constexpr int kTotal = kBarOff + 64;
  static constexpr int kTmemCols = D == 256 ? 512 : 128;
};

template <int D>
__global__ void __launch_bounds__(TC_THREADS, 1) row_solve_tc_kernel(RowParams p) {
  using L = TcLayout<D>;
  constexpr int F4 = D / 32;   // float4 loads per loader lane (its 128 B / 64 B slab of the gathered row)
  constexpr int C = 4 * F4;    // floats per loader lane
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* Mtx = reinterpret_cast<float*>(sm);
  float* rhs_part = reinterpret_cast<float*>(sm + L::kRhsOff);
  float* colk = reinterpret_cast<float*>(sm + L::kColkOff);
  float* ldiag = reinterpret_cast<float*>(sm + L::kLdiagOff);
  float* sol = reinterpret_cast<float*>(sm + L::kSolOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::kBarOff);
  uint64_t* full_bar = bars;       // [2]
  uint64_t* empty_bar = bars + 2;  // [2]
  uint64_t* acc_bar = bars + 4;    // [1]
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = TC_THREADS, NW = TC_THREADS / 32;
  const int mode = p.mode;
  const bool item_side = (mode == RM_SAFER_V);
  float* rhs = Mtx + tri(D);

  if (tid == 0) {
    mbar_init(&full_bar[0], 8);
    mbar_init(&full_bar[1], 8);
    mbar_init(&empty_bar[0], 1);
    mbar_init(&empty_bar[1], 1);
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_LOADER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(L::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sm_addr = smem_u32(sm);

  uint32_t loader_use = 0;          // tiles this loader group has staged so far (all rows)
  uint32_t mma_use[2] = {0u, 0u};   // tiles consumed per stage (all rows)
  uint32_t row_count = 0;

  for (int ri = blockIdx.x; ri < p.num_rows; ri += gridDim.x, ++row_count) {
    const int r = p.order[ri];
    const int beg = p.ptr[r];
    const int n = p.ptr[r + 1] - beg;
    const int xr = p.xmap ? p.xmap[r] : r;
    const int T = (n + KT - 1) / KT;
    int dup_lo = 0, dup_hi = 0;
    if (item_side && n > 128 && (n & 127) != 0) {  // stale tail, safer2.h:200-204 (B-1)
      const int kf = n >> 7;
      dup_lo = 128 * (kf - 1) + (n & 127);
      dup_hi = 128 * kf;
    }

    if (warp < TC_LOADER_WARPS) {
      // ================= loaders: gather, split, transpose into the operand tiles =================
      const int g = warp >> 3, wg = warp & 7;
      const int slab = wg * C;  // first feature this warp handles
      float rhs_acc = 0.f;
      const uint32_t kq = (uint32_t)(lane >> 2), kr = (uint32_t)(lane & 3) << 2;
      for (int t = g; t < T; t += 2) {
        const int e = t * KT + lane;
        const bool valid = e < n;
        float4 v[F4];
        float sq = 0.f, qr = 0.f;
        if (valid) {
          const int c = __ldg(p.col + beg + e);
          float s = 1.f, q = 1.f;
          if (item_side) { const float w = __ldg(p.entry_w + c); s = w; q = w; }
          if (e >= dup_lo && e < dup_hi) s *= 2.f;
          sq = sqrtf(s);
          qr = s > 0.f ? q / sq : 0.f;
          const float4* src = reinterpret_cast<const float4*>(p.E + (size_t)c * D + slab);
#pragma unroll
          for (int j = 0; j < F4; ++j) v[j] = __ldg(src + j);
        } else {
#pragma unroll
          for (int j = 0; j < F4; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (loader_use > 0) mbar_wait(&empty_bar[g], (loader_use - 1) & 1);  // stage free again
        uint8_t* hi_tile = sm + g * L::kStageBytes;
        uint8_t* lo_tile = hi_tile + L::kTileBytes;
        float rv[C];
#pragma unroll
        for (int j = 0; j < F4; ++j) {
          const float x4[4] = {v[j].x * sq, v[j].y * sq, v[j].z * sq, v[j].w * sq};
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            const int mn = slab + 4 * j + t4;
            const float x = x4[t4];
            const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
            const float lo = x - hi;
            const uint32_t off = ((uint32_t)(mn >> 3) << 10) + ((uint32_t)(mn & 7) << 7) +
                                 (((kq ^ (uint32_t)(mn & 7)) & 7u) << 4) + kr;
            *reinterpret_cast<float*>(hi_tile + off) = hi;
            *reinterpret_cast<float*>(lo_tile + off) = lo;
            rv[4 * j + t4] = qr * x;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[g]);
        ++loader_use;
        transpose_reduce<C>(rv, lane);
        rhs_acc += rv[0];
      }
      if (C == 32 || (lane & 1) == 0) rhs_part[g * D + slab + (C == 32 ? lane : (lane >> 1))] = rhs_acc;
    } else {
      // ================= MMA issuer =================
      constexpr uint32_t idesc_n128 = make_idesc_tf32(128);
      constexpr uint32_t idesc_n256 = make_idesc_tf32(256);
      for (int t = 0; t < T; ++t) {
        const int s = t & 1;
        mbar_wait(&full_bar[s], mma_use[s] & 1);
        ++mma_use[s];
        tc_fence_after();
        if (lane == 0) {
          const uint32_t hi_addr = sm_addr + s * L::kStageBytes;
          const uint32_t lo_addr = hi_addr + L::kTileBytes;
#pragma unroll
          for (int ks = 0; ks < KT / 8; ++ks) {
            const uint32_t ko = ks * 32;
            const uint64_t b_hi = make_kmajor_desc(hi_addr + ko);
            const uint64_t b_lo = make_kmajor_desc(lo_addr + ko);
            const uint32_t first = (t == 0 && ks == 0) ? 0u : 1u;
            {  // rows 0..127 x cols 0..127 -> TMEM columns [0,128)
              const uint64_t a_hi = b_hi, a_lo = b_lo;
              umma_tf32(tmem_base, a_hi, b_hi, idesc_n128, first);
              umma_tf32(tmem_base, a_hi, b_lo, idesc_n128, 1u);
              umma_tf32(tmem_base, a_lo, b_hi, idesc_n128, 1u);
            }
            if (D == 256) {  // rows 128..255 x cols 0..255 -> TMEM columns [256,512)
              const uint64_t a_hi = make_kmajor_desc(hi_addr + 128 * 128 + ko);
              const uint64_t a_lo = make_kmajor_desc(lo_addr + 128 * 128 + ko);
              umma_tf32(tmem_base + 256, a_hi, b_hi, idesc_n256, first);
              umma_tf32(tmem_base + 256, a_hi, b_lo, idesc_n256, 1u);
              umma_tf32(tmem_base + 256, a_lo, b_hi, idesc_n256, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (t == T - 1) umma_commit(acc_bar);
        }
        __syncwarp();
      }
    }

    // ================= phase B: everyone =================
    mbar_wait(acc_bar, row_count & 1);
    tc_fence_after();
    __syncthreads();  // rhs_part visible; all operand tiles consumed -> the staging area may be overwritten
    if (warp < 16) {
      // TMEM -> packed lower triangle.  warp w: lane quarter w%4, M block (w/4)%2, column chunks of parity w/8.
      const int q = warp & 3, b = (warp >> 2) & 1, par = warp >> 3;
      if (D == 256 || b == 0) {
        const int i = 128 * b + 32 * q + lane;
        const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + (b ? 256u : 0u);
        const int nchunks = (128 * b + 32 * q + 31) / 32 + 1;  // columns 0 .. 32*nchunks-1 cover j <= i for the whole warp
        float* mrow = Mtx + tri(i);
        for (int ch = par; ch < nchunks; ch += 2) {
          uint32_t u[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
                "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
                "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
              : "r"(tbase + (uint32_t)(32 * ch)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int j = 32 * ch + jj;
            if (j <= i) mrow[j] = __uint_as_float(u[jj]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();

    // ---- per-row scalars (same formulas as the generic kernel) ----
    float reg, weight = 1.f;
    if (mode == RM_IALS) {
      reg = (float)((double)p.reg * pow((double)((float)n + p.uw * (float)p.num_other), (double)p.reg_exp));
    } else if (item_side) {
      reg = p.reg * (p.item_reg[r] + p.alpha * p.uw * (float)p.num_users_total);
    } else {
      reg = p.reg * (1.f + p.uw * (float)p.num_other);
      weight = p.row_w ? p.row_w[r] : 1.f;
    }
    const bool user_form = (mode == RM_SAFER_U);
    const float nf = (float)n;
    for (int i = warp; i < D; i += NW) {
      float* mrow = Mtx + tri(i);
      const float* grow = p.G + (size_t)i * D;
      for (int j = lane; j <= i; j += 32) {
        const float g = __ldg(grow + j);
        const float sij = mrow[j];
        float m;
        if (user_form) {
          m = sij / nf;
          m += p.uw * g;
          m *= weight;
          if (i == j) m += reg;
        } else if (mode == RM_IALS) {
          m = p.uw * g;
          if (i == j) m += reg;
          m += sij;
        } else {
          m = p.uw * g + sij;
          if (i == j) m += reg;
        }
        mrow[j] = m;
      }
    }
    {
      const float sc = user_form ? weight / nf : 1.f;
      for (int k = tid; k < D; k += NT) rhs[k] = (rhs_part[k] + rhs_part[D + k]) * sc;
      if (tid == 0) rhs[D] = 0.f;
    }
    __syncthreads();

    // ---- Cholesky of the augmented system; row D carries rhs -> y ----
    for (int k = 0; k < D; ++k) {
      float pivot = Mtx[tri(k) + k];
      if (!(pivot > 0.f)) {
        if (tid == 0) atomicExch(p.status, 1);
        pivot = 1.f;
      }
      const float l = sqrtf(pivot);
      for (int i = k + 1 + tid; i <= D; i += NT) {
        const float v = Mtx[tri(i) + k] / l;
        Mtx[tri(i) + k] = v;
        colk[i] = v;
      }
      if (tid == 0) ldiag[k] = l;
      __syncthreads();
      for (int i = k + 1 + warp; i <= D; i += NW) {
        const float lik = colk[i];
        float* mrow = Mtx + tri(i);
        const int jmax = min(i, D - 1);
        for (int j = k + 1 + lane; j <= jmax; j += 32) mrow[j] = fmaf(-lik, colk[j], mrow[j]);
      }
      __syncthreads();
    }
    if (warp == 0) {
      for (int j = lane; j < D; j += 32) sol[j] = rhs[j];
      __syncwarp();
      for (int k = D - 1; k >= 0; --k) {
        const float xk = sol[k] / ldiag[k];
        __syncwarp();
        if (lane == 0) sol[k] = xk;
        const float* mrow = Mtx + tri(k);
        for (int j = lane; j < k; j += 32) sol[j] = fmaf(-mrow[j], xk, sol[j]);
        __syncwarp();
      }
    }
    __syncthreads();
    for (int k = tid; k < D; k += NT) p.X[(size_t)xr * D + k] = sol[k];
    __syncthreads();  // Mtx (aliased with the operand stages) is free for the next row's loaders
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_LOADER_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::kTmemCols));
}

}  // namespace

bool row_solve_tc_supported(const RowParams& p) {
  const bool mode_ok = p.mode == RM_IALS || p.mode == RM_SAFER_U || p.mode == RM_SAFER_V;
  return mode_ok && p.cs == 0 && p.bd == p.d && (p.d == 128 || p.d == 256);
}

void launch_row_solve_tc(const RowParams& p, cudaStream_t s, int num_sms, long long* launches) {
  if (p.num_rows <= 0) return;
  if (p.d == 256) {
    const int smem = TcLayout<256>::kTotal + 1024;
    cudaFuncSetAttribute(row_solve_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int grid = num_sms < p.num_rows ? num_sms : p.num_rows;
    row_solve_tc_kernel<256><<<grid, TC_THREADS, smem, s>>>(p);
  } else {
    const int smem = TcLayout<128>::kTotal + 1024;
    cudaFuncSetAttribute(row_solve_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int grid = 2 * num_sms < p.num_rows ? 2 * num_sms : p.num_rows;
    row_solve_tc_kernel<128><<<grid, TC_THREADS, smem, s>>>(p);
  }
  if (launches) ++*launches;
}

}  // namespace frx

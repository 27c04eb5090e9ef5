// Gramian G = E^T diag(w) E on tcgen05 / TMEM, operand rows fed by TMA (d = 128 / 256).
//
// Restates `item_embedding_.transpose() * item_embedding_` (ials.h:321, safer2.h:55,294-295) and the
// weighted `user_embedding.transpose() * w_user_embedding` of StepV (safer2.h:504-509).
//
// E is row-major [n x d], so the contraction index (the row) is the strided one: a TMA tile of E is an
// MN-major operand, which kind::tf32 does not accept (tools/tc_probe.cu).  Pipeline per CTA (one row slab):
//   warp 0      TMA producer: cp.async.bulk.tensor loads [32 rows x 32 floats] boxes, SWIZZLE_128B, into
//               a raw stage (mbarrier complete_tx); out-of-range rows arrive as zeros.
//   warps 2-9   converters: lane = row reads its 128 B slab from the swizzled raw tile (conflict-free),
//               scales by sqrt(w), splits fp32 -> tf32 hi + lo and stores them transposed into the K-major
//               operand tiles [feature][32 rows].
//   warp 1      one thread issues tcgen05.mma kind::tf32 hi*hi + hi*lo + lo*hi (3xTF32, fp32-level accuracy)
//               into TMEM: rows 0-127 x cols 0-127 and rows 128-255 x cols 0-255 (lower tiles).
//   warps 2-9   epilogue: TMEM -> per-CTA partial in global memory; a second kernel adds the partials in a
//               fixed order (deterministic) and mirrors the lower triangle.
#include "frx_kernels.cuh"
#include <cuda.h>
#include <cstdint>
#include <cstdio>

namespace frx {

namespace {

constexpr int GT_THREADS = 320;  // TMA warp, MMA warp, 8 converter/epilogue warps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int D>
struct GtLayout {
  static constexpr int kRawBytes = 32 * D * 4;           // one raw stage: D/32 boxes of 32 rows x 128 B
  static constexpr int kTileBytes = D * 128;             // one operand tile
  static constexpr int kRawOff = 0;
  static constexpr int kOpOff = 2 * kRawBytes;           // two raw stages, then two (hi, lo) operand stages
  static constexpr int kBarOff = kOpOff + 4 * kTileBytes;
  static constexpr int kTotal = kBarOff + 128;
  static constexpr int kTmemCols = D == 256 ? 512 : 128;
};

template <int D>
__global__ void __launch_bounds__(GT_THREADS, 1) gramian_tc_kernel(const __grid_constant__ CUtensorMap tmap, int n,
                                                                 const float* __restrict__ w, float* __restrict__ ws,
                                                                 int chunks_per_cta) {
  using L = GtLayout<D>;
  constexpr int NBOX = D / 32;
  constexpr int C = D / 8;  // floats per converter lane (8 converter warps split the D features)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::kBarOff);
  uint64_t* raw_full = bars;        // [2] TMA bytes landed
  uint64_t* raw_empty = bars + 2;   // [2] converters have read the raw stage
  uint64_t* op_full = bars + 4;     // [2] converters have written the operand stage
  uint64_t* op_empty = bars + 6;    // [2] MMAs that read the operand stage have completed
  uint64_t* acc_full = bars + 8;    // [1]
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], 8);
      mbar_init(&op_full[s], 8);
      mbar_init(&op_empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(L::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sm_addr = smem_u32(sm);

  const int chunk0 = blockIdx.x * chunks_per_cta;
  int nchunks = (n + 31) / 32 - chunk0;
  if (nchunks > chunks_per_cta) nchunks = chunks_per_cta;
  if (nchunks < 0) nchunks = 0;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        if (c >= 2) mbar_wait(&raw_empty[s], ((c >> 1) - 1) & 1);
        mbar_expect_tx(&raw_full[s], L::kRawBytes);
        const int row0 = (chunk0 + c) * 32;
        for (int b = 0; b < NBOX; ++b) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                  sm_addr + L::kRawOff + s * L::kRawBytes + b * 4096),
              "l"(&tmap), "r"(smem_u32(&raw_full[s])), "r"(b * 32), "r"(row0)
              : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_n128 = make_idesc_tf32(128);
    constexpr uint32_t idesc_n256 = make_idesc_tf32(256);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1;
      mbar_wait(&op_full[s], (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t hi_addr = sm_addr + L::kOpOff + s * 2 * L::kTileBytes;
        const uint32_t lo_addr = hi_addr + L::kTileBytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ko = ks * 32;
          const uint64_t b_hi = make_kmajor_desc(hi_addr + ko), b_lo = make_kmajor_desc(lo_addr + ko);
          const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
          umma_tf32(tmem_base, b_hi, b_hi, idesc_n128, first);
          umma_tf32(tmem_base, b_hi, b_lo, idesc_n128, 1u);
          umma_tf32(tmem_base, b_lo, b_hi, idesc_n128, 1u);
          if (D == 256) {
            const uint64_t a_hi = make_kmajor_desc(hi_addr + 128 * 128 + ko), a_lo = make_kmajor_desc(lo_addr + 128 * 128 + ko);
            umma_tf32(tmem_base + 256, a_hi, b_hi, idesc_n256, first);
            umma_tf32(tmem_base + 256, a_hi, b_lo, idesc_n256, 1u);
            umma_tf32(tmem_base + 256, a_lo, b_hi, idesc_n256, 1u);
          }
        }
        umma_commit(&op_empty[s]);
        if (c == nchunks - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ===== converters: raw (row-major, swizzled) -> transposed tf32 hi / lo operand tiles =====
    const int cw = warp - 2;       // 0..7
    const int slab = cw * C;       // first feature of this warp
    const uint32_t kq = (uint32_t)(lane >> 2), kr = (uint32_t)(lane & 3) << 2;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c & 1;
      const int row = (chunk0 + c) * 32 + lane;
      float sq = 0.f;
      if (row < n) sq = w ? sqrtf(__ldg(w + row)) : 1.f;
      mbar_wait(&raw_full[s], (c >> 1) & 1);
      const uint8_t* raw = sm + L::kRawOff + s * L::kRawBytes;
      float4 v[C / 4];
#pragma unroll
      for (int j = 0; j < C / 4; ++j) {
        const int col = slab + 4 * j;
        const int box = col >> 5, ch = (col & 31) >> 2;
        v[j] = *reinterpret_cast<const float4*>(raw + box * 4096 + lane * 128 + ((ch ^ (lane & 7)) << 4));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[s]);
      if (c >= 2) mbar_wait(&op_empty[s], ((c >> 1) - 1) & 1);
      uint8_t* hi_tile = sm + L::kOpOff + s * 2 * L::kTileBytes;
      uint8_t* lo_tile = hi_tile + L::kTileBytes;
#pragma unroll
      for (int j = 0; j < C / 4; ++j) {
        const float x4[4] = {v[j].x * sq, v[j].y * sq, v[j].z * sq, v[j].w * sq};
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          const int mn = slab + 4 * j + t4;
          const float x = x4[t4];
          const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
          const float lo = x - hi;
          const uint32_t off = ((uint32_t)(mn >> 3) << 10) + ((uint32_t)(mn & 7) << 7) + (((kq ^ (uint32_t)(mn & 7)) & 7u) << 4) + kr;
          *reinterpret_cast<float*>(hi_tile + off) = hi;
          *reinterpret_cast<float*>(lo_tile + off) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&op_full[s]);
    }
    // ===== epilogue: TMEM -> this CTA's partial (lower tiles) =====
    float* out = ws + (size_t)blockIdx.x * D * D;
    const int q = warp & 3, b = (warp - 2) >> 2;  // warps 2-5: M block 0, warps 6-9: M block 1
    if (D == 256 || b == 0) {
      if (nchunks > 0) {
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int i = 128 * b + 32 * q + lane;
      const int ncols = b ? 256 : 128;
      const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + (b ? 256u : 0u);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t u[32];
        if (nchunks > 0) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
              "%30,%31}, [%32];"
              : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
                "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
                "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
              : "r"(tbase + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = 0u;
        }
        float4* dst = reinterpret_cast<float4*>(out + (size_t)i * D + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                               __uint_as_float(u[4 * j + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::kTmemCols));
}

// out = sum over CTAs of the lower-tile partials, mirrored to the upper triangle.
__global__ void gramian_tc_reduce_kernel(const float* __restrict__ ws, int nparts, int d, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * d) return;
  const int i = idx / d, j = idx % d;
  if (j > i) return;
  float s = 0.f;
  for (int pt = 0; pt < nparts; ++pt) s += ws[(size_t)pt * d * d + idx];
  out[(size_t)i * d + j] = s;
  out[(size_t)j * d + i] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

}  // namespace

bool gramian_tc_supported(int n, int d, int cs, int bd, int fs, int fd) {
  return (d == 128 || d == 256) && cs == 0 && fs == 0 && bd == d && fd == d && n > 0 && encode_fn() != nullptr;
}

size_t gramian_tc_workspace_floats(int d, int num_sms) { return (size_t)num_sms * d * d; }

int launch_gramian_tc(const float* E, int n, int d, const float* w, float* out, float* workspace, cudaStream_t s,
                      int num_sms, long long* launches) {
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)n};
  cuuint64_t gstr[1] = {(cuuint64_t)d * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(E), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -1;
  const int chunks = (n + 31) / 32;
  int grid = num_sms < chunks ? num_sms : chunks;
  const int per = (chunks + grid - 1) / grid;
  grid = (chunks + per - 1) / per;
  if (d == 256) {
    const int smem = GtLayout<256>::kTotal + 1024;
    cudaFuncSetAttribute(gramian_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gramian_tc_kernel<256><<<grid, GT_THREADS, smem, s>>>(tmap, n, w, workspace, per);
  } else {
    const int smem = GtLayout<128>::kTotal + 1024;
    cudaFuncSetAttribute(gramian_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gramian_tc_kernel<128><<<grid, GT_THREADS, smem, s>>>(tmap, n, w, workspace, per);
  }
  gramian_tc_reduce_kernel<<<(d * d + 255) / 256, 256, 0, s>>>(workspace, grid, d, out);
  if (launches) *launches += 2;
  return 0;
}

}  // namespace frx

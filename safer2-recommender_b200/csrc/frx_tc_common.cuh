// Device helpers shared by the tcgen05 / TMEM row kernels (frx_row_tc.cu, frx_row_wb.cu): mbarriers, UMMA
// issue, shared-memory descriptors, TMEM load/store, packed-FMA sweeps and the per-row scalars.
// Hardware conventions were established with tools/tc_probe.cu / tc_rate.cu / lat_probe.cu on a B200 (see
// the header of frx_row_tc.cu).
#pragma once
#include "frx_kernels.cuh"
#include <cstdint>

namespace frx {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MUFU.RSQ without the denormal fix-up of rsqrtf(); the pivots are normal numbers >= reg.
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32 pairs (FFMA2): a 3-register FFMA issues every other cycle per scheduler, so the in-register
// triangular sweeps are FMA-issue bound; fma.rn.f32x2 does two IEEE fp32 FMAs per issue slot.
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
// (d0, d1) += a2 * (b0, b1)
__device__ __forceinline__ void fma2(float& d0, float& d1, unsigned long long a2, float b0, float b1) {
  asm("{\n\t.reg .b64 rb, rc;\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, %2, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "l"(a2), "f"(b0), "f"(b1));
}
// a[j..j+3] -= l * c for the entries with index > k (k is a compile-time constant after unrolling)
#define FRX_SWEEP4(a, j, k, l, nl2, c)                               \
  do {                                                               \
    if ((j) > (k)) fma2(a[(j)], a[(j) + 1], nl2, c.x, c.y);          \
    else if ((j) + 1 > (k)) a[(j) + 1] = fmaf(-(l), c.y, a[(j) + 1]); \
    if ((j) + 2 > (k)) fma2(a[(j) + 2], a[(j) + 3], nl2, c.z, c.w);  \
    else if ((j) + 3 > (k)) a[(j) + 3] = fmaf(-(l), c.w, a[(j) + 3]); \
  } while (0)

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
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define FRX_TMEM_LD32(u, taddr)                                                                                         \
  asm volatile(                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                         \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29," \
      "%30,%31}, [%32];"                                                                                                \
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),     \
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),          \
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),         \
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])                       \
      : "r"(taddr));                                                                                                    \
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

#define FRX_TMEM_ST32(taddr, u)                                                                                         \
  asm volatile(                                                                                                         \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                   \
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30," \
      "%31,%32};" ::"r"(taddr),                                                                                         \
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),     \
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]),       \
      "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]),       \
      "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])                                                                    \
      : "memory")

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
// kind::tf32, fp32 accumulate, K-major A and B, M = 128; optional negation of A.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N, int a_negate = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_negate & 1) << 13) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// kind::f16 with fp16 A and B (format 0), fp32 accumulate, K-major, M = 128: K = 16 per instruction
__host__ __device__ constexpr uint32_t make_idesc_f16(int N, int a_negate = 0) {
  return (1u << 4) | ((uint32_t)(a_negate & 1) << 13) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of element (row mn, k-chunk c of 4 floats) in a K-major SWIZZLE_128B tile with 128 B rows
__device__ __forceinline__ uint32_t tile_chunk_off(int mn, int c) {
  return ((uint32_t)(mn >> 3) << 10) + ((uint32_t)(mn & 7) << 7) + ((((uint32_t)c ^ (uint32_t)(mn & 7)) & 7u) << 4);
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

// Per-row scalars.  The system handed to the Cholesky is  S + alpha*G + beta*I  with S the tensor-core
// SYRK sum and rhs * bscale on the right: for the user form (safer2.h:143-150) that is the reference's
// w*(S/n + uw*G) + reg*I and rhs*w/n multiplied through by n/w, which leaves the solution unchanged and
// lets alpha*G + beta*I be written into TMEM BEFORE the SYRK accumulates on top of it.
struct RowScalars {
  float alpha, beta, bscale;
};
__device__ __forceinline__ RowScalars row_scalars(const RowParams& p, int r, int n) {
  RowScalars s;
  s.bscale = 1.f;
  if (p.mode == RM_IALS) {  // ials.h:101-105
    s.alpha = p.uw;
    s.beta = (float)((double)p.reg * pow((double)((float)n + p.uw * (float)p.num_other), (double)p.reg_exp));
  } else if (p.mode == RM_SAFER_V) {  // safer2.h:176,206-208
    s.alpha = p.uw;
    s.beta = p.reg * (p.item_reg[r] + p.alpha * p.uw * (float)p.num_users_total);
  } else {
    const float reg = p.reg * (1.f + p.uw * (float)p.num_other);
    const float w = p.row_w ? p.row_w[r] : 1.f;
    if (w > 0.f) {
      s.alpha = (float)n * p.uw;
      s.beta = (float)n * reg / w;
    } else {  // reference system degenerates to reg*I x = 0
      s.alpha = 0.f;
      s.beta = 1.f;
      s.bscale = 0.f;
    }
  }
  return s;
}

}  // namespace tc
}  // namespace frx

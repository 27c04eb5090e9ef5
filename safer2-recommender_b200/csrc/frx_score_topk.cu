// Fused scoring + top-k for the evaluation (recommender.h:109-112, 132-153): scores = Ut * V^T on tcgen05 / TMEM
// with the per-user top-max_k selected in the epilogue -- the nu x num_items score matrix is never written.
//
// The reference scores one held-out user at a time (`item_embedding_ * user_embedding` GEMV, safer2.h:259-262),
// masks the fold-in history to numeric_limits<float>::lowest() (recommender.h:137-140) and takes the top 100 with
// nth_element + stable_sort (:143-153).  Here:
//   split kernel     fp32 -> tf32 hi + lo copies of Ut and V (x = hi + lo, hi = 19 leading bits), once per call;
//   bitmap kernel    one bit per (held-out user, item) of the fold-in history;
//   score_topk       CTA = (block of 128 users, segment of the item range).  Warp 0 streams [128 x 32] / [256 x 32]
//                    operand boxes of the four split arrays with TMA (SWIZZLE_128B, zero fill past the ends) into a
//                    two-stage ring; warp 1 issues kind::tf32 MMAs hi*hi + hi*lo + lo*hi (3xTF32: scores accurate
//                    to ~1e-7 relative, so the ranking is the fp32 ranking up to near-ties), M = 128 users, N = 256
//                    items, into one of two 256-column TMEM accumulators; warps 2-5 (thread = user) read the other
//                    accumulator, apply the history mask and keep the user's best max_k as 64-bit keys
//                    (score key, -item): a candidate enters when it beats the current minimum, which is then
//                    searched again.  The total order on the keys makes the selection deterministic: among equal
//                    scores the lower item id wins (the reference's order among ties is unspecified, B-11);
//   merge + metrics  frx_eval.cu sorts the per-segment candidates of a user and computes Recall / NDCG.
#include "frx_tc_common.cuh"
#include <cuda.h>
#include <cfloat>

namespace frx {

using namespace tc;

namespace {

constexpr int SK_THREADS = 192;   // TMA warp, MMA warp, 4 epilogue warps
constexpr int SK_BM = 128, SK_BN = 256, SK_KC = 32;
constexpr int SK_A_BYTES = SK_BM * 128, SK_B_BYTES = SK_BN * 128;
constexpr int SK_STAGE_BYTES = 2 * SK_A_BYTES + 2 * SK_B_BYTES;  // Ahi, Alo, Bhi, Blo
constexpr int SK_NSTAGE = 2;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void split_tf32_kernel(const float* __restrict__ x, size_t n, float* __restrict__ hi, float* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// bitmap[row][item / 32] |= 1 << (item % 32) for the fold-in history of held-out user `row` (compact index)
__global__ void history_bitmap_kernel(const int* __restrict__ user_ids, int nu, const int* __restrict__ tr_ptr,
                                      const int* __restrict__ tr_col, int words, unsigned* __restrict__ bitmap) {
  const int row = blockIdx.x;
  if (row >= nu) return;
  const int uid = user_ids[row];
  for (int t = tr_ptr[uid] + threadIdx.x; t < tr_ptr[uid + 1]; t += blockDim.x) {
    const int item = tr_col[t];
    atomicOr(bitmap + (size_t)row * words + (item >> 5), 1u << (item & 31));
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1)
    score_topk_kernel(const __grid_constant__ CUtensorMap map_uhi, const __grid_constant__ CUtensorMap map_ulo,
                      const __grid_constant__ CUtensorMap map_vhi, const __grid_constant__ CUtensorMap map_vlo, int nu,
                      int num_items, int d, int user_blocks, int segments, const unsigned* __restrict__ bitmap, int words,
                      int K, int kpad, unsigned long long* __restrict__ cand) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SK_NSTAGE * SK_STAGE_BYTES);
  uint64_t* full_bar = bars;        // [2] TMA bytes landed
  uint64_t* empty_bar = bars + 2;   // [2] the MMAs that read the stage have completed
  uint64_t* acc_full = bars + 4;    // [2] accumulator complete
  uint64_t* acc_empty = bars + 6;   // [2] epilogue warps have read it
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sm_addr = smem_u32(sm);

  const int ub = blockIdx.x % user_blocks, seg = blockIdx.x / user_blocks;
  const int ntiles_all = (num_items + SK_BN - 1) / SK_BN;
  const int t0 = (int)((long long)ntiles_all * seg / segments), t1 = (int)((long long)ntiles_all * (seg + 1) / segments);
  const int ntiles = t1 - t0;
  const int kchunks = d / SK_KC;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t cs = 0;
      for (int t = 0; t < ntiles; ++t) {
        for (int k = 0; k < kchunks; ++k, ++cs) {
          const uint32_t st = cs % SK_NSTAGE, use = cs / SK_NSTAGE;
          if (use > 0) mbar_wait(&empty_bar[st], (use - 1) & 1);
          mbar_expect_tx(&full_bar[st], SK_STAGE_BYTES);
          const uint32_t base = sm_addr + st * SK_STAGE_BYTES;
          tma_load_2d(base, &map_uhi, &full_bar[st], k * SK_KC, ub * SK_BM);
          tma_load_2d(base + SK_A_BYTES, &map_ulo, &full_bar[st], k * SK_KC, ub * SK_BM);
          tma_load_2d(base + 2 * SK_A_BYTES, &map_vhi, &full_bar[st], k * SK_KC, (t0 + t) * SK_BN);
          tma_load_2d(base + 2 * SK_A_BYTES + SK_B_BYTES, &map_vlo, &full_bar[st], k * SK_KC, (t0 + t) * SK_BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_tf32(SK_BN);
    uint32_t cs = 0;
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t buf = (uint32_t)t & 1u, buse = (uint32_t)t >> 1;
      if (buse > 0) mbar_wait(&acc_empty[buf], (buse - 1) & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + 256u * buf;
      for (int k = 0; k < kchunks; ++k, ++cs) {
        const uint32_t st = cs % SK_NSTAGE, use = cs / SK_NSTAGE;
        mbar_wait(&full_bar[st], use & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = sm_addr + st * SK_STAGE_BYTES, a_lo = a_hi + SK_A_BYTES;
          const uint32_t b_hi = a_hi + 2 * SK_A_BYTES, b_lo = b_hi + SK_B_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dah = make_kmajor_desc(a_hi + ks * 32), dal = make_kmajor_desc(a_lo + ks * 32);
            const uint64_t dbh = make_kmajor_desc(b_hi + ks * 32), dbl = make_kmajor_desc(b_lo + ks * 32);
            umma_tf32(d_tmem, dah, dbh, idesc, (k | ks) ? 1u : 0u);
            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
            umma_tf32(d_tmem, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[st]);
          if (k == kchunks - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: thread = user, running top-K of 64-bit keys =====
    const int q = warp & 3;
    const int row = ub * SK_BM + 32 * q + lane;  // compact held-out user index
    const bool live = row < nu;
    unsigned long long* mine = cand + ((size_t)(live ? row : 0) * segments + seg) * kpad;
    unsigned long long minkey = 0ull;
    int minpos = 0;
    if (live)
      for (int j = 0; j < kpad; ++j) mine[j] = 0ull;
    const unsigned* bm = bitmap + (size_t)(live ? row : 0) * words;
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t buf = (uint32_t)t & 1u, buse = (uint32_t)t >> 1;
      mbar_wait(&acc_full[buf], buse & 1);
      tc_fence_after();
      const int item0 = (t0 + t) * SK_BN;
      for (int c = 0; c < SK_BN / 32; ++c) {
        uint32_t u[32];
        FRX_TMEM_LD32(u, tmem_base + ((uint32_t)(32 * q) << 16) + 256u * buf + 32u * (uint32_t)c);
        const int ibase = item0 + 32 * c;
        if (!live || ibase >= num_items) continue;
        const unsigned mbits = bm[ibase >> 5];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int item = ibase + j;
          float sc = __uint_as_float(u[j]);
          if ((mbits >> j) & 1u) sc = -FLT_MAX;  // recommender.h:137-140
          const unsigned long long key = ((unsigned long long)f2key(sc) << 32) | (unsigned)(0xffffffffu - (unsigned)item);
          if (item < num_items && key > minkey) {
            mine[minpos] = key;
            unsigned long long mk = ~0ull;
            int mp = 0;
            for (int e = 0; e < K; ++e) {
              const unsigned long long v = mine[e];
              if (v < mk) { mk = v; mp = e; }
            }
            minkey = mk;
            minpos = mp;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn sk_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
bool make_map(CUtensorMap* m, const float* base, int rows, int d, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)d * 4};
  cuuint32_t box[2] = {(cuuint32_t)SK_KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return sk_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool score_topk_supported(int d, int max_k) { return d % 32 == 0 && d >= 32 && max_k <= FRX_TOPK_PAD && sk_encode_fn() != nullptr; }

int score_topk_segments(int nu, int num_sms) {
  const int ubs = (nu + SK_BM - 1) / SK_BM;
  int s = (num_sms + ubs - 1) / ubs;
  const int max_s = FRX_TOPK_CAND / FRX_TOPK_PAD;
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

size_t score_topk_workspace_bytes(int nu, int num_items, int d, int num_sms) {
  const size_t split = 2 * ((size_t)nu + (size_t)num_items) * d * sizeof(float);
  const size_t words = ((size_t)(num_items + 255) / 256) * 8;
  const size_t bitmap = (size_t)nu * words * sizeof(unsigned);
  const size_t cand = (size_t)nu * score_topk_segments(nu, num_sms) * FRX_TOPK_PAD * sizeof(unsigned long long);
  return split + bitmap + cand + 1024;
}

// Returns 0 on success; *cand_out / *segments_out describe the candidate lists for the merge kernel.
int launch_score_topk(const EvalParams& p, void* workspace, int num_sms, cudaStream_t s, long long* launches,
                      const unsigned long long** cand_out, int* segments_out) {
  const int nu = p.nu, I = p.num_items, d = p.d;
  float* w = reinterpret_cast<float*>(workspace);
  float* uhi = w;
  float* ulo = uhi + (size_t)nu * d;
  float* vhi = ulo + (size_t)nu * d;
  float* vlo = vhi + (size_t)I * d;
  const int words = ((I + 255) / 256) * 8;
  unsigned* bitmap = reinterpret_cast<unsigned*>(vlo + (size_t)I * d);
  const int segments = score_topk_segments(nu, num_sms);
  // 8-byte aligned: all the preceding arrays hold an even number of 4-byte words or are padded by `words`
  uintptr_t cp = reinterpret_cast<uintptr_t>(bitmap + (size_t)nu * words);
  cp = (cp + 7) & ~(uintptr_t)7;
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(cp);
  const int grid1 = num_sms * 8;
  split_tf32_kernel<<<grid1, 256, 0, s>>>(p.Ut, (size_t)nu * d, uhi, ulo);
  split_tf32_kernel<<<grid1, 256, 0, s>>>(p.V, (size_t)I * d, vhi, vlo);
  cudaMemsetAsync(bitmap, 0, (size_t)nu * words * sizeof(unsigned), s);
  history_bitmap_kernel<<<nu, 128, 0, s>>>(p.user_ids, nu, p.tr_ptr, p.tr_col, words, bitmap);
  CUtensorMap muh, mul, mvh, mvl;
  if (!make_map(&muh, uhi, nu, d, SK_BM) || !make_map(&mul, ulo, nu, d, SK_BM) || !make_map(&mvh, vhi, I, d, SK_BN) ||
      !make_map(&mvl, vlo, I, d, SK_BN))
    return -1;
  const int ubs = (nu + SK_BM - 1) / SK_BM;
  const int smem = SK_NSTAGE * SK_STAGE_BYTES + 128 + 1024;
  cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int K = p.max_k < I ? p.max_k : I;
  score_topk_kernel<<<ubs * segments, SK_THREADS, smem, s>>>(muh, mul, mvh, mvl, nu, I, d, ubs, segments, bitmap, words, K,
                                                             FRX_TOPK_PAD, cand);
  if (launches) *launches += 4;
  *cand_out = cand;
  *segments_out = segments;
  return 0;
}

}  // namespace frx

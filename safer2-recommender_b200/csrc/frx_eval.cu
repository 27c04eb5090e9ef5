// Evaluation stage (recommender.h:78-199) and dataset build (dataset.h:71-99).
//
// Evaluation: scores = Ut * V^T computed tile by tile for a chunk of held-out
// users, the user's test_tr history masked to numeric_limits<float>::lowest()
// (recommender.h:137-140), per-user top-max_k by radix select + bitonic sort
// (ties: lower item id first; the reference's nth_element order is
// unspecified, B-11), then Recall@k / NDCG@k with the reference's double
// accumulation (recommender.h:156-181).
#include "frx_kernels.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>

namespace frx {

namespace {

constexpr int ST = 64;  // score tile edge
constexpr int SK = 32;  // k chunk
constexpr int CAND = FRX_TOPK_CAND;

__global__ void __launch_bounds__(256) scores_kernel(const float* __restrict__ Ut, int nu_chunk,
                                                     const float* __restrict__ V, int num_items, int d,
                                                     float* __restrict__ scores) {
  __shared__ float As[ST][SK + 1];
  __shared__ float Bs[ST][SK + 1];
  const int u0 = blockIdx.y * ST, i0 = blockIdx.x * ST;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d; k0 += SK) {
    for (int idx = threadIdx.x; idx < ST * SK; idx += 256) {
      const int rr = idx / SK, kk = idx % SK;
      const int k = k0 + kk;
      As[rr][kk] = (u0 + rr < nu_chunk && k < d) ? Ut[(size_t)(u0 + rr) * d + k] : 0.f;
      Bs[rr][kk] = (i0 + rr < num_items && k < d) ? V[(size_t)(i0 + rr) * d + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < SK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[ty + 16 * a][kk];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[tx + 16 * b][kk];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int u = u0 + ty + 16 * a;
    if (u >= nu_chunk) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + tx + 16 * b;
      if (i < num_items) scores[(size_t)u * num_items + i] = acc[a][b];
    }
  }
}

__device__ __forceinline__ unsigned f2key(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}


// Shared tail of the two top-k front ends: cand[0..CAND) holds 64-bit keys (score key << 32 | ~item), zeros for
// empty slots.  Bitonic sort (descending), then hits against the ground truth and Recall@k / NDCG@k
// (recommender.h:156-181).  The ground-truth size is the number of DISTINCT items of the user in test_te, like
// the reference's std::set (recommender.h:183-188).
__device__ void sort_and_score(const EvalParams& p, int row, int uid, int K, unsigned long long* cand,
                               unsigned char* hit, unsigned* s_distinct) {
  for (int size = 2; size <= CAND; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < CAND / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }
  }
  // ground truth
  const int gb = (uid < p.te_rows) ? p.te_ptr[uid] : 0;
  const int gcount = (uid < p.te_rows) ? p.te_ptr[uid + 1] - gb : 0;
  if (threadIdx.x == 0) *s_distinct = 0;
  __syncthreads();
  {
    unsigned mine = 0;
    for (int g = threadIdx.x; g < gcount; g += blockDim.x) {
      const int it = p.te_col[gb + g];
      bool first = true;
      for (int h = 0; h < g && first; ++h) first = (p.te_col[gb + h] != it);
      mine += first ? 1u : 0u;
    }
    if (mine) atomicAdd(s_distinct, mine);
  }
  for (int r = threadIdx.x; r < K; r += blockDim.x) {
    const int item = (int)(0xffffffffu - (unsigned)(cand[r] & 0xffffffffull));
    if (p.topk) p.topk[(size_t)row * p.max_k + r] = item;
    unsigned char h = 0;
    for (int g = 0; g < gcount; ++g) h |= (p.te_col[gb + g] == item);
    hit[r] = h;
  }
  __syncthreads();
  const int gn = (int)*s_distinct;
  if (threadIdx.x < p.nk) {
    const int k = p.k_list[threadIdx.x];
    float rec = 0.f, nd = 0.f;
    if (gn > 0) {
      double result = 0.0, dcg = 0.0;
      for (int i = 0; i < k && i < K; ++i)
        if (hit[i]) { result += 1.0; dcg += 1.0 / log2(i + 2.0); }
      // recall: result / std::min<float>(k, gt_set.size())   (recommender.h:156-165)
      rec = (float)(result / (double)fminf((float)k, (float)gn));
      double norm = 0.0;
      for (int i = 0; i < min(k, gn); ++i) norm += 1.0 / log2(i + 2.0);
      nd = (float)(dcg / norm);  // recommender.h:168-181
    }
    p.recall[(size_t)row * p.nk + threadIdx.x] = rec;
    p.ndcg[(size_t)row * p.nk + threadIdx.x] = nd;
  }
}

// One CTA per held-out user: mask, top-max_k, metrics.
__global__ void __launch_bounds__(256) topk_metrics_kernel(EvalParams p, int chunk_begin, int chunk_n) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_need;
  __shared__ unsigned long long cand[CAND];
  __shared__ unsigned s_count;
  __shared__ unsigned char hit[CAND];
  __shared__ unsigned s_distinct;
  const int cu = blockIdx.x;
  if (cu >= chunk_n) return;
  const int row = chunk_begin + cu;
  const int uid = p.user_ids[row];
  float* sc = p.scores + (size_t)cu * p.num_items;
  const int I = p.num_items;
  // mask the fold-in history (recommender.h:137-140)
  for (int t = p.tr_ptr[uid] + threadIdx.x; t < p.tr_ptr[uid + 1]; t += blockDim.x) sc[p.tr_col[t]] = -FLT_MAX;
  __syncthreads();
  const int K = min(p.max_k, I);
  // radix select the K-th largest key
  unsigned prefix = 0, mask = 0, need = K;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
      const unsigned k = f2key(sc[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned acc = 0;
      int b = 255;
      for (; b >= 0; --b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      if (b < 0) b = 0;
      s_prefix = prefix | ((unsigned)b << (8 * pass));
      s_need = need - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= 0xffu << (8 * pass);
    __syncthreads();
  }
  const unsigned thr = prefix;  // key of the K-th largest score
  if (threadIdx.x == 0) s_count = 0;
  for (int i = threadIdx.x; i < CAND; i += blockDim.x) cand[i] = 0ull;
  __syncthreads();
  // strictly greater first (fewer than K of them), then ties
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const unsigned k = f2key(sc[i]);
    if (k > thr) {
      const unsigned slot = atomicAdd(&s_count, 1u);
      if (slot < CAND) cand[slot] = ((unsigned long long)k << 32) | (unsigned)(0xffffffffu - (unsigned)i);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const unsigned k = f2key(sc[i]);
    if (k == thr) {
      const unsigned slot = atomicAdd(&s_count, 1u);
      if (slot < CAND) cand[slot] = ((unsigned long long)k << 32) | (unsigned)(0xffffffffu - (unsigned)i);
    }
  }
  __syncthreads();
  sort_and_score(p, row, uid, K, cand, hit, &s_distinct);
}

// One CTA per held-out user: merge the per-segment candidate lists of the fused scoring kernel (frx_score_topk.cu).
__global__ void __launch_bounds__(256) topk_merge_metrics_kernel(EvalParams p, const unsigned long long* __restrict__ lists,
                                                                 int segments, int kpad) {
  __shared__ unsigned long long cand[CAND];
  __shared__ unsigned char hit[CAND];
  __shared__ unsigned s_distinct;
  const int row = blockIdx.x;
  if (row >= p.nu) return;
  const int uid = p.user_ids[row];
  const int n = segments * kpad;
  const unsigned long long* src = lists + (size_t)row * n;
  for (int i = threadIdx.x; i < CAND; i += blockDim.x) cand[i] = i < n ? src[i] : 0ull;
  __syncthreads();
  sort_and_score(p, row, uid, min(p.max_k, p.num_items), cand, hit, &s_distinct);
}

__global__ void row_ptr_kernel(const int* __restrict__ sorted_keys, int n, int nrows, int* __restrict__ ptr) {
  // ptr[r] = first position whose key >= r  (keys sorted ascending)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int prev = (i == 0) ? -1 : sorted_keys[i - 1];
  const int cur = (i == n) ? nrows : sorted_keys[i];
  for (int r = prev + 1; r <= cur && r <= nrows; ++r) ptr[r] = i;
}
__global__ void gather_other_kernel(const int* __restrict__ tup, const int* __restrict__ other, int n,
                                    int* __restrict__ col) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) col[i] = other[tup[i]];
}
__global__ void iota_kernel(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

}  // namespace

void launch_merge_metrics(const EvalParams& p, const unsigned long long* lists, int segments, cudaStream_t s,
                          long long* launches) {
  if (p.nu <= 0) return;
  topk_merge_metrics_kernel<<<p.nu, 256, 0, s>>>(p, lists, segments, FRX_TOPK_PAD);
  if (launches) ++*launches;
}

void launch_evaluate(const EvalParams& p, cudaStream_t s, int num_sms, long long* launches) {
  (void)num_sms;
  for (int c0 = 0; c0 < p.nu; c0 += p.chunk_users) {
    const int cn = (p.nu - c0 < p.chunk_users) ? p.nu - c0 : p.chunk_users;
    dim3 grid((p.num_items + ST - 1) / ST, (cn + ST - 1) / ST);
    scores_kernel<<<grid, 256, 0, s>>>(p.Ut + (size_t)c0 * p.d, cn, p.V, p.num_items, p.d, p.scores);
    topk_metrics_kernel<<<cn, 256, 0, s>>>(p, c0, cn);
    if (launches) *launches += 2;
  }
}

// Stable sort of the tuple indices by row id: row r lists its tuples in file
// order, exactly as by_user_[user].push_back({item, num_tuples_}) builds them.
void build_csr(const int* d_keys, const int* d_other, int n, int nrows, int* ptr, int* col, int* tup,
               cudaStream_t s, long long* launches) {
  int *keys_sorted = nullptr, *iota = nullptr;
  cudaMallocAsync(&keys_sorted, sizeof(int) * (size_t)(n > 0 ? n : 1), s);
  cudaMallocAsync(&iota, sizeof(int) * (size_t)(n > 0 ? n : 1), s);
  if (n > 0) {
    iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(iota, n);
    size_t tmp_bytes = 0;
    int bits = 1;
    while ((1ll << bits) < (long long)nrows + 1) ++bits;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, keys_sorted, iota, tup, n, 0, bits, s);
    void* tmp = nullptr;
    cudaMallocAsync(&tmp, tmp_bytes > 0 ? tmp_bytes : 1, s);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, keys_sorted, iota, tup, n, 0, bits, s);
    cudaFreeAsync(tmp, s);
    gather_other_kernel<<<(n + 255) / 256, 256, 0, s>>>(tup, d_other, n, col);
  }
  row_ptr_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(keys_sorted, n, nrows, ptr);
  cudaFreeAsync(keys_sorted, s);
  cudaFreeAsync(iota, s);
  if (launches) *launches += 4;
}

}  // namespace frx

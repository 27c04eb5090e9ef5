// Dual-form ("Woodbury") row kernel on tcgen05 / TMEM for rows with n <= 128 history entries, D = 128 / 256.
//
// The reference solves, per row,  (alpha*G + beta*I + sum_i s_i e_i e_i^T) x = sum_i q_i e_i  with a d x d
// Cholesky (ials.h:88-144, safer2.h:104-163, safer2.h:166-221).  With G = H T H^T (Householder tridiagonal
// form, frx_eig.cu), the rotated factors Et = E*H and T_r = alpha*T + beta*I = L Dl L^T (bidiagonal L: O(d) per
// row, wb_row_factor_kernel) the push-through identity gives the SAME x as
//     x = H * T_r^-1 * Ft^T * y,     (I + Fh Fh^T) y = t,
//     Ft = diag(sqrt(s)) Et[hist],   Fh = Ft * L^-T * Dl^(-1/2),   t_i = q_i / sqrt(s_i)
// i.e. an n x n SPD system instead of a d x d one: n^2 d SYRK flops instead of n d^2, n^3/3 Cholesky flops
// instead of d^3/3, and 128 TMEM columns per system instead of 384, so that several systems are in flight
// per SM and the latency-bound factorisation of one hides under the tensor-core work of the others.
//
// Work unit = a GROUP of up to four 32-entry slots (128 entries): one row of 97..128 entries, or several
// shorter rows packed side by side (their cross blocks in the 128 x 128 product are never read).
// Roles inside the persistent CTA (one per SM, 18 warps):
//   scheduler warp  takes group ids from a global atomic counter and publishes the group's slot descriptors
//                   in a small shared-memory ring;
//   4 loader warps  warp s = slot s of every group, its 32-feature chunks in order: the chunk of the 32 entries
//                   arrives by line-coalesced cp.async in a swizzled per-warp buffer (next chunk in flight),
//                   is read back with lane = history entry (128 bytes of the rotated factor row), the
//                   bidiagonal forward recurrence w_j = sqrt(s) e_j - l_j w_(j-1) along the features, scaled by
//                   Dl^(-1/2), split into tf32 hi + lo and stored as K-major SWIZZLE_128B operand tiles
//                   [128 entries][32 features] (no transpose: the contraction runs over the features here);
//   MMA warp        C += Fh_chunk Fh_chunk^T as hi*hi + hi*lo + lo*hi (3xTF32), M = N = 128, into one of three
//                   128-column TMEM accumulators; 3 operand stages, mbarrier full/empty pipeline;
//   3 solver sets   (4 warps each, warp w <-> slot w <-> TMEM lanes 32w..32w+31, thread = system row):
//                   blocked right-looking Cholesky of the row's diagonal block(s) with 32-wide panels, forward
//                   substitution fused, factor kept in TMEM, back substitution with inv(L11) per panel, then
//                   g = sum_i sqrt(s_i) y_i Et[c_i]  (second gather, lane = feature) and xt = T_r^-1 g by the two
//                   bidiagonal sweeps; a GEMM with H^T (frx_gemm.cu) takes xt back to the original basis.
#include "frx_tc_common.cuh"

namespace frx {

using namespace tc;

namespace {

constexpr int WB_NACC = 3;     // accumulator slots = solver sets = groups in flight
constexpr int WB_NSTAGE = 3;   // operand stages
constexpr int WB_RING = 4;     // group descriptors in flight
constexpr int WB_MMA_WARP = 4 * WB_NACC;
constexpr int WB_SCHED_WARP = WB_MMA_WARP + 1;
constexpr int WB_LOADER0 = WB_SCHED_WARP + 1;
constexpr int WB_NLOADER = 4;   // loader warp s serves slot s of every group
constexpr int WB_THREADS = (WB_LOADER0 + WB_NLOADER) * 32;

constexpr int WB_TILE_BYTES = 128 * 128;           // [128 entries][32 features] fp32
constexpr int WB_STAGE_BYTES = 2 * WB_TILE_BYTES;  // hi + lo

struct WbRing {  // one published group
  int gid;
  int desc[4];   // (wb row index << 2) | 32-entry chunk of the row, or -1
  int row[4], n[4], beg[4];
  float bscale[4];
};

struct WbSet {  // per solver set / accumulator slot (a multiple of 1024 bytes: the tiles must be 1024-aligned)
  // tf32 hi / lo operand tiles [128 rows][32] (K-major, SWIZZLE_128B) of the L panel for the tensor-core trailing
  // update.  The 4 KB of rows 32w..32w+31 of the hi tile double as the transposed diagonal factor of warp w while
  // the rows below run their triangular solve; after the factorisation the hi tile holds g, l, Dl^(-1/2), xt.
  float tile[2][4096];
  int cidx[128];
  float sqs[128];       // sqrt(s_i)
  float tvec[128];      // rhs t_i
  float coef[128];      // sqrt(s_i) * y_i
  float y1[128];        // forward-substitution result of each panel
  float rd[4][32];      // reciprocal diagonal of each factored block
  float rs[4][32];      // residuals of the panel being back-substituted
  float corr[4][3][32]; // back substitution: [producer slot][target slot < producer] L^T y contributions
};

template <int D>
struct WbLayout {
  static constexpr int KC = D / 32;  // feature chunks per group
  static constexpr int kStagesOff = 0;
  static constexpr int kSetOff = WB_NSTAGE * WB_STAGE_BYTES;
  static constexpr int kRingOff = kSetOff + WB_NACC * (int)sizeof(WbSet);
  static constexpr int kSdOff = ((kRingOff + WB_RING * (int)sizeof(WbRing) + 15) / 16) * 16;
  static constexpr int kGbOff = ((kSdOff + WB_NLOADER * 64 * 4 + 127) / 128) * 128;  // gather buffers, 4 KB per loader warp
  static constexpr int kBarOff = kGbOff + WB_NLOADER * 4096;
  static constexpr int kNumBars = 2 * WB_NSTAGE + 4 * WB_NACC + 2 * WB_RING;
  static constexpr int kTotal = kBarOff + kNumBars * 8;
};

__device__ __forceinline__ void set_barrier(int set) {
  asm volatile("bar.sync %0, 128;" ::"r"(set + 1) : "memory");
}

template <int D>
__global__ void __launch_bounds__(WB_THREADS, 1) row_solve_wb_kernel(RowParams p, WbParams q) {
  using L = WbLayout<D>;
  constexpr int KC = L::KC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WbSet* sets = reinterpret_cast<WbSet*>(sm + L::kSetOff);
  WbRing* ring = reinterpret_cast<WbRing*>(sm + L::kRingOff);
  float* sdS = reinterpret_cast<float*>(sm + L::kSdOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::kBarOff);
  uint64_t* full_bar = bars;                           // [NSTAGE] 4 unit arrivals
  uint64_t* empty_bar = full_bar + WB_NSTAGE;          // [NSTAGE] tcgen05.commit
  uint64_t* acc_full = empty_bar + WB_NSTAGE;          // [NACC] tcgen05.commit after the group's last chunk
  uint64_t* acc_empty = acc_full + WB_NACC;            // [NACC] 4 solver warps
  uint64_t* gq_full = acc_empty + WB_NACC;             // [RING] scheduler
  uint64_t* gq_empty = gq_full + WB_RING;              // [RING] loaders + MMA warp + the 4 warps of the owning set
  uint64_t* acc_meta = gq_empty + WB_RING;             // [NACC] the per-slot arrays (cidx, sqs, tvec) are written
  uint64_t* upd_bar = acc_meta + WB_NACC;              // [NACC] trailing update of the set's system complete
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < WB_NSTAGE; ++i) { mbar_init(&full_bar[i], 4); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < WB_NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); mbar_init(&acc_meta[i], 1); mbar_init(&upd_bar[i], 1); }
    for (int i = 0; i < WB_RING; ++i) { mbar_init(&gq_full[i], 1); mbar_init(&gq_empty[i], WB_NLOADER + 1 + 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WB_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sm_addr = smem_u32(sm);
  // optional cycle counters (FRX_TC_DEBUG): solver set 0 / warp 0 and the MMA warp
  long long tstamp = clock64();
#define WB_LAP(slot) do { if (p.dbg && lane == 0 && (warp == 0 || warp == WB_MMA_WARP)) { const long long now_ = clock64(); atomicAdd(p.dbg + (slot), (unsigned long long)(now_ - tstamp)); tstamp = now_; } } while (0)

  if (warp == WB_SCHED_WARP) {
    // ======================= scheduler =======================
    int sentinels = 0;
    for (uint32_t gseq = 0;; ++gseq) {
      int gid = 0;
      if (lane == 0) {
        gid = sentinels ? q.num_groups : atomicAdd(q.counter, 1);
        if (gid >= q.num_groups) gid = -1;
      }
      gid = __shfl_sync(0xffffffffu, gid, 0);
      const uint32_t rsi = gseq % WB_RING, ruse = gseq / WB_RING;
      if (ruse > 0) mbar_wait(&gq_empty[rsi], (ruse - 1) & 1);
      WbRing& e = ring[rsi];
      if (lane == 0) e.gid = gid;
      if (gid >= 0 && lane < 4) {
        const int desc = __ldg(q.grp_slots + (size_t)gid * 4 + lane);
        e.desc[lane] = desc;
        if (desc >= 0) {
          const int r = __ldg(q.wb_rows + (desc >> 2));
          const int beg = __ldg(p.ptr + r), n = __ldg(p.ptr + r + 1) - beg;
          const RowScalars s = row_scalars(p, r, n);
          e.row[lane] = r; e.n[lane] = n; e.beg[lane] = beg;
          e.bscale[lane] = s.bscale;
        } else {
          e.row[lane] = -1; e.n[lane] = 0; e.beg[lane] = 0;
          e.bscale[lane] = 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&gq_full[rsi]);
      if (gid < 0 && ++sentinels == WB_NACC) break;
    }
  } else if (warp >= WB_LOADER0) {
    // ======================= loaders =======================
    // Loader warp s fills slot s of every group, chunk by chunk in feature order (the bidiagonal recurrence runs
    // along the features).  Every loader warp therefore passes through every use of every stage: the parity
    // waits below are only valid for a waiter that is at most one phase behind the barrier.
    const int lw = warp - WB_LOADER0;
    float* sd = sdS + lw * 64;  // [0,32) l_j, [32,64) Dl_j^(-1/2) of the current chunk
    for (uint32_t gseq = 0;; ++gseq) {
      const uint32_t rsi = gseq % WB_RING;
      mbar_wait(&gq_full[rsi], (gseq / WB_RING) & 1);
      const WbRing& e = ring[rsi];
      if (e.gid < 0) break;
      const uint32_t slot = gseq % WB_NACC, ause = gseq / WB_NACC;
      WbSet& S = sets[slot];
      {
        const int s = lw;
        const int desc = e.desc[s];
        const int n = e.n[s], beg = e.beg[s];
        const int en = 32 * (desc & 3) + lane;  // entry index within the row
        const bool valid = desc >= 0 && en < n;
        const int mn = 32 * s + lane;
        int c = 0;
        float sw = 0.f, qw = 0.f;
        if (valid) {
          c = __ldg(p.col + beg + en);
          sw = 1.f; qw = 1.f;
          if (p.mode == RM_SAFER_V) { const float w = __ldg(p.entry_w + c); sw = w; qw = w; }
        }
        const float sq = sqrtf(sw);
        const float* lrow = q.lsub + (size_t)(desc >= 0 ? (desc >> 2) : 0) * D;
        const float* rrow = q.rsd + (size_t)(desc >= 0 ? (desc >> 2) : 0) * D;
        // The 32-feature chunk of the 32 entries travels by LINE-coalesced cp.async into this warp's gather
        // buffer (8 lanes per 128 B line: with lane = entry every load instruction touched 32 different lines and
        // the L1TEX tag stage ran at 71 %), XOR-swizzled, and is read back with lane = entry; chunk k+1 is in
        // flight while chunk k runs through the recurrence.
        const uint32_t gb = sm_addr + L::kGbOff + (uint32_t)lw * 4096u;
        const int lg = lane >> 3, lch = lane & 7;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        int cidx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cidx[j] = __shfl_sync(0xffffffffu, c, 8 * lg + j);
        auto issue_chunk = [&](int k) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int en = 8 * lg + j;
            const bool ok = (vmask >> en) & 1u;
            const float* srcp = q.Et + (size_t)(ok ? cidx[j] : 0) * D + 32 * k + 4 * lch;
            const uint32_t dst = gb + (uint32_t)en * 128u + ((uint32_t)(lch ^ (en & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(srcp), "r"(ok ? 16 : 0) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float lnext = 0.f, rnext = 0.f;
        if (desc >= 0) {
          lnext = __ldg(lrow + lane);
          rnext = __ldg(rrow + lane);
          issue_chunk(0);
        }
        float wprev = 0.f;  // w_(j-1) of the recurrence, carried across chunks
        for (int k = 0; k < KC; ++k) {
          const uint32_t cs = gseq * KC + (uint32_t)k, st = cs % WB_NSTAGE, use = cs / WB_NSTAGE;
          if (use > 0) mbar_wait(&empty_bar[st], (use - 1) & 1);
          if (desc >= 0) {
            if (k == 0 && ause > 0) mbar_wait(&acc_empty[slot], (ause - 1) & 1);  // per-slot arrays are free again
            __syncwarp();
            sd[lane] = lnext;
            sd[32 + lane] = rnext;
            float4 vc[8];
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(vc[j].x), "=f"(vc[j].y), "=f"(vc[j].z), "=f"(vc[j].w)
                           : "r"(gb + (uint32_t)lane * 128u + ((uint32_t)(j ^ (lane & 7)) << 4))
                           : "memory");
            __syncwarp();  // every lane has its chunk: the buffer may be refilled
            if (k + 1 < KC) {  // next chunk in flight while this one is processed
              lnext = __ldg(lrow + 32 * (k + 1) + lane);
              rnext = __ldg(rrow + 32 * (k + 1) + lane);
              issue_chunk(k + 1);
            }
            __syncwarp();
            uint8_t* hi_tile = sm + st * WB_STAGE_BYTES;
            uint8_t* lo_tile = hi_tile + WB_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 l4 = reinterpret_cast<const float4*>(sd)[j];
              const float4 r4 = reinterpret_cast<const float4*>(sd + 32)[j];
              float x4[4];
              wprev = fmaf(-l4.x, wprev, sq * vc[j].x); x4[0] = wprev * r4.x;
              wprev = fmaf(-l4.y, wprev, sq * vc[j].y); x4[1] = wprev * r4.y;
              wprev = fmaf(-l4.z, wprev, sq * vc[j].z); x4[2] = wprev * r4.z;
              wprev = fmaf(-l4.w, wprev, sq * vc[j].w); x4[3] = wprev * r4.w;
              float hi[4], lo[4];
#pragma unroll
              for (int t4 = 0; t4 < 4; ++t4) {
                hi[t4] = __uint_as_float(__float_as_uint(x4[t4]) & 0xffffe000u);
                lo[t4] = x4[t4] - hi[t4];
              }
              const uint32_t off = tile_chunk_off(mn, j);
              *reinterpret_cast<float4*>(hi_tile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<float4*>(lo_tile + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (k == 0) {
              S.cidx[mn] = c;
              S.sqs[mn] = sq;
              S.tvec[mn] = (valid && sw > 0.f) ? e.bscale[s] * qw / sq : 0.f;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[st]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&gq_empty[rsi]);
    }
  } else if (warp == WB_MMA_WARP) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc = make_idesc_tf32(128);
    for (uint32_t gseq = 0;; ++gseq) {
      const uint32_t rsi = gseq % WB_RING;
      mbar_wait(&gq_full[rsi], (gseq / WB_RING) & 1);
      const int gid = ring[rsi].gid;
      __syncwarp();
      if (lane == 0) mbar_arrive(&gq_empty[rsi]);
      if (gid < 0) break;
      const uint32_t slot = gseq % WB_NACC, ause = gseq / WB_NACC;
      WB_LAP(7);
      if (ause > 0) mbar_wait(&acc_empty[slot], (ause - 1) & 1);
      tc_fence_after();
      WB_LAP(4);
      const uint32_t d_tmem = tmem_base + 128u * slot;
      for (int k = 0; k < KC; ++k) {
        const uint32_t cs = gseq * KC + (uint32_t)k, st = cs % WB_NSTAGE, use = cs / WB_NSTAGE;
        mbar_wait(&full_bar[st], use & 1);
        tc_fence_after();
        WB_LAP(5);
        // chunk 0 carries the per-slot arrays: hand the loaders' writes on to the solver set (acquire above,
        // release here: the ordering is transitive)
        if (k == 0 && lane == 0) mbar_arrive(&acc_meta[slot]);
        if (lane == 0) {
          const uint32_t hi_addr = sm_addr + st * WB_STAGE_BYTES, lo_addr = hi_addr + WB_TILE_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dhi = make_kmajor_desc(hi_addr + ks * 32), dlo = make_kmajor_desc(lo_addr + ks * 32);
            umma_tf32(d_tmem, dhi, dhi, idesc, (k | ks) ? 1u : 0u);
            umma_tf32(d_tmem, dhi, dlo, idesc, 1u);
            umma_tf32(d_tmem, dlo, dhi, idesc, 1u);
          }
          umma_commit(&empty_bar[st]);
          if (k == KC - 1) umma_commit(&acc_full[slot]);
        }
        __syncwarp();
        WB_LAP(6);
      }
    }
  } else {
    // ======================= solver sets =======================
    const int set = warp >> 2, w = warp & 3;
    WbSet& S = sets[set];
    const uint32_t tbase = tmem_base + ((uint32_t)(32 * w) << 16) + 128u * (uint32_t)set;  // block (w, j) at + 32 j
    uint32_t upd_n = 0;  // trailing updates committed by this set so far
    for (uint32_t gseq = (uint32_t)set;; gseq += WB_NACC) {
      const uint32_t rsi = gseq % WB_RING;
      mbar_wait(&gq_full[rsi], (gseq / WB_RING) & 1);
      const WbRing e = ring[rsi];  // private copy: the ring entry is released right away
      __syncwarp();
      if (lane == 0) mbar_arrive(&gq_empty[rsi]);
      if (e.gid < 0) break;
      const uint32_t ause = gseq / WB_NACC;
      int nsteps = 0;
#pragma unroll
      for (int s = 0; s < 4; ++s) nsteps = max(nsteps, e.desc[s] >= 0 ? (e.n[s] + 31) >> 5 : 0);
      int my_desc = -1, my_n = 0;
#pragma unroll
      for (int s = 0; s < 4; ++s) if (s == w) { my_desc = e.desc[s]; my_n = e.n[s]; }
      const bool used = my_desc >= 0;
      const int rel = my_desc & 3;             // my slot's position within its row
      const int s0 = w - rel;                  // first slot of my row
      const int m = (my_n + 31) >> 5;          // slots (panels) of my row
      mbar_wait(&acc_meta[set], ause & 1);
      mbar_wait(&acc_full[set], ause & 1);
      tc_fence_after();
      WB_LAP(0);
      float b_reg = used ? S.tvec[32 * w + lane] : 0.f;  // t_i, then y1_i, then y_i
      float a[32];

      // ---- blocked right-looking Cholesky of (I + Fh Fh^T) restricted to my row's slots: all rows of the group
      //      advance panel by panel in lock step; diagonal blocks and triangular solves in registers (thread =
      //      row), the trailing update C -= L21 L21^T of ALL rows of the group as ONE tensor-core product per
      //      step (3xTF32, a_negate): warps that are not below a panel contribute zero rows ----
      float* tile_hi = S.tile[0];
      uint8_t* tile_hi_b = reinterpret_cast<uint8_t*>(S.tile[0]);
      uint8_t* tile_lo_b = reinterpret_cast<uint8_t*>(S.tile[1]);
      for (int k = 0; k < nsteps; ++k) {
        const int pn = s0 + k;  // slot of the panel being eliminated in my row
        const bool below = used && rel > k;
        if (used && rel == k) {
          // diagonal block: lane = row, column k scaled, published (transposed) and swept; rhs carried along
          uint32_t u[32];
          FRX_TMEM_LD32(u, tbase + 32u * (uint32_t)w);
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(u[j]) + (j == lane ? 1.f : 0.f);  // + I
          float* LdT = tile_hi + 1024 * w;
          float* rd = S.rd[w];
          float rs;
          bool bad_pivot;
          {
            const float akk = __shfl_sync(0xffffffffu, a[0], 0);
            bad_pivot = !(akk > 0.f);
            rs = fast_rsqrt(bad_pivot ? 1.f : akk);
          }
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            const float lk = a[kk] * rs;
            LdT[kk * 32 + lane] = lane >= kk ? lk : 0.f;
            float rs_next = 0.f;
            if (kk + 1 < 32) {
              const float dcand = fmaf(-lk, lk, a[kk + 1]);
              const float akk = __shfl_sync(0xffffffffu, dcand, kk + 1);
              bad_pivot |= !(akk > 0.f);
              rs_next = fast_rsqrt(akk);
              const float lnext = __shfl_sync(0xffffffffu, lk, kk + 1);
              a[kk + 1] = fmaf(-lk, lnext, a[kk + 1]);
            }
            if (lane == kk) rd[kk] = rs;
            a[kk] = lane >= kk ? lk : 0.f;
            const float yk = __shfl_sync(0xffffffffu, b_reg, kk) * rs;
            if (lane == kk) b_reg = yk;
            else if (lane > kk) b_reg = fmaf(-lk, yk, b_reg);
            __syncwarp();
            const unsigned long long nlk2 = pack2(-lk, -lk);
            const float4* col = reinterpret_cast<const float4*>(LdT + kk * 32);
#pragma unroll
            for (int m4 = (kk + 2) / 4; m4 < 8; ++m4) {
              const float4 c = col[m4];
              FRX_SWEEP4(a, 4 * m4, kk + 1, lk, nlk2, c);
            }
            rs = rs_next;
          }
          if (bad_pivot && lane == 0) atomicExch(p.status, 1);
          S.y1[32 * w + lane] = b_reg;
        }
        set_barrier(set);  // A0: the diagonal factors of this step are published
        if (below) {
          // rows below: L21 row by forward substitution against the transposed L11 of slot pn
          uint32_t u[32];
          FRX_TMEM_LD32(u, tbase + 32u * (uint32_t)pn);
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(u[j]);
          const float* LdT = tile_hi + 1024 * pn;
          const float* rd = S.rd[pn];
#pragma unroll
          for (int kk = 0; kk < 32; ++kk) {
            const float l = a[kk] * rd[kk];
            a[kk] = l;
            const unsigned long long nl2 = pack2(-l, -l);
            const float4* col = reinterpret_cast<const float4*>(LdT + kk * 32);
#pragma unroll
            for (int m4 = (kk + 1) / 4; m4 < 8; ++m4) {
              const float4 c = col[m4];
              FRX_SWEEP4(a, 4 * m4, kk, l, nl2, c);
            }
          }
          float dot = 0.f;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 y4 = reinterpret_cast<const float4*>(S.y1 + 32 * pn)[k4];
            dot = fmaf(a[4 * k4], y4.x, dot);
            dot = fmaf(a[4 * k4 + 1], y4.y, dot);
            dot = fmaf(a[4 * k4 + 2], y4.z, dot);
            dot = fmaf(a[4 * k4 + 3], y4.w, dot);
          }
          b_reg -= dot;
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(a[j]);
          FRX_TMEM_ST32(tbase + 32u * (uint32_t)pn, u);  // the factor stays in TMEM for the back substitution
        } else if (used && rel == k) {
          // meanwhile the diagonal warp inverts its L11 (lane c = column c of the inverse) for the back substitution
          const float* LdT = tile_hi + 1024 * w;
          const float* rd = S.rd[w];
          float x[32];
#pragma unroll
          for (int i2 = 0; i2 < 32; ++i2) x[i2] = (i2 == lane) ? 1.f : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            x[j] *= rd[j];
            const float xj = x[j];
            const unsigned long long nx2 = pack2(-xj, -xj);
            const float4* col = reinterpret_cast<const float4*>(LdT + j * 32);
#pragma unroll
            for (int m4 = (j + 1) / 4; m4 < 8; ++m4) {
              const float4 c = col[m4];
              FRX_SWEEP4(x, 4 * m4, j, xj, nx2, c);
            }
          }
          uint32_t u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(x[j]);
          FRX_TMEM_ST32(tbase + 32u * (uint32_t)w, u);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (k + 1 < nsteps) {  // uniform over the set: some row of the group still has a trailing block
          set_barrier(set);    // A1: the transposed factors have been read, their rows of the tile may be rewritten
          {
            const int mn = 32 * w + lane;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              float hi[4], lo[4];
#pragma unroll
              for (int t4 = 0; t4 < 4; ++t4) {
                const float x = below ? a[4 * c4 + t4] : 0.f;
                hi[t4] = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                lo[t4] = x - hi[t4];
              }
              const uint32_t off = tile_chunk_off(mn, c4);
              *reinterpret_cast<float4*>(tile_hi_b + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<float4*>(tile_lo_b + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          set_barrier(set);    // A2: the L panel tile is complete, the factor blocks are in TMEM
          if (w == 0 && lane == 0) {
            tc_fence_after();
            constexpr uint32_t idesc_neg = make_idesc_tf32(128, 1);
            const uint32_t hi_addr = smem_u32(tile_hi_b), lo_addr = smem_u32(tile_lo_b);
            const uint32_t d_tmem = tmem_base + 128u * (uint32_t)set;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t dhi = make_kmajor_desc(hi_addr + ks * 32), dlo = make_kmajor_desc(lo_addr + ks * 32);
              umma_tf32(d_tmem, dhi, dhi, idesc_neg, 1u);
              umma_tf32(d_tmem, dhi, dlo, idesc_neg, 1u);
              umma_tf32(d_tmem, dlo, dhi, idesc_neg, 1u);
            }
            umma_commit(&upd_bar[set]);
          }
          mbar_wait(&upd_bar[set], upd_n & 1);
          ++upd_n;
          tc_fence_after();
        }
      }

      WB_LAP(1);
      // ---- back substitution L^T y = y1, panel by panel from the bottom ----
      for (int k = nsteps - 1; k >= 0; --k) {
        if (used && rel == k) {
          uint32_t u[32];
          FRX_TMEM_LD32(u, tbase + 32u * (uint32_t)w);  // lane c: column c of inv(L11)
          S.rs[w][lane] = b_reg;
          __syncwarp();
          float y0 = 0.f, y1v = 0.f, y2 = 0.f, y3 = 0.f;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 r4 = reinterpret_cast<const float4*>(S.rs[w])[j4];
            y0 = fmaf(__uint_as_float(u[4 * j4]), r4.x, y0);
            y1v = fmaf(__uint_as_float(u[4 * j4 + 1]), r4.y, y1v);
            y2 = fmaf(__uint_as_float(u[4 * j4 + 2]), r4.z, y2);
            y3 = fmaf(__uint_as_float(u[4 * j4 + 3]), r4.w, y3);
          }
          b_reg = (y0 + y1v) + (y2 + y3);
          for (int j = s0; j < w; ++j) {  // contributions L[w,j]^T y_w to the panels on the left
            FRX_TMEM_LD32(u, tbase + 32u * (uint32_t)j);
            float v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(u[c]) * b_reg;
            transpose_reduce<32>(v, lane);
            S.corr[w][j][lane] = v[0];
          }
        }
        set_barrier(set);
        if (used && rel < k && k < m) b_reg -= S.corr[s0 + k][w][lane];
      }
      S.coef[32 * w + lane] = used ? S.sqs[32 * w + lane] * b_reg : 0.f;
      tc_fence_before();
      set_barrier(set);
      WB_LAP(2);

      // ---- g = sum_i coef_i Et[c_i]  (lane = feature; warp w takes features [w*D/4, (w+1)*D/4)) ----
      // the hi tile is free now: g, l, Dl^(-1/2), xt, each [4 row slots][D]
      constexpr int FW = D / 4, F = FW / 32;
      float* gS = S.tile[0];
      float* lS = S.tile[0] + 1024;
      float* rS = S.tile[0] + 2048;
      float* xS = S.tile[0] + 3072;
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        int desc = -1, n = 0;
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) if (s2 == s) { desc = e.desc[s2]; n = e.n[s2]; }
        if (desc < 0 || (desc & 3) != 0) continue;  // rows start at their chunk 0
        const int f0 = w * FW + lane * F;
        float acc[F];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = 0.f;
        const int* ci = S.cidx + 32 * s;
        const float* cf = S.coef + 32 * s;
        int en = 0;
        for (; en + 8 <= n; en += 8) {
          float vv[8][F];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float* src = q.Et + (size_t)ci[en + t] * D + f0;
            if (F == 2) { const float2 x2 = __ldg(reinterpret_cast<const float2*>(src)); vv[t][0] = x2.x; vv[t][F - 1] = x2.y; }
            else if (F == 4) { const float4 x4 = __ldg(reinterpret_cast<const float4*>(src)); vv[t][0] = x4.x; vv[t][1 % F] = x4.y; vv[t][2 % F] = x4.z; vv[t][3 % F] = x4.w; }
            else vv[t][0] = __ldg(src);
          }
#pragma unroll
          for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] = fmaf(cf[en + t], vv[t][f], acc[f]);
        }
        for (; en < n; ++en) {
          const float* src = q.Et + (size_t)ci[en] * D + f0;
#pragma unroll
          for (int f = 0; f < F; ++f) acc[f] = fmaf(cf[en], __ldg(src + f), acc[f]);
        }
        const size_t rowoff = (size_t)(desc >> 2) * D;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          gS[s * D + f0 + f] = acc[f];
          lS[s * D + f0 + f] = __ldg(q.lsub + rowoff + f0 + f);
          rS[s * D + f0 + f] = __ldg(q.rsd + rowoff + f0 + f);
        }
      }
      set_barrier(set);
      // ---- xt = L^-T Dl^-1 L^-1 g: two bidiagonal sweeps, one lane per row ----
      if (w == 0 && lane < 4) {
        int desc = -1;
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) if (s2 == lane) desc = e.desc[s2];
        if (desc >= 0 && (desc & 3) == 0) {
          const float* g = gS + lane * D;
          const float* l = lS + lane * D;
          const float* r = rS + lane * D;
          float* x = xS + lane * D;
          float u = 0.f;
#pragma unroll 4
          for (int j4 = 0; j4 < D / 4; ++j4) {
            const float4 g4 = reinterpret_cast<const float4*>(g)[j4];
            const float4 l4 = reinterpret_cast<const float4*>(l)[j4];
            const float4 r4 = reinterpret_cast<const float4*>(r)[j4];
            float4 o;
            u = fmaf(-l4.x, u, g4.x); o.x = u * (r4.x * r4.x);
            u = fmaf(-l4.y, u, g4.y); o.y = u * (r4.y * r4.y);
            u = fmaf(-l4.z, u, g4.z); o.z = u * (r4.z * r4.z);
            u = fmaf(-l4.w, u, g4.w); o.w = u * (r4.w * r4.w);
            reinterpret_cast<float4*>(x)[j4] = o;
          }
          float xn = 0.f, ln = 0.f;  // x_(j+1), l_(j+1)
#pragma unroll 4
          for (int j4 = D / 4 - 1; j4 >= 0; --j4) {
            float4 o = reinterpret_cast<const float4*>(x)[j4];
            const float4 l4 = reinterpret_cast<const float4*>(l)[j4];
            o.w = fmaf(-ln, xn, o.w);
            o.z = fmaf(-l4.w, o.w, o.z);
            o.y = fmaf(-l4.z, o.z, o.y);
            o.x = fmaf(-l4.y, o.y, o.x);
            xn = o.x; ln = l4.x;
            reinterpret_cast<float4*>(x)[j4] = o;
          }
        }
      }
      set_barrier(set);
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        int desc = -1;
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) if (s2 == s) desc = e.desc[s2];
        if (desc < 0 || (desc & 3) != 0) continue;
        float* dst = q.Xt + (size_t)(desc >> 2) * D;
        for (int f = w * 32 + lane; f < D; f += 128) dst[f] = xS[s * D + f];
      }
      __syncwarp();
      WB_LAP(3);
      if (lane == 0) mbar_arrive(&acc_empty[set]);  // TMEM slot and the per-slot arrays are free
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WB_MMA_WARP)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// Per-row factors of T_r = alpha*T + beta*I = L Dl L^T (L unit lower bidiagonal): lsub[j] = L[j][j-1]
// (lsub[0] = 0) and rsd[j] = Dl_j^(-1/2), for every row of the dual-form path.  Lane = row (the recurrence is
// sequential in j); 32 x 32 tiles go through shared memory so that the rows are written coalesced.
template <int D>
__global__ void __launch_bounds__(128) wb_row_factor_kernel(RowParams p, WbParams q, int nwb) {
  __shared__ float tdS[D], tsS[D];
  __shared__ float tile[4][2][32][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += 128) { tdS[i] = q.tdiag[i]; tsS[i] = q.tsub[i]; }
  __syncthreads();
  const int i0 = (blockIdx.x * 4 + warp) * 32;
  if (i0 >= nwb) return;
  const int i = i0 + lane;
  float alpha = 0.f, beta = 1.f;
  if (i < nwb) {
    const int r = __ldg(q.wb_rows + i);
    const RowScalars s = row_scalars(p, r, __ldg(p.ptr + r + 1) - __ldg(p.ptr + r));
    alpha = s.alpha; beta = s.beta;
  }
  float delta = 1.f;
  bool bad = false;
  for (int c = 0; c < D / 32; ++c) {
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const int j = 32 * c + jj;
      const float ab = alpha * tsS[j];          // tsub[0] = 0
      const float l = j ? ab / delta : 0.f;
      delta = fmaf(alpha, tdS[j], beta) - ab * l;
      bad |= !(delta > 0.f);
      tile[warp][0][lane][jj] = l;
      tile[warp][1][lane][jj] = rsqrtf(delta);
    }
    __syncwarp();
    for (int rr = 0; rr < 32 && i0 + rr < nwb; ++rr) {
      q.lsub[(size_t)(i0 + rr) * D + 32 * c + lane] = tile[warp][0][rr][lane];
      q.rsd[(size_t)(i0 + rr) * D + 32 * c + lane] = tile[warp][1][rr][lane];
    }
    __syncwarp();
  }
  if (bad && i < nwb) atomicExch(p.status, 1);  // alpha*G + beta*I is not positive definite
}

template <int D>
void launch_wb_instance(const RowParams& p, const WbParams& q, int nwb, cudaStream_t s, int num_sms) {
  wb_row_factor_kernel<D><<<(nwb + 127) / 128, 128, 0, s>>>(p, q, nwb);
  const int smem = WbLayout<D>::kTotal + 1024;
  cudaFuncSetAttribute(row_solve_wb_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = num_sms < q.num_groups ? num_sms : q.num_groups;
  row_solve_wb_kernel<D><<<grid, WB_THREADS, smem, s>>>(p, q);
}

}  // namespace

bool row_solve_wb_supported(const RowParams& p) {
  const bool mode_ok = p.mode == RM_IALS || p.mode == RM_SAFER_U || p.mode == RM_SAFER_V;
  return mode_ok && p.cs == 0 && p.bd == p.d && (p.d == 128 || p.d == 256);
}

void launch_row_solve_wb(const RowParams& p, const WbParams& q, int nwb, cudaStream_t s, int num_sms,
                         long long* launches) {
  if (q.num_groups <= 0 || nwb <= 0) return;
  cudaMemsetAsync(q.counter, 0, sizeof(int), s);
  if (p.d == 256) launch_wb_instance<256>(p, q, nwb, s, num_sms);
  else launch_wb_instance<128>(p, q, nwb, s, num_sms);
  if (launches) *launches += 2;
}

}  // namespace frx

// Fused row kernel on 5th-generation tensor cores (tcgen05 + TMEM), D = 128 / 256.
//
// One persistent CTA per SM (16 warps, 128 registers per thread), one row r of the side being solved at
// a time; the d x d system never leaves the SM (DESIGN.md 4.1 has the full description and the measured
// cycle budget):
//   set-up   alpha*G + beta*I of the row is written into the TMEM accumulators BEFORE the SYRK, by the
//            warps that are idle during the back substitution of the previous row (row_scalars,
//            tmem_init_system).
//   phase A  S_r = sum_c s_c e_c e_c^T  (+ rhs = sum_c q_c e_c)
//            The indices / weights of the row are staged in shared memory; 15 loader warps take
//            (32-entry, 32-feature) units round-robin: the unit's 128-byte slabs arrive by line-coalesced
//            cp.async in a swizzled per-warp buffer one unit ahead, are read back as entry PAIRS, scaled
//            by a power of two and written as fp16 hi + lo words into K-major, 128B-swizzled operand tiles
//            [feature][64 entries]; warp 15 issues tcgen05.mma kind::f16 for hi*hi + hi*lo + lo*hi (the
//            11 + 11 significand bits of an error-compensated 3xTF32 product at twice the K per MMA and
//            half the operand bytes: fp32-level accuracy) onto the TMEM accumulators; lower-triangle M
//            blocks only.  Two operand stages, mbarrier full/empty pipeline, tcgen05.commit frees a stage.
//   phase B  right-looking blocked Cholesky in TMEM with 32-wide panels.  Thread i owns matrix row i
//            (tcgen05.ld gives it its 32 panel entries): the diagonal 32x32 block is factored by one
//            warp (columns published through shared memory, shuffle shortcut for the next pivot); the
//            rows below run their triangular solve behind that sweep (mbarrier per 8 columns); the L
//            panel is written (tf32 hi/lo) as K-major operand tiles and the trailing update
//            A22 -= L21 L21^T runs on the tensor cores with the a_negate bit, in two committed batches
//            (the next diagonal block's columns first).  The forward substitution is fused into the
//            sweep; the back substitution is column-oriented over the L panels kept in shared memory.
//   long rows (> FRX_SPLIT_MIN entries) are cut into pieces: MODE 1 dumps per-piece partial sums, MODE 2
//            solves such a row from the sum of its pieces, MODE 0 is the ordinary row.
//   GRAD     the CVaR-MF gradient steps (cvar_mf.h:88-180) take phase A as it is and replace phase B by
//            x - step * (M x - rhs), evaluated from the TMEM-resident SYRK sum.
//
// Hardware conventions were established with tools/tc_probe.cu / tc_rate.cu / lat_probe.cu on a B200:
//  * kind::tf32 and kind::f16 work with K-major operands (SWIZZLE_128B, SBO = 1024 B, K step = +32 B on
//    the descriptor start address); MN-major tf32 operands produce zeros, hence the transpose.
//  * the MMA ignores the low 13 mantissa bits of fp32 inputs (truncation).
//  * tcgen05.ld/st 32x32b: warp w touches TMEM lanes 32*(w%4)..+31, thread = accumulator row.
//  * an M=128 MMA takes 93 / 106 / 170 cycles for N <= 64 / 128 / 256; the issuing thread blocks while
//    the (shallow) tensor queue is full.
// Restates ials.h:88-144, safer2.h:104-163, safer2.h:166-221 (incl. the stale-tail quirk) and cvar_mf.h:88-180.
#include "frx_kernels.cuh"
#include "frx_tc_common.cuh"
#include <cstdint>
#include <cuda_fp16.h>

namespace frx {

using namespace tc;

namespace {

constexpr int TC_LOADER_WARPS = 16;               // 2 groups of 8
constexpr int TC_THREADS = TC_LOADER_WARPS * 32;  // 4 warps per scheduler -> 128 registers per thread
constexpr int TC_MMA_WARP = 15;                   // owns TMEM and issues the trailing updates
constexpr int KT = 32;                                  // panel width = entries per loader unit (lane = entry)
constexpr int KS = 64;                                  // entries per SYRK operand tile: 128 B rows of fp16

template <int D>
struct TcLayout {
  static constexpr int P = D / 32;                        // panels
  static constexpr int kTileBytes = D * 128;              // one operand tile: D rows x 128 B
  static constexpr int kStageBytes = 2 * kTileBytes;      // hi + lo
  static constexpr int kStagingBytes = 2 * kStageBytes;   // phase A: two stages
  // phase B (aliased with the staging area): L-panel operand tiles (hi, lo) + fp32 L panels
  static constexpr int kLstFloats = 32 * (P * D - 16 * P * (P - 1));  // sum_p 32*(D-32p)
  static constexpr int kLstOff = kStageBytes;
  static constexpr int kPhaseB = kLstOff + kLstFloats * 4;
  // phase A: per loader warp one gather buffer (32 lanes x D/8 floats) that cp.async fills one unit ahead; it
  // lies behind the operand stages and is aliased by the L panels of phase B
  static constexpr int kGatherOff = kStagingBytes;
  static constexpr int kGatherWarpBytes = 32 * 128;  // 32 entries x one 128 B slab
  // ... followed by the staged indices and weights of the work item's history entries (also phase A only)
  static constexpr int kIdxOff = kGatherOff + (TC_LOADER_WARPS - 1) * kGatherWarpBytes;
  static constexpr int kPhaseA = kIdxOff + 2 * FRX_STAGE_CAP * 4;
  static constexpr int kBigBytes = kPhaseA > kPhaseB ? kPhaseA : kPhaseB;
  static constexpr int kRhsOff = kBigBytes;               // rhs partials [2][D]
  static constexpr int kLdOff = kRhsOff + 2 * D * 4;      // transposed diagonal factor, double-buffered [2][32][32]
  static constexpr int kWsumOff = kLdOff + 2 * 32 * 32 * 4;  // per-warp column sums [P][32]
  static constexpr int kYOff = kWsumOff + P * 32 * 4;     // y of the current panel [32]
  static constexpr int kRdOff = kYOff + 32 * 4;           // reciprocal diagonal of the current factor, [2][32]
  static constexpr int kBarOff = ((kRdOff + 2 * 32 * 4 + 15) / 16) * 16;
  // G staging for tmem_init_system, 4 KB per warp: aliases operand stage 0 when that holds 64 KB (D = 256)
  static constexpr int kGbufOff = kStageBytes >= TC_LOADER_WARPS * 4096 ? 0 : kBarOff + 128;
  static constexpr int kTotal = kGbufOff == 0 ? kBarOff + 128 : kGbufOff + TC_LOADER_WARPS * 4096;
  static constexpr int kTmemCols = D == 256 ? 512 : 128;
  __host__ __device__ static constexpr int lst_off(int p) { return 32 * (p * D - 16 * p * (p - 1)); }
};

// TMEM <- alpha*G + beta*I on the lower-triangle 32x32 chunks of the d x d accumulator.  Run by the
// `nshare` warps that share TMEM lane quarter `rq` (share id `sid`); `gbuf` is a private 4 KB buffer:
// G rows arrive as coalesced 128 B segments, are stored XOR-swizzled and read back row-per-lane.
// Index of the lower-triangle chunk (M block rb, lane quarter rq, column chunk ch) in a piece dump.
__device__ __forceinline__ int piece_chunk_index(int rb, int rq, int ch) {
  return rb == 0 ? rq * (rq + 1) / 2 + ch : 10 + 5 * rq + rq * (rq - 1) / 2 + ch;
}

// With `pieces` != null the partial SYRK sums of a long row (npieces dumps of `pstride` floats, see the
// piece mode of the kernel) are added in piece order; alpha == beta == 0 and G == null just clears.
template <int D>
__device__ __forceinline__ void tmem_init_system(const float* __restrict__ G, float alpha, float beta, uint32_t tmem_base,
                                                 int rq, int sid, int nshare, float* gbuf, int lane,
                                                 const float* __restrict__ pieces = nullptr, int npieces = 0,
                                                 size_t pstride = 0) {
  constexpr int NB = D / 128;
  const int per_rb0 = rq + 1;                       // chunks with j <= i for M block 0
  const int total = NB == 2 ? 2 * rq + 6 : rq + 1;  // (rq + 1) + (4 + rq + 1)
  float4 gv[8];
  if (G == nullptr && pieces == nullptr) {  // clear only
    uint32_t z[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = 0u;
    for (int c = sid; c < total; c += nshare) {
      const int rb = c < per_rb0 ? 0 : 1, ch = c < per_rb0 ? c : c - per_rb0;
      FRX_TMEM_ST32(tmem_base + ((uint32_t)(32 * rq) << 16) + (rb ? 256u : 0u) + (uint32_t)(32 * ch), z);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    return;
  }
  auto fetch = [&](int c) {
    const int rb = c < per_rb0 ? 0 : 1, ch = c < per_rb0 ? c : c - per_rb0;
    if (G == nullptr) {  // gradient-step variant: the sum of the pieces only
#pragma unroll
      for (int it = 0; it < 8; ++it) gv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      return;
    }
    const float* gblk = G + (size_t)(128 * rb + 32 * rq) * D + 32 * ch;
#pragma unroll
    for (int it = 0; it < 8; ++it) gv[it] = __ldg(reinterpret_cast<const float4*>(gblk + (size_t)(it * 4 + (lane >> 3)) * D) + (lane & 7));
  };
  if (sid < total) fetch(sid);
  for (int c = sid; c < total; c += nshare) {
    const int rb = c < per_rb0 ? 0 : 1, ch = c < per_rb0 ? c : c - per_rb0;
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + (lane >> 3), c16 = lane & 7;
      *reinterpret_cast<float4*>(gbuf + row * 32 + ((c16 ^ (row & 7)) << 2)) = gv[it];
    }
    __syncwarp();
    if (c + nshare < total) fetch(c + nshare);  // next chunk's G is in flight while this one is written
    uint32_t u[32];
    const bool diag_chunk = (4 * rb + rq) == ch;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 g4 = *reinterpret_cast<const float4*>(gbuf + lane * 32 + ((c4 ^ (lane & 7)) << 2));
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int t4 = 0; t4 < 4; ++t4) {
        float m = alpha * gg[t4];
        if (diag_chunk && 4 * c4 + t4 == lane) m += beta;
        u[4 * c4 + t4] = __float_as_uint(m);
      }
    }
    for (int pc = 0; pc < npieces; ++pc) {  // lane = chunk row: 128 contiguous bytes per lane
      const float4* src = reinterpret_cast<const float4*>(pieces + (size_t)pc * pstride +
                                                          (size_t)piece_chunk_index(rb, rq, ch) * 1024 + lane * 32);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 s4 = __ldcs(src + c4);
        u[4 * c4 + 0] = __float_as_uint(__uint_as_float(u[4 * c4 + 0]) + s4.x);
        u[4 * c4 + 1] = __float_as_uint(__uint_as_float(u[4 * c4 + 1]) + s4.y);
        u[4 * c4 + 2] = __float_as_uint(__uint_as_float(u[4 * c4 + 2]) + s4.z);
        u[4 * c4 + 3] = __float_as_uint(__uint_as_float(u[4 * c4 + 3]) + s4.w);
      }
    }
    FRX_TMEM_ST32(tmem_base + ((uint32_t)(32 * rq) << 16) + (rb ? 256u : 0u) + (uint32_t)(32 * ch), u);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// MODE 0: ordinary rows.  MODE 1: the piece launch (phase A only, partial sums dumped to
// RowParams::piece_scratch).  MODE 2: long rows, started from the sum of their pieces (no gather).
// GRAD: the CVaR-MF gradient steps (cvar_mf.h:88-180) instead of a solve: the SYRK sum alone goes to TMEM and
// x <- x - step * (M x - rhs) is evaluated from it (see the gradient-step block below).
template <int D, int MODE, bool GRAD = false>
__global__ void __launch_bounds__(TC_THREADS, 1) row_solve_tc_kernel(RowParams p) {
  using L = TcLayout<D>;
  constexpr bool PIECE = MODE == 1, LONG = MODE == 2;
  constexpr int P = L::P;
  constexpr int F4 = 8;        // float4 columns of a loader unit: a 128 B slab (one line) per history entry
  constexpr int C = 4 * F4;    // features per unit
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzled tiles, by pointer arithmetic on the shared array itself: an
  // integer round trip makes the compiler lose the address space and emit generic LD/ST for every access.
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* opnd_hi = sm;                      // phase B operand tiles (alias stage 0)
  uint8_t* opnd_lo = sm + L::kTileBytes;
  float* Lst = reinterpret_cast<float*>(sm + L::kLstOff);
  float* Ld = reinterpret_cast<float*>(sm + L::kLdOff);
  float* wsum = reinterpret_cast<float*>(sm + L::kWsumOff);
  float* yS = reinterpret_cast<float*>(sm + L::kYOff);
  float* rdiag = reinterpret_cast<float*>(sm + L::kRdOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::kBarOff);
  uint64_t* full_bar = bars;       // [2]
  uint64_t* empty_bar = bars + 2;  // [2]
  uint64_t* acc_bar = bars + 4;    // [1] SYRK accumulators complete
  uint64_t* upd_bar = bars + 5;    // [1] trailing update complete
  uint64_t* col_bar = bars + 6;    // [4] columns 8q..8q+7 of the current diagonal factor are published
  uint64_t* ld_bar = bars + 10;    // [1] the diagonal warp has read its panel rows from TMEM
  uint64_t* upd1_bar = bars + 11;  // [1] first batch of a trailing update (the next panel's columns of its M block)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mode = p.mode;
  const bool item_side = (mode == RM_SAFER_V || mode == RM_CVAR_V);
  const bool is_row_warp = warp < P;   // warp w owns matrix rows 32w .. 32w+31

  if (tid == 0) {
    mbar_init(&full_bar[0], 2 * (D / 32));  // units of a 64-entry tile: 2 entry halves x D / 32 feature slabs
    mbar_init(&full_bar[1], 2 * (D / 32));
    mbar_init(&empty_bar[0], 1);
    mbar_init(&empty_bar[1], 1);
    mbar_init(acc_bar, 1);
    mbar_init(upd_bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&col_bar[q], 1);
    mbar_init(ld_bar, 1);
    mbar_init(upd1_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(L::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sm_addr = smem_u32(sm);
  // TMEM address of (row of this thread's lane quarter, column 0) for the M block this warp belongs to
  const int rq = warp & 3, rb = (D == 256) ? ((warp >> 2) & 1) : 0;
  const uint32_t trow = tmem_base + ((uint32_t)(32 * rq) << 16) + (rb ? 256u : 0u);

  // Power-of-two scale of the fp16 SYRK operands: |sqrt(s) e| <= max|E| * sqrt(2 max w) < 2^ex  ->  sc = 2^(14 - ex).
  // The accumulators then hold sc^2 * S: the rest of the system and the rhs are multiplied by sc^2 as well,
  // which leaves the solution unchanged (exact scaling).
  float sc = 1.f, sc2 = 1.f;
  {
    const float max_e = __uint_as_float(__ldg(p.syrk_absmax));
    const float max_w = __uint_as_float(__ldg(p.syrk_absmax + 1));
    const float bound = (p.mode == RM_SAFER_V || p.mode == RM_CVAR_V) ? max_e * sqrtf(2.f * max_w) : max_e;
    if (bound > 0.f && bound < 3.0e38f) {
      int ex;
      (void)frexpf(bound, &ex);  // bound = m * 2^ex, 0.5 <= m < 1
      ex = max(-96, min(96, 14 - ex));
      sc = exp2f((float)ex);
      sc2 = sc * sc;
    }
  }
  auto scaled_row_scalars = [&](int r_, int n_) {
    RowScalars q = row_scalars(p, r_, n_);
    q.alpha *= sc2; q.beta *= sc2; q.bscale *= sc2;
    return q;
  };
  // alpha*G + beta*I of the first row; for the later rows the warps that are idle during the back
  // substitution of the previous row do it (TMEM is free again by then)
  float* gbuf = reinterpret_cast<float*>(sm + L::kGbufOff) + warp * 1024;
  const int num_work = PIECE ? p.num_pieces : p.num_rows;
  if ((int)blockIdx.x < num_work) {
    if (PIECE) {
      tmem_init_system<D>(nullptr, 0.f, 0.f, tmem_base, warp & 3, warp >> 2, 4, gbuf, lane);
    } else {
      const int r0 = p.order[blockIdx.x];
      RowScalars s0 = scaled_row_scalars(r0, p.ptr[r0 + 1] - p.ptr[r0]);
      if (GRAD) { s0.alpha = 0.f; s0.beta = 0.f; }
      if (LONG) {
        const int n0 = p.ptr[r0 + 1] - p.ptr[r0];
        tmem_init_system<D>(GRAD ? nullptr : p.G, s0.alpha, s0.beta, tmem_base, warp & 3, warp >> 2, 4, gbuf, lane,
                            p.piece_scratch + (size_t)p.row_piece0[r0] * p.piece_stride,
                            (n0 + FRX_PIECE - 1) / FRX_PIECE, p.piece_stride);
      } else {
        tmem_init_system<D>(GRAD ? nullptr : p.G, s0.alpha, s0.beta, tmem_base, warp & 3, warp >> 2, 4, gbuf, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  uint32_t tile_base = 0;           // operand tiles of the rows this CTA has finished
  uint32_t row_count = 0;
  uint32_t upd_count = 0;           // trailing-update commits so far (all rows)
  unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tstamp = clock64();
#define FRX_DBG_LAP(slot) do { if (p.dbg && warp == P - 1 && lane == 0) { const long long now_ = clock64(); dbg_acc[slot] += (unsigned long long)(now_ - tstamp); tstamp = now_; } } while (0)

  // Work items come from a global queue: the first one is blockIdx.x, every later one is taken with an atomic
  // add while the current item is being gathered (items are sorted longest first, so the queue balances the SMs
  // and a CTA that starts late -- its SM was busy with another kernel -- simply takes fewer items).
  __shared__ int next_item_s;
  for (int ri = blockIdx.x; ri < num_work; ++row_count) {
    if (tid == 0) next_item_s = (int)gridDim.x + atomicAdd(p.work_counter, 1);  // read after the next __syncthreads
    // A work item is a row, or (piece mode) one FRX_PIECE-entry piece of a long row.
    const int r = PIECE ? p.piece_row[ri] : p.order[ri];
    const int beg = p.ptr[r];
    const int n = p.ptr[r + 1] - beg;
    const int xr = p.xmap ? p.xmap[r] : r;
    const int e0 = PIECE ? p.piece_off[ri] : 0;             // first entry gathered here
    const int pc_row = LONG ? p.row_piece0[r] : -1;         // first piece of a pre-summed long row
    const int n_here = PIECE ? min(FRX_PIECE, n - e0) : (LONG ? 0 : n);
    const int T = (n_here + KS - 1) / KS;  // SYRK operand tiles
    int dup_lo = 0, dup_hi = 0;
    if (item_side && n > 128 && (n & 127) != 0) {  // stale tail, safer2.h:200-204 (B-1)
      const int kf = n >> 7;
      dup_lo = 128 * (kf - 1) + (n & 127);
      dup_hi = 128 * kf;
    }

    // The indices (and item-side weights) of the work item's entries go to shared memory first: the gather loop
    // below then contains no dependent global load at all (measured: with per-unit index / weight LDGs the loader
    // warps sat on the long scoreboard for a third of their time).  Longer histories than the staging capacity
    // (FRX_NO_SPLIT) read the tail from global memory.
    int* colS = reinterpret_cast<int*>(sm + L::kIdxOff);
    float* wS = reinterpret_cast<float*>(sm + L::kIdxOff + FRX_STAGE_CAP * 4);
    {
      const int ns = min(n_here, FRX_STAGE_CAP);
      for (int i = tid; i < ns; i += TC_THREADS) {
        const int c = __ldg(p.col + beg + e0 + i);
        colS[i] = c;
        if (item_side) wS[i] = p.entry_w_e ? __ldg(p.entry_w_e + beg + e0 + i) : __ldg(p.entry_w + c);
      }
    }
    __syncthreads();
    const uint32_t row_tile0 = tile_base;  // CTA-wide index of this row's first tile: stage = index & 1
    constexpr int SL = D / C;              // feature slabs: (32-entry x 32-feature) units per 32 entries (8 or 4)
    float rhs_acc[SL];
#pragma unroll
    for (int q = 0; q < SL; ++q) rhs_acc[q] = 0.f;
    if (warp != TC_MMA_WARP) {
      // ================= loaders: gather, split, transpose into the operand tiles =================
      // The 15 loader warps take the (tile, feature slab) units of the row round-robin; a stage is full
      // after eight unit arrivals.  (The MMA-issuing thread blocks for most of a tile's MMA time, so it
      // cannot double as a loader without delaying its group's next tile.)
      // Software pipeline of the gather (measured: the SYRK phase was bound first by the latency of the dependent
      // chain col -> entry_w -> E row, then by the instruction count of the conversion, never by the tensor pipe):
      //  * indices and weights come from the staged arrays; the E slab of the NEXT unit travels by cp.async into
      //    this warp's gather buffer (lane = history entry, 4C bytes per lane, 16-byte chunks XOR-swizzled) while
      //    the current unit is converted;
      //  * the buffer is read back with lane (2j + o) taking the ENTRY PAIR (2j, 2j+1) and the float4 columns
      //    of parity o: exactly the distribution the fp16 operand tiles want (a 32-bit word = two adjacent
      //    entries of one feature; even lanes touch feature rows with (f & 4) == 0, odd lanes the others, which
      //    is conflict-free under the 128B swizzle) -- no register exchange at all;
      //  * operands: x = sqrt(s) e scaled by sc, hi = rn_fp16(x), lo = rn_fp16(x - hi) (fp16 carries the 11
      //    significand bits of tf32: hi*hi + hi*lo + lo*hi is at least as accurate as 3xTF32, an MMA covers
      //    K = 16 instead of 8 and the tiles are half as large);
      //  * rhs: q_2j e_2j + q_2j+1 e_2j+1 per lane, then a 15-shuffle transpose-reduce over the 16 lanes of
      //    equal parity leaves one feature of the slab in every lane.
      constexpr int STEP = TC_LOADER_WARPS - 1;
      constexpr int H4 = F4 / 2;  // float4 columns per lane after the read-back
      const uint32_t gwarp = sm_addr + L::kGatherOff + (uint32_t)warp * L::kGatherWarpBytes;
      const int o = lane & 1, pj = lane >> 1;
      // read-back addresses of my pair: entry i -> row i, chunk (2m + o) ^ key(i); at D = 128 (64 B rows) the odd
      // lanes read the odd entry first so that a quarter-warp still covers all 32 banks
      const int i0 = 2 * pj, i1 = 2 * pj + 1;
      const uint32_t key0 = F4 == 8 ? (uint32_t)(i0 & 7) : (uint32_t)(pj & 3);
      const uint32_t key1 = F4 == 8 ? (uint32_t)(i1 & 7) : (uint32_t)(pj & 3);
      const uint32_t grow0 = gwarp + (uint32_t)i0 * (C * 4), grow1 = gwarp + (uint32_t)i1 * (C * 4);
      const int n_units = 2 * SL * T;  // unit u: tile u / (2 SL), entry half (u / SL) & 1, feature slab u % SL
      auto idx_of = [&](int u) { return (u / SL) * KT + lane; };  // my entry of unit u, relative to e0
      auto col_at = [&](int i) { return i < FRX_STAGE_CAP ? colS[i] : __ldg(p.col + beg + e0 + i); };
      auto w_at = [&](int i) {
        return i < FRX_STAGE_CAP ? wS[i] : (p.entry_w_e ? __ldg(p.entry_w_e + beg + e0 + i) : __ldg(p.entry_w + col_at(i)));
      };
      // One cp.async instruction moves F4 lanes x 16 bytes = the whole slab of 32 / F4 entries (measured: with
      // lane = entry every instruction touched 32 different lines and the L1TEX tag stage, 70% busy, was the
      // bound of the whole phase).  Lane group g = lane / F4 takes the entries EG g .. EG g + EG - 1 of the unit
      // over the EG instructions, so its EG indices are contiguous in the staged array.
      constexpr int EG = F4;                   // entries per lane group = cp.async instructions per unit
      const int lg = lane / F4, lch = lane % F4;
      auto issue_gather = [&](int u) {
        const int ebase = (u / SL) * KT;
        const float* slab_src = p.E + (u % SL) * C + 4 * lch;
        int cidx[EG];
        if (ebase + KT <= FRX_STAGE_CAP) {
          const int4* cp4 = reinterpret_cast<const int4*>(colS + ebase + EG * lg);
#pragma unroll
          for (int q4 = 0; q4 < EG / 4; ++q4) {
            const int4 c4 = cp4[q4];
            cidx[4 * q4] = c4.x; cidx[4 * q4 + 1] = c4.y; cidx[4 * q4 + 2] = c4.z; cidx[4 * q4 + 3] = c4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < EG; ++j) {
            const int i = ebase + EG * lg + j;
            cidx[j] = i < n_here ? __ldg(p.col + beg + e0 + i) : 0;
          }
        }
#pragma unroll
        for (int j = 0; j < EG; ++j) {
          const int en = EG * lg + j;  // entry within the unit
          const bool ok = ebase + en < n_here;
          const uint32_t keyn = F4 == 8 ? (uint32_t)(en & 7) : (uint32_t)((en >> 1) & 3);
          const uint32_t dst = gwarp + (uint32_t)en * (C * 4) + ((((uint32_t)lch ^ keyn) & (F4 - 1)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                       "l"(slab_src + (size_t)(ok ? cidx[j] : 0) * D), "r"(ok ? 16 : 0)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      if (warp < n_units) issue_gather(warp);
      for (int u = warp; u < n_units; u += STEP) {
        const int t = u / (2 * SL), w8 = u % SL;
        const int half = (u / SL) & 1;  // which 32 entries of the 64-entry tile
        const int ei = idx_of(u);       // my entry (lane = entry) of this unit
        const int e = e0 + ei;          // ... within the row
        const bool valid = ei < n_here;
        float sq = 0.f, qw = 0.f;       // sqrt(s) * sc and q of my entry
        if (valid) {
          float s_ = 1.f, q_ = 1.f;
          if (item_side) { s_ = w_at(ei); q_ = s_; }
          if (e >= dup_lo && e < dup_hi) s_ *= 2.f;
          sq = sqrtf(s_) * sc;
          qw = q_;
        }
        float4 a[H4], b[H4];  // entries 2j and 2j+1, float4 columns 2m + o
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int m = 0; m < H4; ++m) {
          const uint32_t a0 = grow0 + ((((uint32_t)(2 * m + o)) ^ key0) & (F4 - 1)) * 16u;
          const uint32_t a1 = grow1 + ((((uint32_t)(2 * m + o)) ^ key1) & (F4 - 1)) * 16u;
          if (F4 == 4 && o) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[m].x), "=f"(b[m].y), "=f"(b[m].z), "=f"(b[m].w) : "r"(a1) : "memory");
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[m].x), "=f"(a[m].y), "=f"(a[m].z), "=f"(a[m].w) : "r"(a0) : "memory");
          } else {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[m].x), "=f"(a[m].y), "=f"(a[m].z), "=f"(a[m].w) : "r"(a0) : "memory");
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[m].x), "=f"(b[m].y), "=f"(b[m].z), "=f"(b[m].w) : "r"(a1) : "memory");
          }
#ifdef FRX_EXP_NOLOAD
          a[m] = make_float4(0.01f * e, 0.02f, 0.03f, 0.04f); b[m] = a[m];
#endif
        }
        __syncwarp();  // every lane has its pair: the buffer may be refilled
        if (u + STEP < n_units) issue_gather(u + STEP);  // next unit's slab
        const float sq0 = __shfl_sync(0xffffffffu, sq, i0), sq1 = __shfl_sync(0xffffffffu, sq, i1);
        const float q0 = __shfl_sync(0xffffffffu, qw, i0), q1 = __shfl_sync(0xffffffffu, qw, i1);
        float rv[4 * H4];
#pragma unroll
        for (int m = 0; m < H4; ++m) {
          rv[4 * m + 0] = fmaf(q1, b[m].x, q0 * a[m].x);
          rv[4 * m + 1] = fmaf(q1, b[m].y, q0 * a[m].y);
          rv[4 * m + 2] = fmaf(q1, b[m].z, q0 * a[m].z);
          rv[4 * m + 3] = fmaf(q1, b[m].w, q0 * a[m].w);
        }
        const uint32_t gt = row_tile0 + (uint32_t)t;
        const int st = (int)(gt & 1u);
        const uint32_t use = gt >> 1;  // earlier tiles of this stage
        if (use > 0) mbar_wait(&empty_bar[st], (use - 1) & 1);  // stage free again
        // byte offset of (feature row slab + 4o + t4, word 16 half + j) in the K-major 128B-swizzled tile; + 1024 m
        const uint32_t wd = 16u * (uint32_t)half + (uint32_t)pj;
        const uint32_t wchunk = wd >> 2, wbyte = (wd & 3u) << 2;
        uint8_t* hi_tile = sm + st * L::kStageBytes + ((uint32_t)(w8 * C) >> 3) * 1024u + wbyte;
        uint32_t roff[4];
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          const uint32_t r8 = (uint32_t)(4 * o + t4);
          roff[t4] = (r8 << 7) + ((wchunk ^ r8) << 4);
        }
#pragma unroll
        for (int m = 0; m < H4; ++m) {
          const float xa[4] = {a[m].x * sq0, a[m].y * sq0, a[m].z * sq0, a[m].w * sq0};
          const float xb[4] = {b[m].x * sq1, b[m].y * sq1, b[m].z * sq1, b[m].w * sq1};
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            const __half2 hh = __floats2half2_rn(xa[t4], xb[t4]);  // low half = entry 2j (lower address)
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(xa[t4] - hf.x, xb[t4] - hf.y);
            uint8_t* dst = hi_tile + 1024 * m + roff[t4];
#ifdef FRX_EXP_NOSTORE
            if (hf.x == 123.f) { *reinterpret_cast<__half2*>(dst) = hh; *reinterpret_cast<__half2*>(dst + L::kTileBytes) = ll; }
#else
            *reinterpret_cast<__half2*>(dst) = hh;
            *reinterpret_cast<__half2*>(dst + L::kTileBytes) = ll;
#endif
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[st]);
        {  // column sums over the 16 lanes of equal parity: lane ends with local value index (lane >> 1) & (4 H4 - 1)
          int cnt = 2 * H4;
#pragma unroll
          for (int off = 16; off >= 2; off >>= 1) {
            const bool upper = (lane & off) != 0;
            if (cnt >= 1) {
#pragma unroll
              for (int i = 0; i < 2 * H4; ++i) {
                if (i < cnt) {
                  const float send = upper ? rv[i] : rv[i + cnt];
                  const float recv = __shfl_xor_sync(0xffffffffu, send, off);
                  rv[i] = (upper ? rv[i + cnt] : rv[i]) + recv;
                }
              }
              cnt >>= 1;
            } else {
              rv[0] += __shfl_xor_sync(0xffffffffu, rv[0], off);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < SL; ++q) rhs_acc[q] += (q == w8) ? rv[0] : 0.f;
      }
    } else {
      // ================= MMA issuer: SYRK, tiles in order =================
      constexpr uint32_t idesc_n128 = make_idesc_f16(128);
      constexpr uint32_t idesc_n256 = make_idesc_f16(256);
      for (int t = 0; t < T; ++t) {
        const uint32_t gt = row_tile0 + (uint32_t)t;
        const int st = (int)(gt & 1u);
        mbar_wait(&full_bar[st], (gt >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t hi_addr = sm_addr + st * L::kStageBytes;
          const uint32_t lo_addr = hi_addr + L::kTileBytes;
          const int ksteps = min(KS / 16, (n_here - t * KS + 15) >> 4);  // 16 entries per MMA; the tail is zero-filled
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t ko = ks * 32;
            const uint64_t b_hi = make_kmajor_desc(hi_addr + ko);
            const uint64_t b_lo = make_kmajor_desc(lo_addr + ko);
            {  // rows 0..127 x cols 0..127 -> TMEM columns [0,128); accumulates onto alpha*G + beta*I
              umma_f16(tmem_base, b_hi, b_hi, idesc_n128, 1u);
#ifndef FRX_EXP_HIONLY
              umma_f16(tmem_base, b_hi, b_lo, idesc_n128, 1u);
              umma_f16(tmem_base, b_lo, b_hi, idesc_n128, 1u);
#endif
            }
            if (D == 256) {  // rows 128..255 x cols 0..255 -> TMEM columns [256,512)
              const uint64_t a_hi = make_kmajor_desc(hi_addr + 128 * 128 + ko);
              const uint64_t a_lo = make_kmajor_desc(lo_addr + 128 * 128 + ko);
              umma_f16(tmem_base + 256, a_hi, b_hi, idesc_n256, 1u);
#ifndef FRX_EXP_HIONLY
              umma_f16(tmem_base + 256, a_hi, b_lo, idesc_n256, 1u);
              umma_f16(tmem_base + 256, a_lo, b_hi, idesc_n256, 1u);
#endif
            }
          }
          umma_commit(&empty_bar[st]);
          if (t == T - 1) umma_commit(acc_bar);
        }
        __syncwarp();
      }
      if (T == 0 && lane == 0) mbar_arrive(acc_bar);  // pre-summed row: nothing to accumulate
    }
    tile_base += (uint32_t)T;

    // ================= phase B =================
    mbar_wait(acc_bar, row_count & 1);  // all operand tiles consumed -> the staging area may be reused
    tc_fence_after();
    {
      // per-warp rhs partials go through the (now idle) staging area and are summed in a fixed order
      float* part = reinterpret_cast<float*>(sm) + warp * D;
      // the lane's feature within a slab (see the transpose-reduce of the loaders): local value k of parity o is
      // feature 8 (k >> 2) + 4 o + (k & 3); at D = 128 lanes (lane >> 1) & 1 hold duplicates
      const int kloc = C == 32 ? (lane >> 1) : ((lane >> 2) & 7);
      const int fidx = 8 * (kloc >> 2) + 4 * (lane & 1) + (kloc & 3);
      if (warp != TC_MMA_WARP && (C == 32 || ((lane >> 1) & 1) == 0)) {
#pragma unroll
        for (int q = 0; q < SL; ++q) part[q * C + fidx] = rhs_acc[q];
      }
    }
    __syncthreads();
    const int ri_next = next_item_s;  // stable until thread 0 writes it again at the top of the next item
    FRX_DBG_LAP(0);  // phase A (gather + SYRK)

    const RowScalars rs_row = scaled_row_scalars(r, n);
    float b_reg = 0.f;  // this row-thread's rhs element, then y_i, then x_i
    if (is_row_warp) {
      const int i = 32 * warp + lane;
      const float* part = reinterpret_cast<const float*>(sm);
#pragma unroll
      for (int w = 0; w < TC_LOADER_WARPS - 1; ++w) b_reg += part[w * D + i];
      if (LONG) {
        const int npc = (n + FRX_PIECE - 1) / FRX_PIECE;
        for (int pc = 0; pc < npc; ++pc)
          b_reg += p.piece_scratch[(size_t)(pc_row + pc) * p.piece_stride + (p.piece_stride - D) + i];
      }
      if (!PIECE && !GRAD) b_reg *= rs_row.bscale;
    }
    if (PIECE) {
      // ---- piece mode: dump the partial sums (lower chunks + rhs) and clear TMEM for the next piece ----
      float* dst = p.piece_scratch + (size_t)ri * p.piece_stride;
      if (is_row_warp) dst[(p.piece_stride - D) + 32 * warp + lane] = b_reg;
      {
        const int rq_ = warp & 3, sid = warp >> 2;
        const int per_rb0 = rq_ + 1, total = D == 256 ? 2 * rq_ + 6 : rq_ + 1;
        uint32_t z[32];
        for (int c = sid; c < total; c += 4) {
          const int rb_ = c < per_rb0 ? 0 : 1, ch = c < per_rb0 ? c : c - per_rb0;
          const uint32_t ta = tmem_base + ((uint32_t)(32 * rq_) << 16) + (rb_ ? 256u : 0u) + (uint32_t)(32 * ch);
          uint32_t u[32];
          FRX_TMEM_LD32(u, ta);
          float4* out = reinterpret_cast<float4*>(dst + (size_t)piece_chunk_index(rb_, rq_, ch) * 1024 + lane * 32);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            __stcs(out + c4, make_float4(__uint_as_float(u[4 * c4]), __uint_as_float(u[4 * c4 + 1]),
                                         __uint_as_float(u[4 * c4 + 2]), __uint_as_float(u[4 * c4 + 3])));
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0u;
          FRX_TMEM_ST32(ta, z);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      ri = ri_next;
      continue;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FRX_DBG_LAP(1);  // rhs

    if (GRAD) {
      // ---- CVaR-MF gradient step (cvar_mf.h:88-180): x <- x - step * (M x - rhs) with the reference's
      //      half-updated matrix M (B-3): its lower triangle carries the rank updates, the Gramian term is
      //      complete, so  M x = cW * tril(S) x + cG * G x + reg * x.  tril(S) x comes from the TMEM-resident SYRK
      //      sum (thread = row, panel by panel), G x from a GEMM over all rows ahead of the launch (RowParams::Xg).
      //      User form (B-4: weight := stepsize, step := z_u): cW = s / n, cG = s * uw, rhs * s / n;
      //      item form: cW = 1, cG = uw, step = s.
      float* xS = wsum;  // [D]
      float x_i = 0.f;
      if (is_row_warp) {
        x_i = p.Xread[(size_t)xr * D + 32 * warp + lane];
        xS[32 * warp + lane] = x_i;
      }
      __syncthreads();
      if (is_row_warp) {
        const int i = 32 * warp + lane;
        float acc0 = 0.f, acc1 = 0.f;
        for (int pn = 0; pn <= warp; ++pn) {
          uint32_t u[32];
          FRX_TMEM_LD32(u, trow + 32u * (uint32_t)pn);
          const int lim = pn < warp ? 32 : lane + 1;  // columns j <= i
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 x4 = reinterpret_cast<const float4*>(xS + 32 * pn)[c4];
            if (4 * c4 + 0 < lim) acc0 = fmaf(__uint_as_float(u[4 * c4 + 0]), x4.x, acc0);
            if (4 * c4 + 1 < lim) acc1 = fmaf(__uint_as_float(u[4 * c4 + 1]), x4.y, acc1);
            if (4 * c4 + 2 < lim) acc0 = fmaf(__uint_as_float(u[4 * c4 + 2]), x4.z, acc0);
            if (4 * c4 + 3 < lim) acc1 = fmaf(__uint_as_float(u[4 * c4 + 3]), x4.w, acc1);
          }
        }
        const float tril = (acc0 + acc1) / sc2;
        const float gx = __ldg(p.Xg + (size_t)xr * D + i);
        float mx, rhs_i, step;
        if (mode == RM_CVAR_U) {
          const float reg = p.reg * (1.f + p.uw * (float)p.num_other);  // safer2.h:418-421
          const float wgt = p.stepsize / (float)n;
          mx = fmaf(wgt, tril, fmaf(p.stepsize * p.uw, gx, reg * x_i));
          rhs_i = b_reg * wgt;
          step = p.row_w ? p.row_w[r] : 1.f;
        } else {
          const float reg = p.reg * (p.item_reg[r] + p.alpha * p.uw * (float)p.num_users_total);  // safer2.h:426-432
          mx = tril + fmaf(p.uw, gx, reg * x_i);
          rhs_i = b_reg;
          step = p.stepsize;
        }
        p.X[(size_t)xr * D + i] = x_i - step * (mx - rhs_i);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // TMEM for this CTA's next item: zeros, or the sum of the pieces of a long row
      if (ri_next < p.num_rows) {
        if (LONG) {
          const int rn = p.order[ri_next];
          const int nn = p.ptr[rn + 1] - p.ptr[rn];
          tmem_init_system<D>(nullptr, 0.f, 0.f, tmem_base, warp & 3, warp >> 2, 4, gbuf, lane,
                              p.piece_scratch + (size_t)p.row_piece0[rn] * p.piece_stride,
                              (nn + FRX_PIECE - 1) / FRX_PIECE, p.piece_stride);
        } else {
          tmem_init_system<D>(nullptr, 0.f, 0.f, tmem_base, warp & 3, warp >> 2, 4, gbuf, lane);
        }
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      FRX_DBG_LAP(5);
      ri = ri_next;
      continue;
    }

    // ---- blocked Cholesky, 32-wide panels ----
#pragma unroll 1
    for (int pn = 0; pn < P; ++pn) {
      const int c0 = 32 * pn, c1 = c0 + 32;
      if (pn > 0) {
        // The previous trailing update is issued in two batches: the diagonal warp only needs the first one
        // (this panel's 32 columns of its own M block) and starts factoring while the rest still runs.
        mbar_wait(warp == pn ? upd1_bar : upd_bar, (upd_count - 1) & 1);
        tc_fence_after();
      }
      FRX_DBG_LAP(2);  // exposed wait for the trailing update
      float a[32];
      if (is_row_warp && warp >= pn) {
        // TMEM reads run at 64 B/clk: the diagonal warp (critical path) goes first, the rows below follow
        if (warp > pn) mbar_wait(ld_bar, (uint32_t)(pn & 1));
        uint32_t u[32];
        FRX_TMEM_LD32(u, trow + (uint32_t)c0);
        if (warp == pn && lane == 0) mbar_arrive(ld_bar);
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(u[j]);
      }
      float* Lp = Lst + L::lst_off(pn);  // panel storage: row (i - c0) -> 32 floats; rows 0..31 hold inv(L11)
      float* LdT = Ld + (pn & 1) * 1024;  // L11 transposed: LdT[k * 32 + m] = L11[m][k] (double-buffered by panel parity)
      float* rd = rdiag + (pn & 1) * 32;  // 1 / L11[k][k]
      if (warp == pn) {
        // -- diagonal block: right-looking Cholesky, lane = row.  Column k is scaled, published to shared
        //    memory with one store and read back as broadcast float4s for the rank-1 update; the next pivot
        //    (lane k+1 computes it from its own column-k entry) and the next column take shuffles so that
        //    the shared-memory round trip is off the dependency chain; the rhs is carried along as an
        //    extra column so that y = inv(L11) b comes out of the same sweep.  (Measured: the pivot chain
        //    FMUL -> FFMA -> SHFL -> MUFU.RSQ -> FMUL is ~52 cycles, one step ~96 with the in-order issue
        //    of the update; publishing the pivot ROW instead of the column is slower.) --
        long long dgt0 = 0;
        if (p.dbg && lane == 0) dgt0 = clock64();
        float rs;
        bool bad_pivot;  // warp-uniform; reported once after the sweep so that the loop stays branch-free
        {
          float akk = __shfl_sync(0xffffffffu, a[0], 0);
          bad_pivot = !(akk > 0.f);
          rs = fast_rsqrt(bad_pivot ? 1.f : akk);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float lk = a[k] * rs;          // L[i][k] of this lane's row (lane k: sqrt(pivot))
          // column k of the factor goes to shared memory with ONE store (it is the transposed factor the
          // rows below need anyway); every lane then reads it back as broadcast float4s
          LdT[k * 32 + lane] = lane >= k ? lk : 0.f;
          float rs_next = 0.f;
          if (k + 1 < 32) {
            const float dcand = fmaf(-lk, lk, a[k + 1]);  // valid in lane k+1: next pivot
            const float akk = __shfl_sync(0xffffffffu, dcand, k + 1);
            bad_pivot |= !(akk > 0.f);                    // off the dependency chain; a bad pivot yields NaNs + status
            rs_next = fast_rsqrt(akk);
            // the next column is needed first: its update takes the shuffle shortcut instead of the
            // shared-memory round trip, which then has two steps of slack
            const float lnext = __shfl_sync(0xffffffffu, lk, k + 1);  // L[k+1][k]
            a[k + 1] = fmaf(-lk, lnext, a[k + 1]);
          }
          if (lane == k) rd[k] = rs;
          a[k] = lane >= k ? lk : 0.f;         // strict upper part of the factor is zero
          const float yk = __shfl_sync(0xffffffffu, b_reg, k) * rs;  // y_k = b_k / L_kk
          if (lane == k) b_reg = yk;
          else if (lane > k) b_reg = fmaf(-lk, yk, b_reg);
          __syncwarp();
          const unsigned long long nlk2 = pack2(-lk, -lk);
          const float4* col = reinterpret_cast<const float4*>(LdT + k * 32);
#pragma unroll
          for (int m4 = (k + 2) / 4; m4 < 8; ++m4) {
            const float4 c = col[m4];  // L[j][k], j = 4*m4 .. 4*m4+3
            FRX_SWEEP4(a, 4 * m4, k + 1, lk, nlk2, c);
          }
          rs = rs_next;
          if ((k & 7) == 7) {
            // the rows below run their triangular solve behind this sweep, eight columns at a time
            if (k == 31) yS[lane] = b_reg;  // y_i of this row
            __syncwarp();
            if (lane == 0) mbar_arrive(&col_bar[k >> 3]);
          }
        }
        if (bad_pivot && lane == 0) atomicExch(p.status, 1);
        if (p.dbg && lane == 0) atomicAdd(p.dbg + 10, (unsigned long long)(clock64() - dgt0));
      }
      FRX_DBG_LAP(3);  // panel load (+ the diagonal block factor in the diagonal warp)
      if (is_row_warp && warp > pn) {
        // -- rows below: L21 row by forward substitution against L11 (column sweep, in place) --
        long long trt0 = 0;
        if (p.dbg && warp == P - 1 && lane == 0) trt0 = clock64();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          // one arrival per panel on each column barrier and P is even: the phase parity is the panel parity
          if ((k & 7) == 0) mbar_wait(&col_bar[k >> 3], (uint32_t)(pn & 1));
          const float l = a[k] * rd[k];
          a[k] = l;
          const unsigned long long nl2 = pack2(-l, -l);
          const float4* col = reinterpret_cast<const float4*>(LdT + k * 32);
#pragma unroll
          for (int m4 = (k + 1) / 4; m4 < 8; ++m4) {
            const float4 c = col[m4];
            FRX_SWEEP4(a, 4 * m4, k, l, nl2, c);
          }
        }
        const int i = 32 * warp + lane;
        float4* dst = reinterpret_cast<float4*>(Lp + (size_t)(i - c0) * 32);
        float dot = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 y4 = reinterpret_cast<const float4*>(yS)[k4];
          dot = fmaf(a[4 * k4], y4.x, dot);
          dot = fmaf(a[4 * k4 + 1], y4.y, dot);
          dot = fmaf(a[4 * k4 + 2], y4.z, dot);
          dot = fmaf(a[4 * k4 + 3], y4.w, dot);
          dst[(k4 ^ (lane & 7)) & 7] = make_float4(a[4 * k4], a[4 * k4 + 1], a[4 * k4 + 2], a[4 * k4 + 3]);
        }
        b_reg -= dot;  // forward substitution: b_i -= L21[i][:] . y_p
        if (p.dbg && warp == P - 1 && lane == 0) atomicAdd(p.dbg + 11, (unsigned long long)(clock64() - trt0));
      }
      FRX_DBG_LAP(6);  // wait for the diagonal warp + triangular solve (this warp)
      if (c1 < D) {
        if (is_row_warp && warp >= pn) {
          // -- L panel rows as K-major tf32 hi/lo operand tiles (the previous update must be done reading them) --
          if (pn > 0 && warp == pn) mbar_wait(upd_bar, (upd_count - 1) & 1);
          const int i = 32 * warp + lane;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float hi[4], lo[4];
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              const float x = a[4 * c4 + t4];
              hi[t4] = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
              lo[t4] = x - hi[t4];
            }
            const uint32_t off = tile_chunk_off(i, c4);
            *reinterpret_cast<float4*>(opnd_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(opnd_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        FRX_DBG_LAP(7);  // operand tiles (this warp)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        FRX_DBG_LAP(4);  // TRSM + operand tiles
        if (warp == TC_MMA_WARP) {
          tc_fence_after();
          if (lane == 0) {
            // A22 -= L21 L21^T on columns [c1, limit) of every M block that still has rows >= c1, in two
            // batches: (1) the next panel's 32 columns of the M block that holds the next diagonal block,
            // (2) everything else.  (An M=128 MMA costs the same 93 cycles for N = 32 and N = 64, so the split
            // adds tensor time, but that is hidden under the next diagonal factor.)
            const uint32_t hi_addr = sm_addr, lo_addr = sm_addr + L::kTileBytes;
            const bool blk1 = D == 256 && c1 >= 128;  // M block of rows c1 .. c1+31
            const uint32_t a_off = blk1 ? 128u * 128u : 0u, d_off = blk1 ? 256u : 0u;
            constexpr uint32_t idesc32 = make_idesc_tf32(32, 1);
#pragma unroll
            for (int ks = 0; ks < KT / 8; ++ks) {
              const uint32_t ko = ks * 32;
              const uint64_t b_hi = make_kmajor_desc(hi_addr + (uint32_t)c1 * 128 + ko);
              const uint64_t b_lo = make_kmajor_desc(lo_addr + (uint32_t)c1 * 128 + ko);
              const uint64_t a_hi = make_kmajor_desc(hi_addr + a_off + ko), a_lo = make_kmajor_desc(lo_addr + a_off + ko);
              umma_tf32(tmem_base + d_off + (uint32_t)c1, a_hi, b_hi, idesc32, 1u);
              umma_tf32(tmem_base + d_off + (uint32_t)c1, a_hi, b_lo, idesc32, 1u);
              umma_tf32(tmem_base + d_off + (uint32_t)c1, a_lo, b_hi, idesc32, 1u);
            }
            umma_commit(upd1_bar);
            const int c2 = c1 + 32;
            const int lim = blk1 ? 256 : 128;  // column limit of that M block
#pragma unroll
            for (int ks = 0; ks < KT / 8; ++ks) {
              const uint32_t ko = ks * 32;
              if (c2 < lim) {  // the remaining columns of the same M block
                const uint32_t idesc = make_idesc_tf32(lim - c2, 1);
                const uint64_t b_hi = make_kmajor_desc(hi_addr + (uint32_t)c2 * 128 + ko);
                const uint64_t b_lo = make_kmajor_desc(lo_addr + (uint32_t)c2 * 128 + ko);
                const uint64_t a_hi = make_kmajor_desc(hi_addr + a_off + ko), a_lo = make_kmajor_desc(lo_addr + a_off + ko);
                umma_tf32(tmem_base + d_off + (uint32_t)c2, a_hi, b_hi, idesc, 1u);
                umma_tf32(tmem_base + d_off + (uint32_t)c2, a_hi, b_lo, idesc, 1u);
                umma_tf32(tmem_base + d_off + (uint32_t)c2, a_lo, b_hi, idesc, 1u);
              }
              if (D == 256 && !blk1) {  // rows 128..255, all columns from c1
                const uint32_t idesc = make_idesc_tf32(256 - c1, 1);
                const uint64_t b_hi = make_kmajor_desc(hi_addr + (uint32_t)c1 * 128 + ko);
                const uint64_t b_lo = make_kmajor_desc(lo_addr + (uint32_t)c1 * 128 + ko);
                const uint64_t a_hi = make_kmajor_desc(hi_addr + 128 * 128 + ko);
                const uint64_t a_lo = make_kmajor_desc(lo_addr + 128 * 128 + ko);
                umma_tf32(tmem_base + 256 + (uint32_t)c1, a_hi, b_hi, idesc, 1u);
                umma_tf32(tmem_base + 256 + (uint32_t)c1, a_hi, b_lo, idesc, 1u);
                umma_tf32(tmem_base + 256 + (uint32_t)c1, a_lo, b_hi, idesc, 1u);
              }
            }
            umma_commit(upd_bar);
          }
          __syncwarp();
        }
        ++upd_count;
      }
      if (warp == pn) {
        // -- off the critical path (the tensor cores are busy with the trailing update, the other warps wait for
        //    it): inv(L11) by columns, lane c solves L x = e_c with a column sweep; kept for the back substitution --
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
#pragma unroll
        for (int i2 = 0; i2 < 32; ++i2)
          Lp[i2 * 32 + ((((lane >> 2) ^ (i2 & 7)) & 7) << 2) + (lane & 3)] = x[i2];  // lane c holds column c
      }
    }
    tc_fence_before();  // the last panel has been read from TMEM
    __syncthreads();

    // ---- back substitution L^T x = y, panel by panel from the bottom; b_reg: y_i -> x_i ----
    // Thread k owns COLUMN k of L here: element (i, k) of a stored panel row sits at word
    // i*32 + swz(k, i), so the 32 lanes of a warp read one row conflict-free.  Warp pn first applies
    // inv(L11)^T to its residuals (32 terms), publishes x, then the warps above subtract the 32 new terms.
    float* xS = wsum;  // x_i of the solved panels, [D]
    if (!is_row_warp) {
      // meanwhile the other warps write alpha*G + beta*I of this CTA's NEXT row into the idle TMEM
      const int rin = ri_next;
      if (rin < p.num_rows) {
        const int rn = p.order[rin];
        const int nn = p.ptr[rn + 1] - p.ptr[rn];
        const RowScalars sn = scaled_row_scalars(rn, nn);
        constexpr int nshare = (TC_LOADER_WARPS - P) / 4;
        tc_fence_after();
        if (LONG)
          tmem_init_system<D>(p.G, sn.alpha, sn.beta, tmem_base, warp & 3, (warp - P) >> 2, nshare, gbuf, lane,
                              p.piece_scratch + (size_t)p.row_piece0[rn] * p.piece_stride,
                              (nn + FRX_PIECE - 1) / FRX_PIECE, p.piece_stride);
        else
          tmem_init_system<D>(p.G, sn.alpha, sn.beta, tmem_base, warp & 3, (warp - P) >> 2, nshare, gbuf, lane);
      }
    }
#pragma unroll 1
    for (int pn = P - 1; is_row_warp && pn >= 0; --pn) {
      if (warp == pn) {
        const float* Lp = Lst + L::lst_off(pn);
        yS[lane] = b_reg;  // residual r_j of this panel
        __syncwarp();
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 r4 = reinterpret_cast<const float4*>(yS)[j4];
          const int j = 4 * j4;  // inv(L11)[j][lane]; exact zeros above the diagonal
          x0 = fmaf(Lp[(j + 0) * 32 + ((((lane >> 2) ^ ((j + 0) & 7)) & 7) << 2) + (lane & 3)], r4.x, x0);
          x1 = fmaf(Lp[(j + 1) * 32 + ((((lane >> 2) ^ ((j + 1) & 7)) & 7) << 2) + (lane & 3)], r4.y, x1);
          x2 = fmaf(Lp[(j + 2) * 32 + ((((lane >> 2) ^ ((j + 2) & 7)) & 7) << 2) + (lane & 3)], r4.z, x2);
          x3 = fmaf(Lp[(j + 3) * 32 + ((((lane >> 2) ^ ((j + 3) & 7)) & 7) << 2) + (lane & 3)], r4.w, x3);
        }
        b_reg = (x0 + x1) + (x2 + x3);
        xS[32 * pn + lane] = b_reg;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(P * 32) : "memory");  // row warps only
      if (warp < pn) {
        // rows [32 pn, 32 pn + 32) of this warp's own column panel
        const float* Lw = Lst + L::lst_off(warp) + (size_t)(32 * (pn - warp)) * 32;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int t4 = 0; t4 < 8; ++t4) {
          const float4 x4 = reinterpret_cast<const float4*>(xS + 32 * pn)[t4];
          const int t = 4 * t4;  // (32 pn + t) & 7 == t & 7
          s0 = fmaf(Lw[(t + 0) * 32 + ((((lane >> 2) ^ ((t + 0) & 7)) & 7) << 2) + (lane & 3)], x4.x, s0);
          s1 = fmaf(Lw[(t + 1) * 32 + ((((lane >> 2) ^ ((t + 1) & 7)) & 7) << 2) + (lane & 3)], x4.y, s1);
          s2 = fmaf(Lw[(t + 2) * 32 + ((((lane >> 2) ^ ((t + 2) & 7)) & 7) << 2) + (lane & 3)], x4.z, s2);
          s3 = fmaf(Lw[(t + 3) * 32 + ((((lane >> 2) ^ ((t + 3) & 7)) & 7) << 2) + (lane & 3)], x4.w, s3);
        }
        b_reg -= (s0 + s1) + (s2 + s3);
      }
    }
    if (is_row_warp) p.X[(size_t)xr * D + 32 * warp + lane] = b_reg;
    tc_fence_before();
    __syncthreads();  // the aliased shared memory and TMEM are free for the next row
    FRX_DBG_LAP(5);  // back substitution + store
    ri = ri_next;
  }

  if (p.dbg && warp == P - 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) atomicAdd(p.dbg + i + (i >= 6 ? 2 : 0), dbg_acc[i]);
  }
  if (p.dbg && tid == 0) atomicAdd(p.dbg + 6, (unsigned long long)row_count);
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::kTmemCols));
}

}  // namespace

size_t row_solve_tc_piece_floats(int d) { return (size_t)(d == 256 ? 36 : 10) * 1024 + (size_t)d; }

bool row_solve_tc_supported(const RowParams& p) {
  const bool mode_ok = p.mode == RM_IALS || p.mode == RM_SAFER_U || p.mode == RM_SAFER_V || p.mode == RM_CVAR_U ||
                       p.mode == RM_CVAR_V;
  return mode_ok && p.cs == 0 && p.bd == p.d && (p.d == 128 || p.d == 256);
}

template <int D, int MODE, bool GRAD>
static void launch_tc_instance(const RowParams& p, int work, cudaStream_t s, int num_sms) {
  cudaMemsetAsync(p.work_counter, 0, sizeof(int), s);
  const int smem = TcLayout<D>::kTotal + 1024;
  cudaFuncSetAttribute(row_solve_tc_kernel<D, MODE, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = num_sms < work ? num_sms : work;
  row_solve_tc_kernel<D, MODE, GRAD><<<grid, TC_THREADS, smem, s>>>(p);
}

template <int D>
static void launch_tc_dim(const RowParams& p, int work, cudaStream_t s, int num_sms) {
  const bool grad = p.mode == RM_CVAR_U || p.mode == RM_CVAR_V;
  if (p.piece_mode == 1) launch_tc_instance<D, 1, false>(p, work, s, num_sms);  // the piece sums are the same
  else if (p.piece_mode == 2) { if (grad) launch_tc_instance<D, 2, true>(p, work, s, num_sms); else launch_tc_instance<D, 2, false>(p, work, s, num_sms); }
  else { if (grad) launch_tc_instance<D, 0, true>(p, work, s, num_sms); else launch_tc_instance<D, 0, false>(p, work, s, num_sms); }
}

// p.piece_mode: 0 ordinary rows (p.order / p.num_rows), 1 pieces (p.piece_* / p.num_pieces),
// 2 pre-summed long rows (p.order / p.num_rows, every row has pieces).
void launch_row_solve_tc(const RowParams& p, cudaStream_t s, int num_sms, long long* launches) {
  const int work = p.piece_mode == 1 ? p.num_pieces : p.num_rows;
  if (work <= 0) return;
  if (p.d == 256) launch_tc_dim<256>(p, work, s, num_sms);
  else launch_tc_dim<128>(p, work, s, num_sms);
  if (launches) ++*launches;
}

}  // namespace frx

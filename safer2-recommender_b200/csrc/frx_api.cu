// C ABI (include/frecsys_b200.h) and the host-side sequencing of the epoch
// stages.  Each frx_model_* entry point restates the stage ORDER of the
// reference method it replaces (cited inline); the arithmetic is in the kernels.
#include "../../include/frecsys_b200.h"
#include "frx_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#ifdef FRX_WITH_NCCL
#include <nccl.h>
#endif

using namespace frx;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(FRX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define RC0(x) do { int rc0__ = (x); if (rc0__) return rc0__; } while (0)

struct StageTimer {
  std::string name;
  cudaEvent_t a, b;
};

struct frx_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;  // device->host copies that overlap the rest of an epoch (frx_model_train_to_host)
  cudaEvent_t copy_ev = nullptr;
  // Side stream of the dual-form row path: tridiagonalisation + rotation of the Gramian basis, per-row factors,
  // group kernel and the back rotation run beside the direct row kernel of the main stream (the persistent CTAs of
  // the two kernels cannot share an SM, so the dual-form CTAs fill the SMs the direct kernel's queue drains).
  cudaStream_t side_stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  bool side_pending = false;  // work on the side stream that the main stream has not waited for yet
  int ensure_side() {
    if (side_stream) return 0;
    CK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&join_ev, cudaEventDisableTiming));
    return 0;
  }
  // side stream continues after everything enqueued on the main stream so far
  int fork_side() {
    int rc = ensure_side();
    if (rc) return rc;
    CK(cudaEventRecord(fork_ev, stream));
    CK(cudaStreamWaitEvent(side_stream, fork_ev, 0));
    return 0;
  }
  // main stream continues after everything enqueued on the side stream so far
  int join_side() {
    if (!side_pending) return 0;
    CK(cudaEventRecord(join_ev, side_stream));
    CK(cudaStreamWaitEvent(stream, join_ev, 0));
    side_pending = false;
    return 0;
  }
  int num_sms = 148;
  long long launches = 0;
  float* gram_ws = nullptr;
  size_t gram_ws_floats = 0;
  float* row_scratch = nullptr;
  size_t row_scratch_floats = 0;
  float* wb_scratch = nullptr;  // rotated solutions of the dual-form row path
  size_t wb_scratch_floats = 0;
  int* wb_counter = nullptr;    // its work queue ([0]) and those of the direct row-kernel launches ([1..3])
  double* dws = nullptr;   // xi partials / mean partials
  int* status_dev = nullptr;
  bool profiling = false;
  std::vector<StageTimer> timers;
  size_t timers_used = 0;
  int rank = 0, world = 1;
  long long collectives = 0;
#ifdef FRX_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  int ensure_gram_ws(size_t floats) {
    if (floats <= gram_ws_floats) return 0;
    if (gram_ws) cudaFree(gram_ws);
    gram_ws = nullptr;
    gram_ws_floats = 0;
    CK(cudaMalloc(&gram_ws, floats * sizeof(float)));
    gram_ws_floats = floats;
    return 0;
  }
  int ensure_row_scratch(size_t floats) {
    if (floats <= row_scratch_floats) return 0;
    if (row_scratch) cudaFree(row_scratch);
    row_scratch = nullptr;
    row_scratch_floats = 0;
    CK(cudaMalloc(&row_scratch, floats * sizeof(float)));
    row_scratch_floats = floats;
    return 0;
  }
  int ensure_wb_scratch(size_t floats) {
    if (floats <= wb_scratch_floats) return 0;
    if (wb_scratch) cudaFree(wb_scratch);
    wb_scratch = nullptr;
    wb_scratch_floats = 0;
    CK(cudaMalloc(&wb_scratch, floats * sizeof(float)));
    wb_scratch_floats = floats;
    return 0;
  }
  void stage_begin(const char* name) {
    if (!profiling) return;
    if (timers_used == timers.size()) {
      StageTimer t;
      cudaEventCreate(&t.a);
      cudaEventCreate(&t.b);
      timers.push_back(t);
    }
    timers[timers_used].name = name;
    cudaEventRecord(timers[timers_used].a, stream);
  }
  void stage_end() {
    if (!profiling) return;
    cudaEventRecord(timers[timers_used].b, stream);
    ++timers_used;
  }
};

struct Csr {
  int nrows = 0;
  int* ptr = nullptr;
  int* col = nullptr;
  int* tup = nullptr;
  int* order = nullptr;  // this rank's non-empty rows, longest first
  int num_order = 0;
  int distinct = 0;
  std::vector<int> h_ptr;
  std::vector<int> rank_begin;  // row-id range per rank [world+1]
  // long rows of this rank cut into pieces for the tensor-core row kernel (RowParams::piece_*)
  int num_pieces = 0;
  int num_long = 0;  // the first num_long entries of `order` are the rows cut into pieces
  int* piece_row = nullptr;
  int* piece_off = nullptr;
  int* row_piece0 = nullptr;
  // FRX_LOSS_CHUNK-entry chunks of this rank's rows for the two-pass user loss (LossParams::chunk_*)
  int num_chunks = 0;
  int* chunk_row = nullptr;
  int* chunk_off = nullptr;
  float* resid = nullptr;  // [nnz]
  mutable float* ew = nullptr;  // [nnz] per-entry weights of the tensor-core item half-step (allocated on first use)
  // dual-form row path: the rows of `order` with at most FRX_WB_MAX entries (its tail: order is longest first)
  // packed into groups of four 32-entry slots (WbParams::grp_slots)
  int num_direct = 0;      // order[0, num_direct) have more than FRX_WB_MAX entries
  int wb_num_groups = 0;
  int* wb_groups = nullptr;
};

struct frx_dataset {
  frx_context* ctx = nullptr;
  int num_tuples = 0, max_user = -1, max_item = -1;
  Csr by_user, by_item;
  // evaluation helpers (built lazily): ascending ids of non-empty users and id -> compact row
  std::vector<int> h_user_ids;
  int* xmap = nullptr;
  int* user_ids_dev = nullptr;
};

// Tridiagonal form G = H T H^T of a Gramian and the fixed-side factors rotated into that basis (dual-form row path).
struct Basis {
  float *H = nullptr, *HT = nullptr, *tdiag = nullptr, *tsub = nullptr, *Et = nullptr;
  size_t et_rows = 0;
  bool valid = false;           // enqueued on the side stream for the current G / E
  cudaEvent_t done = nullptr;   // recorded on the side stream after the rotation
};

static void free_basis(Basis& b) {
  cudaFree(b.H); cudaFree(b.HT); cudaFree(b.tdiag); cudaFree(b.tsub); cudaFree(b.Et);
  if (b.done) cudaEventDestroy(b.done);
  b = Basis();
}

struct frx_model {
  frx_context* ctx = nullptr;
  Basis basisV;    // of the cached item Gramian G and V (user half-steps of SAFER2 / ERM-MF, fold-in evaluation)
  Basis basisTmp;  // of a Gramian computed inside a half-step (iALS)
  // SetPrintResidualStats (safer2.h:323-328): snapshots of U, V, z taken before the stages that change them and
  // the squared differences per primal-dual iteration [3 * iterations]: U, V, z
  bool residual_stats = false;
  float *snapU = nullptr, *snapV = nullptr, *snapZ = nullptr;
  double* resid_dev = nullptr;
  int resid_count = 0;
  bool G_dirty = false;  // V was overwritten from the host: G = V^T V must be recomputed before it is read
  float* early_U_host = nullptr;  // frx_model_train_to_host: where U / V go as soon as their half-step is final
  float* early_V_host = nullptr;
  bool early_U_done = false, early_V_done = false;
  frx_config cfg;
  int num_users = 0, num_items = 0;
  float *U = nullptr, *V = nullptr, *G = nullptr, *Gz = nullptr;
  float *z = nullptr, *loss = nullptr, *hist_size = nullptr, *item_reg = nullptr, *norm_w = nullptr,
        *quad = nullptr, *Uprev = nullptr, *pred = nullptr;
  size_t pred_cap = 0;
  float* scal = nullptr;  // device: [0]=prev_xi [1]=mean z [2]=mean z*loss
  int xi_calls = 0;
  std::vector<std::vector<int>> last_snr;
  int* snr_dev = nullptr;
  size_t snr_cap = 0;
  int* snr_host[2] = {nullptr, nullptr};
  size_t snr_host_cap[2] = {0, 0};
  cudaEvent_t snr_ev[2] = {nullptr, nullptr};
  int snr_slot = 0;
  bool is_pp() const { return cfg.model == FRX_IALSPP || cfg.model == FRX_SAFER2PP; }
  bool is_ials_family() const { return cfg.model == FRX_IALS || cfg.model == FRX_IALSPP; }
};

extern "C" const char* frx_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int frx_context_create(int device, void* cuda_stream, frx_context** out) {
  if (!out) return fail(FRX_ERR_ARG, "out is null");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(FRX_ERR_CUDA, "no CUDA device (%s): frecsys_b200 has no CPU fallback",
                cudaGetErrorString(e));
  CK(cudaSetDevice(device));
  frx_context* c = new frx_context();
  c->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
  } else {
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  CK(cudaMalloc(&c->dws, sizeof(double) * (xi_partials_doubles(c->num_sms) + 512)));
  CK(cudaMalloc(&c->status_dev, sizeof(int)));
  CK(cudaMemsetAsync(c->status_dev, 0, sizeof(int), c->stream));
  CK(cudaMalloc(&c->wb_counter, 8 * sizeof(int)));  // [4], [5]: launch_absmax
  *out = c;
  return FRX_OK;
}

extern "C" void frx_context_destroy(frx_context* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
#ifdef FRX_WITH_NCCL
  if (c->comm) ncclCommDestroy(c->comm);
#endif
  for (auto& t : c->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  cudaFree(c->gram_ws);
  cudaFree(c->row_scratch);
  cudaFree(c->wb_scratch);
  cudaFree(c->wb_counter);
  cudaFree(c->dws);
  cudaFree(c->status_dev);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->copy_ev) cudaEventDestroy(c->copy_ev);
  if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
  if (c->fork_ev) cudaEventDestroy(c->fork_ev);
  if (c->join_ev) cudaEventDestroy(c->join_ev);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int frx_context_sync(frx_context* c) {
  { const int rc_ = c->join_side(); if (rc_) return rc_; }
  CK(cudaStreamSynchronize(c->stream));
  int st = 0;
  CK(cudaMemcpyAsync(&st, c->status_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (st != 0) {
    cudaMemsetAsync(c->status_dev, 0, sizeof(int), c->stream);
    cudaStreamSynchronize(c->stream);
    return fail(FRX_ERR_NUMERIC, "non-positive Cholesky pivot (reference asserts at safer2.h:160)");
  }
  return FRX_OK;
}
extern "C" void* frx_context_stream(frx_context* c) { return (void*)c->stream; }
extern "C" long long frx_context_launch_count(frx_context* c) { return c->launches; }
extern "C" int frx_context_set_profiling(frx_context* c, int on) {
  c->profiling = on != 0;
  c->timers_used = 0;
  return FRX_OK;
}
extern "C" int frx_context_stage_times(frx_context* c, char* names_buf, int buf_len, float* ms, int max_n) {
  CK(cudaStreamSynchronize(c->stream));
  std::string names;
  int n = 0;
  for (size_t i = 0; i < c->timers_used && n < max_n; ++i, ++n) {
    float t = 0.f;
    cudaEventElapsedTime(&t, c->timers[i].a, c->timers[i].b);
    ms[n] = t;
    if (i) names += ";";
    names += c->timers[i].name;
  }
  if (names_buf && buf_len > 0) {
    strncpy(names_buf, names.c_str(), buf_len - 1);
    names_buf[buf_len - 1] = 0;
  }
  return n;
}

extern "C" int frx_comm_unique_id(void* out128) {
#ifdef FRX_WITH_NCCL
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return fail(FRX_ERR_COMM, "ncclGetUniqueId failed");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out128, &id, 128);
  return FRX_OK;
#else
  (void)out128;
  return fail(FRX_ERR_COMM, "built without NCCL");
#endif
}
extern "C" int frx_context_init_comm(frx_context* c, int rank, int world, const void* id128) {
  if (world <= 1) { c->rank = 0; c->world = 1; return FRX_OK; }
#ifdef FRX_WITH_NCCL
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  CK(cudaSetDevice(c->device));
  ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) return fail(FRX_ERR_COMM, "ncclCommInitRank: %s", ncclGetErrorString(r));
  c->rank = rank;
  c->world = world;
  return FRX_OK;
#else
  (void)rank; (void)id128;
  return fail(FRX_ERR_COMM, "built without NCCL");
#endif
}

// ---------------------------------------------------------------------------
// dataset
// ---------------------------------------------------------------------------
// Cost model of one row for the multi-GPU partition: history length + a fixed per-row share for the
// d x d factorisation, in units of history entries.
static const int kRowUnit = 480;  // d = 256 tensor-core kernel: ~62K cycles fixed vs ~128 cycles per history entry

// Contiguous row ranges balanced on the cost of a row in units of history entries: history length + row_unit for
// a row of the direct d x d kernel, a flat short_unit for a row of at most FRX_WB_MAX entries (dual-form kernel:
// its cost is the n x n factorisation of a 32..128-entry group, nearly independent of n).  row_unit < 0 keeps the
// built-in constants.  Pure host code (no GPU needed).
static const int kShortUnit = 150;  // ~19K cycles per dual-form row vs ~128 cycles per history entry
extern "C" int frx_partition_rows(const int* ptr, int nrows, int world, int row_unit, int* rank_begin) {
  if (!ptr || !rank_begin || nrows < 0 || world < 1) return fail(FRX_ERR_ARG, "bad partition arguments");
  const bool builtin = row_unit < 0;
  if (builtin) row_unit = kRowUnit;
  auto cost = [&](int r) -> long long {
    const int n = ptr[r + 1] - ptr[r];
    if (n == 0) return 0;
    return (builtin && n <= FRX_WB_MAX) ? kShortUnit : (long long)n + row_unit;
  };
  for (int k = 0; k <= world; ++k) rank_begin[k] = nrows;
  rank_begin[0] = 0;
  long long total = 0;
  for (int r = 0; r < nrows; ++r) total += cost(r);
  long long done = 0;
  int k = 1;
  for (int r = 0; r < nrows && k < world; ++r) {
    done += cost(r);
    while (k < world && done * world >= total * k) rank_begin[k++] = r + 1;
  }
  return FRX_OK;
}

// Host -> device upload of a small index array on the context stream (the kernels that read it run there;
// the host vector must stay alive until the caller's stream synchronisation).
static int upload_ints(frx_context* c, int** dev, const std::vector<int>& h) {
  CK(cudaMalloc(dev, sizeof(int) * std::max<size_t>(1, h.size())));
  if (!h.empty()) CK(cudaMemcpyAsync(*dev, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, c->stream));
  return FRX_OK;
}

static int finish_csr(frx_context* c, Csr& m, const int* cost_other_dim) {
  (void)cost_other_dim;
  m.h_ptr.resize(m.nrows + 1);
  CK(cudaMemcpyAsync(m.h_ptr.data(), m.ptr, sizeof(int) * (m.nrows + 1), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // Row-id ranges per rank; rank k owns [rank_begin[k], rank_begin[k+1]).
  m.rank_begin.assign(c->world + 1, m.nrows);
  frx_partition_rows(m.h_ptr.data(), m.nrows, c->world, -1, m.rank_begin.data());
  std::vector<int> order;
  m.distinct = 0;
  for (int r = 0; r < m.nrows; ++r)
    if (m.h_ptr[r + 1] > m.h_ptr[r]) ++m.distinct;
  for (int r = m.rank_begin[c->rank]; r < m.rank_begin[c->rank + 1]; ++r)
    if (m.h_ptr[r + 1] > m.h_ptr[r]) order.push_back(r);
  auto len = [&](int r) { return m.h_ptr[r + 1] - m.h_ptr[r]; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len(a) > len(b); });
  m.num_order = (int)order.size();
  RC0(upload_ints(c, &m.order, order));
  // pieces of the long rows (a popular item of ML-20M has ~70K entries: one CTA would need milliseconds)
  std::vector<int> prow, poff, p0(std::max(1, m.nrows), -1);
  for (int r : order) {
    const int n = len(r);
    if (n <= FRX_SPLIT_MIN) break;  // longest first
    ++m.num_long;
    p0[r] = (int)prow.size();
    for (int off = 0; off < n; off += FRX_PIECE) { prow.push_back(r); poff.push_back(off); }
  }
  std::vector<int> crow, coff;
  for (int r : order) {
    const int n = len(r);
    for (int off = 0; off < n; off += FRX_LOSS_CHUNK) { crow.push_back(r); coff.push_back(off); }
  }
  m.num_chunks = (int)crow.size();
  if (m.num_chunks) {
    RC0(upload_ints(c, &m.chunk_row, crow));
    RC0(upload_ints(c, &m.chunk_off, coff));
    CK(cudaMalloc(&m.resid, sizeof(float) * (size_t)std::max(1, m.h_ptr[m.nrows])));
  }
  m.num_pieces = (int)prow.size();
  if (m.num_pieces) {
    RC0(upload_ints(c, &m.piece_row, prow));
    RC0(upload_ints(c, &m.piece_off, poff));
    RC0(upload_ints(c, &m.row_piece0, p0));
  }
  // Groups of the dual-form row path: the rows with at most FRX_WB_MAX entries take ceil(n / 32) consecutive
  // 32-entry slots of a 4-slot group; first-fit decreasing (order is longest first), heaviest groups first.
  m.num_direct = 0;
  while (m.num_direct < m.num_order && len(order[m.num_direct]) > FRX_WB_MAX) ++m.num_direct;
  std::vector<int> groups;          // [g][4]
  std::vector<int> open_by_free[4]; // groups with 1..3 free slots
  for (int i = m.num_direct; i < m.num_order; ++i) {
    const int need = (len(order[i]) + 31) / 32;
    int g = -1;
    for (int fr = need; fr <= 3 && g < 0; ++fr)
      if (!open_by_free[fr].empty()) { g = open_by_free[fr].back(); open_by_free[fr].pop_back(); }
    int used = 0;
    if (g < 0) {
      g = (int)groups.size() / 4;
      groups.insert(groups.end(), 4, -1);
    } else {
      while (used < 4 && groups[(size_t)g * 4 + used] >= 0) ++used;
    }
    for (int k = 0; k < need; ++k) groups[(size_t)g * 4 + used + k] = ((i - m.num_direct) << 2) | k;
    const int fr = 4 - used - need;
    if (fr > 0) open_by_free[fr].push_back(g);
  }
  m.wb_num_groups = (int)groups.size() / 4;
  if (m.wb_num_groups) RC0(upload_ints(c, &m.wb_groups, groups));
  CK(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
  return FRX_OK;
}

extern "C" int frx_dataset_create(frx_context* c, int n, const int* users, const int* items,
                                  frx_dataset** out) {
  if (!c || !out || n < 0 || (n > 0 && (!users || !items))) return fail(FRX_ERR_ARG, "bad dataset arguments");
  CK(cudaSetDevice(c->device));
  frx_dataset* d = new frx_dataset();
  d->ctx = c;
  d->num_tuples = n;
  for (int t = 0; t < n; ++t) {  // dataset.h:90-91
    if (users[t] < 0 || items[t] < 0) { delete d; return fail(FRX_ERR_ARG, "negative id at tuple %d", t); }
    d->max_user = std::max(d->max_user, users[t]);
    d->max_item = std::max(d->max_item, items[t]);
  }
  int *du = nullptr, *di = nullptr;
  const size_t nb = sizeof(int) * (size_t)std::max(1, n);
  CK(cudaMalloc(&du, nb));
  CK(cudaMalloc(&di, nb));
  if (n) {
    CK(cudaMemcpyAsync(du, users, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(di, items, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  Csr* sides[2] = {&d->by_user, &d->by_item};
  for (int sd = 0; sd < 2; ++sd) {
    Csr& m = *sides[sd];
    m.nrows = (sd == 0 ? d->max_user : d->max_item) + 1;
    CK(cudaMalloc(&m.ptr, sizeof(int) * (size_t)(m.nrows + 1)));
    CK(cudaMalloc(&m.col, nb));
    CK(cudaMalloc(&m.tup, nb));
    build_csr(sd == 0 ? du : di, sd == 0 ? di : du, n, m.nrows, m.ptr, m.col, m.tup, c->stream, &c->launches);
    CK(cudaGetLastError());
    int rc = finish_csr(c, m, nullptr);
    if (rc) return rc;
  }
  CK(cudaFree(du));
  CK(cudaFree(di));
  *out = d;
  return FRX_OK;
}

extern "C" void frx_dataset_destroy(frx_dataset* d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  for (Csr* m : {&d->by_user, &d->by_item}) {
    cudaFree(m->ptr); cudaFree(m->col); cudaFree(m->tup); cudaFree(m->order);
    cudaFree(m->piece_row); cudaFree(m->piece_off); cudaFree(m->row_piece0);
    cudaFree(m->chunk_row); cudaFree(m->chunk_off); cudaFree(m->resid); cudaFree(m->ew);
    cudaFree(m->wb_groups);
  }
  cudaFree(d->xmap);
  cudaFree(d->user_ids_dev);
  delete d;
}

extern "C" int frx_dataset_info(frx_dataset* d, int* out5) {
  out5[0] = d->max_user; out5[1] = d->max_item; out5[2] = d->num_tuples;
  out5[3] = d->by_user.distinct; out5[4] = d->by_item.distinct;
  return FRX_OK;
}

extern "C" int frx_dataset_get_csr(frx_dataset* d, int by_item, int nrows, int* ptr, int* ids, int* tup) {
  Csr& m = by_item ? d->by_item : d->by_user;
  CK(cudaStreamSynchronize(d->ctx->stream));
  for (int r = 0; r <= nrows; ++r) ptr[r] = m.h_ptr[std::min(r, m.nrows)];
  if (d->num_tuples) {
    CK(cudaMemcpy(ids, m.col, sizeof(int) * (size_t)d->num_tuples, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tup, m.tup, sizeof(int) * (size_t)d->num_tuples, cudaMemcpyDeviceToHost));
  }
  return FRX_OK;
}

static int ensure_eval_maps(frx_dataset* d) {
  if (d->xmap) return FRX_OK;
  Csr& m = d->by_user;
  std::vector<int> xmap(std::max(1, m.nrows), -1);
  for (int r = 0; r < m.nrows; ++r)
    if (m.h_ptr[r + 1] > m.h_ptr[r]) {
      xmap[r] = (int)d->h_user_ids.size();
      d->h_user_ids.push_back(r);
    }
  RC0(upload_ints(d->ctx, &d->xmap, xmap));
  RC0(upload_ints(d->ctx, &d->user_ids_dev, d->h_user_ids));
  CK(cudaStreamSynchronize(d->ctx->stream));  // xmap goes out of scope
  return FRX_OK;
}

// ---------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------
#ifdef FRX_WITH_NCCL
#define NCK(call)                                                                                \
  do {                                                                                           \
    ncclResult_t r_ = (call);                                                                    \
    if (r_ != ncclSuccess) return fail(FRX_ERR_COMM, "%s failed: %s", #call, ncclGetErrorString(r_)); \
  } while (0)
#endif

// All-gather of row blocks: rank k owns rows [rank_begin[k], rank_begin[k+1]) of X (row_floats floats
// each) and every rank ends with all of them (collectives C2/C3/C4).  In place.  The blocks are
// balanced on work, not on row count, so this is not an ncclAllGather: every rank sends its block to every peer
// and receives theirs, one group of point-to-point transfers through the NVSwitch (FRX_ALLGATHER=bcast selects
// the round-1 formulation, a group of broadcasts, for comparison).
static int allgather_rows(frx_context* c, float* X, size_t row_floats, const std::vector<int>& rank_begin) {
  if (c->world <= 1) return FRX_OK;
#ifdef FRX_WITH_NCCL
  static const bool use_bcast = []() { const char* e = getenv("FRX_ALLGATHER"); return e && !strcmp(e, "bcast"); }();
  NCK(ncclGroupStart());
  if (use_bcast) {
    for (int k = 0; k < c->world; ++k) {
      const size_t b = rank_begin[k], e = rank_begin[k + 1];
      if (e > b) {
        float* blk = X + b * row_floats;
        NCK(ncclBroadcast(blk, blk, (e - b) * row_floats, ncclFloat, k, c->comm, c->stream));
      }
    }
  } else {
    const size_t mb = rank_begin[c->rank], me = rank_begin[c->rank + 1];
    for (int s = 1; s < c->world; ++s) {
      const int to = (c->rank + s) % c->world, from = (c->rank - s + c->world) % c->world;
      if (me > mb) NCK(ncclSend(X + mb * row_floats, (me - mb) * row_floats, ncclFloat, to, c->comm, c->stream));
      const size_t b = rank_begin[from], e = rank_begin[from + 1];
      if (e > b) NCK(ncclRecv(X + b * row_floats, (e - b) * row_floats, ncclFloat, from, c->comm, c->stream));
    }
  }
  NCK(ncclGroupEnd());
  ++c->collectives;
  return FRX_OK;
#else
  (void)X; (void)row_floats; (void)rank_begin;
  return fail(FRX_ERR_COMM, "built without NCCL");
#endif
}

// Sum of the per-rank partial Gramians (collective C1).
static int allreduce_sum(frx_context* c, float* buf, size_t count) {
  if (c->world <= 1) return FRX_OK;
#ifdef FRX_WITH_NCCL
  NCK(ncclAllReduce(buf, buf, count, ncclFloat, ncclSum, c->comm, c->stream));
  ++c->collectives;
  return FRX_OK;
#else
  (void)buf; (void)count;
  return fail(FRX_ERR_COMM, "built without NCCL");
#endif
}

// out[cs:cs+bd, fs:fs+fd] = E[:, cs:cs+bd]^T diag(w) E[:, fs:fs+fd].  With several ranks each one
// contracts an even share of the rows and the d x d (or B x d) partials are all-reduced.
static int gramian_into(frx_model* m, const float* E, int n, int cs, int bd, int fs, int fd,
                        const float* w, float* out) {
  frx_context* c = m->ctx;
  const int d = m->cfg.dim;
  int b = 0, e = n;
  if (c->world > 1) {
    b = (int)((long long)n * c->rank / c->world);
    e = (int)((long long)n * (c->rank + 1) / c->world);
  }
  static const bool disable_tc = getenv("FRX_DISABLE_TC") != nullptr;
  if (!disable_tc && gramian_tc_supported(e - b, d, cs, bd, fs, fd)) {
    int rc = c->ensure_gram_ws(gramian_tc_workspace_floats(d, c->num_sms));
    if (rc) return rc;
    if (launch_gramian_tc(E + (size_t)b * d, e - b, d, w ? w + b : nullptr, out, c->gram_ws, c->stream, c->num_sms,
                          &c->launches) != 0)
      return fail(FRX_ERR_CUDA, "cuTensorMapEncodeTiled failed for the Gramian operand");
  } else {
    int rc = c->ensure_gram_ws(gramian_workspace_floats(e - b, bd, fd, c->num_sms));
    if (rc) return rc;
    launch_gramian(E + (size_t)b * d, e - b, d, cs, bd, fs, fd, w ? w + b : nullptr, out + (size_t)cs * d + fs, d,
                   c->gram_ws, c->gram_ws_floats, c->stream, c->num_sms, &c->launches);
  }
  CK(cudaGetLastError());
  if (c->world > 1) {
    if (fs != 0 || fd != d) return fail(FRX_ERR_ARG, "sharded Gramian needs full-width strips");
    return allreduce_sum(c, out + (size_t)cs * d, (size_t)bd * d);
  }
  return FRX_OK;
}

static int reset_state(frx_model* m) {
  // constructor tail: safer2.h:55-59 — G_V = V^T V, z = alpha, loss = 0, hist = 0, item_reg = 0, xi = 0
  frx_context* c = m->ctx;
  int rc = gramian_into(m, m->V, m->num_items, 0, m->cfg.dim, 0, m->cfg.dim, nullptr, m->G);
  if (rc) return rc;
  m->basisV.valid = false;
  m->G_dirty = false;
  launch_fill(m->z, m->num_users, m->cfg.alpha, c->stream, &c->launches);
  CK(cudaMemsetAsync(m->loss, 0, sizeof(float) * m->num_users, c->stream));
  CK(cudaMemsetAsync(m->hist_size, 0, sizeof(float) * m->num_users, c->stream));
  CK(cudaMemsetAsync(m->item_reg, 0, sizeof(float) * m->num_items, c->stream));
  CK(cudaMemsetAsync(m->scal, 0, sizeof(float) * 4, c->stream));
  m->xi_calls = 0;
  return FRX_OK;
}

extern "C" int frx_model_create(frx_context* c, const frx_config* cfg, int num_users, int num_items,
                                frx_model** out) {
  if (!c || !cfg || !out || num_users <= 0 || num_items <= 0 || cfg->dim <= 0 || cfg->dim > 1024)
    return fail(FRX_ERR_ARG, "bad model arguments (dim must be in [1,1024])");
  if (cfg->model < 0 || cfg->model > FRX_SAFER2PP) return fail(FRX_ERR_ARG, "unknown model kind %d", cfg->model);
  CK(cudaSetDevice(c->device));
  frx_model* m = new frx_model();
  m->ctx = c;
  m->cfg = *cfg;
  if (m->cfg.block_size <= 0) m->cfg.block_size = 64;
  m->num_users = num_users;
  m->num_items = num_items;
  const size_t d = cfg->dim;
  CK(cudaMalloc(&m->U, sizeof(float) * num_users * d));
  CK(cudaMalloc(&m->V, sizeof(float) * num_items * d));
  CK(cudaMalloc(&m->G, sizeof(float) * d * d));
  CK(cudaMalloc(&m->Gz, sizeof(float) * d * d));
  CK(cudaMalloc(&m->z, sizeof(float) * num_users));
  CK(cudaMalloc(&m->loss, sizeof(float) * num_users));
  CK(cudaMalloc(&m->hist_size, sizeof(float) * num_users));
  CK(cudaMalloc(&m->norm_w, sizeof(float) * num_users));
  CK(cudaMalloc(&m->quad, sizeof(float) * num_users));
  CK(cudaMalloc(&m->item_reg, sizeof(float) * num_items));
  CK(cudaMalloc(&m->scal, sizeof(float) * 4));
  if (cfg->model == FRX_CVAR_MF) CK(cudaMalloc(&m->Uprev, sizeof(float) * num_users * d));
  CK(cudaMemsetAsync(m->U, 0, sizeof(float) * num_users * d, c->stream));
  CK(cudaMemsetAsync(m->V, 0, sizeof(float) * num_items * d, c->stream));
  for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&m->snr_ev[i], cudaEventDisableTiming));
  int rc = reset_state(m);
  if (rc) return rc;
  *out = m;
  return FRX_OK;
}

extern "C" void frx_model_destroy(frx_model* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  m->ctx->join_side();
  cudaStreamSynchronize(m->ctx->stream);
  for (float* p : {m->U, m->V, m->G, m->Gz, m->z, m->loss, m->hist_size, m->norm_w, m->quad, m->item_reg,
                   m->scal, m->Uprev, m->pred})
    cudaFree(p);
  cudaFree(m->snr_dev);
  cudaFree(m->snapU); cudaFree(m->snapV); cudaFree(m->snapZ); cudaFree(m->resid_dev);
  free_basis(m->basisV);
  free_basis(m->basisTmp);
  for (int i = 0; i < 2; ++i) {
    if (m->snr_host[i]) cudaFreeHost(m->snr_host[i]);
    if (m->snr_ev[i]) cudaEventDestroy(m->snr_ev[i]);
  }
  delete m;
}

extern "C" int frx_model_set_factors(frx_model* m, const float* U, const float* V) {
  frx_context* c = m->ctx;
  { const int jrc_ = c->join_side(); if (jrc_) return jrc_; }
  const size_t d = m->cfg.dim;
  if (U) CK(cudaMemcpyAsync(m->U, U, sizeof(float) * m->num_users * d, cudaMemcpyHostToDevice, c->stream));
  if (V) CK(cudaMemcpyAsync(m->V, V, sizeof(float) * m->num_items * d, cudaMemcpyHostToDevice, c->stream));
  return reset_state(m);
}

extern "C" int frx_model_upload_factors(frx_model* m, const float* U, const float* V) {
  frx_context* c = m->ctx;
  { const int jrc_ = c->join_side(); if (jrc_) return jrc_; }
  CK(cudaSetDevice(c->device));
  const size_t d = m->cfg.dim;
  if (U) CK(cudaMemcpyAsync(m->U, U, sizeof(float) * m->num_users * d, cudaMemcpyHostToDevice, c->stream));
  if (V) CK(cudaMemcpyAsync(m->V, V, sizeof(float) * m->num_items * d, cudaMemcpyHostToDevice, c->stream));
  if (V) m->G_dirty = true;  // item_gramian_ == V^T V is an invariant at Train() entry (safer2.h:294)
  return FRX_OK;
}

extern "C" int frx_model_init_factors(frx_model* m, unsigned seed) {
  // Recommender::init_matrix, recommender.h:61-67; order U then V, safer2.h:50-54.
  const size_t d = m->cfg.dim;
  std::vector<float> U((size_t)m->num_users * d), V((size_t)m->num_items * d);
  const float adjusted_stdev = m->cfg.stdev / std::sqrt((double)m->cfg.dim);
  std::mt19937 gen{seed};
  {
    std::normal_distribution<float> dist(0, adjusted_stdev);
    for (auto& x : U) x = dist(gen);
  }
  {
    std::normal_distribution<float> dist(0, adjusted_stdev);
    for (auto& x : V) x = dist(gen);
  }
  int rc = frx_model_set_factors(m, U.data(), V.data());
  if (rc) return rc;
  CK(cudaStreamSynchronize(m->ctx->stream));  // host vectors go out of scope
  return FRX_OK;
}

extern "C" int frx_model_get_factors(frx_model* m, float* U, float* V) {
  frx_context* c = m->ctx;
  const size_t d = m->cfg.dim;
  if (U) CK(cudaMemcpyAsync(U, m->U, sizeof(float) * m->num_users * d, cudaMemcpyDeviceToHost, c->stream));
  if (V) CK(cudaMemcpyAsync(V, m->V, sizeof(float) * m->num_items * d, cudaMemcpyDeviceToHost, c->stream));
  return frx_context_sync(c);
}

// Row-sharded host<->device legs for a multi-rank job whose factors live in host memory: every rank
// moves only the rows it owns (the ranges of frx_partition_rows over `train`), the rest travels over
// NVLink.  With one rank these are frx_model_upload_factors / frx_model_get_factors.
extern "C" int frx_model_upload_factors_sharded(frx_model* m, frx_dataset* train, const float* U, const float* V) {
  frx_context* c = m->ctx;
  { const int jrc_ = c->join_side(); if (jrc_) return jrc_; }
  CK(cudaSetDevice(c->device));
  if (c->world <= 1) return frx_model_upload_factors(m, U, V);
  const size_t d = m->cfg.dim;
  const Csr* side[2] = {&train->by_user, &train->by_item};
  const float* host[2] = {U, V};
  float* dev[2] = {m->U, m->V};
  for (int k = 0; k < 2; ++k) {
    if (!host[k]) continue;
    const size_t b = side[k]->rank_begin[c->rank], e = side[k]->rank_begin[c->rank + 1];
    if (e > b)
      CK(cudaMemcpyAsync(dev[k] + b * d, host[k] + b * d, sizeof(float) * (e - b) * d, cudaMemcpyHostToDevice, c->stream));
    // rows past the last id that occurs in `train` belong to no rank: every rank takes them from the host
    const size_t last = side[k]->nrows, total = k == 0 ? m->num_users : m->num_items;
    if (total > last)
      CK(cudaMemcpyAsync(dev[k] + last * d, host[k] + last * d, sizeof(float) * (total - last) * d, cudaMemcpyHostToDevice, c->stream));
    { const int rc_ = allgather_rows(c, dev[k], d, side[k]->rank_begin); if (rc_) return rc_; }
  }
  if (V) m->G_dirty = true;
  return FRX_OK;
}

extern "C" int frx_model_get_factors_sharded(frx_model* m, frx_dataset* train, float* U, float* V) {
  frx_context* c = m->ctx;
  if (c->world <= 1) return frx_model_get_factors(m, U, V);
  const size_t d = m->cfg.dim;
  const Csr* side[2] = {&train->by_user, &train->by_item};
  float* host[2] = {U, V};
  const float* dev[2] = {m->U, m->V};
  for (int k = 0; k < 2; ++k) {
    if (!host[k]) continue;
    const size_t b = side[k]->rank_begin[c->rank], e = side[k]->rank_begin[c->rank + 1];
    if (e > b)
      CK(cudaMemcpyAsync(host[k] + b * d, dev[k] + b * d, sizeof(float) * (e - b) * d, cudaMemcpyDeviceToHost, c->stream));
    const size_t last = side[k]->nrows, total = k == 0 ? m->num_users : m->num_items;
    if (total > last && c->rank == c->world - 1)  // the unowned tail goes with the last rank
      CK(cudaMemcpyAsync(host[k] + last * d, dev[k] + last * d, sizeof(float) * (total - last) * d, cudaMemcpyDeviceToHost, c->stream));
  }
  return frx_context_sync(c);
}

static int ensure_pred(frx_model* m, size_t n) {
  if (n <= m->pred_cap) return FRX_OK;
  if (m->pred) cudaFree(m->pred);
  m->pred = nullptr;
  m->pred_cap = 0;
  CK(cudaMalloc(&m->pred, sizeof(float) * std::max<size_t>(1, n)));
  m->pred_cap = n;
  return FRX_OK;
}

// ---- stage helpers -----------------------------------------------------------------
struct RowCall {
  const Csr* rows;
  const float* E;
  int num_other;
  float* X;
  const float* Xread;
  const int* xmap;
  const float* G;
  const float* entry_w;
  const float* row_w;
  int mode;
  int cs, bd;
  float* pred;
  const Basis* basis = nullptr;  // eigenbasis of G with E rotated into it, or null: no dual-form path
};

// G = H T H^T and Et = E * H, enqueued on the SIDE stream after everything the main stream holds so far (G and E
// are final there); consumers run on the side stream too (run_rows), the main stream joins after them.
static int compute_basis(frx_model* m, Basis& b, const float* G, const float* E, int rows) {
  frx_context* c = m->ctx;
  const size_t d = m->cfg.dim;
  if (!b.H) {
    CK(cudaMalloc(&b.H, sizeof(float) * d * d));
    CK(cudaMalloc(&b.HT, sizeof(float) * d * d));
    CK(cudaMalloc(&b.tdiag, sizeof(float) * d));
    CK(cudaMalloc(&b.tsub, sizeof(float) * d));
    CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
  }
  if (b.et_rows < (size_t)rows) {
    RC0(c->join_side());
    CK(cudaStreamSynchronize(c->stream));  // nobody reads the old buffer any more
    cudaFree(b.Et);
    b.Et = nullptr;
    CK(cudaMalloc(&b.Et, sizeof(float) * (size_t)rows * d));
    b.et_rows = rows;
  }
  RC0(c->fork_side());
  if (launch_sym_tridiag(G, (int)d, b.H, b.HT, b.tdiag, b.tsub, c->side_stream, &c->launches) != 0)
    return fail(FRX_ERR_CUDA, "launch of sym_tridiag_kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  launch_rows_gemm(E, rows, (int)d, b.H, b.Et, nullptr, nullptr, c->side_stream, &c->launches);
  CK(cudaGetLastError());
  CK(cudaEventRecord(b.done, c->side_stream));
  c->side_pending = true;
  b.valid = true;
  return FRX_OK;
}

static bool wb_enabled(const frx_model* m) {
  static const bool off = getenv("FRX_DISABLE_WB") != nullptr || getenv("FRX_DISABLE_TC") != nullptr;
  return !off && sym_tridiag_supported(m->cfg.dim);
}

static int run_rows(frx_model* m, const RowCall& rc_) {
  frx_context* c = m->ctx;
  RowParams p;
  memset(&p, 0, sizeof p);
  p.ptr = rc_.rows->ptr; p.col = rc_.rows->col; p.tup = rc_.rows->tup;
  p.order = rc_.rows->order; p.num_rows = rc_.rows->num_order;
  p.E = rc_.E; p.d = m->cfg.dim; p.num_other = rc_.num_other;
  p.cs = rc_.cs; p.bd = rc_.bd;
  p.X = rc_.X; p.Xread = rc_.Xread ? rc_.Xread : rc_.X; p.xmap = rc_.xmap;
  p.G = rc_.G; p.entry_w = rc_.entry_w; p.row_w = rc_.row_w; p.item_reg = m->item_reg;
  p.pred = rc_.pred; p.mode = rc_.mode;
  p.uw = m->cfg.uobs_weight; p.reg = m->cfg.reg; p.reg_exp = m->cfg.reg_exp; p.alpha = m->cfg.alpha;
  p.stepsize = m->cfg.stepsize; p.num_users_total = m->num_users;
  p.status = c->status_dev;
  // --use_cg: the reference's iterative solvers, with their stopping rule, in the generic kernel
  if (m->cfg.use_cg && (p.mode == RM_IALS || p.mode == RM_SAFER_U || p.mode == RM_SAFER_V) &&
      (m->cfg.model == FRX_IALS || m->cfg.model == FRX_SAFER2 || m->cfg.model == FRX_ERM_MF)) {
    p.solver = m->cfg.model == FRX_ERM_MF ? 2 : 1;
    p.cg_tol = m->cfg.cg_tol;
    p.cg_max_it = m->cfg.cg_max_it;
  }
  static const bool disable_tc = getenv("FRX_DISABLE_TC") != nullptr;
  if (!disable_tc && p.solver == 0 && row_solve_tc_supported(p)) {
    static const bool tc_debug = getenv("FRX_TC_DEBUG") != nullptr;
    unsigned long long* dbg = nullptr;
    if (tc_debug) {
      CK(cudaMalloc(&dbg, 16 * sizeof(unsigned long long)));
      CK(cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), c->stream));
      p.dbg = dbg;
    }
    static const bool no_split = getenv("FRX_NO_SPLIT") != nullptr;
    // bound of the SYRK operands (fp16 hi/lo pairs scaled by a power of two): max |E|, max entry weight
    launch_absmax(p.E, (size_t)p.num_other * p.d, p.entry_w, (size_t)p.num_other,
                  reinterpret_cast<unsigned*>(c->wb_counter + 4), c->stream, c->num_sms, &c->launches);
    p.syrk_absmax = reinterpret_cast<const unsigned*>(c->wb_counter + 4);
    if (p.entry_w && !rc_.rows->h_ptr.empty()) {
      // the weight of every history entry of this rank's rows, aligned with col (staged with the indices)
      const Csr* rw = rc_.rows;
      if (!rw->ew) CK(cudaMalloc(&rw->ew, sizeof(float) * (size_t)std::max(1, rw->h_ptr[rw->nrows])));
      const int rb = rc_.xmap ? 0 : rw->rank_begin[c->rank], re = rc_.xmap ? rw->nrows : rw->rank_begin[c->rank + 1];
      launch_entry_weights(p.col, p.entry_w, (size_t)rw->h_ptr[rb], (size_t)rw->h_ptr[re], rw->ew, c->stream, &c->launches);
      p.entry_w_e = rw->ew;
    }
    if (p.mode == RM_CVAR_U || p.mode == RM_CVAR_V) {
      // gradient steps: G x for every row of the side being updated, as one GEMM ahead of the row kernel
      const int nx = rc_.rows->nrows;
      int r = c->ensure_wb_scratch((size_t)nx * p.d);
      if (r) return r;
      launch_rows_gemm(p.Xread, nx, p.d, p.G, c->wb_scratch, nullptr, nullptr, c->stream, &c->launches);
      CK(cudaGetLastError());
      p.Xg = c->wb_scratch;
    }
    // rows with at most FRX_WB_MAX entries go to the dual-form kernel when the eigenbasis of G is at hand
    const bool use_wb = rc_.basis && rc_.basis->valid && rc_.rows->wb_num_groups > 0 && row_solve_wb_supported(p);
    const int direct_rows = use_wb ? rc_.rows->num_direct : rc_.rows->num_order;
    int direct_sms = c->num_sms;
    if (use_wb) {
      // dual-form rows on the side stream (ordered after the basis there and after the main stream's work so far)
      RC0(c->fork_side());
      const int nwb = rc_.rows->num_order - rc_.rows->num_direct;
      int r = c->ensure_wb_scratch((size_t)3 * nwb * p.d);
      if (r) return r;
      WbParams q;
      q.grp_slots = rc_.rows->wb_groups;
      q.wb_rows = rc_.rows->order + rc_.rows->num_direct;
      q.num_groups = rc_.rows->wb_num_groups;
      q.Et = rc_.basis->Et;
      q.tdiag = rc_.basis->tdiag;
      q.tsub = rc_.basis->tsub;
      q.Xt = c->wb_scratch;
      q.lsub = c->wb_scratch + (size_t)nwb * p.d;
      q.rsd = c->wb_scratch + (size_t)2 * nwb * p.d;
      q.counter = c->wb_counter;
      RowParams pw = p;
      pw.order = rc_.rows->order;
      pw.num_rows = rc_.rows->num_order;
      if (tc_debug) CK(cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), c->side_stream));
      launch_row_solve_wb(pw, q, nwb, c->side_stream, c->num_sms, &c->launches);
      CK(cudaGetLastError());
      if (tc_debug) {
        unsigned long long h[16];
        CK(cudaMemcpyAsync(h, dbg, sizeof h, cudaMemcpyDeviceToHost, c->side_stream));
        CK(cudaStreamSynchronize(c->side_stream));
        const double g = (double)q.num_groups;
        fprintf(stderr, "[frx wb] mode=%d groups=%d rows=%d | cycles/group (summed over SMs): solver set0: wait=%.0f chol=%.0f backsub=%.0f gather+sweeps=%.0f (x3 sets) | mma warp: wait_slot=%.0f wait_stage=%.0f issue=%.0f ring=%.0f\n",
                p.mode, q.num_groups, nwb, 3 * h[0] / g, 3 * h[1] / g, 3 * h[2] / g, 3 * h[3] / g, h[4] / g, h[5] / g, h[6] / g, h[7] / g);
        CK(cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), c->side_stream));
        CK(cudaStreamSynchronize(c->side_stream));
      }
      // back to the original basis: X[row] = xt * H^T
      launch_rows_gemm(c->wb_scratch, nwb, p.d, rc_.basis->HT, p.X, q.wb_rows, p.xmap, c->side_stream, &c->launches);
      CK(cudaGetLastError());
      c->side_pending = true;
      // While the 8-CTA cluster of the tridiagonalisation is still running it needs its SMs: the direct kernel's
      // work queue does not care how many CTAs serve it.
      if (cudaEventQuery(rc_.basis->done) != cudaSuccess) direct_sms = std::max(8, c->num_sms - 8);
      (void)cudaGetLastError();
    }
    p.num_rows = direct_rows;
    if (rc_.rows->num_pieces > 0 && !no_split) {
      // first launch: partial sums of the pieces of the long rows
      p.piece_stride = row_solve_tc_piece_floats(p.d);
      int r = c->ensure_row_scratch(p.piece_stride * (size_t)rc_.rows->num_pieces);
      if (r) return r;
      p.piece_scratch = c->row_scratch;
      p.piece_row = rc_.rows->piece_row; p.piece_off = rc_.rows->piece_off; p.row_piece0 = rc_.rows->row_piece0;
      p.num_pieces = rc_.rows->num_pieces;
      p.piece_mode = 1;
      p.work_counter = c->wb_counter + 1;
      launch_row_solve_tc(p, c->stream, direct_sms, &c->launches);
      CK(cudaGetLastError());
      // second launch: the long rows, each started from the sum of its pieces
      p.piece_mode = 2;
      p.num_rows = rc_.rows->num_long;
      p.work_counter = c->wb_counter + 2;
      launch_row_solve_tc(p, c->stream, direct_sms, &c->launches);
      CK(cudaGetLastError());
      // the ordinary rows follow
      p.piece_mode = 0;
      p.order = rc_.rows->order + rc_.rows->num_long;
      p.num_rows = direct_rows - rc_.rows->num_long;
    }
    p.work_counter = c->wb_counter + 3;
    launch_row_solve_tc(p, c->stream, direct_sms, &c->launches);
    CK(cudaGetLastError());
    if (use_wb) RC0(c->join_side());
    if (tc_debug) {
      unsigned long long h[16];
      CK(cudaMemcpyAsync(h, dbg, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      cudaFree(dbg);
      const double rows = h[6] ? (double)h[6] : 1.0;
      fprintf(stderr, "[frx tc] mode=%d rows=%llu cycles/row: gather+syrk=%.0f assemble=%.0f upd_wait=%.0f diag=%.0f trsm+tiles=%.0f backsub=%.0f\n",
              p.mode, h[6], h[0] / rows, h[1] / rows, h[2] / rows, h[3] / rows, h[4] / rows, h[5] / rows);
      fprintf(stderr, "[frx tc]   (timeline of the last row-warp) wait_diag+trsm=%.0f opnd_tiles=%.0f | diag warps: factor=%.0f | last row-warp pure trsm=%.0f\n",
              h[8] / rows, h[9] / rows, h[10] / rows, h[11] / rows);
    }
    return FRX_OK;
  }
  const size_t per = row_solve_generic_scratch_floats(p.bd, p.solver);
  if (per) {
    const int g = row_solve_generic_grid(p.num_rows, c->num_sms);
    int r = c->ensure_row_scratch(per * (size_t)g);
    if (r) return r;
    p.scratch = c->row_scratch;
    p.scratch_stride = per;
  }
  launch_row_solve_generic(p, c->stream, c->num_sms, &c->launches);
  CK(cudaGetLastError());
  return FRX_OK;
}

// Row solve on this rank's rows, then the all-gather that makes the updated factor block visible
// to every rank before the next half-step.
static int run_rows_sharded(frx_model* m, const RowCall& rc_) {
  int rc = run_rows(m, rc_);
  if (rc) return rc;
  if (m->ctx->world > 1 && !rc_.xmap) {
    m->ctx->stage_end();  // the caller's stage times the kernel; the exchange is reported separately
    m->ctx->stage_begin("allgather_rows");
    return allgather_rows(m->ctx, rc_.X, (size_t)m->cfg.dim, rc_.rows->rank_begin);
  }
  return FRX_OK;
}

static int stage_item_gramian(frx_model* m) {
  RC0(m->ctx->join_side());  // a basis computation of the previous G may still be reading it
  m->ctx->stage_begin("gramian_V");
  int rc = gramian_into(m, m->V, m->num_items, 0, m->cfg.dim, 0, m->cfg.dim, nullptr, m->G);
  m->ctx->stage_end();
  m->basisV.valid = false;
  m->G_dirty = false;
  return rc;
}

// The cached item Gramian follows V when V was overwritten from the host (frx_model_upload_factors*).
static int refresh_item_gramian(frx_model* m) {
  if (!m->G_dirty) return FRX_OK;
  return stage_item_gramian(m);
}

// Eigenbasis of the cached item Gramian + V rotated into it, for the dual-form path of the user half-steps.
static const Basis* item_basis(frx_model* m, int* rc) {
  *rc = FRX_OK;
  if (!wb_enabled(m)) return nullptr;
  if (!m->basisV.valid) *rc = compute_basis(m, m->basisV, m->G, m->V, m->num_items);
  return *rc ? nullptr : &m->basisV;
}

static int stage_user_loss(frx_model* m, frx_dataset* ds, const float* G, const float* pred) {
  frx_context* c = m->ctx;
  c->stage_begin("quadform");
  LossParams p;
  memset(&p, 0, sizeof p);
  p.ptr = ds->by_user.ptr; p.col = ds->by_user.col; p.tup = ds->by_user.tup;
  p.order = ds->by_user.order; p.num_rows = ds->by_user.num_order;
  p.U = m->U; p.V = m->V; p.d = m->cfg.dim; p.G = G; p.pred = pred;
  p.beta = m->cfg.uobs_weight; p.halve = m->is_ials_family() ? 0 : 1;
  p.quad = m->quad; p.loss = m->loss;
  p.resid = ds->by_user.resid; p.chunk_row = ds->by_user.chunk_row; p.chunk_off = ds->by_user.chunk_off;
  p.num_chunks = ds->by_user.num_chunks;
  launch_quadform(p, ds->by_user.rank_begin[c->rank], ds->by_user.rank_begin[c->rank + 1], c->stream, &c->launches);
  c->stage_end();
  c->stage_begin("user_loss");
  launch_user_loss_rows(p, c->stream, c->num_sms, &c->launches);
  CK(cudaGetLastError());
  int rc = FRX_OK;
  if (c->world > 1) {
    c->stage_end();
    c->stage_begin("allgather_loss");
    rc = allgather_rows(c, m->loss, 1, ds->by_user.rank_begin);  // C4: every rank needs all losses for xi / z
  }
  c->stage_end();
  return rc;
}

static int stage_weights(frx_model* m) {
  frx_context* c = m->ctx;
  c->stage_begin("weights");
  const int kind = m->cfg.model == FRX_CVAR_MF ? 2 : (m->cfg.use_epanechnikov ? 1 : 0);
  // safer2.h iterates data.by_user() (users with history); safer2pp.h:847-856 updates all users.
  const int only_hist = m->cfg.model == FRX_SAFER2PP ? 0 : 1;
  launch_user_weights(m->loss, m->hist_size, m->num_users, m->z, m->scal, m->cfg.bandwidth, kind, only_hist,
                      c->stream, &c->launches);
  c->stage_end();
  CK(cudaGetLastError());
  return FRX_OK;
}

static int stage_means(frx_model* m) {
  frx_context* c = m->ctx;
  launch_weight_means(m->z, m->loss, m->num_users, m->scal + 1, c->dws + xi_partials_doubles(c->num_sms),
                      c->stream, &c->launches);
  CK(cudaGetLastError());
  return FRX_OK;
}

// ComputeXi, safer2.h:716-742.  SNR indices are drawn on the host with the
// std:: classes of safer2.h:728-736 (bit-exact by construction) and uploaded.
static int stage_xi(frx_model* m, bool from_mean) {
  frx_context* c = m->ctx;
  c->stage_begin("xi");
  XiParams p;
  memset(&p, 0, sizeof p);
  p.loss = m->loss; p.num_users = m->num_users; p.iters = m->cfg.xi_iterations;
  p.alpha = m->cfg.alpha; p.bandwidth = m->cfg.bandwidth; p.epanechnikov = m->cfg.use_epanechnikov;
  p.start_from_mean = from_mean ? 1 : 0;
  p.xi_io = m->scal; p.partials = c->dws;
  m->last_snr.clear();
  if (m->cfg.use_snr) {
    const int num_samples = (int)((float)m->num_users * m->cfg.sampling_ratio);  // B-10
    const size_t total = (size_t)num_samples * (size_t)std::max(0, p.iters);
    const int slot = m->snr_slot;
    m->snr_slot ^= 1;
    if (total > m->snr_host_cap[slot]) {
      if (m->snr_host[slot]) cudaFreeHost(m->snr_host[slot]);
      CK(cudaMallocHost(&m->snr_host[slot], sizeof(int) * total));
      m->snr_host_cap[slot] = total;
    } else {
      CK(cudaEventSynchronize(m->snr_ev[slot]));  // previous upload from this pinned buffer is done
    }
    if (total > m->snr_cap) {
      CK(cudaStreamSynchronize(c->stream));
      if (m->snr_dev) cudaFree(m->snr_dev);
      CK(cudaMalloc(&m->snr_dev, sizeof(int) * total));
      m->snr_cap = total;
    }
    for (int t = 0; t < p.iters; ++t) {
      std::mt19937 rng(m->cfg.snr_seed + 1000u * (unsigned)m->xi_calls + (unsigned)t);
      std::uniform_int_distribution<int> uni(0, m->num_users - 1);
      int* dst = m->snr_host[slot] + (size_t)t * num_samples;
      for (int j = 0; j < num_samples; j++) dst[j] = uni(rng);
      m->last_snr.emplace_back(dst, dst + num_samples);
    }
    if (total) {
      CK(cudaMemcpyAsync(m->snr_dev, m->snr_host[slot], sizeof(int) * total, cudaMemcpyHostToDevice, c->stream));
      CK(cudaEventRecord(m->snr_ev[slot], c->stream));
    }
    p.snr_idx = m->snr_dev;
    p.n_samples = num_samples;
  }
  ++m->xi_calls;
  if (launch_xi_newton(p, c->stream, c->num_sms, &c->launches) != 0)
    return fail(FRX_ERR_CUDA, "cooperative launch of xi_newton_kernel failed: %s",
                cudaGetErrorString(cudaGetLastError()));
  c->stage_end();
  return FRX_OK;
}

static int stage_xi_exact(frx_model* m) {
  frx_context* c = m->ctx;
  c->stage_begin("xi_exact");
  launch_exact_quantile(m->loss, m->num_users, m->cfg.alpha, m->scal, nullptr, c->stream, &c->launches);
  c->stage_end();
  CK(cudaGetLastError());
  return FRX_OK;
}

// The next user half-step (and the fold-in evaluation) reads the basis of the item Gramian that has just been
// computed: start it on the side stream now, under the loss / xi stages and the direct rows of the next epoch.
static int prefetch_item_basis(frx_model* m, frx_dataset* ds) {
  if (ds->by_user.wb_num_groups <= 0) return FRX_OK;
  int brc;
  item_basis(m, &brc);
  return brc;
}

// StepU of SAFER2 / ERM-MF (safer2.h:437-490) on the model's own users.
static int stage_step_u(frx_model* m, frx_dataset* ds) {
  const Basis* basis = nullptr;
  if (m->cfg.model != FRX_CVAR_MF && ds->by_user.wb_num_groups > 0) {
    int brc;
    basis = item_basis(m, &brc);
    if (brc) return brc;
  }
  m->ctx->stage_begin("step_U");
  RowCall rc{&ds->by_user, m->V, m->num_items, m->U, nullptr, nullptr, m->G, nullptr, m->z,
             RM_SAFER_U, 0, m->cfg.dim, nullptr};
  rc.basis = basis;
  if (m->cfg.model == FRX_CVAR_MF) rc.mode = RM_CVAR_U;
  int r = run_rows_sharded(m, rc);
  m->ctx->stage_end();
  return r;
}

// StepV (safer2.h:493-555): w = z/|hist|, G = U^T diag(z) U over all users, per-item ProjectV.
static int stage_step_v(frx_model* m, frx_dataset* ds, const float* users) {
  frx_context* c = m->ctx;
  c->stage_begin("gramian_Uz");
  launch_norm_weights(m->z, m->hist_size, m->num_users, m->norm_w, c->stream, &c->launches);
  int r = gramian_into(m, users, m->num_users, 0, m->cfg.dim, 0, m->cfg.dim, m->z, m->Gz);
  c->stage_end();
  if (r) return r;
  c->stage_begin("step_V");
  RowCall rc{&ds->by_item, users, m->num_users, m->V, nullptr, nullptr, m->Gz, m->norm_w, nullptr,
             RM_SAFER_V, 0, m->cfg.dim, nullptr};
  if (m->cfg.model == FRX_CVAR_MF) rc.mode = RM_CVAR_V;
  r = run_rows_sharded(m, rc);
  c->stage_end();
  return r;
}

// IALSRecommender::Step (ials.h:317-365): Gramian of the fixed side, then per-row Project.
static int stage_ials_step(frx_model* m, frx_dataset* ds, bool user_side, float* X, const int* xmap,
                           const Csr* rows_override) {
  frx_context* c = m->ctx;
  const float* other = user_side ? m->V : m->U;
  const int n_other = user_side ? m->num_items : m->num_users;
  c->stage_begin(user_side ? "gramian_V" : "gramian_U");
  int r = gramian_into(m, other, n_other, 0, m->cfg.dim, 0, m->cfg.dim, nullptr, m->Gz);
  c->stage_end();
  if (r) return r;
  const Csr* rows = rows_override ? rows_override : (user_side ? &ds->by_user : &ds->by_item);
  const Basis* basis = nullptr;
  if (wb_enabled(m) && rows->wb_num_groups > 0) {
    r = compute_basis(m, m->basisTmp, m->Gz, other, n_other);
    if (r) return r;
    basis = &m->basisTmp;
  }
  c->stage_begin(user_side ? "step_U" : "step_V");
  RowCall rc{rows, other, n_other, X, nullptr, xmap, m->Gz, nullptr, nullptr, RM_IALS, 0, m->cfg.dim, nullptr};
  rc.basis = basis;
  r = run_rows_sharded(m, rc);
  c->stage_end();
  return r;
}

// One block sweep of a ++ model on one side (ialspp.h:351-424, safer2pp.h:448-609).
static int stage_block(frx_model* m, const Csr* rows, bool user_side, float* X, const int* xmap, int bs,
                       int be, bool safer, const float* row_w) {
  frx_context* c = m->ctx;
  const int d = m->cfg.dim;
  const float* other = user_side ? m->V : m->U;
  const int n_other = user_side ? m->num_items : m->num_users;
  const bool weighted = safer && !user_side;  // safer2pp.h:534-544: (z o U_B)^T U
  if (c->world > 1 && !xmap) {
    // Several ranks: the tuple-indexed prediction cache is replicated in memory, but only the entries of the rows
    // a rank is about to solve have to be current.  They are recomputed from the (all-gathered) factors: the
    // reference's incremental cache (ialspp.h:136-143) holds the same dot products up to fp32 rounding.
    c->stage_begin("predict_rows");
    launch_predict_chunks(rows->ptr, rows->col, rows->tup, rows->chunk_row, rows->chunk_off, rows->num_chunks,
                          user_side ? m->U : m->V, nullptr, other, d, m->pred, c->stream, c->num_sms, &c->launches);
    c->stage_end();
    CK(cudaGetLastError());
  }
  c->stage_begin(user_side ? "block_gramian_V" : "block_gramian_U");
  if (weighted) launch_norm_weights(m->z, m->hist_size, m->num_users, m->norm_w, c->stream, &c->launches);
  // strip G[bs:be, 0:d] = E_B^T (w o) E   (contains the B x B block, the reference's TODO ialspp.h:360-361)
  int r = gramian_into(m, other, n_other, bs, be - bs, 0, d, weighted ? m->z : nullptr, m->Gz);
  c->stage_end();
  if (r) return r;
  c->stage_begin(user_side ? "block_U" : "block_V");
  RowCall rc{rows, other, n_other, X, nullptr, xmap, m->Gz, weighted ? m->norm_w : nullptr, row_w,
             safer ? (user_side ? RM_PP_SAFER_U : RM_PP_SAFER_V) : RM_PP_IALS, bs, be - bs, m->pred};
  r = run_rows(m, rc);
  c->stage_end();
  if (r) return r;
  if (c->world > 1 && !xmap) {
    // every rank needs the updated block before the other side's sweep (whole rows: the B columns are strided)
    c->stage_begin("allgather_rows");
    r = allgather_rows(c, X, (size_t)d, rows->rank_begin);
    c->stage_end();
  }
  return r;
}

static int stage_predict(frx_model* m, const Csr* rows, const float* U, const int* xmap) {
  frx_context* c = m->ctx;
  c->stage_begin("predict");
  launch_predict_chunks(rows->ptr, rows->col, rows->tup, rows->chunk_row, rows->chunk_off, rows->num_chunks, U, xmap,
                        m->V, m->cfg.dim, m->pred, c->stream, c->num_sms, &c->launches);
  c->stage_end();
  CK(cudaGetLastError());
  return FRX_OK;
}

#define RC(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

extern "C" int frx_model_initialize(frx_model* m, frx_dataset* ds) {
  frx_context* c = m->ctx;
  { const int jrc_ = c->join_side(); if (jrc_) return jrc_; }
  CK(cudaSetDevice(c->device));
  if (m->is_ials_family()) return FRX_OK;  // run_model.cc:246-257
  if (ds->by_user.nrows > m->num_users || ds->by_item.nrows > m->num_items)
    return fail(FRX_ERR_ARG, "dataset ids exceed the model's num_users/num_items");
  RC(refresh_item_gramian(m));
  const float* pred = nullptr;
  if (m->cfg.model == FRX_SAFER2PP) {
    RC(ensure_pred(m, ds->num_tuples));
    RC(stage_predict(m, &ds->by_user, m->U, nullptr));
    pred = m->pred;
  }
  RC(stage_user_loss(m, ds, m->G, pred));  // safer2.h:820-821
  if (m->cfg.model == FRX_SAFER2 || m->cfg.model == FRX_SAFER2PP) RC(stage_xi(m, /*from_mean=*/true));
  // cvar_mf.h:713 drops its local prev_xi (B-7): xi stays 0.
  launch_hist_and_item_reg(ds->by_user.ptr, ds->by_user.nrows, m->hist_size, ds->by_item.ptr, ds->by_item.col,
                           ds->by_item.nrows, m->item_reg, c->stream, &c->launches);
  CK(cudaGetLastError());
  return FRX_OK;
}

// U is final for this epoch once the last user half-step (and its all-gather) is on the stream: its
// device->host copy (this rank's rows) runs on the copy stream under the item half-step.
static int maybe_download(frx_model* m, frx_dataset* ds, bool item_side) {
  float* host = item_side ? m->early_V_host : m->early_U_host;
  bool& done = item_side ? m->early_V_done : m->early_U_done;
  if (!host || done) return FRX_OK;
  frx_context* c = m->ctx;
  if (!c->copy_stream) {
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->copy_ev, cudaEventDisableTiming));
  }
  const size_t d = m->cfg.dim;
  const Csr& side = item_side ? ds->by_item : ds->by_user;
  const size_t total = item_side ? (size_t)m->num_items : (size_t)m->num_users;
  const float* dev = item_side ? m->V : m->U;
  size_t b = 0, e = total;
  if (c->world > 1) {
    b = side.rank_begin[c->rank];
    e = c->rank == c->world - 1 ? total : (size_t)side.rank_begin[c->rank + 1];
  }
  CK(cudaEventRecord(c->copy_ev, c->stream));
  CK(cudaStreamWaitEvent(c->copy_stream, c->copy_ev, 0));
  if (e > b)
    CK(cudaMemcpyAsync(host + b * d, dev + b * d, sizeof(float) * (e - b) * d, cudaMemcpyDeviceToHost, c->copy_stream));
  done = true;
  return FRX_OK;
}
static int maybe_download_U(frx_model* m, frx_dataset* ds) { return maybe_download(m, ds, false); }
static int maybe_download_V(frx_model* m, frx_dataset* ds) { return maybe_download(m, ds, true); }

// Residual statistics: snapshot(which) before a stage, residual(which, slot) after it.  which: 0 U, 1 V, 2 z.
static const int kMaxResid = 64;
static int resid_snapshot(frx_model* m, int which) {
  if (!m->residual_stats) return FRX_OK;
  frx_context* c = m->ctx;
  const size_t d = m->cfg.dim;
  float** snap = which == 0 ? &m->snapU : which == 1 ? &m->snapV : &m->snapZ;
  const float* src = which == 0 ? m->U : which == 1 ? m->V : m->z;
  const size_t n = which == 0 ? (size_t)m->num_users * d : which == 1 ? (size_t)m->num_items * d : (size_t)m->num_users;
  if (!*snap) CK(cudaMalloc(snap, sizeof(float) * n));
  if (!m->resid_dev) CK(cudaMalloc(&m->resid_dev, sizeof(double) * 3 * kMaxResid));
  CK(cudaMemcpyAsync(*snap, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, c->stream));
  return FRX_OK;
}
static int resid_record(frx_model* m, int which, int iter) {
  if (!m->residual_stats || iter >= kMaxResid) return FRX_OK;
  frx_context* c = m->ctx;
  const size_t d = m->cfg.dim;
  const float* snap = which == 0 ? m->snapU : which == 1 ? m->snapV : m->snapZ;
  const float* cur = which == 0 ? m->U : which == 1 ? m->V : m->z;
  const size_t n = which == 0 ? (size_t)m->num_users * d : which == 1 ? (size_t)m->num_items * d : (size_t)m->num_users;
  launch_sqdiff(cur, snap, n, m->resid_dev + 3 * iter + which, c->stream, c->num_sms, &c->launches);
  CK(cudaGetLastError());
  m->resid_count = std::max(m->resid_count, iter + 1);
  return FRX_OK;
}

extern "C" int frx_model_set_residual_stats(frx_model* m, int on) {
  m->residual_stats = on != 0;
  return FRX_OK;
}

// U / V / z residual norms of the last Train(), one triple per primal-dual iteration, as the reference logs them
// (safer2.h:323-328, erm_mf.h:297-300, cvar_mf.h:322-326, ialspp.h:257-260, safer2pp.h:346-350).  iALS reports
// 0, 0 (its Step returns a constant 0, ials.h:363-364), CVaR-MF 0 for U (cvar_mf.h:472-473).  Returns the count.
extern "C" int frx_model_get_residuals(frx_model* m, float* out, int max_triples) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  const int n = std::min(m->resid_count, max_triples);
  if (n <= 0) return 0;
  std::vector<double> h(3 * (size_t)n);
  CK(cudaMemcpyAsync(h.data(), m->resid_dev, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 3 * n; ++i) out[i] = (float)std::sqrt(h[i]);
  if (m->cfg.model == FRX_IALS) for (int i = 0; i < n; ++i) out[3 * i] = out[3 * i + 1] = 0.f;
  if (m->cfg.model == FRX_CVAR_MF) for (int i = 0; i < n; ++i) out[3 * i] = 0.f;
  return n;
}

extern "C" int frx_model_train(frx_model* m, frx_dataset* ds) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  if (ds->by_user.nrows > m->num_users || ds->by_item.nrows > m->num_items)
    return fail(FRX_ERR_ARG, "dataset ids exceed the model's num_users/num_items");
  c->timers_used = 0;
  RC(refresh_item_gramian(m));
  m->resid_count = 0;
  if (m->residual_stats) {
    if (!m->resid_dev) CK(cudaMalloc(&m->resid_dev, sizeof(double) * 3 * kMaxResid));
    CK(cudaMemsetAsync(m->resid_dev, 0, sizeof(double) * 3 * kMaxResid, c->stream));
  }
  const int d = m->cfg.dim, B = m->cfg.block_size;
  switch (m->cfg.model) {
    case FRX_IALS:  // ials.h:187-224
      RC(resid_snapshot(m, 0));
      RC(stage_ials_step(m, ds, true, m->U, nullptr, nullptr));
      RC(resid_record(m, 0, 0));
      RC(maybe_download_U(m, ds));
      RC(resid_snapshot(m, 1));
      RC(stage_ials_step(m, ds, false, m->V, nullptr, nullptr));
      RC(resid_record(m, 1, 0));
      RC(maybe_download_V(m, ds));
      RC(stage_item_gramian(m));  // ComputeUserLoss recomputes the Gramian, ials.h:371
      RC(stage_user_loss(m, ds, m->G, nullptr));
      break;
    case FRX_IALSPP:  // ialspp.h:208-261
      RC(ensure_pred(m, ds->num_tuples));
      RC(stage_predict(m, &ds->by_user, m->U, nullptr));
      RC(resid_snapshot(m, 0));
      RC(resid_snapshot(m, 1));
      for (int start = 0; start < d; start += B) {
        const int end = std::min(start + B, d);
        RC(stage_block(m, &ds->by_user, true, m->U, nullptr, start, end, false, nullptr));
        RC(stage_block(m, &ds->by_item, false, m->V, nullptr, start, end, false, nullptr));
      }
      RC(resid_record(m, 0, 0));  // sum over the blocks of |delta|^2 = |U_new - U_old|^2 (ialspp.h:218-259)
      RC(resid_record(m, 1, 0));
      break;
    case FRX_ERM_MF:  // erm_mf.h:257-301
      RC(resid_snapshot(m, 0));
      RC(stage_step_u(m, ds));
      RC(resid_record(m, 0, 0));
      RC(maybe_download_U(m, ds));
      RC(resid_snapshot(m, 1));
      RC(stage_step_v(m, ds, m->U));
      RC(resid_record(m, 1, 0));
      RC(maybe_download_V(m, ds));
      RC(stage_item_gramian(m));
      RC(prefetch_item_basis(m, ds));
      RC(stage_user_loss(m, ds, m->G, nullptr));
      RC(stage_means(m));
      break;
    case FRX_CVAR_MF:  // cvar_mf.h:276-330
      RC(resid_snapshot(m, 2));
      RC(stage_weights(m));
      RC(resid_record(m, 2, 0));
      CK(cudaMemcpyAsync(m->Uprev, m->U, sizeof(float) * (size_t)m->num_users * d, cudaMemcpyDeviceToDevice,
                         c->stream));  // cvar_mf.h:282
      RC(stage_step_u(m, ds));
      RC(resid_snapshot(m, 1));
      RC(stage_step_v(m, ds, m->Uprev));
      RC(resid_record(m, 1, 0));
      RC(stage_item_gramian(m));
      RC(stage_user_loss(m, ds, m->G, nullptr));
      RC(stage_means(m));
      RC(stage_xi_exact(m));
      break;
    case FRX_SAFER2:  // safer2.h:266-334
      for (int t = 0; t < m->cfg.pd_iterations; ++t) {
        RC(resid_snapshot(m, 2));
        RC(stage_weights(m));
        RC(resid_record(m, 2, t));
        RC(resid_snapshot(m, 0));
        RC(stage_step_u(m, ds));
        RC(resid_record(m, 0, t));
        if (t == m->cfg.pd_iterations - 1) RC(maybe_download_U(m, ds));
        RC(resid_snapshot(m, 1));
        RC(stage_step_v(m, ds, m->U));
        RC(resid_record(m, 1, t));
        if (t == m->cfg.pd_iterations - 1) RC(maybe_download_V(m, ds));
        RC(stage_item_gramian(m));
        RC(prefetch_item_basis(m, ds));
        RC(stage_user_loss(m, ds, m->G, nullptr));
        RC(stage_means(m));
      }
      RC(stage_xi(m, false));
      break;
    case FRX_SAFER2PP:  // safer2pp.h:288-355
      RC(ensure_pred(m, ds->num_tuples));
      RC(stage_predict(m, &ds->by_user, m->U, nullptr));
      for (int t = 0; t < m->cfg.pd_iterations; ++t) {
        RC(resid_snapshot(m, 2));
        RC(stage_weights(m));
        RC(resid_record(m, 2, t));
        RC(resid_snapshot(m, 0));
        RC(resid_snapshot(m, 1));
        for (int start = 0; start < d; start += B) {
          const int end = std::min(start + B, d);
          RC(stage_block(m, &ds->by_user, true, m->U, nullptr, start, end, true, m->z));
          RC(stage_block(m, &ds->by_item, false, m->V, nullptr, start, end, true, nullptr));
        }
        RC(resid_record(m, 0, t));
        RC(resid_record(m, 1, t));
        RC(stage_item_gramian(m));
        if (c->world > 1) RC(stage_predict(m, &ds->by_user, m->U, nullptr));  // the item sweeps of other ranks moved V
        RC(stage_user_loss(m, ds, m->G, m->pred));
        RC(stage_means(m));
      }
      RC(stage_xi(m, false));
      break;
  }
  return FRX_OK;
}

// One Train() epoch for a caller whose factors live in host memory: on return U and V (this rank's rows
// when there are several ranks, as frx_model_get_factors_sharded) are in the host arrays.  The copy of U
// overlaps the item half-step for the models whose user factors are final after the user half-step.
extern "C" int frx_model_train_to_host(frx_model* m, frx_dataset* ds, float* U, float* V) {
  frx_context* c = m->ctx;
  m->early_U_host = U;
  m->early_V_host = V;
  m->early_U_done = m->early_V_done = false;
  int rc = frx_model_train(m, ds);
  const bool u_done = m->early_U_done, v_done = m->early_V_done;
  m->early_U_host = m->early_V_host = nullptr;
  m->early_U_done = m->early_V_done = false;
  if (rc) return rc;
  rc = frx_model_get_factors_sharded(m, ds, u_done ? nullptr : U, v_done ? nullptr : V);  // synchronises the main stream
  if (rc) return rc;
  if (u_done || v_done) CK(cudaStreamSynchronize(c->copy_stream));
  return FRX_OK;
}

extern "C" int frx_model_stage(frx_model* m, frx_dataset* ds, int stage) {
  frx_context* c = m->ctx;
  { const int jrc_ = c->join_side(); if (jrc_) return jrc_; }
  CK(cudaSetDevice(c->device));
  RC(refresh_item_gramian(m));
  const int d = m->cfg.dim;
  switch (stage) {
    case 0: return stage_weights(m);
    case 1: return stage_step_u(m, ds);
    case 2:
      if (m->cfg.model == FRX_CVAR_MF) {
        CK(cudaMemcpyAsync(m->Uprev, m->U, sizeof(float) * (size_t)m->num_users * d, cudaMemcpyDeviceToDevice,
                           c->stream));
        return stage_step_v(m, ds, m->Uprev);
      }
      return stage_step_v(m, ds, m->U);
    case 3: return stage_item_gramian(m);
    case 4:
      if (m->is_ials_family()) RC(stage_item_gramian(m));
      return stage_user_loss(m, ds, m->G, nullptr);
    case 5: return m->cfg.model == FRX_CVAR_MF ? stage_xi_exact(m) : stage_xi(m, false);
    case 6: return stage_ials_step(m, ds, true, m->U, nullptr, nullptr);
    case 7: return stage_ials_step(m, ds, false, m->V, nullptr, nullptr);
  }
  return fail(FRX_ERR_ARG, "unknown stage %d", stage);
}

extern "C" int frx_model_get_state(frx_model* m, float* z, float* loss, float* hist_size, float* item_reg,
                                   float* scalars, float* gramian) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  const size_t d = m->cfg.dim;
  if (gramian) RC(refresh_item_gramian(m));
  if (scalars) RC(stage_means(m));
  if (z) CK(cudaMemcpyAsync(z, m->z, sizeof(float) * m->num_users, cudaMemcpyDeviceToHost, c->stream));
  if (loss) CK(cudaMemcpyAsync(loss, m->loss, sizeof(float) * m->num_users, cudaMemcpyDeviceToHost, c->stream));
  if (hist_size) CK(cudaMemcpyAsync(hist_size, m->hist_size, sizeof(float) * m->num_users, cudaMemcpyDeviceToHost, c->stream));
  if (item_reg) CK(cudaMemcpyAsync(item_reg, m->item_reg, sizeof(float) * m->num_items, cudaMemcpyDeviceToHost, c->stream));
  if (gramian) CK(cudaMemcpyAsync(gramian, m->G, sizeof(float) * d * d, cudaMemcpyDeviceToHost, c->stream));
  float sc[4] = {0, 0, 0, 0};
  if (scalars) CK(cudaMemcpyAsync(sc, m->scal, sizeof(float) * 4, cudaMemcpyDeviceToHost, c->stream));
  RC(frx_context_sync(c));
  if (scalars) { scalars[0] = sc[0]; scalars[1] = sc[2]; scalars[2] = sc[1]; }
  return FRX_OK;
}

extern "C" int frx_model_set_state(frx_model* m, const float* z, const float* loss, float xi) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  if (z) CK(cudaMemcpyAsync(m->z, z, sizeof(float) * m->num_users, cudaMemcpyHostToDevice, c->stream));
  if (loss) CK(cudaMemcpyAsync(m->loss, loss, sizeof(float) * m->num_users, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(m->scal, &xi, sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FRX_OK;
}

// Checkpoint (the reference has none).  Everything an epoch reads that is not recomputed
// from the data set each time: factors, dual weights, per-user loss, history sizes, item regularisation
// sums, xi and the running means, and the ComputeXi call counter (it seeds the SNR subsamples).  The item
// Gramian is recomputed from V on load (same kernel, same bits).  Replicated state: with several ranks
// every rank writes / reads the same content.
static const char kCkptMagic[8] = {'F', 'R', 'X', 'C', 'K', 'P', 'T', '1'};

extern "C" int frx_model_save(frx_model* m, const char* path) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  const size_t d = m->cfg.dim, nu = m->num_users, ni = m->num_items;
  std::vector<float> buf(nu * d + ni * d + 3 * nu + ni + 4);
  float* p = buf.data();
  const struct { const float* dev; size_t n; } parts[] = {{m->U, nu * d}, {m->V, ni * d}, {m->z, nu}, {m->loss, nu},
                                                          {m->hist_size, nu}, {m->item_reg, ni}, {m->scal, 4}};
  for (auto& q : parts) {
    CK(cudaMemcpyAsync(p, q.dev, sizeof(float) * q.n, cudaMemcpyDeviceToHost, c->stream));
    p += q.n;
  }
  CK(cudaStreamSynchronize(c->stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(FRX_ERR_ARG, "cannot open %s for writing", path);
  const int hdr[5] = {m->cfg.model, m->cfg.dim, m->num_users, m->num_items, m->xi_calls};
  bool ok = fwrite(kCkptMagic, 1, 8, f) == 8 && fwrite(hdr, sizeof(int), 5, f) == 5 &&
            fwrite(buf.data(), sizeof(float), buf.size(), f) == buf.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? FRX_OK : fail(FRX_ERR_ARG, "short write to %s", path);
}

extern "C" int frx_model_load(frx_model* m, const char* path) {
  frx_context* c = m->ctx;
  { const int jrc_ = c->join_side(); if (jrc_) return jrc_; }
  CK(cudaSetDevice(c->device));
  FILE* f = fopen(path, "rb");
  if (!f) return fail(FRX_ERR_ARG, "cannot open %s", path);
  char magic[8];
  int hdr[5];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kCkptMagic, 8) != 0 || fread(hdr, sizeof(int), 5, f) != 5) {
    fclose(f);
    return fail(FRX_ERR_ARG, "%s is not a frecsys_b200 checkpoint", path);
  }
  if (hdr[0] != m->cfg.model || hdr[1] != m->cfg.dim || hdr[2] != m->num_users || hdr[3] != m->num_items) {
    fclose(f);
    return fail(FRX_ERR_ARG, "checkpoint %s is for model %d dim %d %d x %d, this model is %d dim %d %d x %d", path, hdr[0],
                hdr[1], hdr[2], hdr[3], m->cfg.model, m->cfg.dim, m->num_users, m->num_items);
  }
  const size_t d = m->cfg.dim, nu = m->num_users, ni = m->num_items;
  std::vector<float> buf(nu * d + ni * d + 3 * nu + ni + 4);
  const bool ok = fread(buf.data(), sizeof(float), buf.size(), f) == buf.size();
  fclose(f);
  if (!ok) return fail(FRX_ERR_ARG, "checkpoint %s is truncated", path);
  const float* p = buf.data();
  const struct { float* dev; size_t n; } parts[] = {{m->U, nu * d}, {m->V, ni * d}, {m->z, nu}, {m->loss, nu},
                                                    {m->hist_size, nu}, {m->item_reg, ni}, {m->scal, 4}};
  for (auto& q : parts) {
    CK(cudaMemcpyAsync(q.dev, p, sizeof(float) * q.n, cudaMemcpyHostToDevice, c->stream));
    p += q.n;
  }
  int rc = gramian_into(m, m->V, m->num_items, 0, m->cfg.dim, 0, m->cfg.dim, nullptr, m->G);
  if (rc) return rc;
  m->basisV.valid = false;
  m->G_dirty = false;
  m->xi_calls = hdr[4];
  CK(cudaStreamSynchronize(c->stream));  // buf goes out of scope
  return FRX_OK;
}

extern "C" int frx_model_last_snr(frx_model* m, int* n_iters, int* n_samples, int* out) {
  *n_iters = (int)m->last_snr.size();
  *n_samples = m->last_snr.empty() ? 0 : (int)m->last_snr[0].size();
  if (out)
    for (size_t t = 0; t < m->last_snr.size(); ++t)
      std::copy(m->last_snr[t].begin(), m->last_snr[t].end(), out + t * (size_t)*n_samples);
  return FRX_OK;
}

// PrintLosses / ComputeLosses numbers (safer2.h:337-413, ials.h:226-305).
extern "C" int frx_model_compute_stats(frx_model* m, frx_dataset* ds, double* out6) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  const int d = m->cfg.dim;
  const int nu = m->num_users, ni = m->num_items;
  // observed loss: per-user sum of squared residuals in double
  float* tmp_loss = nullptr;
  double* obs = nullptr;
  CK(cudaMalloc(&tmp_loss, sizeof(float) * nu));
  CK(cudaMalloc(&obs, sizeof(double) * nu));
  CK(cudaMemsetAsync(obs, 0, sizeof(double) * nu, c->stream));
  LossParams p;
  memset(&p, 0, sizeof p);
  p.ptr = ds->by_user.ptr; p.col = ds->by_user.col; p.tup = ds->by_user.tup;
  p.order = ds->by_user.order; p.num_rows = ds->by_user.num_order;
  p.U = m->U; p.V = m->V; p.d = d; p.G = m->G; p.beta = 0.f; p.halve = 0;
  p.quad = m->quad; p.loss = tmp_loss; p.obs_sq = obs;
  p.resid = ds->by_user.resid; p.chunk_row = ds->by_user.chunk_row; p.chunk_off = ds->by_user.chunk_off;
  p.num_chunks = ds->by_user.num_chunks;
  launch_user_loss(p, 0, nu, c->stream, c->num_sms, &c->launches);
  // everything is reduced on the device (one 8-double read-back): [0] observed, [1] reg, [2] user norms,
  // [3] item norms, [4] sum(G_U o G_V), [5] sum of the per-user losses, [6], [7] scratch of the item pass
  double* acc = nullptr;
  CK(cudaMalloc(&acc, sizeof(double) * 8));
  CK(cudaMemsetAsync(acc, 0, sizeof(double) * 8, c->stream));
  launch_sum_d(obs, (size_t)nu, acc + 0, c->stream, c->num_sms, &c->launches);
  const float uw = m->cfg.uobs_weight;
  const bool ials = m->is_ials_family();
  // users: reg_u = lambda (n_u + uw I)^nu (ials.h:310-315) or lambda (1 + uw I) (safer2.h:418-421)
  launch_reg_sums(m->U, ds->by_user.ptr, ds->by_user.nrows, d, ials ? 0 : 1, m->cfg.reg, m->cfg.reg_exp, uw, m->cfg.alpha, ni,
                  nullptr, acc + 6, c->stream, c->num_sms, &c->launches);
  CK(cudaMemcpyAsync(acc + 2, acc + 7, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(acc + 1, acc + 6, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemsetAsync(acc + 6, 0, sizeof(double) * 2, c->stream));
  // items: reg_v = lambda (n_v + uw U)^nu or lambda (item_reg_v + alpha uw U) (safer2.h:426-432)
  launch_reg_sums(m->V, ds->by_item.ptr, ds->by_item.nrows, d, ials ? 0 : 2, m->cfg.reg, m->cfg.reg_exp, uw, m->cfg.alpha, nu,
                  m->item_reg, acc + 6, c->stream, c->num_sms, &c->launches);
  CK(cudaMemcpyAsync(acc + 3, acc + 7, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  RC(gramian_into(m, m->U, nu, 0, d, 0, d, nullptr, m->Gz));
  float* Gtmp = nullptr;
  CK(cudaMalloc(&Gtmp, sizeof(float) * (size_t)d * d));
  RC(gramian_into(m, m->V, ni, 0, d, 0, d, nullptr, Gtmp));
  launch_dot_sum(m->Gz, Gtmp, (size_t)d * d, acc + 4, c->stream, c->num_sms, &c->launches);
  launch_dot_sum(m->loss, nullptr, (size_t)nu, acc + 5, c->stream, c->num_sms, &c->launches);
  CK(cudaGetLastError());
  double h[8];
  CK(cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  RC(frx_context_sync(c));
  cudaFree(tmp_loss); cudaFree(obs); cudaFree(Gtmp); cudaFree(acc);
  const double o = h[0], reg = h[1] + h[6], ru = h[2], ri = h[3], unobs = h[4];
  out6[1] = o / ds->num_tuples;
  out6[2] = unobs / ni / nu;
  out6[3] = reg;
  out6[4] = ru / nu;
  out6[5] = ri / ni;
  out6[0] = ials ? o + uw * unobs + reg : h[5];
  return FRX_OK;
}

// EvaluateDataset, safer2.h:225-263 and siblings + recommender.h:78-199.
extern "C" int frx_model_evaluate(frx_model* m, frx_dataset* tr, frx_dataset* te, const int* k_list, int nk,
                                  int* user_ids, float* recall, float* ndcg, int* topk, float* folded) {
  frx_context* c = m->ctx;
  CK(cudaSetDevice(c->device));
  RC(ensure_eval_maps(tr));
  const int nu = (int)tr->h_user_ids.size();
  if (!recall) return nu;
  if (c->world > 1) return fail(FRX_ERR_ARG, "evaluate: run on a single-rank context");
  if (tr->by_item.nrows > m->num_items) return fail(FRX_ERR_ARG, "test items exceed the model's num_items");
  RC(refresh_item_gramian(m));
  const int d = m->cfg.dim, B = m->cfg.block_size;
  int max_k = 0;
  for (int i = 0; i < nk; ++i) max_k = std::max(max_k, k_list[i]);
  if (nk <= 0 || nk > 32 || max_k <= 0 || max_k > 512) return fail(FRX_ERR_ARG, "bad k_list");
  float* Ut = nullptr;
  CK(cudaMalloc(&Ut, sizeof(float) * (size_t)std::max(1, nu) * d));
  CK(cudaMemsetAsync(Ut, 0, sizeof(float) * (size_t)std::max(1, nu) * d, c->stream));  // MatrixXf::Zero, safer2.h:230
  switch (m->cfg.model) {
    case FRX_IALS:  // ials.h:169-174
      RC(stage_ials_step(m, tr, true, Ut, tr->xmap, nullptr));
      break;
    case FRX_ERM_MF:
    case FRX_CVAR_MF:
    case FRX_SAFER2: {  // safer2.h:246-252: StepU with weight 1 and the cached item_gramian_
      RowCall rc{&tr->by_user, m->V, m->num_items, Ut, nullptr, tr->xmap, m->G, nullptr, nullptr,
                 RM_SAFER_U, 0, d, nullptr};
      if (tr->by_user.wb_num_groups > 0) {
        int brc;
        rc.basis = item_basis(m, &brc);
        RC(brc);
      }
      RC(run_rows(m, rc));
      break;
    }
    case FRX_IALSPP:
    case FRX_SAFER2PP: {  // ialspp.h:152-195, safer2pp.h:223-275: 8 user-only block-sweep epochs
      float* saved_pred = m->pred;
      size_t saved_cap = m->pred_cap;
      m->pred = nullptr; m->pred_cap = 0;
      RC(ensure_pred(m, tr->num_tuples));
      CK(cudaMemsetAsync(m->pred, 0, sizeof(float) * (size_t)std::max(1, tr->num_tuples), c->stream));
      for (int e = 0; e < 8; ++e) {
        RC(stage_predict(m, &tr->by_user, Ut, tr->xmap));
        for (int start = 0; start < d; start += B) {
          const int end = std::min(start + B, d);
          RC(stage_block(m, &tr->by_user, true, Ut, tr->xmap, start, end, m->cfg.model == FRX_SAFER2PP, nullptr));
        }
      }
      CK(cudaStreamSynchronize(c->stream));
      cudaFree(m->pred);
      m->pred = saved_pred; m->pred_cap = saved_cap;
      break;
    }
  }
  EvalParams p;
  memset(&p, 0, sizeof p);
  p.Ut = Ut; p.V = m->V; p.nu = nu; p.num_items = m->num_items; p.d = d;
  p.user_ids = tr->user_ids_dev;
  p.tr_ptr = tr->by_user.ptr; p.tr_col = tr->by_user.col;
  p.te_ptr = te->by_user.ptr; p.te_col = te->by_user.col; p.te_rows = te->by_user.nrows;
  int* d_k = nullptr; int* d_topk = nullptr; float *d_rec = nullptr, *d_ndcg = nullptr, *d_scores = nullptr;
  CK(cudaMalloc(&d_k, sizeof(int) * nk));
  CK(cudaMemcpyAsync(d_k, k_list, sizeof(int) * nk, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMalloc(&d_topk, sizeof(int) * (size_t)std::max(1, nu) * max_k));
  CK(cudaMalloc(&d_rec, sizeof(float) * (size_t)std::max(1, nu) * nk));
  CK(cudaMalloc(&d_ndcg, sizeof(float) * (size_t)std::max(1, nu) * nk));
  p.k_list = d_k; p.nk = nk; p.max_k = max_k;
  p.topk = d_topk; p.recall = d_rec; p.ndcg = d_ndcg;
  static const bool no_fused = getenv("FRX_DISABLE_TC") != nullptr || getenv("FRX_DISABLE_FUSED_EVAL") != nullptr;
  void* d_ws = nullptr;
  c->stage_begin("evaluate");
  if (!no_fused && nu > 0 && score_topk_supported(d, max_k)) {
    // fused tcgen05 scoring + top-k: no score matrix (recommender.h:109-112, 132-153)
    CK(cudaMalloc(&d_ws, score_topk_workspace_bytes(nu, m->num_items, d, c->num_sms)));
    const unsigned long long* lists = nullptr;
    int segments = 0;
    if (launch_score_topk(p, d_ws, c->num_sms, c->stream, &c->launches, &lists, &segments) != 0)
      return fail(FRX_ERR_CUDA, "cuTensorMapEncodeTiled failed for the scoring operands");
    CK(cudaGetLastError());
    launch_merge_metrics(p, lists, segments, c->stream, &c->launches);
  } else {
    // any dimension / max_k > 128: tiled scores into a chunk buffer, then per-user radix select
    size_t chunk = ((size_t)512 << 20) / (sizeof(float) * (size_t)m->num_items);
    chunk = std::max<size_t>(1, std::min<size_t>(chunk, (size_t)std::max(1, nu)));
    CK(cudaMalloc(&d_scores, sizeof(float) * chunk * m->num_items));
    p.scores = d_scores; p.chunk_users = (int)chunk;
    launch_evaluate(p, c->stream, c->num_sms, &c->launches);
  }
  c->stage_end();
  CK(cudaGetLastError());
  if (nu) {
    CK(cudaMemcpyAsync(recall, d_rec, sizeof(float) * (size_t)nu * nk, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(ndcg, d_ndcg, sizeof(float) * (size_t)nu * nk, cudaMemcpyDeviceToHost, c->stream));
    if (topk) CK(cudaMemcpyAsync(topk, d_topk, sizeof(int) * (size_t)nu * max_k, cudaMemcpyDeviceToHost, c->stream));
    if (folded) CK(cudaMemcpyAsync(folded, Ut, sizeof(float) * (size_t)nu * d, cudaMemcpyDeviceToHost, c->stream));
    if (user_ids) std::copy(tr->h_user_ids.begin(), tr->h_user_ids.end(), user_ids);
  }
  int rc = frx_context_sync(c);
  cudaFree(Ut); cudaFree(d_k); cudaFree(d_topk); cudaFree(d_rec); cudaFree(d_ndcg); cudaFree(d_scores); cudaFree(d_ws);
  if (rc) return rc;
  return nu;
}

extern "C" int frx_gramian(frx_context* c, const float* E, int n, int d, const float* w, float* out) {
  CK(cudaSetDevice(c->device));
  float *dE = nullptr, *dw = nullptr, *dG = nullptr;
  CK(cudaMalloc(&dE, sizeof(float) * (size_t)n * d));
  CK(cudaMalloc(&dG, sizeof(float) * (size_t)d * d));
  CK(cudaMemcpyAsync(dE, E, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice, c->stream));
  if (w) {
    CK(cudaMalloc(&dw, sizeof(float) * (size_t)n));
    CK(cudaMemcpyAsync(dw, w, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  static const bool disable_tc = getenv("FRX_DISABLE_TC") != nullptr;
  if (!disable_tc && gramian_tc_supported(n, d, 0, d, 0, d)) {
    int rc = c->ensure_gram_ws(gramian_tc_workspace_floats(d, c->num_sms));
    if (rc) return rc;
    if (launch_gramian_tc(dE, n, d, dw, dG, c->gram_ws, c->stream, c->num_sms, &c->launches) != 0)
      return fail(FRX_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  } else {
    int rc = c->ensure_gram_ws(gramian_workspace_floats(n, d, d, c->num_sms));
    if (rc) return rc;
    launch_gramian(dE, n, d, 0, d, 0, d, dw, dG, d, c->gram_ws, c->gram_ws_floats, c->stream, c->num_sms, &c->launches);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dG, sizeof(float) * (size_t)d * d, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(dE); cudaFree(dw); cudaFree(dG);
  return FRX_OK;
}

// G = H T H^T of a symmetric d x d matrix (d = 128 / 256) with the cluster Householder kernel; host in / out.
extern "C" int frx_sym_tridiag(frx_context* c, const float* G, int d, float* H, float* tdiag, float* tsub) {
  CK(cudaSetDevice(c->device));
  if (!sym_tridiag_supported(d)) return fail(FRX_ERR_ARG, "frx_sym_tridiag: d must be 128 or 256");
  float *dG = nullptr, *dH = nullptr, *dHT = nullptr, *dd = nullptr, *ds = nullptr;
  const size_t n2 = (size_t)d * d;
  CK(cudaMalloc(&dG, sizeof(float) * n2));
  CK(cudaMalloc(&dH, sizeof(float) * n2));
  CK(cudaMalloc(&dHT, sizeof(float) * n2));
  CK(cudaMalloc(&dd, sizeof(float) * d));
  CK(cudaMalloc(&ds, sizeof(float) * d));
  CK(cudaMemcpyAsync(dG, G, sizeof(float) * n2, cudaMemcpyHostToDevice, c->stream));
  const int rc = launch_sym_tridiag(dG, d, dH, dHT, dd, ds, c->stream, &c->launches);
  if (rc == 0) {
    CK(cudaMemcpyAsync(H, dH, sizeof(float) * n2, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(tdiag, dd, sizeof(float) * d, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(tsub, ds, sizeof(float) * d, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(dG); cudaFree(dH); cudaFree(dHT); cudaFree(dd); cudaFree(ds);
  if (rc) return fail(FRX_ERR_CUDA, "launch of sym_tridiag_kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  return FRX_OK;
}

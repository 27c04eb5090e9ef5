// Device-side interfaces of the frecsys_b200 CUDA library (sm_100a only).
// Kernel launchers are plain C++ functions taking a stream; frx_api.cu
// sequences them into the reference's Train()/EvaluateDataset() stages.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace frx {

// Which reference projection a row-solve launch restates.
enum RowMode : int {
  RM_IALS = 0,        // IALSRecommender::Project            ials.h:88-144
  RM_SAFER_U = 1,     // SAFER2/ERM ProjectU, CVaR ProjectU_eval  safer2.h:104-163
  RM_SAFER_V = 2,     // SAFER2/ERM ProjectV (stale tail B-1)     safer2.h:166-221
  RM_CVAR_U = 3,      // CVaRMF ProjectU gradient step (B-3,B-4)  cvar_mf.h:88-134
  RM_CVAR_V = 4,      // CVaRMF ProjectV gradient step (B-1,B-3)  cvar_mf.h:136-180
  RM_PP_IALS = 5,     // IALSpp ProjectBlock                  ialspp.h:85-145
  RM_PP_SAFER_U = 6,  // SAFER2pp ProjectU                    safer2pp.h:97-159
  RM_PP_SAFER_V = 7   // SAFER2pp ProjectV                    safer2pp.h:161-216
};

struct RowParams {
  // sparse rows of the side being solved (CSR by user or CSC by item), file order
  const int* ptr;
  const int* col;
  const int* tup;
  const int* order;  // non-empty row ids, longest history first
  int num_rows;      // entries of `order` to process
  // the fixed side
  const float* E;  // [num_other x d]
  int d;
  int num_other;
  int cs, bd;  // column block [cs, cs+bd); full-dimension solves use cs=0, bd=d
  // the side being solved
  float* X;            // [.. x d]
  const float* Xread;  // where the current row is read from (CVaR V step reads pre-update copies)
  const int* xmap;     // row id -> row of X (fold-in evaluation), or null
  const float* G;      // [d x d] (weighted) Gramian of the fixed side
  const float* Xg;     // tensor-core CVaR-MF gradient steps: Xread * G for all rows ([.. x d]), else null
  const float* entry_w;   // per fixed-side-row weight w_c = z_c/n_c (item side) or null
  const float* entry_w_e; // the same weights per history ENTRY (aligned with col: launch_entry_weights), or null
  const float* row_w;     // per-row dual weight z_u (user side) or null (= 1)
  const float* item_reg;  // item side: sum_{u in hist} 1/n_u
  float* pred;            // ++ prediction cache, indexed by tuple
  int mode;
  float uw, reg, reg_exp, alpha, stepsize;
  int num_users_total;  // ItemRegularizationValue's num_choices
  float* scratch;       // per-CTA global matrices when the system does not fit in shared memory
  size_t scratch_stride;
  int use_smem_matrix;
  int* status;  // set non-zero on a non-positive pivot
  // --use_cg (generic kernel only): 0 LLT, 1 Eigen::ConjugateGradient<Lower> (ials.h:133-138, safer2.h:152-157),
  // 2 Eigen::BiCGSTAB on the unsymmetrised full matrix (erm_mf.h:139-145, SURVEY B-5)
  int solver;
  float cg_tol;
  int cg_max_it;
  int* work_counter;  // tensor-core row kernel: work queue of the launch (zeroed by the launcher)
  // tensor-core row kernel: bit patterns of max|E| and max entry weight (launch_absmax); the SYRK operands are
  // fp16 hi + lo pairs of the entries scaled by a power of two derived from this bound
  const unsigned* syrk_absmax;
  unsigned long long* dbg;  // optional per-phase cycle counters (FRX_TC_DEBUG), else null
  // Long rows (tensor-core path): a row with more than FRX_SPLIT_MIN entries is cut into pieces of
  // FRX_PIECE entries whose partial SYRK sums are produced by a first launch (piece_mode = 1, one
  // work item per piece) into piece_scratch; the main launch then starts such a row from the sum of
  // its pieces instead of gathering it, so that no single CTA serialises a 80K-entry history.
  const int* piece_row;    // [num_pieces] row id
  const int* piece_off;    // [num_pieces] first entry of the piece within the row
  const int* row_piece0;   // [rows] first piece of the row, or -1 (null when nothing is split)
  int num_pieces;
  int piece_mode;
  float* piece_scratch;    // [num_pieces][piece_stride]: lower 32x32 chunks (row-major), then the rhs partial
  size_t piece_stride;
};
// (The piece length also bounds the accumulation chain inside the tensor cores, whose fp32 accumulator rounds
// toward zero: 64 K-steps of 16 entries per TMEM pass keep that bias near 4e-6; the pieces are added with IEEE
// fp32 adds.)
constexpr int FRX_STAGE_CAP = 1792;  // history entries of a work item whose indices / weights the tensor-core row kernel stages
constexpr int FRX_SPLIT_MIN = FRX_STAGE_CAP;
constexpr int FRX_PIECE = 1024;
size_t row_solve_tc_piece_floats(int d);  // piece_stride for dimension d

void launch_row_solve_generic(const RowParams& p, cudaStream_t s, int num_sms, long long* launches);
size_t row_solve_generic_scratch_floats(int bd, int solver = 0);  // per-CTA scratch (0 if shared memory suffices)
int row_solve_generic_grid(int num_rows, int num_sms);

// tcgen05 / TMEM path (frx_row_tc.cu): full-dimension solves with d = 128 or 256.
bool row_solve_tc_supported(const RowParams& p);
// ew[e] = w[col[e]] for the entries e in [e_begin, e_end)
void launch_entry_weights(const int* col, const float* w, size_t e_begin, size_t e_end, float* ew, cudaStream_t s,
                          long long* launches);
// out[0] = bits(max |E[i]|), out[1] = bits(max |w[i]|) (0 without w); non-negative floats order like uints
void launch_absmax(const float* E, size_t n, const float* w, size_t nw, unsigned* out, cudaStream_t s, int num_sms,
                   long long* launches);
void launch_row_solve_tc(const RowParams& p, cudaStream_t s, int num_sms, long long* launches);

// Dual-form row path (frx_row_wb.cu) for rows with at most FRX_WB_MAX entries: needs the tridiagonal form
// G = H T H^T of the Gramian (frx_eig.cu) and the fixed-side factors rotated into that basis (frx_gemm.cu).
constexpr int FRX_WB_MAX = 128;
struct WbParams {
  const int* grp_slots;  // [num_groups][4]: (index into wb_rows << 2) | 32-entry chunk of that row, or -1
  const int* wb_rows;    // row ids solved by this path
  int num_groups;
  const float* Et;       // [num_other x d] = E * H
  const float* tdiag;    // [d] diagonal of T
  const float* tsub;     // [d] tsub[j] = T[j][j-1], tsub[0] = 0
  float* lsub;           // [num wb rows x d] scratch: bidiagonal factor of alpha*T + beta*I per row
  float* rsd;            // [num wb rows x d] scratch: its Dl^(-1/2)
  float* Xt;             // [num wb rows x d] rotated solutions (the caller applies H^T)
  int* counter;          // work queue (zeroed by the launcher)
};
bool row_solve_wb_supported(const RowParams& p);
void launch_row_solve_wb(const RowParams& p, const WbParams& q, int num_wb_rows, cudaStream_t s, int num_sms,
                         long long* launches);

// G = H T H^T for the symmetric d x d Gramian (d = 128 / 256), T tridiagonal: tdiag[d], tsub[d] (tsub[j] =
// T[j][j-1], tsub[0] = 0), H and HT = H^T row-major, all fp32.  Returns 0 on success.
bool sym_tridiag_supported(int d);
int launch_sym_tridiag(const float* G, int d, float* H, float* HT, float* tdiag, float* tsub, cudaStream_t s,
                       long long* launches);

// C[out(i)][:] = A[i][:] * B, A [M x d], B [d x d] row-major, out(i) = c_map[c_rows[i]] (either may be null).
void launch_rows_gemm(const float* A, int M, int d, const float* B, float* C, const int* c_rows, const int* c_map,
                      cudaStream_t s, long long* launches);

// out[bd x fd] = sum_r w[r] * E[r][cs+i] * E[r][fs+j]; two-stage, deterministic.
void launch_gramian(const float* E, int n, int d, int cs, int bd, int fs, int fd, const float* w,
                    float* out, int ld_out, float* workspace, size_t workspace_floats,
                    cudaStream_t s, int num_sms, long long* launches);
size_t gramian_workspace_floats(int n, int bd, int fd, int num_sms);

// tcgen05 + TMA Gramian (frx_gramian_tc.cu): full d x d, d = 128 or 256.  Returns 0 on success.
bool gramian_tc_supported(int n, int d, int cs, int bd, int fs, int fd);
size_t gramian_tc_workspace_floats(int d, int num_sms);
int launch_gramian_tc(const float* E, int n, int d, const float* w, float* out, float* workspace, cudaStream_t s,
                      int num_sms, long long* launches);

struct LossParams {
  const int* ptr;
  const int* col;
  const int* tup;
  const int* order;
  int num_rows;
  const float* U;
  const float* V;
  int d;
  const float* G;       // item Gramian
  const float* pred;    // SAFER2++: cached predictions (else null)
  float beta;           // unobserved weight
  int halve;            // SAFER2-family /2 (safer2.h:99); iALS family not (ials.h:70-86)
  float* quad;          // [num_users] scratch: u^T G u
  float* loss;          // [num_users] out (rows without history untouched)
  double* obs_sq;       // optional [num_users]: sum (pred-1)^2 in double (stats)
  // Two-pass form (used when resid != null and pred == null): pass 1 writes the residual of every
  // history entry, one warp per FRX_LOSS_CHUNK-entry chunk (chunk_row / chunk_off), so that a 6K-entry
  // history is spread over many warps; pass 2 adds them per user in history order exactly like the
  // reference's running float sum.
  float* resid;          // [nnz] scratch, indexed like col
  const int* chunk_row;
  const int* chunk_off;
  int num_chunks;
};
constexpr int FRX_LOSS_CHUNK = 256;
void launch_quadform(const LossParams& p, int user_begin, int user_end, cudaStream_t s, long long* launches);
void launch_user_loss_rows(const LossParams& p, cudaStream_t s, int num_sms, long long* launches);
void launch_user_loss(const LossParams& p, int user_begin, int user_end, cudaStream_t s, int num_sms,
                      long long* launches);

// pred[t] = v_i . u for every tuple (PredictDataset ialspp.h:469-517).
void launch_predict(const int* ptr, const int* col, const int* tup, const int* order, int num_rows,
                    const float* U, const int* xmap, const float* V, int d, float* pred,
                    cudaStream_t s, long long* launches);
// the same over FRX_LOSS_CHUNK-entry chunks of the rows (balanced: long rows are spread over many warps)
void launch_predict_chunks(const int* ptr, const int* col, const int* tup, const int* chunk_row, const int* chunk_off,
                           int num_chunks, const float* U, const int* xmap, const float* V, int d, float* pred,
                           cudaStream_t s, int num_sms, long long* launches);

// z_u update (safer2.h:745-794, safer2pp.h:839-862, cvar_mf.h:597-642).
// kind: 0 gaussian, 1 epanechnikov, 2 indicator.  only_with_history: skip rows with hist_size==0.
void launch_user_weights(const float* loss, const float* hist_size, int num_users, float* z,
                         const float* xi_dev, float bandwidth, int kind, int only_with_history,
                         cudaStream_t s, long long* launches);

// Newton / Armijo xi estimation (safer2.h:652-742) in one cooperative kernel.
// xi_io: device scalar, in = start value, out = result.  If start_from_mean,
// the start value is mean(loss) (Initialize, safer2.h:822).
struct XiParams {
  const float* loss;
  int num_users;
  const int* snr_idx;  // [iters x n_samples] or null
  int n_samples;
  int iters;
  float alpha, bandwidth;
  int epanechnikov;
  int start_from_mean;
  float* xi_io;
  double* partials;  // [2][grid][3]
  unsigned* barrier; // [2] zero-initialised
};
int launch_xi_newton(const XiParams& p, cudaStream_t s, int num_sms, long long* launches);
size_t xi_partials_doubles(int num_sms);

// Exact quantile (cvar_mf.h:582-595): -vals[Q] where vals = sorted(-loss), Q = (size_t)(n*alpha).
void launch_exact_quantile(const float* loss, int n, float alpha, float* xi_out, unsigned* hist_ws,
                           cudaStream_t s, long long* launches);

// mean(z), mean(z*loss) in double -> scalars_out[0..1] (float)
void launch_weight_means(const float* z, const float* loss, int n, float* out2, double* ws,
                         cudaStream_t s, long long* launches);

// Initialize(): hist_size[u] = n_u; item_reg[v] = sum_{u in hist(v)} 1/n_u (safer2.h:826-837)
void launch_hist_and_item_reg(const int* uptr, int num_users_ds, float* hist_size, const int* iptr,
                              const int* icol, int num_items_ds, float* item_reg, cudaStream_t s,
                              long long* launches);

// norm_w[u] = z[u] / hist_size[u] (safer2.h:502-503)
void launch_norm_weights(const float* z, const float* hist_size, int n, float* out, cudaStream_t s,
                         long long* launches);

void launch_fill(float* p, size_t n, float v, cudaStream_t s, long long* launches);
// PrintLosses / ComputeLosses sums on the device (safer2.h:337-413, ials.h:226-305): out2[0] += sum_r |x_r|^2 * reg_r,
// out2[1] += sum_r |x_r|^2 over the rows with history; kind 0 iALS, 1 SAFER2-family user, 2 SAFER2-family item.
void launch_reg_sums(const float* X, const int* ptr, int rows, int d, int kind, float reg, float reg_exp, float uw,
                     float alpha, int num_other, const float* item_reg, double* out2, cudaStream_t s, int num_sms,
                     long long* launches);
void launch_dot_sum(const float* a, const float* b, size_t n, double* out, cudaStream_t s, int num_sms, long long* launches);
void launch_sum_d(const double* a, size_t n, double* out, cudaStream_t s, int num_sms, long long* launches);
// *out += sum (a[i] - b[i])^2 (double); residual statistics (safer2.h:475-478, 550-553, 789-792)
void launch_sqdiff(const float* a, const float* b, size_t n, double* out, cudaStream_t s, int num_sms, long long* launches);

// Evaluation (recommender.h:78-199): scores = Ut * V^T tile by tile, history
// mask, per-user top-k, Recall/NDCG.
struct EvalParams {
  const float* Ut;  // [nu x d] folded-in users
  const float* V;   // [num_items x d]
  int nu, num_items, d;
  const int* user_ids;               // [nu] ascending ids
  const int* tr_ptr; const int* tr_col;  // test_tr CSR (history to mask)
  const int* te_ptr; const int* te_col;  // test_te CSR (ground truth)
  int te_rows;                       // rows available in te_ptr
  const int* k_list; int nk; int max_k;
  float* scores;                     // [chunk_users x num_items] workspace
  int chunk_users;
  int* topk;                         // [nu x max_k]
  float* recall; float* ndcg;        // [nu x nk]
};
void launch_evaluate(const EvalParams& p, cudaStream_t s, int num_sms, long long* launches);

// Fused scoring + top-k (frx_score_topk.cu): tcgen05 Ut * V^T with the per-user top-max_k kept in the epilogue;
// no score matrix.  d % 32 == 0 and max_k <= FRX_TOPK_PAD.  The candidate lists ([nu][segments][FRX_TOPK_PAD]
// 64-bit keys) go to launch_merge_metrics, which sorts them and computes Recall / NDCG.
constexpr int FRX_TOPK_PAD = 128;
constexpr int FRX_TOPK_CAND = 1024;
bool score_topk_supported(int d, int max_k);
size_t score_topk_workspace_bytes(int nu, int num_items, int d, int num_sms);
int launch_score_topk(const EvalParams& p, void* workspace, int num_sms, cudaStream_t s, long long* launches,
                      const unsigned long long** cand_out, int* segments_out);
void launch_merge_metrics(const EvalParams& p, const unsigned long long* lists, int segments, cudaStream_t s,
                          long long* launches);

// Dataset build: stable sort of tuple ids by row id -> ptr/col/tup (dataset.h:86-91).
void build_csr(const int* d_keys, const int* d_other, int n, int nrows, int* ptr, int* col, int* tup,
               cudaStream_t s, long long* launches);

}  // namespace frx

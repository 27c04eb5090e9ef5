/* frecsys_b200 — C ABI of the B200-native ALS hot path of frecsys
 * (riktor/safer2-recommender).
 *
 * The reference has no FFI: its "plugin boundary" is the C++ class
 * frecsys::Recommender (include/frecsys/recommender.h:40-130) and its six
 * subclasses, constructed by tools/run_model.cc:43-123.  The host-side mirror
 * of those classes lives in include/frecsys/ (same class names, constructor
 * argument orders and methods); each of their methods forwards to one entry
 * point below.  Every function is synchronous on return unless noted, takes
 * plain pointers/sizes, returns 0 on success or a negative frx_status, and
 * leaves a message retrievable with frx_last_error().  There is no CPU
 * fallback: without a CUDA device every compute call fails with
 * FRX_ERR_CUDA.
 */
#ifndef FRECSYS_B200_H_
#define FRECSYS_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct frx_context frx_context; /* one GPU: device, stream, scratch, optional NCCL comm */
typedef struct frx_dataset frx_dataset; /* device-resident CSR (by user) + CSC (by item) of one Dataset */
typedef struct frx_model frx_model;     /* factors + per-user state of one recommender */

enum frx_status {
  FRX_OK = 0,
  FRX_ERR_CUDA = -1,     /* CUDA runtime / launch failure (message has the CUDA error) */
  FRX_ERR_ARG = -2,      /* bad argument */
  FRX_ERR_NUMERIC = -3,  /* non-SPD system or NaN (reference: assert safer2.h:160, exit safer2.h:399-404) */
  FRX_ERR_COMM = -4      /* NCCL failure */
};

/* --model_name values, tools/run_model.cc:206-213 */
enum frx_model_kind {
  FRX_IALS = 0, FRX_IALSPP = 1, FRX_ERM_MF = 2, FRX_CVAR_MF = 3, FRX_SAFER2 = 4, FRX_SAFER2PP = 5
};

/* The run_model flags that reach the model constructors (run_model.cc:43-123). */
typedef struct frx_config {
  int model;             /* frx_model_kind                      --model_name */
  int dim;               /*                                     --dim */
  float reg;             /*                                     --l2_reg */
  float reg_exp;         /* iALS / iALS++ only                  --l2_reg_exp */
  float uobs_weight;     /*                                     --uobs_weight */
  float stdev;           /*                                     --stdev */
  float alpha;           /*                                     --alpha */
  float bandwidth;       /* SAFER2 / SAFER2++                   --bandwidth */
  float stepsize;        /* CVaR-MF                             --stepsize */
  int xi_iterations;     /*                                     --xi_iterations */
  int pd_iterations;     /*                                     --pd_iterations */
  int use_epanechnikov;  /*                                     --use_epanechnikov */
  int use_snr;           /*                                     --use_snr */
  float sampling_ratio;  /*                                     --sampling_ratio */
  int use_cg;            /* accepted; the CUDA path always solves by Cholesky (--use_cg) */
  float cg_tol;          /*                                     --cg_error_tolerance */
  int cg_max_it;         /*                                     --cg_max_iterations */
  int block_size;        /* iALS++ / SAFER2++                   --block_size */
  unsigned snr_seed;     /* base seed of the SNR index draws (reference: random_device, safer2.h:728) */
} frx_config;

const char* frx_last_error(void);

/* ---- context ------------------------------------------------------------ */
/* `cuda_stream` may be NULL (the library creates its own stream) or a
 * cudaStream_t owned by the caller (e.g. torch's current stream). */
int frx_context_create(int device, void* cuda_stream, frx_context** out);
void frx_context_destroy(frx_context* ctx);
int frx_context_sync(frx_context* ctx);
void* frx_context_stream(frx_context* ctx);
/* Multi-GPU (row-sharded data parallel): one context per rank.
 * rank 0 calls frx_comm_unique_id and ships the 128 bytes to every rank. */
int frx_comm_unique_id(void* out128);
int frx_context_init_comm(frx_context* ctx, int rank, int world_size, const void* unique_id128);
/* Host-only: the contiguous row ranges the ranks own, balanced on sum(history length + row_unit) over the non-empty
 * rows; row_unit < 0 selects the built-in cost model of frx_dataset_create (direct rows: length + 480, rows of at
 * most 128 entries -- the dual-form kernel -- a flat 150);
 * ptr[nrows+1] is a CSR row pointer, rank_begin receives world+1 entries. */
int frx_partition_rows(const int* ptr, int nrows, int world, int row_unit, int* rank_begin);

/* ---- dataset: replaces frecsys::Dataset's by_user_/by_item_ build ---------
 * (include/frecsys/dataset.h:71-99).  Input is the tuple list in FILE order;
 * the library builds, on the device, the CSR by user and the CSC by item whose
 * rows list (other_id, tuple_index) in file order — bit-identical to the
 * reference's hash-of-vectors (a stable sort by row id). */
int frx_dataset_create(frx_context* ctx, int num_tuples, const int* users, const int* items,
                       frx_dataset** out);
void frx_dataset_destroy(frx_dataset* ds);
/* out5 = max_user, max_item, num_tuples, distinct users, distinct items (dataset.h:94-98) */
int frx_dataset_info(frx_dataset* ds, int* out5);
/* Download one orientation: ptr[nrows+1], ids[num_tuples], tup[num_tuples]. */
int frx_dataset_get_csr(frx_dataset* ds, int by_item, int nrows, int* ptr, int* ids, int* tup);

/* ---- model: replaces the six Recommender subclasses ------------------------ */
/* Constructor (e.g. SAFER2Recommender safer2.h:37-77): allocates factors and
 * state; factors are zero until init/set. */
int frx_model_create(frx_context* ctx, const frx_config* cfg, int num_users, int num_items,
                     frx_model** out);
void frx_model_destroy(frx_model* m);
/* Recommender::init_matrix (recommender.h:61-67) with an explicit seed: host
 * std::mt19937 + std::normal_distribution<float>, U then V, then uploaded. */
int frx_model_init_factors(frx_model* m, unsigned seed);
/* Host row-major fp32 [num_users x dim], [num_items x dim]; either may be NULL.
 * Setting factors also resets z=alpha, loss=0, xi=0 and G_V = V^T V as the
 * constructors do (safer2.h:55-59). */
int frx_model_set_factors(frx_model* m, const float* U, const float* V);
int frx_model_get_factors(frx_model* m, float* U, float* V);
/* Overwrite the factors WITHOUT touching the per-user state (asynchronous: the host
 * buffers must stay valid until frx_context_sync; use pinned memory).  This is the
 * per-epoch host->device leg of a caller that keeps its factors in host memory. */
int frx_model_upload_factors(frx_model* m, const float* U, const float* V);
/* Row-sharded variants for a multi-rank job (one process per GPU) whose factors live in host memory:
 * each rank copies only the rows it owns under frx_partition_rows over `train` (by user for U, by item
 * for V); after the upload the blocks are all-gathered over NCCL so every rank holds the full factors,
 * after the download the caller's host array holds this rank's rows (other rows untouched).  With one
 * rank they equal the calls above.  The reference keeps U, V in one process' memory (recommender.h);
 * this is the per-process view of that array. */
int frx_model_upload_factors_sharded(frx_model* m, frx_dataset* train, const float* U, const float* V);
int frx_model_get_factors_sharded(frx_model* m, frx_dataset* train, float* U, float* V);
/* Train() (below) followed by frx_model_get_factors_sharded, for a caller that keeps its factors in host
 * memory like the reference does (recommender.h): the device->host copies of U and V are started as soon as
 * their last half-step of the epoch is final and run under the rest of the epoch.  Synchronous. */
int frx_model_train_to_host(frx_model* m, frx_dataset* train, float* U, float* V);
/* Checkpoint / resume (the reference has none): factors, dual weights, per-user loss, history
 * sizes, item regularisation sums, xi, running means and the ComputeXi call counter (SNR seeds).  A model
 * created with the same kind, dim and sizes and loaded from the file continues bit-identically; no
 * Initialize() call is needed (or wanted: it would recompute xi from the mean loss, safer2.h:822). */
int frx_model_save(frx_model* m, const char* path);
int frx_model_load(frx_model* m, const char* path);
/* Initialize(const Dataset&) — safer2.h:819-838, safer2pp.h, erm_mf.h:573-587,
 * cvar_mf.h:710-726; a no-op for iALS / iALS++ (run_model.cc:246-257). */
int frx_model_initialize(frx_model* m, frx_dataset* train);
/* Train(const Dataset&): one epoch — safer2.h:266-334, ials.h:187-224,
 * ialspp.h:208-261, erm_mf.h:257-301, cvar_mf.h:276-330, safer2pp.h:288-355.
 * Asynchronous: returns after enqueueing; any getter or frx_context_sync waits. */
int frx_model_train(frx_model* m, frx_dataset* train);
/* One stage of the epoch, for per-kernel parity tests and profiling:
 *  0 ComputeUserWeights(prev_xi)  1 StepU  2 StepV  3 G_V = V^T V
 *  4 ComputeUserLoss  5 xi = ComputeXi  6 iALS user Step  7 iALS item Step */
int frx_model_stage(frx_model* m, frx_dataset* train, int stage);
/* Any output may be NULL.  scalars[3] = prev_xi, last weighted loss, mean z. */
int frx_model_get_state(frx_model* m, float* z, float* loss, float* hist_size, float* item_reg,
                        float* scalars, float* gramian);
int frx_model_set_state(frx_model* m, const float* z, const float* loss, float xi);
/* SetPrintTrainStats + the numbers PrintLosses/ComputeLosses log
 * (safer2.h:337-413, ials.h:226-305): out6 = Loss, Loss_observed,
 * Loss_unobserved, Loss_reg, Loss_reg(user), Loss_reg(item). */
int frx_model_compute_stats(frx_model* m, frx_dataset* train, double* out6);
/* SetPrintResidualStats(bool) (safer2.h:804-806 and siblings): when on, Train() also records, per primal-dual
 * iteration, the norms the reference logs as "U residual / V residual / z residual" (safer2.h:323-328,475-478,
 * 550-553,789-792): |U_new - U_old| over the solved rows, |V_new - V_old|_F, |z_new - z_old|.
 * frx_model_get_residuals copies up to max_triples triples (U, V, z) of the last Train() and returns their
 * count; iALS reports 0, 0 (ials.h:363-364) and CVaR-MF 0 for U (cvar_mf.h:472-473) like the reference. */
int frx_model_set_residual_stats(frx_model* m, int on);
int frx_model_get_residuals(frx_model* m, float* out, int max_triples);
/* The SNR indices drawn by the last ComputeXi: [n_iters x n_samples]. */
int frx_model_last_snr(frx_model* m, int* n_iters, int* n_samples, int* out);
/* EvaluateDataset(k_list, alpha_list, data=test_tr, eval_by_user=test_te) —
 * safer2.h:225-263 + recommender.h:78-199: fold-in solve of the held-out users
 * from their test_tr history, scores V*u, history mask, top-max_k, Recall@k and
 * NDCG@k.  Rows follow ascending user id of test_tr.  Returns the number of
 * evaluated users (call with recall == NULL to size the outputs).
 * user_ids[nu], recall/ndcg[nu*nk], topk[nu*max_k] (nullable), folded[nu*dim] (nullable). */
int frx_model_evaluate(frx_model* m, frx_dataset* test_tr, frx_dataset* test_te, const int* k_list,
                       int nk, int* user_ids, float* recall, float* ndcg, int* topk, float* folded);

/* ---- instrumentation ---------------------------------------------------------- */
/* Number of kernels this library has launched since the context was created. */
long long frx_context_launch_count(frx_context* ctx);
/* Per-stage device time of the last frx_model_train, by CUDA events on the
 * context's stream (enabled with frx_context_set_profiling).  names is a
 * ';'-separated list written into buf; ms[] receives up to max_n values. */
int frx_context_set_profiling(frx_context* ctx, int on);
int frx_context_stage_times(frx_context* ctx, char* names_buf, int buf_len, float* ms, int max_n);
/* Standalone Gramian G = E^T diag(w) E for tests / microbenchmarks (host in/out). */
int frx_gramian(frx_context* ctx, const float* E, int n, int d, const float* w, float* out);
/* Standalone tridiagonal reduction G = H T H^T (d = 128 or 256; host in/out) with the kernel the dual-form row
 * path runs on the Gramian: H row-major [d x d] orthogonal, tdiag[d] = diag(T), tsub[d] with tsub[j] = T[j][j-1]
 * (tsub[0] = 0).  Replaces nothing in the reference (it has no such step): the dual form solves the same
 * systems as Eigen::LLT in safer2.h:159-161 / ials.h:140-142. */
int frx_sym_tridiag(frx_context* ctx, const float* G, int d, float* H, float* tdiag, float* tsub);

#ifdef __cplusplus
}
#endif
#endif /* FRECSYS_B200_H_ */

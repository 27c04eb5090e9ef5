// C ABI over the CPU oracle (frecsys_oracle.hpp) for ctypes — TEST
// INFRASTRUCTURE ONLY (see the header of frecsys_oracle.hpp).  Loaded by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
#include "frecsys_oracle.hpp"

using namespace oracle;

extern "C" {

// Mirrors oracle::Config field by field (plain C layout for ctypes).
struct OrcConfig {
  int model, dim;
  float reg, reg_exp, uobs_weight, stdev, alpha, bandwidth, stepsize;
  int xi_iterations, pd_iterations, use_epanechnikov, use_snr;
  float sampling_ratio;
  int use_cg;
  float cg_tol;
  int cg_max_it, block_size;
  unsigned snr_seed;
};

static Config ToConfig(const OrcConfig* c) {
  Config k;
  k.model = c->model; k.dim = c->dim; k.reg = c->reg; k.reg_exp = c->reg_exp;
  k.uobs_weight = c->uobs_weight; k.stdev = c->stdev; k.alpha = c->alpha;
  k.bandwidth = c->bandwidth; k.stepsize = c->stepsize;
  k.xi_iterations = c->xi_iterations; k.pd_iterations = c->pd_iterations;
  k.use_epanechnikov = c->use_epanechnikov; k.use_snr = c->use_snr;
  k.sampling_ratio = c->sampling_ratio; k.use_cg = c->use_cg; k.cg_tol = c->cg_tol;
  k.cg_max_it = c->cg_max_it; k.block_size = c->block_size; k.snr_seed = c->snr_seed;
  return k;
}

void* orc_dataset_from_csv(const char* path) { return new Dataset(Dataset::FromCsv(path)); }
void* orc_dataset_from_tuples(const int* users, const int* items, int n) {
  return new Dataset(Dataset::FromTuples(users, items, n));
}
void orc_dataset_free(void* d) { delete (Dataset*)d; }
void orc_dataset_info(void* dv, int* out5) {
  Dataset* d = (Dataset*)dv;
  out5[0] = d->max_user; out5[1] = d->max_item; out5[2] = d->num_tuples;
  out5[3] = d->distinct_users; out5[4] = d->distinct_items;
}
// Flat export of by_user (by_item=0) or by_item (=1): ptr[nrows+1], ids[nnz], tup[nnz].
void orc_dataset_csr(void* dv, int by_item, int nrows, int* ptr, int* ids, int* tup) {
  Dataset* d = (Dataset*)dv;
  const std::vector<SpVector>& rows = by_item ? d->by_item : d->by_user;
  int k = 0;
  for (int r = 0; r < nrows; ++r) {
    ptr[r] = k;
    if (r < (int)rows.size())
      for (auto& p : rows[r]) { ids[k] = p.first; tup[k] = p.second; ++k; }
  }
  ptr[nrows] = k;
}
// Tuple list in file order (users[t], items[t]).
void orc_dataset_tuples(void* dv, int* users, int* items) {
  Dataset* d = (Dataset*)dv;
  for (size_t u = 0; u < d->by_user.size(); ++u)
    for (auto& p : d->by_user[u]) { users[p.second] = (int)u; items[p.second] = p.first; }
}

void* orc_model_create(const OrcConfig* c, int num_users, int num_items, unsigned init_seed) {
  return new Model(ToConfig(c), num_users, num_items, init_seed);
}
void orc_model_free(void* m) { delete (Model*)m; }
void orc_model_set_factors(void* mv, const float* U, const float* V) { ((Model*)mv)->SetFactors(U, V); }
void orc_model_get_factors(void* mv, float* U, float* V) {
  Model* m = (Model*)mv;
  if (U) std::copy(m->U.a.begin(), m->U.a.end(), U);
  if (V) std::copy(m->V.a.begin(), m->V.a.end(), V);
}
void orc_model_initialize(void* mv, void* d) { ((Model*)mv)->Initialize(*(Dataset*)d); }
void orc_model_train(void* mv, void* d) { ((Model*)mv)->Train(*(Dataset*)d); }
void orc_model_set_print_train_stats(void* mv, int on) { ((Model*)mv)->print_train_stats = on != 0; }
// scalars: [0]=prev_xi, [1]=last weighted loss, [2]=mean dual weight
void orc_model_get_state(void* mv, float* z, float* loss, float* hist_size, float* item_reg,
                         float* scalars, float* gramian) {
  Model* m = (Model*)mv;
  if (z) std::copy(m->dual_weight.begin(), m->dual_weight.end(), z);
  if (loss) std::copy(m->user_loss.begin(), m->user_loss.end(), loss);
  if (hist_size) std::copy(m->user_history_size.begin(), m->user_history_size.end(), hist_size);
  if (item_reg) std::copy(m->item_reg.begin(), m->item_reg.end(), item_reg);
  if (scalars) { scalars[0] = m->prev_xi; scalars[1] = m->last_weighted_loss; scalars[2] = m->GetMeanWeight(); }
  if (gramian) std::copy(m->item_gramian.a.begin(), m->item_gramian.a.end(), gramian);
}
void orc_model_set_state(void* mv, const float* z, const float* loss, float xi) {
  Model* m = (Model*)mv;
  if (z) std::copy(z, z + m->num_users, m->dual_weight.begin());
  if (loss) std::copy(loss, loss + m->num_users, m->user_loss.begin());
  m->prev_xi = xi;
}
void orc_model_get_stats(void* mv, double* out6) {
  const LossStats& s = ((Model*)mv)->last_stats;
  out6[0] = s.loss; out6[1] = s.loss_observed; out6[2] = s.loss_unobserved; out6[3] = s.loss_reg;
  out6[4] = s.loss_reg_user; out6[5] = s.loss_reg_item;
}
void orc_model_compute_stats(void* mv, void* d, double* out6) {
  Model* m = (Model*)mv;
  m->last_stats = m->ComputeStats(*(Dataset*)d);
  orc_model_get_stats(mv, out6);
}
// Number of SNR index vectors drawn by the last ComputeXi and their common length.
void orc_model_last_snr(void* mv, int* n_iters, int* n_samples, int* out) {
  Model* m = (Model*)mv;
  *n_iters = (int)m->last_snr_indices.size();
  *n_samples = m->last_snr_indices.empty() ? 0 : (int)m->last_snr_indices[0].size();
  if (out)
    for (size_t t = 0; t < m->last_snr_indices.size(); ++t)
      std::copy(m->last_snr_indices[t].begin(), m->last_snr_indices[t].end(),
                out + t * (size_t)*n_samples);
}

// Stage-level entry points for per-kernel parity tests.
//  0: ComputeUserWeights(prev_xi)   1: StepU (model's user step)   2: StepV
//  3: item_gramian = V^T V          4: ComputeUserLoss             5: prev_xi = ComputeXi
//  6: iALS user Step                7: iALS item Step
void orc_model_stage(void* mv, void* dv, int stage) {
  Model* m = (Model*)mv;
  Dataset& d = *(Dataset*)dv;
  switch (stage) {
    case 0: m->ComputeUserWeights(d, m->prev_xi); break;
    case 1:
      if (m->cfg.model == kCVaRMF) m->StepU_CVaR(d);
      else m->StepU(d.by_user, &m->U, nullptr, m->V, m->item_gramian, m->dual_weight.data());
      break;
    case 2:
      if (m->cfg.model == kCVaRMF) { Mat up = m->U; m->StepV_CVaR(d, up); }
      else m->StepV(d, m->U, &m->V);
      break;
    case 3: m->item_gramian = Gramian(m->V); break;
    case 4: {
      Mat G = m->is_ials_family() ? Gramian(m->V) : m->item_gramian;
      m->ComputeUserLoss(d, G, nullptr);
      break;
    }
    case 5:
      m->prev_xi = m->cfg.model == kCVaRMF ? m->ComputeXiExact(m->user_loss)
                                           : m->ComputeXi(m->user_loss, m->prev_xi, m->cfg.xi_iterations);
      break;
    case 6: m->StepIals(d.by_user, &m->U, nullptr, m->V); break;
    case 7: m->StepIals(d.by_item, &m->V, nullptr, m->U); break;
  }
}

// EvaluateDataset.  user_ids[nu], recall/ndcg[nu*nk], topk[nu*max_k] (nullable),
// folded[nu*dim] (nullable).  Returns nu (call with null outputs to size).
int orc_model_evaluate(void* mv, void* trv, void* tev, const int* k_list, int nk, int* user_ids,
                       float* recall, float* ndcg, int* topk, float* folded) {
  Model* m = (Model*)mv;
  Dataset& tr = *(Dataset*)trv;
  if (!recall) return tr.distinct_users;
  std::vector<int> ks(k_list, k_list + nk);
  Mat UE;
  EvalResult r = m->Evaluate(tr, *(Dataset*)tev, ks, topk != nullptr, &UE);
  std::copy(r.user_ids.begin(), r.user_ids.end(), user_ids);
  std::copy(r.recall.a.begin(), r.recall.a.end(), recall);
  std::copy(r.ndcg.a.begin(), r.ndcg.a.end(), ndcg);
  if (topk) std::copy(r.topk.begin(), r.topk.end(), topk);
  if (folded) std::copy(UE.a.begin(), UE.a.end(), folded);
  return (int)r.user_ids.size();
}

void orc_metric_cvar(const float* ms, int n, const float* alphas, int na, float* out) {
  std::vector<float> r = MetricCVaR(std::vector<float>(ms, ms + n), std::vector<float>(alphas, alphas + na));
  std::copy(r.begin(), r.end(), out);
}

void orc_gramian(const float* E, int n, int d, const float* w, float* out) {
  Mat M(n, d);
  std::copy(E, E + (size_t)n * d, M.a.begin());
  Mat G = Gramian(M, w);
  std::copy(G.a.begin(), G.a.end(), out);
}

void orc_init_factors(int nu, int ni, int d, float stdev, unsigned seed, float* U, float* V) {
  Mat MU(nu, d), MV(ni, d);
  InitFactors(&MU, &MV, stdev, seed);
  std::copy(MU.a.begin(), MU.a.end(), U);
  std::copy(MV.a.begin(), MV.a.end(), V);
}

int orc_num_threads() { return NumThreads(); }

// ---- hooks for the CPU emulation of the row-sharded epoch (tests/test_dist_cpu.py) ----
void orc_model_set_range(void* mv, int lo, int hi) { ((Model*)mv)->row_lo = lo; ((Model*)mv)->row_hi = hi; }
void orc_model_put_factors(void* mv, const float* U, const float* V) {  // overwrite without resetting the state
  Model* m = (Model*)mv;
  if (U) std::copy(U, U + m->U.a.size(), m->U.a.begin());
  if (V) std::copy(V, V + m->V.a.size(), m->V.a.begin());
}
void orc_model_set_item_gramian(void* mv, const float* G) {
  Model* m = (Model*)mv;
  std::copy(G, G + m->item_gramian.a.size(), m->item_gramian.a.begin());
}
void orc_model_set_gz_override(void* mv, const float* G) {
  Model* m = (Model*)mv;
  m->use_gz_override = G != nullptr;
  if (G) {
    m->gz_override = Mat(m->cfg.dim, m->cfg.dim);
    std::copy(G, G + m->gz_override.a.size(), m->gz_override.a.begin());
  }
}

// One block sweep of a ++ model on the rows in [row_lo, row_hi) of one side (ialspp.h:351-424, safer2pp.h:448-609)
// with the caller's prediction cache, and the refresh of that cache for the same rows from the current factors:
// the two steps every rank of the multi-GPU block solvers performs (csrc/frx_api.cu stage_block).
void orc_model_block_step(void* mv, void* dv, int item_side, int bs, int be, float* pred) {
  Model* m = (Model*)mv;
  Dataset& d = *(Dataset*)dv;
  std::vector<float> p(pred, pred + d.num_tuples);
  const bool safer = m->cfg.model == kSAFER2pp;
  if (!item_side) m->StepBlockUserSide(d.by_user, &m->U, nullptr, m->V, bs, be, &p, safer, safer ? m->dual_weight.data() : nullptr);
  else if (safer) m->StepBlockV_Safer(d, bs, be, &p);
  else m->StepBlockUserSide(d.by_item, &m->V, nullptr, m->U, bs, be, &p, false, nullptr);
  std::copy(p.begin(), p.end(), pred);
}
void orc_model_predict_rows(void* mv, void* dv, int item_side, float* pred) {
  Model* m = (Model*)mv;
  Dataset& d = *(Dataset*)dv;
  const std::vector<SpVector>& rows = item_side ? d.by_item : d.by_user;
  const Mat& X = item_side ? m->V : m->U;
  const Mat& E = item_side ? m->U : m->V;
  for (int r = 0; r < (int)rows.size(); ++r) {
    if (!m->InRange(r)) continue;
    for (const auto& ir : rows[r]) pred[ir.second] = (float)Dot(E.row(ir.first), X.row(r), E.cols);
  }
}

}  // extern "C"

// frecsys::Recommender and the device-backed base of the six models.
//
// Same public interface as the reference's abstract class (include/frecsys/recommender.h:40-130):
// Score / EvaluateDataset / Train / SetPrint*Stats.  The reference implements the stages with
// Eigen + std::thread inside each subclass; here every subclass forwards to the C ABI of the CUDA
// library (include/frecsys_b200.h) through DeviceRecommender, which owns the frx_model and caches
// the device-resident form of each Dataset it is given (the reference passes the Dataset on every
// call and keeps no reference, recommender.h:55 — we key the cache on the object's address and
// tuple count).  There is no CPU fallback: if the library or a GPU is missing, construction throws.
#pragma once
#include <algorithm>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "frecsys/dataset.h"
#include "frecsys/evaluation.h"
#include "frecsys/logging.h"
#include "frecsys/types.h"
#include "frecsys_b200.h"

namespace frecsys {

class Recommender {
public:
  virtual ~Recommender() {}

  virtual VectorXf Score(const int user_id, const SpVector& user_history) { return VectorXf::Zero(1); }

  // Fold-in + ranking metrics (recommender.h:201-211 and the per-model overrides).
  virtual EvaluationResult EvaluateDataset(const VectorXi& k_list, const VectorXf& alpha_list, const Dataset& data,
                                           const SpMatrix& eval_by_user) = 0;

  virtual void Train(const Dataset& dataset) {}
  virtual void SetPrintTrainStats(const bool print_trainstats) {}
  virtual void SetPrintResidualStats(const bool print_residualstats) {}
  virtual void SetPrintVarStats(const bool print_varstats) {}
};

namespace detail {

inline void check(int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string("frecsys_b200: ") + what + ": " + frx_last_error());
}

// One process-wide GPU context (device from FRECSYS_DEVICE, default 0).
inline frx_context* default_context() {
  static frx_context* ctx = [] {
    frx_context* c = nullptr;
    const char* dev = std::getenv("FRECSYS_DEVICE");
    check(frx_context_create(dev ? std::atoi(dev) : 0, nullptr, &c), "frx_context_create");
    return c;
  }();
  return ctx;
}

class DeviceRecommender : public Recommender {
public:
  DeviceRecommender(const frx_config& cfg, int num_users, int num_items)
      : cfg_(cfg), num_users_(num_users), num_items_(num_items), item_embedding_(num_items, cfg.dim) {
    ctx_ = default_context();
    check(frx_model_create(ctx_, &cfg_, num_users, num_items, &model_), "frx_model_create");
    // The reference seeds its mt19937 from std::random_device (safer2.h:51-52); FRECSYS_INIT_SEED pins it.
    const char* seed = std::getenv("FRECSYS_INIT_SEED");
    unsigned s = seed ? (unsigned)std::strtoul(seed, nullptr, 10) : std::random_device{}();
    check(frx_model_init_factors(model_, s), "frx_model_init_factors");
  }
  ~DeviceRecommender() override {
    for (auto& kv : datasets_) frx_dataset_destroy(kv.second.handle);
    frx_model_destroy(model_);
  }
  DeviceRecommender(const DeviceRecommender&) = delete;
  DeviceRecommender& operator=(const DeviceRecommender&) = delete;

  VectorXf Score(const int user_id, const SpVector& user_history) override {
    throw("Function 'Score' is not implemented");  // as every reference model (safer2.h:79-82)
  }

  void Train(const Dataset& data) override {
    frx_dataset* ds = device_dataset(data);
    if (print_trainstats_ && !stats_after_train()) PrintLosses(ds);  // safer2.h:267 (before the update)
    check(frx_model_train(model_, ds), "frx_model_train");
    if (print_trainstats_ && stats_after_train()) PrintLosses(ds);   // ials.h:203 (after both steps)
    // the reference's Train() is synchronous: the epoch is complete (and a non-positive pivot has been
    // reported, safer2.h:160) when it returns, so `Timer: Train=` in run_model measures the whole epoch
    check(frx_context_sync(ctx_), "frx_model_train");
    after_train();
  }

  EvaluationResult EvaluateDataset(const VectorXi& k_list, const VectorXf& alpha_list, const Dataset& data,
                                   const SpMatrix& eval_by_user) override {
    frx_dataset* tr = device_dataset(data);
    // ground truth: tuples of eval_by_user (order is irrelevant for the metrics)
    std::vector<int> gu, gi;
    for (const auto& kv : eval_by_user)
      for (const auto& ir : kv.second) {
        gu.push_back(kv.first);
        gi.push_back(ir.first);
      }
    frx_dataset* te = nullptr;
    check(frx_dataset_create(ctx_, (int)gu.size(), gu.data(), gi.data(), &te), "frx_dataset_create(eval)");
    const int nk = k_list.size();
    int nu = frx_model_evaluate(model_, tr, te, k_list.data(), nk, nullptr, nullptr, nullptr, nullptr, nullptr);
    check(nu, "frx_model_evaluate");
    // the reference sizes the result by eval_by_user.size() and indexes rows by the position of the
    // user in `data` (recommender.h:88-90,115-117); rows here follow ascending user id of `data`.
    MatrixXf recall = MatrixXf::Zero(nu, nk), ndcg = MatrixXf::Zero(nu, nk);
    std::vector<int> ids(nu);
    int rc = frx_model_evaluate(model_, tr, te, k_list.data(), nk, ids.data(), recall.data(), ndcg.data(), nullptr, nullptr);
    frx_dataset_destroy(te);
    check(rc, "frx_model_evaluate");
    EvaluationResult result = {k_list, alpha_list, recall, ndcg};
    return result;
  }

  void SetPrintTrainStats(const bool v) override { print_trainstats_ = v; }
  void SetPrintResidualStats(const bool v) override {
    print_residualstats_ = v;
    check(frx_model_set_residual_stats(model_, v ? 1 : 0), "frx_model_set_residual_stats");
  }
  void SetPrintVarStats(const bool v) override { print_varstats_ = v; }

  // ials.h:410-412 etc.: the only factor accessor of the reference.
  const MatrixXf& item_embedding() const {
    check(frx_model_get_factors(model_, nullptr, const_cast<float*>(item_embedding_.data())), "frx_model_get_factors");
    return item_embedding_;
  }
  // Parity-harness hooks (the reference has none).
  void SetFactors(const MatrixXf& U, const MatrixXf& V) { check(frx_model_set_factors(model_, U.data(), V.data()), "set_factors"); }
  void GetFactors(MatrixXf* U, MatrixXf* V) const {
    check(frx_model_get_factors(model_, U ? U->data() : nullptr, V ? V->data() : nullptr), "get_factors");
  }
  // Checkpoint / resume (the reference has none): see frx_model_save / frx_model_load.
  void SaveCheckpoint(const std::string& path) const { check(frx_model_save(model_, path.c_str()), "frx_model_save"); }
  void LoadCheckpoint(const std::string& path) { check(frx_model_load(model_, path.c_str()), "frx_model_load"); }
  frx_model* handle() const { return model_; }

protected:
  struct Scalars { float xi, weighted_loss, mean_weight; };
  Scalars scalars() const {
    float sc[3];
    check(frx_model_get_state(model_, nullptr, nullptr, nullptr, nullptr, sc, nullptr), "frx_model_get_state");
    return {sc[0], sc[1], sc[2]};
  }
  void initialize_on_device(const Dataset& data) { check(frx_model_initialize(model_, device_dataset(data)), "frx_model_initialize"); }
  // "VaR: .. CVaR: .." and "Min: .. Mean: .. Max: .." of --print_var_stats (safer2.h:303-319, safer2pp.h),
  // restated on the host from the per-user loss and dual weights of the finished epoch, with the reference's
  // types: Q = size * alpha as a float, nth_element of -loss at (ptrdiff_t)Q, float running sum over i <= Q,
  // divided by Q.  (With pd_iterations > 1 the reference prints inside every primal-dual iteration; here the
  // epoch is one call, so the line is printed once, for the last iteration.)
  void PrintVarStats(float alpha) const {
    std::vector<float> z(num_users_), loss(num_users_);
    check(frx_model_get_state(model_, z.data(), loss.data(), nullptr, nullptr, nullptr, nullptr), "frx_model_get_state");
    std::vector<float> vals;
    vals.reserve(loss.size());
    for (size_t i = 0; i < loss.size(); i++) vals.push_back(-loss[i]);
    if (vals.empty()) return;
    auto const Q = vals.size() * alpha;
    std::nth_element(vals.begin(), vals.begin() + (std::ptrdiff_t)Q, vals.end());
    float acc = 0;
    for (int i = 0; i <= Q && i < (int)vals.size(); i++) acc += -vals[i];
    LOG(INFO) << "VaR: " << -vals[(int)Q] << " CVaR: " << acc / Q;
    float mn = z[0], mx = z[0];
    float sum = 0;  // Eigen's mean() of a VectorXf: float accumulation
    for (float v : z) { mn = std::min(mn, v); mx = std::max(mx, v); sum += v; }
    char buf[128];
    std::snprintf(buf, sizeof buf, "Min: %.3f, Mean: %.3f, Max: %.3f", mn, sum / (float)z.size(), mx);
    LOG(INFO) << buf;
  }
  // "U residual: .., V residual: ..[, z residual: ..]" of --print_residual_stats (safer2.h:323-328, ials.h:220-223,
  // ialspp.h:257-260, erm_mf.h:297-300, cvar_mf.h:322-326, safer2pp.h:346-350): norms of the change of U, V and
  // the dual weights, computed on the device; one line per primal-dual iteration.
  void PrintResidualStats(bool with_z) const {
    float r[3 * 64];
    const int n = frx_model_get_residuals(model_, r, 64);
    check(n, "frx_model_get_residuals");
    for (int t = 0; t < n; ++t) {
      char buf[160];
      if (with_z) std::snprintf(buf, sizeof buf, "U residual: %.9g, V residual: %.9g, z residual: %.9g", r[3 * t], r[3 * t + 1], r[3 * t + 2]);
      else std::snprintf(buf, sizeof buf, "U residual: %.9g, V residual: %.9g", r[3 * t], r[3 * t + 1]);
      LOG(INFO) << buf;
    }
  }
  virtual bool stats_after_train() const { return false; }
  virtual void after_train() { if (print_residualstats_) PrintResidualStats(false); }  // ials.h:220-223, ialspp.h:257-260

  // "Loss=... Loss_observed=..." and "Time=" lines (safer2.h:405-412, ials.h:297-304).
  void PrintLosses(frx_dataset* ds) {
    auto t0 = std::chrono::steady_clock::now();
    double s[6];
    check(frx_model_compute_stats(model_, ds, s), "frx_model_compute_stats");
    auto t1 = std::chrono::steady_clock::now();
    if (s[0] != s[0]) {  // NaN (safer2.h:399-404)
      LOG(ERROR) << "!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!";
      LOG(ERROR) << "NaN is detected!!";
      LOG(ERROR) << "!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!!";
      std::exit(0);
    }
    char buf[256];
    std::snprintf(buf, sizeof buf,
                  "Loss=%.2f Loss_observed=%.2f Loss_unobserved=%.2f Loss_reg=%.2f Loss_reg (user)=%.2f Loss_reg (item)=%.2f",
                  s[0], s[1], s[2], s[3], s[4], s[5]);
    LOG(INFO) << buf;
    LOG(INFO) << "Time=" << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
  }

  frx_dataset* device_dataset(const Dataset& data) {
    auto it = datasets_.find(&data);
    if (it != datasets_.end() && it->second.num_tuples == data.num_tuples()) return it->second.handle;
    if (it != datasets_.end()) {
      frx_dataset_destroy(it->second.handle);
      datasets_.erase(it);
    }
    frx_dataset* h = nullptr;
    check(frx_dataset_create(ctx_, data.num_tuples(), data.users().data(), data.items().data(), &h), "frx_dataset_create");
    datasets_[&data] = {h, data.num_tuples()};
    return h;
  }

  frx_config cfg_;
  int num_users_, num_items_;
  frx_context* ctx_ = nullptr;
  frx_model* model_ = nullptr;
  bool print_trainstats_ = false, print_residualstats_ = false, print_varstats_ = false;  // B-12: defined defaults
  mutable MatrixXf item_embedding_;

private:
  struct Cached { frx_dataset* handle; int num_tuples; };
  std::unordered_map<const Dataset*, Cached> datasets_;
};

inline frx_config base_config(int model, int dim, float reg, float uobs_weight, float stdev, float alpha) {
  frx_config c{};
  c.model = model; c.dim = dim; c.reg = reg; c.reg_exp = 1.0f; c.uobs_weight = uobs_weight; c.stdev = stdev;
  c.alpha = alpha; c.bandwidth = 1.0f; c.stepsize = 0.1f; c.xi_iterations = 5; c.pd_iterations = 1;
  c.use_epanechnikov = 0; c.use_snr = 0; c.sampling_ratio = 0.1f; c.use_cg = 0; c.cg_tol = 1e-10f; c.cg_max_it = 100;
  c.block_size = 64;
  c.snr_seed = std::random_device{}();  // the reference reseeds from random_device (safer2.h:728)
  if (const char* s = std::getenv("FRECSYS_SNR_SEED")) c.snr_seed = (unsigned)std::strtoul(s, nullptr, 10);
  return c;
}

}  // namespace detail
}  // namespace frecsys

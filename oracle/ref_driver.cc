// ref_driver — runs the REFERENCE's own, unmodified headers (/root/reference/include/frecsys/*.h)
// against the API shim in oracle/eigen_shim (Eigen / glog are absent from the image) and dumps the
// trained state, so that the oracle restatement can be checked against the reference's control flow.
// TEST INFRASTRUCTURE ONLY; built by `make -C oracle _ref` into oracle/_ref/ and only in a container
// where /root/reference exists.  The golden files it produces are committed under tests/golden/.
//
//   ref_driver <model> <train.csv> <test_tr.csv> <test_te.csv> <out.bin> key=value ...
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>

#include <fmt/core.h>
#include <glog/logging.h>
#include "Eigen/Core"
#include "Eigen/Dense"
#include <Eigen/IterativeLinearSolvers>

// The reference keeps its factors private and has no setter; the harness has to
// inject identical initial factors into the reference, the oracle and the CUDA path.
#define private public
#include "frecsys/cvar_mf.h"
#include "frecsys/erm_mf.h"
#include "frecsys/ials.h"
#include "frecsys/ialspp.h"
#include "frecsys/safer2.h"
#include "frecsys/safer2pp.h"
#undef private

namespace oracle {
void InitFactorsRaw(float* U, size_t nU, float* V, size_t nV, int dim, float stdev, unsigned seed);
}

static std::map<std::string, std::string> kv;
static float F(const char* k, float d) { return kv.count(k) ? (float)std::atof(kv[k].c_str()) : d; }
static int I(const char* k, int d) { return kv.count(k) ? std::atoi(kv[k].c_str()) : d; }

template <class M>
static void inject(M* m, unsigned seed, float stdev, bool has_gramian) {
  oracle::InitFactorsRaw(m->user_embedding_.data(), m->user_embedding_.size(), m->item_embedding_.data(),
                         m->item_embedding_.size(), (int)m->user_embedding_.cols(), stdev, seed);
  (void)has_gramian;
}

static void write_vec(std::ofstream& o, const float* p, size_t n) {
  uint64_t nn = n;
  o.write((const char*)&nn, 8);
  o.write((const char*)p, 4 * n);
}

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: ref_driver model train tr te out.bin k=v...\n"); return 2; }
  const std::string model = argv[1];
  for (int i = 6; i < argc; ++i) {
    std::string a = argv[i];
    size_t e = a.find('=');
    if (e != std::string::npos) kv[a.substr(0, e)] = a.substr(e + 1);
  }
  google::ShimQuiet() = I("quiet", 1) != 0;
  frecsys::Dataset train(argv[2]), test_tr(argv[3]), test_te(argv[4]);
  const int nu = train.max_user() + 1, ni = train.max_item() + 1;
  const int dim = I("dim", 8), epochs = I("epochs", 1);
  const unsigned seed = (unsigned)I("init_seed", 1);
  const float reg = F("reg", 0.002f), uw = F("uobs_weight", 0.1f), stdev = F("stdev", 0.1f), alpha = F("alpha", 0.3f);
  const float reg_exp = F("reg_exp", 1.0f), bw = F("bandwidth", 1.0f), step = F("stepsize", 0.1f);
  const int xi_it = I("xi_iterations", 5), pd_it = I("pd_iterations", 1), epan = I("use_epanechnikov", 0);
  const int block = I("block_size", 64), use_cg = I("use_cg", 0);

  frecsys::Recommender* rec = nullptr;
  std::vector<float> mean_weights;
  const frecsys::MatrixXf* Up = nullptr; const frecsys::MatrixXf* Vp = nullptr;
  const frecsys::VectorXf* zp = nullptr; const frecsys::VectorXf* lp = nullptr;
  float* xip = nullptr;
  auto train_loop = [&](auto* m, bool has_mean_weight) {
    for (int e = 0; e < epochs; ++e) {
      m->Train(train);
      if constexpr (requires { m->GetMeanWeight(); }) { if (has_mean_weight) mean_weights.push_back(m->GetMeanWeight()); }
    }
  };
  if (model == "ials") {
    auto* m = new frecsys::IALSRecommender(dim, nu, ni, reg, reg_exp, uw, stdev, alpha, use_cg, 1e-10f, 100);
    m->SetPrintTrainStats(false); m->SetPrintResidualStats(false); m->SetPrintVarStats(false);
    inject(m, seed, stdev, false);
    train_loop(m, false);
    rec = m; Up = &m->user_embedding_; Vp = &m->item_embedding_; lp = &m->user_loss_;
  } else if (model == "ialspp") {
    auto* m = new frecsys::IALSppRecommender(dim, nu, ni, reg, reg_exp, uw, stdev, alpha, block);
    m->SetPrintTrainStats(false); m->SetPrintResidualStats(false); m->SetPrintVarStats(false);
    inject(m, seed, stdev, false);
    train_loop(m, false);
    rec = m; Up = &m->user_embedding_; Vp = &m->item_embedding_;
  } else if (model == "erm_mf") {
    auto* m = new frecsys::ERMMFRecommender(dim, nu, ni, reg, uw, stdev, alpha, use_cg, 1e-10f, 100);
    m->SetPrintTrainStats(false); m->SetPrintResidualStats(false); m->SetPrintVarStats(false);
    inject(m, seed, stdev, true);
    m->item_gramian_ = m->item_embedding_.transpose() * m->item_embedding_;
    m->Initialize(train);
    train_loop(m, false);
    rec = m; Up = &m->user_embedding_; Vp = &m->item_embedding_; zp = &m->dual_weight_; lp = &m->user_loss_;
  } else if (model == "cvar_mf") {
    auto* m = new frecsys::CVaRMFRecommender(dim, nu, ni, reg, uw, alpha, step, stdev);
    m->SetPrintTrainStats(false); m->SetPrintResidualStats(false); m->SetPrintVarStats(false);
    inject(m, seed, stdev, true);
    m->item_gramian_ = m->item_embedding_.transpose() * m->item_embedding_;
    m->Initialize(train);
    train_loop(m, false);
    rec = m; Up = &m->user_embedding_; Vp = &m->item_embedding_; zp = &m->dual_weight_; lp = &m->user_loss_; xip = &m->prev_xi_;
  } else if (model == "safer2") {
    auto* m = new frecsys::SAFER2Recommender(dim, nu, ni, reg, uw, bw, alpha, stdev, xi_it, pd_it, epan, false, 0.1f, use_cg, 1e-10f, 100);
    m->SetPrintTrainStats(false); m->SetPrintResidualStats(false); m->SetPrintVarStats(false);
    inject(m, seed, stdev, true);
    m->item_gramian_ = m->item_embedding_.transpose() * m->item_embedding_;
    m->Initialize(train);
    train_loop(m, true);
    rec = m; Up = &m->user_embedding_; Vp = &m->item_embedding_; zp = &m->dual_weight_; lp = &m->user_loss_; xip = &m->prev_xi_;
  } else if (model == "safer2pp") {
    auto* m = new frecsys::SAFER2ppRecommender(dim, nu, ni, reg, uw, bw, alpha, stdev, xi_it, pd_it, epan, false, 0.1f, block);
    m->SetPrintTrainStats(false); m->SetPrintResidualStats(false); m->SetPrintVarStats(false);
    inject(m, seed, stdev, true);
    m->item_gramian_ = m->item_embedding_.transpose() * m->item_embedding_;
    m->Initialize(train);
    train_loop(m, true);
    rec = m; Up = &m->user_embedding_; Vp = &m->item_embedding_; zp = &m->dual_weight_; lp = &m->user_loss_; xip = &m->prev_xi_;
  } else {
    std::fprintf(stderr, "unknown model %s\n", model.c_str());
    return 2;
  }
  // evaluation exactly as tools/run_model.cc:30-41
  Eigen::VectorXi k_list = Eigen::VectorXi::Zero(5);
  Eigen::VectorXf alpha_list = Eigen::VectorXf::Zero(9);
  k_list << 5, 10, 20, 50, 100;
  alpha_list << 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9;
  frecsys::EvaluationResult metrics = rec->EvaluateDataset(k_list, alpha_list, test_tr, test_te.by_user());
  frecsys::VectorXf rec_mean = metrics.recall.colwise().mean(), ndcg_mean = metrics.ndcg.colwise().mean();
  frecsys::VectorXf ndcg20_cvar = metrics.cvar(metrics.ndcg.transpose().row(2));

  std::ofstream o(argv[5], std::ios::binary);
  write_vec(o, Up->data(), Up->size());
  write_vec(o, Vp->data(), Vp->size());
  if (zp) write_vec(o, zp->data(), zp->size()); else write_vec(o, nullptr, 0);
  if (lp) write_vec(o, lp->data(), lp->size()); else write_vec(o, nullptr, 0);
  float xi = xip ? *xip : 0.f;
  write_vec(o, &xi, 1);
  write_vec(o, mean_weights.data(), mean_weights.size());
  write_vec(o, rec_mean.data(), rec_mean.size());
  write_vec(o, ndcg_mean.data(), ndcg_mean.size());
  write_vec(o, ndcg20_cvar.data(), ndcg20_cvar.size());
  std::printf("%s dim=%d epochs=%d: ndcg@20=%.6f rec@20=%.6f xi=%.6f\n", model.c_str(), dim, epochs, ndcg_mean[2], rec_mean[2], xi);
  return 0;
}

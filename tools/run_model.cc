// run_model — the reference's experimentation CLI (tools/run_model.cc:125-274) on the CUDA path:
// same flags and defaults (run_model.cc:128-230), same model construction (run_model.cc:43-123), same
// epoch loop and log lines.  Bool flags take a value (`--use_snr 1`), as with the reference's CLI11 setup.
// Extra (not in the reference): --init_seed N, --snr_seed N pin the RNG seeds; --device N selects the GPU.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>

#include "frecsys/cvar_mf.h"
#include "frecsys/erm_mf.h"
#include "frecsys/ials.h"
#include "frecsys/ialspp.h"
#include "frecsys/safer2.h"
#include "frecsys/safer2pp.h"

namespace {

struct Flags {
  std::map<std::string, std::string> v = {
      {"print_evaluation_stats", "0"}, {"dim", "8"}, {"uobs_weight", "0.1"}, {"l2_reg", "0.002"},
      {"l2_reg_exp", "1.0"}, {"stdev", "0.1"}, {"print_train_stats", "1"}, {"print_test_results", "0"},
      {"print_residual_stats", "0"}, {"print_var_stats", "0"}, {"cg_error_tolerance", "1e-10"},
      {"cg_max_iterations", "100"}, {"use_cg", "0"}, {"block_size", "64"}, {"alpha", "0.3"},
      {"bandwidth", "1.0"}, {"stepsize", "0.1"}, {"xi_iterations", "5"}, {"sampling_ratio", "0.1"},
      {"pd_iterations", "1"}, {"use_epanechnikov", "0"}, {"use_snr", "0"}, {"epoch", "50"},
      {"model_name", ""}, {"train_data", ""}, {"test_train_data", ""}, {"test_test_data", ""},
      {"init_seed", ""}, {"snr_seed", ""}, {"device", ""}};
  std::map<std::string, std::string> shorts = {{"d", "dim"}, {"r", "l2_reg"}, {"s", "stdev"}, {"e", "epoch"}, {"n", "model_name"}};

  bool parse(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i], key, val;
      if (a.rfind("--", 0) == 0) key = a.substr(2);
      else if (a.rfind("-", 0) == 0 && shorts.count(a.substr(1))) key = shorts[a.substr(1)];
      else { std::fprintf(stderr, "unexpected argument %s\n", a.c_str()); return false; }
      size_t eq = key.find('=');
      if (eq != std::string::npos) { val = key.substr(eq + 1); key = key.substr(0, eq); }
      else if (i + 1 < argc) val = argv[++i];
      else { std::fprintf(stderr, "--%s needs a value\n", key.c_str()); return false; }
      if (!v.count(key)) { std::fprintf(stderr, "unknown option --%s\n", key.c_str()); return false; }
      v[key] = val;
    }
    for (const char* req : {"model_name", "train_data", "test_train_data", "test_test_data"})
      if (v[req].empty()) { std::fprintf(stderr, "--%s is required\n", req); return false; }
    for (const char* f : {"train_data", "test_train_data", "test_test_data"})
      if (!std::ifstream(v[f]).good()) { std::fprintf(stderr, "--%s: File does not exist: %s\n", f, v[f].c_str()); return false; }
    return true;
  }
  int i(const char* k) { return std::atoi(v[k].c_str()); }
  float f(const char* k) { return (float)std::atof(v[k].c_str()); }
  bool b(const char* k) {
    std::string s = v[k];
    std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    return s == "1" || s == "true" || s == "on" || s == "yes";
  }
};

void evaluate(int epoch, frecsys::Recommender* recommender, frecsys::Dataset& exclude, frecsys::Dataset& test) {
  frecsys::VectorXi k_list = {5, 10, 20, 50, 100};  // run_model.cc:33-36
  frecsys::VectorXf alpha_list = {0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f, 0.9f};
  frecsys::EvaluationResult metrics = recommender->EvaluateDataset(k_list, alpha_list, exclude, test.by_user());
  LOG(INFO) << "Epoch " << epoch << ":";
  metrics.show();
}

frecsys::Recommender* get_model(const std::string model_name, const int num_users, const int num_items, Flags& fl) {
  frecsys::Recommender* recommender = nullptr;  // run_model.cc:43-123
  if (model_name == "ials") {
    recommender = new frecsys::IALSRecommender(fl.i("dim"), num_users, num_items, fl.f("l2_reg"), fl.f("l2_reg_exp"),
                                               fl.f("uobs_weight"), fl.f("stdev"), fl.f("alpha"), fl.b("use_cg"),
                                               fl.f("cg_error_tolerance"), fl.i("cg_max_iterations"));
  } else if (model_name == "ialspp") {
    recommender = new frecsys::IALSppRecommender(fl.i("dim"), num_users, num_items, fl.f("l2_reg"), fl.f("l2_reg_exp"),
                                                 fl.f("uobs_weight"), fl.f("stdev"), fl.f("alpha"), fl.i("block_size"));
  } else if (model_name == "safer2") {
    recommender = new frecsys::SAFER2Recommender(
        fl.i("dim"), num_users, num_items, fl.f("l2_reg"), fl.f("uobs_weight"), fl.f("bandwidth"), fl.f("alpha"),
        fl.f("stdev"), fl.i("xi_iterations"), fl.i("pd_iterations"), fl.b("use_epanechnikov"), fl.b("use_snr"),
        fl.f("sampling_ratio"), fl.b("use_cg"), fl.f("cg_error_tolerance"), fl.i("cg_max_iterations"));
  } else if (model_name == "safer2pp") {
    recommender = new frecsys::SAFER2ppRecommender(
        fl.i("dim"), num_users, num_items, fl.f("l2_reg"), fl.f("uobs_weight"), fl.f("bandwidth"), fl.f("alpha"),
        fl.f("stdev"), fl.i("xi_iterations"), fl.i("pd_iterations"), fl.b("use_epanechnikov"), fl.b("use_snr"),
        fl.f("sampling_ratio"), fl.i("block_size"));
  } else if (model_name == "erm_mf") {
    recommender = new frecsys::ERMMFRecommender(fl.i("dim"), num_users, num_items, fl.f("l2_reg"), fl.f("uobs_weight"),
                                                fl.f("stdev"), fl.f("alpha"), fl.b("use_cg"), fl.f("cg_error_tolerance"),
                                                fl.i("cg_max_iterations"));
  } else if (model_name == "cvar_mf") {
    recommender = new frecsys::CVaRMFRecommender(fl.i("dim"), num_users, num_items, fl.f("l2_reg"), fl.f("uobs_weight"),
                                                 fl.f("alpha"), fl.f("stepsize"), fl.f("stdev"));
  } else {
    std::fprintf(stderr, "--model_name: %s not in [ials, ialspp, safer2, safer2pp, cvar_mf, erm_mf]\n", model_name.c_str());
    std::exit(105);
  }
  recommender->SetPrintResidualStats(fl.b("print_residual_stats"));
  recommender->SetPrintVarStats(fl.b("print_var_stats"));
  recommender->SetPrintTrainStats(fl.b("print_train_stats"));
  return recommender;
}

}  // namespace

int main(int argc, char* argv[]) {
  Flags fl;
  if (!fl.parse(argc, argv)) return 105;
  if (!fl.v["init_seed"].empty()) setenv("FRECSYS_INIT_SEED", fl.v["init_seed"].c_str(), 1);
  if (!fl.v["snr_seed"].empty()) setenv("FRECSYS_SNR_SEED", fl.v["snr_seed"].c_str(), 1);
  if (!fl.v["device"].empty()) setenv("FRECSYS_DEVICE", fl.v["device"].c_str(), 1);
  std::string model_name = fl.v["model_name"];
  std::transform(model_name.begin(), model_name.end(), model_name.begin(), ::tolower);  // CLI::ignore_case
  const int epochs = fl.i("epoch");

  frecsys::Dataset train(fl.v["train_data"]);
  frecsys::Dataset test_tr(fl.v["test_train_data"]);
  frecsys::Dataset test_te(fl.v["test_test_data"]);

  frecsys::Recommender* recommender = get_model(model_name, train.max_user() + 1, train.max_item() + 1, fl);
  setbuf(stdout, NULL);

  // run_model.cc:246-257
  if (model_name == "cvar_mf") ((frecsys::CVaRMFRecommender*)recommender)->Initialize(train);
  if (model_name == "safer2") ((frecsys::SAFER2Recommender*)recommender)->Initialize(train);
  if (model_name == "safer2pp") ((frecsys::SAFER2ppRecommender*)recommender)->Initialize(train);
  if (model_name == "erm_mf") ((frecsys::ERMMFRecommender*)recommender)->Initialize(train);
  for (int epoch = 0; epoch < epochs; ++epoch) {
    auto time_train_start = std::chrono::steady_clock::now();
    recommender->Train(train);
    auto time_train_end = std::chrono::steady_clock::now();
    uint64_t train_time = std::chrono::duration_cast<std::chrono::milliseconds>(time_train_end - time_train_start).count();
    LOG(INFO) << "Epoch: " << epoch << ", Timer: Train=" << train_time;
    if (fl.b("print_evaluation_stats")) evaluate(epoch, recommender, test_tr, test_te);
  }
  LOG(INFO) << "Validation Results";
  evaluate(epochs, recommender, test_tr, test_te);
  return 0;
}

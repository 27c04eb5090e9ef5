// frecsys::CVaRMFRecommender — reference: include/frecsys/cvar_mf.h:34-64 (exact-quantile CVaR, gradient steps).
#pragma once
#include "frecsys/recommender.h"

namespace frecsys {

class CVaRMFRecommender : public detail::DeviceRecommender {
public:
  CVaRMFRecommender(int embedding_dim, int num_users, int num_items, float reg, float unobserved_weight, float alpha,
                    float stepsize, float stdev)
      : DeviceRecommender(make(embedding_dim, reg, unobserved_weight, alpha, stepsize, stdev), num_users, num_items) {}

  void Initialize(const Dataset& data) { initialize_on_device(data); }  // cvar_mf.h:710-726

protected:
  void after_train() override {
    const Scalars s = scalars();
    LOG(INFO) << "Weighted Loss: " << s.weighted_loss;  // cvar_mf.h:301-302
    LOG(INFO) << "Mean weights: " << s.mean_weight;     // cvar_mf.h:303
    LOG(INFO) << "Exact Quantile:" << s.xi;             // cvar_mf.h:592
    if (print_residualstats_) PrintResidualStats(true);  // cvar_mf.h:322-326
    LOG(INFO) << "Xi:" << s.xi;                         // cvar_mf.h:327
  }

private:
  static frx_config make(int dim, float reg, float uw, float alpha, float stepsize, float stdev) {
    frx_config c = detail::base_config(FRX_CVAR_MF, dim, reg, uw, stdev, alpha);
    c.stepsize = stepsize;
    return c;
  }
};

}  // namespace frecsys

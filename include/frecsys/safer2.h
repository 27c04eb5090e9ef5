// frecsys::SAFER2Recommender — reference: include/frecsys/safer2.h:35-77.  Same constructor argument
// order, Train / Initialize / EvaluateDataset / GetMeanWeight; the epoch (safer2.h:266-334) runs as
// CUDA stages sequenced by frx_model_train.
#pragma once
#include "frecsys/recommender.h"

namespace frecsys {

class SAFER2Recommender : public detail::DeviceRecommender {
public:
  SAFER2Recommender(int embedding_dim, int num_users, int num_items, float reg, float unobserved_weight,
                    float bandwidth, float alpha, float stdev, int xi_iterations, int pd_iterations,
                    bool use_epanechnikov, bool use_snr, float sampling_ratio, bool use_cg, float cg_error_tolerance,
                    int cg_max_iterations)
      : DeviceRecommender(make(FRX_SAFER2, embedding_dim, reg, unobserved_weight, bandwidth, alpha, stdev, xi_iterations,
                               pd_iterations, use_epanechnikov, use_snr, sampling_ratio, use_cg, cg_error_tolerance,
                               cg_max_iterations, 64),
                          num_users, num_items) {}

  // safer2.h:819-838
  void Initialize(const Dataset& data) {
    initialize_on_device(data);
    LOG(INFO) << "Initial Xi:" << scalars().xi;
  }
  // safer2.h:815-817
  float GetMeanWeight() const { return scalars().mean_weight; }

  static frx_config make(int model, int dim, float reg, float uw, float bandwidth, float alpha, float stdev, int xi_it,
                         int pd_it, bool epan, bool snr, float ratio, bool use_cg, float tol, int max_it, int block) {
    frx_config c = detail::base_config(model, dim, reg, uw, stdev, alpha);
    c.bandwidth = bandwidth; c.xi_iterations = xi_it; c.pd_iterations = pd_it; c.use_epanechnikov = epan; c.use_snr = snr;
    c.sampling_ratio = ratio; c.use_cg = use_cg; c.cg_tol = tol; c.cg_max_it = max_it; c.block_size = block;
    return c;
  }

protected:
  void after_train() override {
    const Scalars s = scalars();
    LOG(INFO) << "Weighted Loss: " << s.weighted_loss;  // safer2.h:300-301 (last primal-dual iteration)
    if (print_varstats_) PrintVarStats(cfg_.alpha);  // safer2.h:303-319
    if (print_residualstats_) PrintResidualStats(true);  // safer2.h:323-328
    LOG(INFO) << "Xi:" << s.xi;                         // safer2.h:332
  }
};

}  // namespace frecsys

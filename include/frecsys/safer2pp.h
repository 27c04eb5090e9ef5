// frecsys::SAFER2ppRecommender — reference: include/frecsys/safer2pp.h:34-76 (SAFER2 with the iALS++
// block-subspace solver).
#pragma once
#include "frecsys/safer2.h"

namespace frecsys {

class SAFER2ppRecommender : public detail::DeviceRecommender {
public:
  SAFER2ppRecommender(int embedding_dim, int num_users, int num_items, float reg, float unobserved_weight,
                      float bandwidth, float alpha, float stdev, int xi_iterations, int pd_iterations,
                      bool use_epanechnikov, bool use_snr, float sampling_ratio, int block_size)
      : DeviceRecommender(SAFER2Recommender::make(FRX_SAFER2PP, embedding_dim, reg, unobserved_weight, bandwidth, alpha,
                                                  stdev, xi_iterations, pd_iterations, use_epanechnikov, use_snr,
                                                  sampling_ratio, false, 1e-10f, 100, block_size),
                          num_users, num_items) {}

  void Initialize(const Dataset& data) {
    initialize_on_device(data);
    LOG(INFO) << "Initial Xi:" << scalars().xi;
  }
  float GetMeanWeight() const { return scalars().mean_weight; }

protected:
  void after_train() override {
    const Scalars s = scalars();
    LOG(INFO) << "Weighted Loss: " << s.weighted_loss;
    if (print_varstats_) PrintVarStats(cfg_.alpha);  // safer2pp.h
    if (print_residualstats_) PrintResidualStats(true);  // safer2pp.h:346-350
    LOG(INFO) << "Xi:" << s.xi;
  }
};

}  // namespace frecsys

// frecsys::ERMMFRecommender — reference: include/frecsys/erm_mf.h:35-70 (SAFER2's loss with z == alpha).
#pragma once
#include "frecsys/recommender.h"

namespace frecsys {

class ERMMFRecommender : public detail::DeviceRecommender {
public:
  ERMMFRecommender(int embedding_dim, int num_users, int num_items, float reg, float unobserved_weight, float stdev,
                   float alpha, bool use_cg, float cg_error_tolerance, int cg_max_iterations)
      : DeviceRecommender(make(embedding_dim, reg, unobserved_weight, stdev, alpha, use_cg, cg_error_tolerance,
                               cg_max_iterations),
                          num_users, num_items) {}

  void Initialize(const Dataset& data) { initialize_on_device(data); }  // erm_mf.h:573-587

protected:
  void after_train() override {
    LOG(INFO) << "Weighted Loss: " << scalars().weighted_loss;  // erm_mf.h:277-278
    if (print_residualstats_) PrintResidualStats(false);       // erm_mf.h:297-300
  }

private:
  static frx_config make(int dim, float reg, float uw, float stdev, float alpha, bool use_cg, float tol, int max_it) {
    frx_config c = detail::base_config(FRX_ERM_MF, dim, reg, uw, stdev, alpha);
    c.use_cg = use_cg; c.cg_tol = tol; c.cg_max_it = max_it;
    return c;
  }
};

}  // namespace frecsys

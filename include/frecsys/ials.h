// frecsys::IALSRecommender — same constructor and methods as the reference (include/frecsys/ials.h:37-66),
// executed by the CUDA library.
#pragma once
#include "frecsys/recommender.h"

namespace frecsys {

class IALSRecommender : public detail::DeviceRecommender {
public:
  IALSRecommender(int embedding_dim, int num_users, int num_items, float reg, float reg_exp, float unobserved_weight,
                  float stdev, float alpha, bool use_cg, float cg_error_tolerance, int cg_max_iterations)
      : DeviceRecommender(make(embedding_dim, reg, reg_exp, unobserved_weight, stdev, alpha, use_cg, cg_error_tolerance,
                               cg_max_iterations),
                          num_users, num_items) {}

protected:
  bool stats_after_train() const override { return true; }  // ComputeLosses runs after both steps (ials.h:203)

private:
  static frx_config make(int dim, float reg, float reg_exp, float uw, float stdev, float alpha, bool use_cg, float tol,
                         int max_it) {
    frx_config c = detail::base_config(FRX_IALS, dim, reg, uw, stdev, alpha);
    c.reg_exp = reg_exp; c.use_cg = use_cg; c.cg_tol = tol; c.cg_max_it = max_it;
    return c;
  }
};

}  // namespace frecsys

// frecsys::IALSppRecommender — reference: include/frecsys/ialspp.h:37-61 (block-subspace iALS++).
#pragma once
#include "frecsys/recommender.h"

namespace frecsys {

class IALSppRecommender : public detail::DeviceRecommender {
public:
  IALSppRecommender(int embedding_dim, int num_users, int num_items, float reg, float reg_exp, float unobserved_weight,
                    float stdev, float alpha, int block_size)
      : DeviceRecommender(make(embedding_dim, reg, reg_exp, unobserved_weight, stdev, alpha, block_size), num_users,
                          num_items) {}

protected:
  bool stats_after_train() const override { return true; }  // ialspp.h:241

private:
  static frx_config make(int dim, float reg, float reg_exp, float uw, float stdev, float alpha, int block_size) {
    frx_config c = detail::base_config(FRX_IALSPP, dim, reg, uw, stdev, alpha);
    c.reg_exp = reg_exp; c.block_size = block_size;
    return c;
  }
};

}  // namespace frecsys

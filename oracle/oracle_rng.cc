// Oracle RNG pieces (TEST INFRASTRUCTURE ONLY), compiled with plain -O2 and
// -ffp-contract=off so std::normal_distribution<float> produces the same bits on
// every host regardless of -march.  Restates recommender.h:61-67.
#include <cmath>
#include <cstddef>
#include <random>

namespace oracle {
void InitFactorsRaw(float* U, size_t nU, float* V, size_t nV, int dim, float stdev, unsigned seed) {
  const float adjusted = stdev / std::sqrt((double)dim);  // float / double sqrt(int), safer2.h:50
  std::mt19937 gen{seed};
  {
    std::normal_distribution<float> d(0, adjusted);
    for (size_t i = 0; i < nU; ++i) U[i] = d(gen);
  }
  {
    std::normal_distribution<float> d(0, adjusted);
    for (size_t i = 0; i < nV; ++i) V[i] = d(gen);
  }
}
}  // namespace oracle

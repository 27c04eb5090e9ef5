// Evaluation result types — same fields and reporting as the reference
// (include/frecsys/evaluation.h:30-103): per-user Recall@k / NDCG@k matrices, column means and the
// lower-tail CVaR of each metric at the quantiles in alpha_list, logged in the reference's format.
#pragma once

#include <algorithm>
#include <cstdio>
#include <sstream>
#include <string>
#include <vector>

#include "frecsys/logging.h"
#include "frecsys/types.h"

namespace frecsys {

struct UserEvaluationResult {
  const VectorXf recall;
  const VectorXf ndcg;
};

struct EvaluationResult {
  const VectorXi k_list;
  const VectorXf alpha_list;
  const MatrixXf recall;
  const MatrixXf ndcg;

  // "{name}@{k}={value:.4f}" joined by spaces (evaluation.h:43-58).
  std::string format(std::string measure_name, VectorXf measurements) const {
    std::stringstream ss;
    for (int i = 0; i < k_list.size(); i++) {
      char buf[64];
      std::snprintf(buf, sizeof buf, "%s@%d=%.4f", measure_name.c_str(), k_list[i], measurements[i]);
      ss << buf;
      if (i != k_list.size() - 1) ss << " ";
    }
    return ss.str();
  }

  // evaluation.h:61-81
  void show() const {
    LOG(INFO) << format("Mean Rec", recall.colwise_mean());
    LOG(INFO) << format("Mean NDCG", ndcg.colwise_mean());
    std::vector<VectorXf> ndcg_cvar, rec_cvar;
    for (int i = 0; i < k_list.size(); i++) {
      ndcg_cvar.push_back(cvar(ndcg.col(i)));
      rec_cvar.push_back(cvar(recall.col(i)));
    }
    for (int a = 0; a < alpha_list.size(); a++) {
      VectorXf r(k_list.size()), n(k_list.size());
      for (int i = 0; i < k_list.size(); i++) {
        r[i] = rec_cvar[i][a];
        n[i] = ndcg_cvar[i][a];
      }
      char name[64];
      std::snprintf(name, sizeof name, "Rec CVaR (q=%.2f)", alpha_list[a]);
      LOG(INFO) << format(name, r);
      std::snprintf(name, sizeof name, "NDCG CVaR (q=%.2f)", alpha_list[a]);
      LOG(INFO) << format(name, n);
    }
  }

  // Lower-tail CVaR of one metric column at every alpha (evaluation.h:83-102).
  VectorXf cvar(VectorXf measurements) const {
    std::vector<float> ms(measurements.data(), measurements.data() + measurements.rows());
    std::sort(ms.begin(), ms.end());
    int counter = 0;
    VectorXf cvars = VectorXf::Zero(alpha_list.size());
    float accs = 0;
    for (size_t i = 0; i < ms.size(); i++) {
      accs += ms.at(i);
      for (int j = counter; j < alpha_list.size(); j++) {
        int pos = ms.size() * alpha_list[j];
        if (pos == (int)i) {
          cvars[counter] = accs / (i + 1);
          counter++;
        }
      }
    }
    return cvars;
  }
};

}  // namespace frecsys

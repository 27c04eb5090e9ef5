// frecsys::Dataset — same interface and semantics as the reference (include/frecsys/dataset.h:24-99):
// a `uid,sid` CSV with a header line becomes by_user_[uid] / by_item_[sid] lists of
// (other id, tuple index) in FILE order.  In addition the tuple list is kept flat so the CUDA
// library can build its device-resident CSR/CSC from it (frx_dataset_create).
#pragma once

#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "frecsys/logging.h"
#include "frecsys/types.h"

namespace frecsys {

class Dataset {
public:
  explicit Dataset(const std::string& filename);
  // Builds a Dataset from tuple arrays (synthetic data, tests).
  Dataset(const int* users, const int* items, int n);
  const SpMatrix& by_user() const { return by_user_; }
  const SpMatrix& by_item() const { return by_item_; }
  const int max_user() const { return max_user_; }
  const int max_item() const { return max_item_; }
  const int num_tuples() const { return num_tuples_; }
  // Flat tuple list in file order (users()[t], items()[t]).
  const std::vector<int>& users() const { return users_; }
  const std::vector<int>& items() const { return items_; }

private:
  void add(int user, int item) {  // dataset.h:86-91
    by_user_[user].push_back({item, num_tuples_});
    by_item_[item].push_back({user, num_tuples_});
    users_.push_back(user);
    items_.push_back(item);
    max_user_ = std::max(max_user_, user);
    max_item_ = std::max(max_item_, item);
    ++num_tuples_;
  }
  void log_summary() const {  // dataset.h:94-98
    LOG(INFO) << "max_user=" << max_user() << "\tmax_item=" << max_item() << "\tdistinct user=" << by_user_.size()
              << "\tdistinct item=" << by_item_.size() << "\tnum_tuples=" << num_tuples();
  }
  SpMatrix by_user_;
  SpMatrix by_item_;
  std::vector<int> users_, items_;
  int max_user_ = -1;
  int max_item_ = -1;
  int num_tuples_ = 0;
};

inline Dataset::Dataset(const std::string& filename) {
  std::ifstream infile(filename);
  std::string line;
  // Discard header (the reference asserts on it, dataset.h:80).
  if (!std::getline(infile, line)) throw std::runtime_error("frecsys::Dataset: cannot read " + filename);
  while (std::getline(infile, line)) {
    int pos = line.find(',');
    int user = std::atoi(line.substr(0, pos).c_str());
    int item = std::atoi(line.substr(pos + 1).c_str());
    add(user, item);
  }
  log_summary();
}

inline Dataset::Dataset(const int* users, const int* items, int n) {
  for (int t = 0; t < n; ++t) add(users[t], items[t]);
  log_summary();
}

}  // namespace frecsys

// Boundary value types of the frecsys interface (reference: include/frecsys/types.h:23-31).
// The reference aliases Eigen types; Eigen is not a dependency of this build, so these are small
// self-contained row-major value types exposing the members the Recommender / Dataset /
// EvaluationResult interface and its callers (tools/run_model.cc, tests/*_test.cc) use.
#pragma once

#include <algorithm>
#include <cassert>
#include <initializer_list>
#include <unordered_map>
#include <utility>
#include <vector>

namespace frecsys {

template <typename T>
class VectorT {
public:
  VectorT() {}
  explicit VectorT(int n) : v_(n) {}
  VectorT(std::initializer_list<T> l) : v_(l) {}
  static VectorT Zero(int n) { VectorT r(n); std::fill(r.v_.begin(), r.v_.end(), T(0)); return r; }
  static VectorT Ones(int n) { VectorT r(n); std::fill(r.v_.begin(), r.v_.end(), T(1)); return r; }
  int size() const { return (int)v_.size(); }
  int rows() const { return (int)v_.size(); }
  T& operator[](int i) { return v_[i]; }
  const T& operator[](int i) const { return v_[i]; }
  T& operator()(int i) { return v_[i]; }
  const T& operator()(int i) const { return v_[i]; }
  T* data() { return v_.data(); }
  const T* data() const { return v_.data(); }
  T maxCoeff() const { return *std::max_element(v_.begin(), v_.end()); }
  T minCoeff() const { return *std::min_element(v_.begin(), v_.end()); }
  T sum() const { T s = 0; for (const T& x : v_) s += x; return s; }
  T mean() const { return v_.empty() ? T(0) : (T)(sum() / (T)v_.size()); }
private:
  std::vector<T> v_;
};
typedef VectorT<float> VectorXf;
typedef VectorT<int> VectorXi;

// Row-major float matrix (types.h:25-27).
class MatrixXf {
public:
  MatrixXf() : rows_(0), cols_(0) {}
  MatrixXf(int r, int c) : rows_(r), cols_(c), v_((size_t)r * c) {}
  static MatrixXf Zero(int r, int c) { MatrixXf m(r, c); std::fill(m.v_.begin(), m.v_.end(), 0.f); return m; }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  size_t size() const { return v_.size(); }
  float& operator()(int i, int j) { return v_[(size_t)i * cols_ + j]; }
  float operator()(int i, int j) const { return v_[(size_t)i * cols_ + j]; }
  float* data() { return v_.data(); }
  const float* data() const { return v_.data(); }
  float* row(int i) { return v_.data() + (size_t)i * cols_; }
  const float* row(int i) const { return v_.data() + (size_t)i * cols_; }
  VectorXf col(int j) const { VectorXf r(rows_); for (int i = 0; i < rows_; ++i) r[i] = (*this)(i, j); return r; }
  // recall.colwise().mean() of the reference callers (evaluation.h:62-63)
  VectorXf colwise_mean() const {
    VectorXf r = VectorXf::Zero(cols_);
    for (int j = 0; j < cols_; ++j) {
      float s = 0.f;
      for (int i = 0; i < rows_; ++i) s += (*this)(i, j);
      r[j] = rows_ ? s / rows_ : 0.f;
    }
    return r;
  }
private:
  int rows_, cols_;
  std::vector<float> v_;
};

using SpVector = std::vector<std::pair<int, int>>;       // (other id, tuple index), types.h:30
using SpMatrix = std::unordered_map<int, SpVector>;      // types.h:31
}  // namespace frecsys

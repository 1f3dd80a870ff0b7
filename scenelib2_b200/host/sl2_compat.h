// sl2_compat.h — the host shim is written against Eigen3 / OpenCV types like the reference.
// When those libraries are installed (a real SceneLib2 build) they are used directly; when they
// are absent (this container) the minimal stand-ins below provide just the storage / indexing
// surface the shim touches, so that the shim and the headless driver still compile and run.
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#if defined(SL2_USE_REAL_EIGEN_OPENCV)
#include <Eigen/Eigen>
#include <opencv2/opencv.hpp>
#else
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace Eigen {
class MatrixXd {  // column-major like Eigen's default
 public:
  MatrixXd() : r_(0), c_(0) {}
  MatrixXd(int r, int c) : r_(r), c_(c), a_((size_t)r * c, 0.0) {}
  void resize(int r, int c) { r_ = r; c_ = c; a_.assign((size_t)r * c, 0.0); }
  void setZero() { std::fill(a_.begin(), a_.end(), 0.0); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double &operator()(int i, int j) { return a_[(size_t)i + (size_t)j * r_]; }
  double operator()(int i, int j) const { return a_[(size_t)i + (size_t)j * r_]; }
  double *data() { return a_.data(); }
  const double *data() const { return a_.data(); }
  double trace() const { double t = 0; for (int i = 0; i < r_ && i < c_; ++i) t += (*this)(i, i); return t; }
 private:
  int r_, c_;
  std::vector<double> a_;
};
class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(int n) : a_((size_t)n, 0.0) {}
  void resize(int n) { a_.assign((size_t)n, 0.0); }
  void setZero() { std::fill(a_.begin(), a_.end(), 0.0); }
  int size() const { return (int)a_.size(); }
  double &operator()(int i) { return a_[i]; }
  double operator()(int i) const { return a_[i]; }
  double &operator[](int i) { return a_[i]; }
  double operator[](int i) const { return a_[i]; }
  double *data() { return a_.data(); }
  const double *data() const { return a_.data(); }
 private:
  std::vector<double> a_;
};
struct Vector2d : VectorXd { Vector2d() : VectorXd(2) {} };
struct Vector3d : VectorXd { Vector3d() : VectorXd(3) {} };
struct Matrix2d : MatrixXd { Matrix2d() : MatrixXd(2, 2) {} };
}  // namespace Eigen
#define CV_8UC1 0
namespace cv {
struct Size { int width, height; };
class Mat {  // 8-bit single channel only
 public:
  unsigned char *data;
  int rows, cols;
  size_t step;
  Mat() : data(nullptr), rows(0), cols(0), step(0) {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), step((size_t)c) {
    own_.reset(new std::vector<unsigned char>((size_t)r * c, 0));
    data = own_->data();
  }
  Mat(int r, int c, int /*type*/, void *ext, size_t stp = 0)
      : data((unsigned char *)ext), rows(r), cols(c), step(stp ? stp : (size_t)c) {}
  Size size() const { return Size{cols, rows}; }
  bool empty() const { return data == nullptr; }
 private:
  std::shared_ptr<std::vector<unsigned char>> own_;
};
}  // namespace cv
#endif

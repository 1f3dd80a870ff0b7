// ORACLE — TEST INFRASTRUCTURE ONLY.  8-bit single-channel cv::Mat stand-in for the reference sources
// compiled into _ref/libsl2refmodels.so (monoslam.cpp, feature.cpp, improc/*.cpp): pixel storage with
// shared ownership, at<T>(row, col), size(), type(), clone(), a PGM reader for cv::imread(path, 0) and a
// no-op cv::imwrite.  No image arithmetic.
#ifndef SL2_ORACLE_OPENCV_ARITH_STUB
#define SL2_ORACLE_OPENCV_ARITH_STUB
#include <math.h>

#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_64FC1 6

namespace cv {
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
class Mat {
 public:
  unsigned char *data;
  int rows, cols;
  Mat() : data(nullptr), rows(0), cols(0), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type) { alloc(0.0); }
  Mat(int r, int c, int type, void *ext) : data((unsigned char *)ext), rows(r), cols(c), type_(type) {}
  Mat(Size s, int type, double fill) : rows(s.height), cols(s.width), type_(type) { alloc(fill); }
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  bool empty() const { return data == nullptr; }
  Mat clone() const {
    Mat m(rows, cols, type_);
    std::memcpy(m.data, data, (size_t)rows * cols * esz());
    return m;
  }
  template <class T>
  T &at(int r, int c) { return reinterpret_cast<T *>(data)[(size_t)r * cols + c]; }
  template <class T>
  const T &at(int r, int c) const { return reinterpret_cast<const T *>(data)[(size_t)r * cols + c]; }

 private:
  size_t esz() const { return type_ == CV_64FC1 ? 8 : 1; }
  void alloc(double fill) {
    own_.reset(new std::vector<unsigned char>((size_t)rows * cols * esz()));
    data = own_->data();
    if (type_ == CV_64FC1) {
      double *d = reinterpret_cast<double *>(data);
      for (size_t i = 0; i < (size_t)rows * cols; ++i) d[i] = fill;
    } else {
      std::memset(data, (int)fill, (size_t)rows * cols);
    }
  }
  int type_;
  std::shared_ptr<std::vector<unsigned char>> own_;
};

inline Mat imread(const std::string &path, int /*flags*/) {  // binary PGM (P5) only
  std::ifstream f(path.c_str(), std::ios::binary);
  std::string magic;
  if (!f || !(f >> magic) || magic != "P5") return Mat();
  int vals[3], got = 0;
  while (got < 3) {
    f >> std::ws;
    if (f.peek() == '#') {
      std::string c;
      std::getline(f, c);
      continue;
    }
    if (!(f >> vals[got++])) return Mat();
  }
  f.get();
  Mat m(vals[1], vals[0], CV_8UC1);
  f.read(reinterpret_cast<char *>(m.data), (std::streamsize)vals[0] * vals[1]);
  return m;
}
inline bool imwrite(const std::string &, const Mat &) { return true; }
}  // namespace cv
#endif

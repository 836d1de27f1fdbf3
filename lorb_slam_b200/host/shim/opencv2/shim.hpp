// Minimal stand-in for the handful of OpenCV types the hot-path operators touch.
// ONLY for building and testing the drop-in layer in an image without
// OpenCV-C++ (this one).  In the real LORB-SLAM tree the genuine <opencv2/...>
// headers are used and this directory is not on the include path
// (INTEGRATION.md).
#ifndef LORB_CV_SHIM_HPP
#define LORB_CV_SHIM_HPP
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
};
template <typename T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point_<float> Point2f;
typedef Point3_<float> Point3f;

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};

// Dense row-major matrix with shared storage; enough of cv::Mat for
// at<>/ptr<>/row/clone/push_back/rows/cols/empty.
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r;
    cols = c;
    type_ = type;
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * esz(), 0);
    off_ = 0;
    step = (size_t)c * esz();
  }
  int type() const { return type_; }
  // InputArray / OutputArray surface used by the drop-in ORBextractor
  Mat getMat() const { return *this; }
  void release() { *this = Mat(); }
  size_t step = 0;  // bytes per row (rows are dense in this stand-in)
  bool empty() const { return rows == 0 || cols == 0 || !buf_; }
  size_t elemSize() const { return esz(); }
  template <typename T>
  T* ptr(int r = 0) {
    return reinterpret_cast<T*>(buf_->data() + off_ + (size_t)r * cols * esz());
  }
  template <typename T>
  const T* ptr(int r = 0) const {
    return reinterpret_cast<const T*>(buf_->data() + off_ + (size_t)r * cols * esz());
  }
  template <typename T>
  T& at(int i) { return ptr<T>()[i]; }
  template <typename T>
  const T& at(int i) const { return ptr<T>()[i]; }
  template <typename T>
  T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T>
  const T& at(int r, int c) const { return ptr<T>(r)[c]; }
  Mat row(int r) const {
    Mat m;
    m.rows = 1;
    m.cols = cols;
    m.type_ = type_;
    m.buf_ = buf_;
    m.off_ = off_ + (size_t)r * cols * esz();
    return m;
  }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    std::memcpy(m.buf_->data(), buf_->data() + off_, (size_t)rows * cols * esz());
    return m;
  }
  void push_back(const Mat& r) {
    if (r.empty()) return;
    if (empty()) {
      *this = r.clone();
      return;
    }
    auto nb = std::make_shared<std::vector<unsigned char>>((size_t)(rows + r.rows) * cols * esz());
    std::memcpy(nb->data(), buf_->data() + off_, (size_t)rows * cols * esz());
    std::memcpy(nb->data() + (size_t)rows * cols * esz(), r.buf_->data() + r.off_,
                (size_t)r.rows * cols * esz());
    buf_ = nb;
    off_ = 0;
    rows += r.rows;
  }

 private:
  size_t esz() const { return type_ == CV_8U ? 1 : 4; }
  int type_ = CV_8U;
  std::shared_ptr<std::vector<unsigned char>> buf_;
  size_t off_ = 0;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif

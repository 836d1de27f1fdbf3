// Stand-in for the OpenCV 3 types and functions that the reference's hot-path sources
// (src/matcher.cpp, src/frame.cpp, src/map_point.cpp, src/bundle_adjust.cpp, src/map.cpp)
// mention, so that those files can be compiled UNMODIFIED, from where they lie under
// /root/reference, into oracle/_ref/libref.so in an image without OpenCV-C++.
//
// TEST INFRASTRUCTURE ONLY (oracle/ rules: imported by tests/, smoke() and bench.py's CPU legs).
// Nothing in the product (lorb_slam_b200/) includes this header.
//
// What is implemented and what is only declared:
//  * implemented, and pinned bit-exactly against cv2 4.13 golden vectors
//    (tests/golden/cvshim_golden.npz, tests/test_ref_build.py): float matrix product /
//    sum / difference (sequential fp32, k ascending, no FMA), Mat::inv() of a 4x4 float matrix
//    (partial-pivot LU in fp32, OpenCV's LUImpl order), cv::Rodrigues vector -> matrix
//    (double arithmetic, rounded to the output type), cv::norm of a Point3f (double),
//    BFMatcher(NORM_HAMMING, crossCheck) (3.4+/4.x mutual semantics).
//  * implemented, plain: views (row/col/rowRange/colRange share storage), clone, copyTo,
//    push_back, convertTo, eye/zeros/ones, comma initialiser, L1/L2 norms.
//  * declared only (link-time lazy, never called on the hot path; calling one aborts at the
//    PLT): imshow, waitKey, ORB::create, FeatureDetector/DescriptorExtractor, matrix -> vector
//    Rodrigues.
#ifndef LORB_ORACLE_CVSHIM_HPP
#define LORB_ORACLE_CVSHIM_HPP
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <set>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6

typedef unsigned char uchar;

// cvRound: round to nearest, ties to even (OpenCV: _mm_cvtsd_si32 / lrint)
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

using ::cvRound;
using ::cvFloor;
using ::cvCeil;
using ::uchar;

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

// cv::fastAtan2 (degrees, polynomial of core/src/mathfuncs_core): restated without FMA and pinned
// bit-exactly against cv2.fastAtan2 (tests/golden/cvshim_golden.npz)
inline float fastAtan2(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / CV_PI), p3 = -0.3258083974640975f * (float)(180 / CV_PI),
              p5 = 0.1555786518463281f * (float)(180 / CV_PI), p7 = -0.04432655554792128f * (float)(180 / CV_PI);
  const float ax = std::abs(x), ay = std::abs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

template <typename T>
struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
};
typedef Rect_<int> Rect;

enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };

// ------------------------------------------------------------------ points
template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <typename S>
  Point_& operator*=(S s) {
    x = (T)(x * s);
    y = (T)(y * s);
    return *this;
  }
};
template <typename T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
  // OpenCV: saturate_cast<T>(x*pt.x + y*pt.y + z*pt.z), evaluated in T
  T dot(const Point3_& p) const { return (T)(x * p.x + y * p.y + z * p.z); }
};
template <typename T>
inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) {
  return Point3_<T>((T)(a.x - b.x), (T)(a.y - b.y), (T)(a.z - b.z));
}
template <typename T>
inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) {
  return Point3_<T>((T)(a.x + b.x), (T)(a.y + b.y), (T)(a.z + b.z));
}
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point3_<float> Point3f;

// OpenCV: std::sqrt((double)x*x + (double)y*y + (double)z*z)
template <typename T>
inline double norm(const Point3_<T>& p) {
  return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z);
}

struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
  static void convert(const std::vector<KeyPoint>& kps, std::vector<Point2f>& pts) {
    pts.resize(kps.size());
    for (size_t i = 0; i < kps.size(); i++) pts[i] = kps[i].pt;
  }
};

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = FLT_MAX;
};

// --------------------------------------------------------------------- Mat
template <typename T>
class Mat_;
template <typename T>
class MatCommaInitializer_;

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size sz, int type) { create(sz.height, sz.width, type); }
  template <typename T>
  explicit Mat(const Point3_<T>& p) {
    create(3, 1, sizeof(T) == 4 ? CV_32F : CV_64F);
    at<T>(0) = p.x;
    at<T>(1) = p.y;
    at<T>(2) = p.z;
  }
  void create(int r, int c, int type) {
    rows = r;
    cols = c;
    type_ = type;
    step = (size_t)c * esz(type);
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step + 8, 0);
    off_ = 0;
  }
  // Mat::zeros is a MatExpr in OpenCV: ASSIGNING it to a matrix that already has that shape zeroes
  // the matrix in place (Mat::create keeps the buffer), it does not rebind the header --
  // computeDescriptors (src/ORBextractor.cpp:1081) relies on that to write into a rowRange view.
  struct ZerosExpr {
    int r, c, type;
    operator Mat() const { return Mat(r, c, type); }
  };
  static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
  Mat(const ZerosExpr& z) { create(z.r, z.c, z.type); }
  Mat& operator=(const ZerosExpr& z) {
    if (buf_ && rows == z.r && cols == z.c && type_ == z.type) {
      for (int i = 0; i < rows; i++) std::memset(raw(i), 0, (size_t)cols * esz(type_));
    } else {
      create(z.r, z.c, z.type);
    }
    return *this;
  }
  static Mat ones(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) m.set(i, j, 1.0);
    return m;
  }
  static Mat eye(int r, int c, int type) {
    Mat m(r, c, type);
    for (int i = 0; i < std::min(r, c); i++) m.set(i, i, 1.0);
    return m;
  }
  size_t step = 0;  // bytes per row (public like cv::Mat::step)
  size_t step1() const { return step / esz(type_); }
  Size size() const { return Size(cols, rows); }
  Mat operator()(const Rect& r) const { return view(r.y, r.y + r.height, r.x, r.x + r.width); }
  Mat getMat() const { return *this; }
  void release() { *this = Mat(); }
  uchar* ptr(int r = 0) { return raw(r); }
  const uchar* ptr(int r = 0) const { return raw(r); }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || cols == 0 || !buf_; }
  size_t elemSize() const { return esz(type_); }
  bool isContinuous() const { return step == (size_t)cols * esz(type_); }

  unsigned char* raw(int r) const { return buf_->data() + off_ + (size_t)r * step; }
  template <typename T>
  T* ptr(int r = 0) { return reinterpret_cast<T*>(raw(r)); }
  template <typename T>
  const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(raw(r)); }
  // single-index access follows OpenCV: row vector -> column i, otherwise row i (of a column vector)
  template <typename T>
  T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  template <typename T>
  const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
  template <typename T>
  T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T>
  const T& at(int r, int c) const { return ptr<T>(r)[c]; }

  Mat view(int r0, int r1, int c0, int c1) const {
    Mat m;
    m.rows = r1 - r0;
    m.cols = c1 - c0;
    m.type_ = type_;
    m.step = step;
    m.buf_ = buf_;
    m.off_ = off_ + (size_t)r0 * step + (size_t)c0 * esz(type_);
    return m;
  }
  Mat row(int r) const { return view(r, r + 1, 0, cols); }
  Mat col(int c) const { return view(0, rows, c, c + 1); }
  Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
  Mat colRange(int a, int b) const { return view(0, rows, a, b); }

  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; r++) std::memcpy(m.raw(r), raw(r), (size_t)cols * esz(type_));
    return m;
  }
  // OutputArray semantics: a destination of the right size and type is written in place
  // (that is how the reference fills sub-blocks of mTcw), otherwise it is re-allocated.
  void copyTo(const Mat& dst_) const {
    Mat& dst = const_cast<Mat&>(dst_);
    if (dst.empty() || dst.rows != rows || dst.cols != cols || dst.type_ != type_) {
      dst = clone();
      return;
    }
    for (int r = 0; r < rows; r++) std::memmove(dst.raw(r), raw(r), (size_t)cols * esz(type_));
  }
  void convertTo(Mat& dst, int type) const {
    Mat out(rows, cols, type);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) out.set(r, c, get(r, c));
    dst = out;
  }
  // amortised growth like cv::Mat::push_back (reserve 1.5x), so that the reference's
  // row-by-row descriptor gathering (src/matcher.cpp:30) costs what it costs with OpenCV
  void push_back(const Mat& m) {
    if (m.empty()) return;
    if (empty()) {
      *this = m.clone();
      return;
    }
    assert(m.cols == cols && m.type_ == type_);
    const size_t need = off_ + (size_t)(rows + m.rows) * step + 8;
    if (!isContinuous() || buf_.use_count() != 1 || buf_->capacity() < need) {
      auto nb = std::make_shared<std::vector<unsigned char>>();
      nb->reserve(std::max(need, (size_t)((rows + m.rows) * 3 / 2 + 4) * cols * esz(type_) + 8));
      nb->resize((size_t)(rows + m.rows) * cols * esz(type_) + 8);
      for (int r = 0; r < rows; r++) std::memcpy(nb->data() + (size_t)r * cols * esz(type_), raw(r), (size_t)cols * esz(type_));
      buf_ = nb;
      off_ = 0;
      step = (size_t)cols * esz(type_);
    } else {
      if (buf_->size() < need) buf_->resize(need);  // within capacity: no reallocation
    }
    for (int r = 0; r < m.rows; r++) std::memcpy(raw(rows + r), m.raw(r), (size_t)cols * esz(type_));
    rows += m.rows;
  }
  Mat t() const {
    Mat out(cols, rows, type_);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++)
        std::memcpy(out.raw(c) + (size_t)r * esz(type_), raw(r) + (size_t)c * esz(type_),
                    esz(type_));
    return out;
  }
  inline Mat inv() const;

  // generic element access in double (exact for u8 / f32 / f64 / i32)
  double get(int r, int c) const {
    switch (type_) {
      case CV_8U: return ptr<unsigned char>(r)[c];
      case CV_32S: return ptr<int>(r)[c];
      case CV_32F: return ptr<float>(r)[c];
      default: return ptr<double>(r)[c];
    }
  }
  void set(int r, int c, double v) {
    switch (type_) {
      case CV_8U: ptr<unsigned char>(r)[c] = (unsigned char)std::lrint(std::min(255.0, std::max(0.0, v))); break;
      case CV_32S: ptr<int>(r)[c] = (int)std::lrint(v); break;
      case CV_32F: ptr<float>(r)[c] = (float)v; break;
      default: ptr<double>(r)[c] = v; break;
    }
  }
  static size_t esz(int type) { return type == CV_8U ? 1 : (type == CV_64F ? 8 : 4); }

 private:
  int type_ = CV_8U;
  size_t off_ = 0;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

template <typename T>
struct DataType_ { enum { type = CV_64F }; };
template <>
struct DataType_<float> { enum { type = CV_32F }; };
template <>
struct DataType_<int> { enum { type = CV_32S }; };
template <>
struct DataType_<unsigned char> { enum { type = CV_8U }; };

template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DataType_<T>::type) {}
};

// (Mat_<T>(r,c) << a, b, c): values are converted to T one by one, row-major
template <typename T>
class MatCommaInitializer_ {
 public:
  explicit MatCommaInitializer_(const Mat_<T>& m) : m_(m), k_(0) {}
  template <typename T2>
  MatCommaInitializer_& operator,(T2 v) {
    put((T)v);
    return *this;
  }
  void put(T v) {
    m_.template at<T>(k_ / m_.cols, k_ % m_.cols) = v;
    k_++;
  }
  operator Mat_<T>() const { return m_; }

 private:
  Mat_<T> m_;
  int k_;
};
template <typename T, typename T2>
inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v) {
  MatCommaInitializer_<T> ci(m);
  ci.put((T)v);
  return ci;
}

// ------------------------------------------------------------ arithmetic
// cv::gemm on small float matrices: every output element is a sequential fp32 sum over k
// ascending, products rounded separately (no FMA; this file is compiled -ffp-contract=off).
// An expression A*B+C is evaluated by OpenCV as gemm(A,B,1,C,1): the product first, then + C,
// which is what the eager operators below do (pinned: tests/golden/gemm_golden.npz).
inline Mat operator*(const Mat& a, const Mat& b) {
  assert(a.cols == b.rows && a.type() == b.type());
  Mat out(a.rows, b.cols, a.type());
  if (a.type() == CV_32F) {
    for (int i = 0; i < a.rows; i++)
      for (int j = 0; j < b.cols; j++) {
        float s = a.at<float>(i, 0) * b.at<float>(0, j);
        for (int k = 1; k < a.cols; k++) {
          const float p = a.at<float>(i, k) * b.at<float>(k, j);
          s = s + p;
        }
        out.at<float>(i, j) = s;
      }
  } else {
    for (int i = 0; i < a.rows; i++)
      for (int j = 0; j < b.cols; j++) {
        double s = 0;
        for (int k = 0; k < a.cols; k++) s += a.get(i, k) * b.get(k, j);
        out.set(i, j, s);
      }
  }
  return out;
}
template <typename F>
inline Mat cvshim_zip(const Mat& a, const Mat& b, F f) {
  assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
  Mat out(a.rows, a.cols, a.type());
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) {
      if (a.type() == CV_32F)
        out.at<float>(i, j) = f(a.at<float>(i, j), b.at<float>(i, j));
      else
        out.set(i, j, f(a.get(i, j), b.get(i, j)));
    }
  return out;
}
struct cvshim_add {
  float operator()(float x, float y) const { return x + y; }
  double operator()(double x, double y) const { return x + y; }
};
struct cvshim_sub {
  float operator()(float x, float y) const { return x - y; }
  double operator()(double x, double y) const { return x - y; }
};
inline Mat operator+(const Mat& a, const Mat& b) { return cvshim_zip(a, b, cvshim_add()); }
inline Mat operator-(const Mat& a, const Mat& b) { return cvshim_zip(a, b, cvshim_sub()); }
inline Mat cvshim_scale(const Mat& a, double s) {
  Mat out(a.rows, a.cols, a.type());
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) {
      if (a.type() == CV_32F)
        out.at<float>(i, j) = (float)(a.at<float>(i, j) * s); /* OpenCV scales in double */
      else
        out.set(i, j, a.get(i, j) * s);
    }
  return out;
}
inline Mat operator-(const Mat& a) { return cvshim_scale(a, -1.0); }
inline Mat operator*(double s, const Mat& a) { return cvshim_scale(a, s); }
inline Mat operator*(const Mat& a, double s) { return cvshim_scale(a, s); }
inline Mat operator/(const Mat& a, double s) { return cvshim_scale(a, 1.0 / s); }

inline double norm(const Mat& a, int type = NORM_L2) {
  double s = 0;
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) {
      const double v = a.get(i, j);
      s += type == NORM_L1 ? std::fabs(v) : v * v;
    }
  return type == NORM_L1 ? s : std::sqrt(s);
}
inline double norm(const Mat& a, const Mat& b, int type = NORM_L2) {
  double s = 0;
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) {
      const double v = a.get(i, j) - b.get(i, j);
      s += type == NORM_L1 ? std::fabs(v) : v * v;
    }
  return type == NORM_L1 ? s : std::sqrt(s);
}

// cv::invert(DECOMP_LU) for n > 3: LU with partial pivoting on a copy, applied to the identity,
// in the matrix' own precision (OpenCV hal LUImpl: d = -1/pivot, rows updated by alpha = a*d,
// back-substitution s -= a*b; s / pivot).
template <typename T>
inline bool cvshim_lu_inv(T* A, int n, T* B, T eps) {
  for (int i = 0; i < n; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++)
      if (std::abs(A[j * n + i]) > std::abs(A[k * n + i])) k = j;
    if (std::abs(A[k * n + i]) < eps) return false;
    if (k != i) {
      for (int j = i; j < n; j++) std::swap(A[i * n + j], A[k * n + j]);
      for (int j = 0; j < n; j++) std::swap(B[i * n + j], B[k * n + j]);
    }
    const T d = -1 / A[i * n + i];
    for (int j = i + 1; j < n; j++) {
      const T alpha = A[j * n + i] * d;
      for (int c = i + 1; c < n; c++) {
        const T p = alpha * A[i * n + c];
        A[j * n + c] = A[j * n + c] + p;
      }
      for (int c = 0; c < n; c++) {
        const T p = alpha * B[i * n + c];
        B[j * n + c] = B[j * n + c] + p;
      }
    }
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < n; j++) {
      T s = B[i * n + j];
      for (int k = i + 1; k < n; k++) {
        const T p = A[i * n + k] * B[k * n + j];
        s = s - p;
      }
      B[i * n + j] = s / A[i * n + i];
    }
  return true;
}
inline Mat Mat::inv() const {
  assert(rows == cols && rows >= 4 && "only the general LU branch (n > 3) is restated");
  const int n = rows;
  Mat out = Mat::eye(n, n, type_);
  if (type_ == CV_32F) {
    std::vector<float> A((size_t)n * n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) A[i * n + j] = at<float>(i, j);
    if (!cvshim_lu_inv<float>(A.data(), n, out.ptr<float>(), FLT_EPSILON * 10))
      return Mat::zeros(n, n, type_);
  } else {
    std::vector<double> A((size_t)n * n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) A[i * n + j] = get(i, j);
    if (!cvshim_lu_inv<double>(A.data(), n, out.ptr<double>(), DBL_EPSILON * 100))
      return Mat::zeros(n, n, type_);
  }
  return out;
}

// cv::Rodrigues, rotation vector (3x1 or 1x3, float or double) -> 3x3 matrix of the same type.
// OpenCV (calib3d): double arithmetic; theta = |r|; below DBL_EPSILON the identity; otherwise
// R = c*I + (1-c)*r r^T + s*[r]x with r normalised by 1/theta, stored to the output type.
inline void Rodrigues(const Mat& src, Mat& dst) {
  if (!((src.rows == 3 && src.cols == 1) || (src.rows == 1 && src.cols == 3))) {
    std::fprintf(stderr, "cvshim: Rodrigues matrix->vector is not on the hot path\n");
    std::abort();
  }
  double r[3];
  for (int i = 0; i < 3; i++) r[i] = src.type() == CV_32F ? (double)src.at<float>(i) : src.at<double>(i);
  const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  double R[9];
  if (theta < DBL_EPSILON) {
    for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
  } else {
    const double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c;
    const double itheta = theta ? 1. / theta : 0.;
    r[0] *= itheta;
    r[1] *= itheta;
    r[2] *= itheta;
    const double rrt[9] = {r[0] * r[0], r[0] * r[1], r[0] * r[2], r[0] * r[1], r[1] * r[1],
                           r[1] * r[2], r[0] * r[2], r[1] * r[2], r[2] * r[2]};
    const double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * rx[k];
  }
  Mat out(3, 3, src.type());
  for (int k = 0; k < 9; k++) out.set(k / 3, k % 3, R[k]);
  dst = out;
}

// ------------------------------------------------------------- BFMatcher
// cv::BFMatcher(NORM_HAMMING, crossCheck).match(query, train): forward = first strict minimum
// over train, backward = first strict minimum over query, keep i iff backward[forward[i]] == i
// (OpenCV >= 3.4 semantics, pinned against cv2 4.13: tests/golden/bf_golden.npz), ascending i.
class BFMatcher {
 public:
  BFMatcher(int normType = NORM_L2, bool crossCheck = false) : cross_(crossCheck) { (void)normType; }
  void match(const Mat& q, const Mat& t, std::vector<DMatch>& out) const {
    out.clear();
    const int N = q.rows, M = t.rows;
    if (N == 0 || M == 0) return;
    std::vector<int> fwd(N, -1), fd(N, INT_MAX), bwd(M, -1), bd(M, INT_MAX);
    for (int i = 0; i < N; i++) {
      const uint64_t* a = reinterpret_cast<const uint64_t*>(q.raw(i));
      for (int j = 0; j < M; j++) {
        const uint64_t* b = reinterpret_cast<const uint64_t*>(t.raw(j));
        int d = 0;
        for (int w = 0; w < q.cols / 8; w++) d += __builtin_popcountll(a[w] ^ b[w]);
        if (d < fd[i]) { fd[i] = d; fwd[i] = j; }
        if (d < bd[j]) { bd[j] = d; bwd[j] = i; }
      }
    }
    for (int i = 0; i < N; i++) {
      if (cross_ && bwd[fwd[i]] != i) continue;
      DMatch m;
      m.queryIdx = i;
      m.trainIdx = fwd[i];
      m.imgIdx = 0;
      m.distance = (float)fd[i];
      out.push_back(m);
    }
  }

 private:
  bool cross_;
};

// ------------------------------------------------- declared, never defined
template <typename T>
struct Ptr {
  T* p = nullptr;
  Ptr() {}
  Ptr(T* q) : p(q) {}
  template <typename U>
  Ptr(const Ptr<U>& o) : p(o.p) {}
  T* operator->() const { return p; }
};
struct Feature2D {
  void detect(const Mat& img, std::vector<KeyPoint>& kps);
  void compute(const Mat& img, std::vector<KeyPoint>& kps, Mat& desc);
};
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;
struct ORB : Feature2D {
  static Ptr<ORB> create();
};
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};
void imshow(const std::string& name, const Mat& img);
int waitKey(int delay = 0);

}  // namespace cv
#endif

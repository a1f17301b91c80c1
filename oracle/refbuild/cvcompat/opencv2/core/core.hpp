// =====================================================================================
// OpenCV-compat layer for building the UNMODIFIED reference sources in an image without
// the OpenCV SDK (oracle/refbuild/build.sh -> oracle/_ref/).  TEST INFRASTRUCTURE ONLY.
//
// It declares exactly the cv:: surface that
//   /root/reference/introspective_ORB_SLAM/src/ORBextractor.cc  (whole file) and
//   /root/reference/introspective_ORB_SLAM/src/Frame.cc:758-932 (ComputeStereoMatches)
// touch.  Containers (Mat with ROI + shared storage, KeyPoint, Point_, Size, Rect,
// InputArray/OutputArray proxies, the zeros/ones initialiser expression) are written
// here; the pixel primitives (resize, GaussianBlur, FAST, fastAtan2) are NOT restated a
// second time: cvcompat_impl.cpp forwards them to the functions of
// oracle/ivslam_oracle.cpp that tests/test_oracle_vs_cv2.py pins byte-for-byte to this
// image's cv2 4.13.  So  _ref  =  reference control flow + float code as the reference's
// authors wrote it  +  cv2-pinned pixel arithmetic.
//
// Semantics that matter for the reference and are reproduced on purpose:
//   * Mat::create is a no-op when size and type already match (so resize/copyMakeBorder
//     write INTO the ROI of the padded buffer, ORBextractor.cc:1311-1313);
//   * `Mat& m = ...; m = Mat::zeros(r,c,t)` fills the existing storage when it matches
//     (MatOp_Initializer::assign), so computeDescriptors (ORBextractor.cc:1218) writes
//     into the caller's descriptor rows and does not re-seat the header;
//   * cvRound = cvtss2si/cvtsd2si (round-half-even); cvFloor/cvCeil as in fast_math.hpp;
//   * KeyPoint has cv::KeyPoint's 28-byte layout; FAST emits (x, y, 7, -1, score, 0, -1).
// =====================================================================================
#pragma once
#include <algorithm>
#include <cassert>
#include <stdexcept>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <emmintrin.h>

#define CV_VERSION_MAJOR 4
#define CV_VERSION_MINOR 13
#define CV_PI 3.1415926535897932384626433832795

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5

typedef unsigned char uchar;

// OpenCV's CV_Assert is active in Release builds and throws cv::Exception; here: std::runtime_error.
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

inline int cvRound(double value) { return _mm_cvtsd_si32(_mm_set_sd(value)); }
inline int cvRound(float value) { return _mm_cvtss_si32(_mm_set_ss(value)); }
inline int cvRound(int value) { return value; }
inline int cvFloor(double value) { int i = (int)value; return i - (i > value); }
inline int cvFloor(float value) { int i = (int)value; return i - (i > value); }
inline int cvCeil(double value) { int i = (int)value; return i + (i < value); }
inline int cvCeil(float value) { int i = (int)value; return i + (i < value); }

namespace cv {

using ::uchar;

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }   // saturate_cast<float> is the identity
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {} };
struct Scalar { double val[4]; double operator[](int i) const { return val[i]; } };

struct KeyPoint {
  Point2f pt; float size; float angle; float response; int octave; int class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

// Mat::zeros / Mat::ones (a MatExpr in OpenCV): only ever assigned to a Mat or combined as `Mat - scalar*ones`.
struct MatInit { int rows, cols, type; double value; };
inline MatInit operator*(double s, const MatInit& m) { return MatInit{m.rows, m.cols, m.type, m.value * s}; }

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;          // bytes per row (MatStep converts to size_t in OpenCV)
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size sz, int type) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((uchar*)ext), type_(type) {}
  Mat(const MatInit& e) { *this = e; }

  int type() const { return type_; }
  size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
  size_t step1() const { return step / elemSize(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    store_ = std::shared_ptr<uchar>(new uchar[std::max<size_t>((size_t)r * step, 1)], std::default_delete<uchar[]>());
    data = store_.get();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { store_.reset(); data = nullptr; rows = cols = 0; step = 0; }

  Mat& operator=(const MatInit& e) {
    create(e.rows, e.cols, e.type);
    for (int y = 0; y < rows; ++y) {
      if (type_ == CV_32F) { float* p = ptr<float>(y); for (int x = 0; x < cols; ++x) p[x] = (float)e.value; }
      else std::memset(ptr(y), (int)e.value, cols);
    }
    return *this;
  }
  static MatInit zeros(int r, int c, int type) { return MatInit{r, c, type, 0.0}; }
  static MatInit ones(int r, int c, int type) { return MatInit{r, c, type, 1.0}; }

  Mat rowRange(int a, int b) const { CV_Assert(0 <= a && a <= b && b <= rows); Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { CV_Assert(0 <= a && a <= b && b <= cols); Mat m = *this; m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }

  Mat clone() const {
    Mat m; m.create(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memcpy(m.ptr(y), ptr(y), (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memmove(dst.ptr(y), ptr(y), (size_t)cols * elemSize());
  }
  // only CV_8U -> CV_32F is used (Frame.cc:856,873).  `m.convertTo(m, CV_32F)` on a ROI header allocates, like OpenCV.
  void convertTo(Mat& dst, int rtype) const {
    CV_Assert(type_ == CV_8U && rtype == CV_32F);
    Mat src = *this, out;
    out.create(src.rows, src.cols, CV_32F);
    for (int y = 0; y < src.rows; ++y) { const uchar* s = src.ptr(y); float* d = out.ptr<float>(y); for (int x = 0; x < src.cols; ++x) d[x] = (float)s[x]; }
    dst = out;
  }

  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
  template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }

 private:
  std::shared_ptr<uchar> store_;
  int type_ = 0;
};

// `IL - IL.at<float>(w,w) * Mat::ones(...)` (Frame.cc:857,874): CV_32F, exact for the integer-valued inputs
inline Mat operator-(const Mat& a, const MatInit& e) {
  CV_Assert(a.type() == CV_32F && e.rows == a.rows && e.cols == a.cols);
  Mat out; out.create(a.rows, a.cols, CV_32F);
  const float v = (float)e.value;
  for (int y = 0; y < a.rows; ++y) { const float* s = a.ptr<float>(y); float* d = out.ptr<float>(y); for (int x = 0; x < a.cols; ++x) d[x] = s[x] - v; }
  return out;
}

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_->empty(); }
  Mat getMat() const { return *m_; }
  int type() const { return m_->type(); }
 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void release() const { m_->release(); }
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  Mat getMat() const { return *m_; }
  Mat& ref() const { return *m_; }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

float fastAtan2(float y, float x);
Scalar sum(InputArray src);
double norm(InputArray a, InputArray b, int normType);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);

class KeyPointsFilter {
 public:
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv

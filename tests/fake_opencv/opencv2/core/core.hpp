// Minimal stand-in for <opencv2/core/core.hpp>: just enough of cv::Mat / cv::KeyPoint / InputArray / OutputArray to
// COMPILE and exercise shim/*.cc in an image without the OpenCV SDK.  Test infrastructure only (tests/test_shim.py).
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32FC1 5

namespace cv {

struct Point2f { float x, y; };

struct KeyPoint {
  Point2f pt; float size; float angle; float response; int octave; int class_id;
};

class _OutputArray;

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((unsigned char*)ext), type_(type) {}
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == 5 ? 4 : 1); }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && store_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    store_ = std::shared_ptr<unsigned char>(new unsigned char[(size_t)r * step + 1], std::default_delete<unsigned char[]>());
    data = store_.get();
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
  void release() { store_.reset(); data = nullptr; rows = cols = 0; }
  void copyTo(const _OutputArray& o) const;
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
  }
 private:
  std::shared_ptr<unsigned char> store_;
  int type_ = 0;
};

// In OpenCV these are proxy classes; references to Mat are enough for the shim's use.
class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_->empty(); }
  Mat getMat() const { return *m_; }
 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void release() const { m_->release(); }
  Mat& ref() const { return *m_; }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline void Mat::copyTo(const _OutputArray& o) const { copyTo(o.ref()); }

}  // namespace cv

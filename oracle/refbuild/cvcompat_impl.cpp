// Implementation half of the OpenCV-compat layer (see cvcompat/opencv2/core/core.hpp).  TEST INFRASTRUCTURE ONLY.
//
// Pixel primitives forward to oracle/ivslam_oracle.cpp (linked into the same shared object), whose
// resize / GaussianBlur / FAST / fastAtan2 are pinned byte-for-byte to cv2 4.13 by tests/test_oracle_vs_cv2.py.
// The functions below that are written out (copyMakeBorder, sum, norm, KeyPointsFilter::retainBest) follow the
// published OpenCV algorithms; retainBest uses the real libstdc++ std::nth_element / std::partition exactly as
// OpenCV's features2d/src/keypoint.cpp does.
#include <opencv2/core/core.hpp>

extern "C" {
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw, int dh, size_t dstride);
void orc_gauss7_u8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride);
const int* orc_fast9_tl(const uint8_t* img, int w, int h, size_t stride, int th, int nms, int* n);
float orc_fast_atan2(float y, float x);
}

namespace cv {

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

Scalar sum(InputArray _src) {
  Mat src = _src.getMat();
  CV_Assert(src.type() == CV_8U);
  unsigned long long s = 0;   // OpenCV accumulates 8-bit sums in int blocks then double: exact either way
  for (int y = 0; y < src.rows; ++y) { const uchar* p = src.ptr(y); for (int x = 0; x < src.cols; ++x) s += p[x]; }
  return Scalar{{(double)s, 0, 0, 0}};
}

double norm(InputArray _a, InputArray _b, int normType) {
  Mat a = _a.getMat(), b = _b.getMat();
  CV_Assert(normType == NORM_L1 && a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
  double s = 0;   // normDiffL1_<float,double>
  for (int y = 0; y < a.rows; ++y) {
    const float *p = a.ptr<float>(y), *q = b.ptr<float>(y);
    for (int x = 0; x < a.cols; ++x) s += std::abs(p[x] - q[x]);
  }
  return s;
}

static int border_interpolate(int p, int len, int borderType) {
  if ((unsigned)p < (unsigned)len) return p;
  CV_Assert(borderType == BORDER_REFLECT_101);
  if (len == 1) return 0;
  do { if (p < 0) p = -p; else p = 2 * len - 2 - p; } while ((unsigned)p >= (unsigned)len);
  return p;
}

// The reference only ever uses BORDER_REFLECT_101 (ORBextractor.cc:1313-1319, 1345-1353); a non-isolated source that is
// a sub-matrix is treated as isolated (the 19-px frame it would fill is never read on this path).
void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType) {
  Mat src = _src.getMat();
  borderType &= ~BORDER_ISOLATED;
  CV_Assert(src.type() == CV_8U);
  _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
  Mat dst = _dst.getMat();
  std::vector<int> tab(dst.cols);
  for (int x = 0; x < dst.cols; ++x) tab[x] = border_interpolate(x - left, src.cols, borderType);
  for (int y = 0; y < src.rows; ++y) {          // interior rows first (source may alias the destination's interior)
    uchar* d = dst.ptr(y + top);
    const uchar* s = src.ptr(y);
    if (d + left != s) std::memmove(d + left, s, src.cols);
    for (int x = 0; x < left; ++x) d[x] = d[left + tab[x]];
    for (int x = left + src.cols; x < dst.cols; ++x) d[x] = d[left + tab[x]];
  }
  for (int y = 0; y < top; ++y) std::memcpy(dst.ptr(y), dst.ptr(top + border_interpolate(y - top, src.rows, borderType)), dst.cols);
  for (int y = top + src.rows; y < dst.rows; ++y) std::memcpy(dst.ptr(y), dst.ptr(top + border_interpolate(y - top, src.rows, borderType)), dst.cols);
}

void resize(InputArray _src, OutputArray _dst, Size dsize, double fx, double fy, int interpolation) {
  Mat src = _src.getMat();
  CV_Assert(src.type() == CV_8U && interpolation == INTER_LINEAR && fx == 0 && fy == 0 && dsize.width > 0 && dsize.height > 0);
  _dst.create(dsize.height, dsize.width, src.type());
  Mat dst = _dst.getMat();
  orc_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  Mat src = _src.getMat();
  CV_Assert(src.type() == CV_8U && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
  _dst.create(src.rows, src.cols, src.type());
  Mat dst = _dst.getMat();
  orc_gauss7_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step);   // reads all of src before it writes dst
}

void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  Mat img = _img.getMat();
  CV_Assert(img.type() == CV_8U);
  keypoints.clear();
  int n = 0;
  const int* c = orc_fast9_tl(img.data, img.cols, img.rows, img.step, threshold, nonmaxSuppression ? 1 : 0, &n);
  for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)c[3 * i], (float)c[3 * i + 1], 7.f, -1, (float)c[3 * i + 2]));
}

namespace {
struct KeypointResponseGreaterThanOrEqualToThreshold {
  explicit KeypointResponseGreaterThanOrEqualToThreshold(float v) : value(v) {}
  bool operator()(const KeyPoint& kpt) const { return kpt.response >= value; }
  float value;
};
struct KeypointResponseGreater {
  bool operator()(const KeyPoint& kp1, const KeyPoint& kp2) const { return kp1.response > kp2.response; }
};
}  // namespace

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int n_points) {
  if (n_points >= 0 && keypoints.size() > (size_t)n_points) {
    if (n_points == 0) { keypoints.clear(); return; }
    std::nth_element(keypoints.begin(), keypoints.begin() + n_points - 1, keypoints.end(), KeypointResponseGreater());
    float ambiguous_response = keypoints[n_points - 1].response;
    std::vector<KeyPoint>::const_iterator new_end =
        std::partition(keypoints.begin() + n_points, keypoints.end(), KeypointResponseGreaterThanOrEqualToThreshold(ambiguous_response));
    keypoints.resize(new_end - keypoints.begin());
  }
}

}  // namespace cv

// =====================================================================================
// oracle/_ref C API: the UNMODIFIED reference hot path behind plain C entry points.
// TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
// --impl reference legs).  Nothing under iv_slam_b200/, shim/ or include/ links or loads it.
//
// What is compiled here, none of it edited:
//   * /root/reference/introspective_ORB_SLAM/src/ORBextractor.cc — the whole file, as its own
//     translation unit, with the reference's own include/ORBextractor.h (see build.sh);
//   * Frame::ComputeStereoMatches (src/Frame.cc:758-932), ORBmatcher::TH_HIGH/TH_LOW
//     (src/ORBmatcher.cc:37-39) and ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1698-1716),
//     cut out verbatim by extract_reference.py and #included below into stub classes that
//     declare only the members those bodies touch, with the types include/Frame.h:163-313 and
//     include/ORBmatcher.h:44,93-95 give them.
// OpenCV is replaced by cvcompat/ (containers written there, pixel primitives forwarded to
// the cv2-pinned ones of oracle/ivslam_oracle.cpp).
// =====================================================================================
#include <opencv2/core/core.hpp>

#include <atomic>
#include <cstdio>
#include <thread>
#include <utility>
#include <vector>

#include "ORBextractor.h"   // the reference's header (-I /root/reference/introspective_ORB_SLAM/include)

using namespace std;   // Frame.h / ORBmatcher.cc rely on it (src/ORBmatcher.cc:32)

namespace ORB_SLAM2 {

class ORBmatcher {   // include/ORBmatcher.h:44,93-95
 public:
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;
};

class Frame {   // include/Frame.h:163,190,206,209,216,221,255,256,263,312,313
 public:
  void ComputeStereoMatches();
  ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
  float mbf;
  float mb;
  int N;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
  std::vector<float> mvuRight;
  std::vector<float> mvDepth;
  cv::Mat mDescriptors, mDescriptorsRight;
  vector<float> mvScaleFactors;
  vector<float> mvInvScaleFactors;
};

#include "gen/ORBmatcher_constants.inc"
#include "gen/Frame_ComputeStereoMatches.inc"
#include "gen/ORBmatcher_DescriptorDistance.inc"

}  // namespace ORB_SLAM2

namespace {

// protected members of the reference class, reachable from a derived class without touching the source
struct RefExtractor : ORB_SLAM2::ORBextractor {
  using ORB_SLAM2::ORBextractor::ORBextractor;
  const std::vector<int>& perLevel() const { return mnFeaturesPerLevel; }
  const std::vector<int>& uMax() const { return umax; }
};

int extract(RefExtractor& e, const uint8_t* img, int w, int h, size_t stride, const uint8_t* cost, size_t cost_stride,
            cv::KeyPoint* kps, uint8_t* desc, int cap, int* n_out) {
  cv::Mat image(h, w, CV_8UC1, (void*)img, stride), mask, descriptors;
  if (cost) mask = cv::Mat(h, w, CV_8UC1, (void*)cost, cost_stride);
  std::vector<cv::KeyPoint> keys;
  e(image, mask, keys, descriptors);
  const int n = (int)keys.size();
  *n_out = n;
  if (n > cap) return -3;
  for (int i = 0; i < n; ++i) kps[i] = keys[i];
  for (int i = 0; i < n; ++i) std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
  return 0;
}

int stereo(RefExtractor& L, RefExtractor& R, const cv::KeyPoint* kL, int N, const uint8_t* dL, const cv::KeyPoint* kR, int Nr,
           const uint8_t* dR, float mbf, float mb, float* uRight, float* depth) {
  if (N <= 0) return 0;   // the reference indexes an empty vector here (Frame.cc:919); nothing to report
  ORB_SLAM2::Frame F;
  F.mpORBextractorLeft = &L; F.mpORBextractorRight = &R;
  F.mbf = mbf; F.mb = mb; F.N = N;
  F.mvKeys.assign(kL, kL + N);
  F.mvKeysRight.assign(kR, kR + Nr);
  F.mDescriptors = cv::Mat(N, 32, CV_8U, (void*)dL, 32);
  F.mDescriptorsRight = cv::Mat(Nr, 32, CV_8U, (void*)dR, 32);
  F.mvScaleFactors = L.GetScaleFactors();            // Frame.cc:99-100
  F.mvInvScaleFactors = L.GetInverseScaleFactors();
  F.ComputeStereoMatches();
  for (int i = 0; i < N; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
  return 0;
}

template <typename Fn> int guarded(Fn&& fn) {
  try { return fn(); }
  catch (const std::exception& ex) { std::fprintf(stderr, "oracle/_ref: %s\n", ex.what()); return -1; }
}

}  // namespace

extern "C" {

// 1 when this object was built with FP contraction allowed (the reference's as-built flags), 0 for -ffp-contract=off
int ref_fp_contract() {
  return REF_FP_CONTRACT;   // set by build.sh
}

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int enableIntrospection) {
  return new RefExtractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, enableIntrospection != 0);
}
void ref_extractor_destroy(void* h) { delete (RefExtractor*)h; }

int ref_features_per_level(void* h, int* out) { auto& v = ((RefExtractor*)h)->perLevel(); for (size_t i = 0; i < v.size(); ++i) out[i] = v[i]; return (int)v.size(); }
int ref_umax(void* h, int* out) { auto& v = ((RefExtractor*)h)->uMax(); for (size_t i = 0; i < v.size(); ++i) out[i] = v[i]; return (int)v.size(); }
int ref_scale_factors(void* h, float* out) { auto v = ((RefExtractor*)h)->GetScaleFactors(); for (size_t i = 0; i < v.size(); ++i) out[i] = v[i]; return (int)v.size(); }

int ref_extract(void* h, const uint8_t* img, int w, int hgt, size_t stride, const uint8_t* cost, size_t cost_stride,
                void* kps, uint8_t* desc, int cap, int* n_out) {
  return guarded([&] { return extract(*(RefExtractor*)h, img, w, hgt, stride, cost, cost_stride, (cv::KeyPoint*)kps, desc, cap, n_out); });
}

// the public pyramid members (include/ORBextractor.h:91-92); which: 0 mvImagePyramid, 2 mvQualityImagePyramid
int ref_level_size(void* h, int level, int which, int* w, int* hgt) {
  RefExtractor* e = (RefExtractor*)h;
  const cv::Mat& m = which == 2 ? e->mvQualityImagePyramid[level] : e->mvImagePyramid[level];
  *w = m.cols; *hgt = m.rows;
  return m.empty() ? -1 : 0;
}
int ref_get_level(void* h, int level, int which, uint8_t* dst, size_t dstride) {
  RefExtractor* e = (RefExtractor*)h;
  const cv::Mat& m = which == 2 ? e->mvQualityImagePyramid[level] : e->mvImagePyramid[level];
  if (m.empty()) return -1;
  for (int y = 0; y < m.rows; ++y) std::memcpy(dst + (size_t)y * dstride, m.ptr(y), m.cols);
  return 0;
}

// Frame::ComputeStereoMatches on caller-supplied keypoints/descriptors; the pyramids are the ones the two extractors
// hold from their last operator() call.  `mb` is the reference's member (maxD = mbf/mb is computed inside, Frame.cc:789).
int ref_stereo_match(void* left, void* right, const void* kL, int N, const uint8_t* dL, const void* kR, int Nr, const uint8_t* dR,
                     float mbf, float mb, float* uRight, float* depth) {
  return guarded([&] { return stereo(*(RefExtractor*)left, *(RefExtractor*)right, (const cv::KeyPoint*)kL, N, dL, (const cv::KeyPoint*)kR, Nr, dR, mbf, mb, uRight, depth); });
}

// One stereo frame with the reference's threading (Frame.cc:115-125: two extraction threads, then matching on the caller).
int ref_stereo_frame(void* left, void* right, const uint8_t* imgL, const uint8_t* imgR, int w, int hgt, size_t stride,
                     const uint8_t* cost, size_t cost_stride, float mbf, float mb, int cap,
                     void* kL, uint8_t* dL, int* nL, void* kR, uint8_t* dR, int* nR, float* uRight, float* depth, int threads) {
  int rcL = 0, rcR = 0;
  auto runL = [&] { rcL = ref_extract(left, imgL, w, hgt, stride, cost, cost_stride, kL, dL, cap, nL); };
  auto runR = [&] { rcR = ref_extract(right, imgR, w, hgt, stride, nullptr, 0, kR, dR, cap, nR); };
  if (threads >= 2) { std::thread tl(runL), tr(runR); tl.join(); tr.join(); }
  else { runL(); runR(); }
  if (rcL) return rcL;
  if (rcR) return rcR;
  return ref_stereo_match(left, right, kL, *nL, dL, kR, *nR, dR, mbf, mb, uRight, depth);
}

// Frame-parallel batch for CPU timing: `workers` threads, each with its own extractor pair, pulling frames from a shared
// counter; every frame runs as ref_stereo_frame with threads=1.
// `costs` (nullable): n cost-maps, same layout as the images, applied to the left eye of an introspection-enabled extractor
// (the right extractor never weights, Tracking.cc:182-183).
int ref_stereo_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int n,
                     const uint8_t* imgsL, const uint8_t* imgsR, int w, int hgt, size_t stride,
                     float mbf, float mb, int workers, int* nL_out, int* nMatched_out, const uint8_t* costs) {
  std::atomic<int> next(0), err(0);
  auto work = [&] {
    RefExtractor eL(nfeatures, scaleFactor, nlevels, iniTh, minTh, costs != nullptr), eR(nfeatures, scaleFactor, nlevels, iniTh, minTh, false);
    const int cap = nfeatures + 64;
    std::vector<cv::KeyPoint> kL(cap), kR(cap);
    std::vector<uint8_t> dL((size_t)cap * 32), dR((size_t)cap * 32);
    std::vector<float> uR(cap), dep(cap);
    for (;;) {
      const int f = next.fetch_add(1);
      if (f >= n) break;
      int nl = 0, nr = 0;
      int rc = ref_stereo_frame(&eL, &eR, imgsL + (size_t)f * hgt * stride, imgsR + (size_t)f * hgt * stride, w, hgt, stride,
                                costs ? costs + (size_t)f * hgt * stride : nullptr, stride, mbf, mb, cap, kL.data(), dL.data(), &nl, kR.data(), dR.data(), &nr,
                                uR.data(), dep.data(), 1);
      if (rc) { err = rc; break; }
      int m = 0;
      for (int i = 0; i < nl; ++i) m += uR[i] >= 0;
      nL_out[f] = nl; nMatched_out[f] = m;
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < workers; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
  return err.load();
}

}  // extern "C"

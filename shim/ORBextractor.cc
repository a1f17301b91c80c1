/* shim/ORBextractor.cc — ORB_SLAM2::ORBextractor on the B200 C ABI (include/ivslam_gpu.h).
 * Replaces introspective_ORB_SLAM/src/ORBextractor.cc (constructor :411-476, operator() :1224-1296). */
#include "ORBextractor.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "ivslam_gpu.h"

namespace ORB_SLAM2 {

static std::atomic<int> g_device{-1};     // -1: not set, read IVSLAM_DEVICE

void ORBextractor::SetDevice(int device) { g_device.store(device); }
int ORBextractor::GetDevice() {
  int d = g_device.load();
  if (d < 0) {
    const char* e = std::getenv("IVSLAM_DEVICE");
    d = e ? std::atoi(e) : 0;
  }
  return d;
}

static void check(int rc, const char* what) {
  if (rc != IVG_OK)
    throw std::runtime_error(std::string(what) + ": " + ivg_strerror(rc) + " " + ivg_last_cuda_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST,
                           bool enableIntrospection)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
      minThFAST(_minThFAST), benableIntrospection(enableIntrospection) {
  check(ivg_extractor_create(&mHandle, GetDevice(), nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST,
                             enableIntrospection ? 1 : 0), "ivg_extractor_create");
  ivg_set_graph_mode(mHandle, 1);   // one frame at a time: replay the kernel sequence as one CUDA graph (0.186 vs 0.198 ms per stereo frame)
  mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
  ivg_get_scale_table(mHandle, 0, mvScaleFactor.data());
  ivg_get_scale_table(mHandle, 1, mvInvScaleFactor.data());
  ivg_get_scale_table(mHandle, 2, mvLevelSigma2.data());
  ivg_get_scale_table(mHandle, 3, mvInvLevelSigma2.data());
  mvImagePyramid.resize(nlevels);
  mvQualityImagePyramid.resize(nlevels);
  static_assert(sizeof(cv::KeyPoint) == sizeof(ivg_keypoint), "ivg_keypoint must mirror cv::KeyPoint");
}

ORBextractor::~ORBextractor() { ivg_extractor_destroy(mHandle); }

void ORBextractor::SetKeypointMode(int mode) { check(ivg_extractor_set_mode(mHandle, mode), "ivg_extractor_set_mode"); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
  if (_image.empty()) return;                                   // src/ORBextractor.cc:1227-1228
  cv::Mat image = _image.getMat();
  cv::Mat mask;
  bqualityScoresAvailable = !_mask.empty() && benableIntrospection;   // :1231
  if (bqualityScoresAvailable) mask = _mask.getMat();

  const int cap = ivg_max_keypoints(mHandle);
  _keypoints.resize(cap);
  cv::Mat desc(cap, 32, CV_8U);
  int n = 0;
  check(ivg_extract(mHandle, image.data, image.cols, image.rows, image.step,
                    bqualityScoresAvailable ? mask.data : nullptr, bqualityScoresAvailable ? (size_t)mask.step : 0,
                    reinterpret_cast<ivg_keypoint*>(_keypoints.data()), desc.data, cap, &n), "ivg_extract");
  _keypoints.resize(n);
  if (n == 0) _descriptors.release();                           // :1255-1256
  else desc.rowRange(0, n).copyTo(_descriptors);
}

void ORBextractor::SetRectifyMaps(const cv::Mat& M1, const cv::Mat& M2) {
  if (M1.empty() || M2.empty()) { check(ivg_set_rectify_maps(mHandle, nullptr, nullptr, 0, 0, 0), "ivg_set_rectify_maps"); return; }
  if (M1.type() != CV_32FC1 || M2.type() != CV_32FC1 || M1.rows != M2.rows || M1.cols != M2.cols || M1.step != M2.step)
    throw std::runtime_error("SetRectifyMaps: CV_32FC1 maps of equal size expected (initUndistortRectifyMap(..., CV_32F, ...))");
  check(ivg_set_rectify_maps(mHandle, reinterpret_cast<const float*>(M1.data), reinterpret_cast<const float*>(M2.data), M1.cols, M1.rows,
                             M1.step / sizeof(float)), "ivg_set_rectify_maps");
}

void ORBextractor::ExtractRaw(const cv::Mat& raw, const cv::Mat& cost, bool rgb, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
  if (raw.empty()) return;
  bqualityScoresAvailable = !cost.empty() && benableIntrospection;
  check(ivg_upload_batch_raw(mHandle, 1, raw.data, raw.cols, raw.rows, raw.step, raw.step * raw.rows, raw.channels(), rgb ? 1 : 0,
                             bqualityScoresAvailable ? cost.data : nullptr, bqualityScoresAvailable ? (size_t)cost.step : 0,
                             bqualityScoresAvailable ? (size_t)cost.step * cost.rows : 0), "ivg_upload_batch_raw");
  check(ivg_run_batch(mHandle), "ivg_run_batch");
  const int cap = ivg_max_keypoints(mHandle);
  _keypoints.resize(cap);
  cv::Mat desc(cap, 32, CV_8U);
  int n = 0;
  check(ivg_download_batch(mHandle, reinterpret_cast<ivg_keypoint*>(_keypoints.data()), desc.data, cap, &n), "ivg_download_batch");
  check(ivg_sync(mHandle), "ivg_sync");
  _keypoints.resize(n);
  if (n == 0) _descriptors.release();
  else desc.rowRange(0, n).copyTo(_descriptors);
}

void ORBextractor::SyncPyramidsToHost() {
  for (int l = 0; l < nlevels; ++l) {
    int w = 0, h = 0;
    check(ivg_level_size(mHandle, l, &w, &h), "ivg_level_size");
    mvImagePyramid[l].create(h, w, CV_8UC1);
    check(ivg_get_pyramid_level(mHandle, 0, l, 0, mvImagePyramid[l].data, mvImagePyramid[l].step), "ivg_get_pyramid_level");
    if (bqualityScoresAvailable) {
      mvQualityImagePyramid[l].create(h, w, CV_8UC1);
      check(ivg_get_pyramid_level(mHandle, 0, l, 2, mvQualityImagePyramid[l].data, mvQualityImagePyramid[l].step), "ivg_get_pyramid_level");
    }
  }
}

}  // namespace ORB_SLAM2

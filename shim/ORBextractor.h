/* shim/ORBextractor.h — source-compatible replacement for introspective_ORB_SLAM/include/ORBextractor.h.
 *
 * Same namespace, class name, constructor, operator(), getters and public pyramid members as the reference
 * (introspective_ORB_SLAM/include/ORBextractor.h:54-128), implemented on the C ABI of include/ivslam_gpu.h.
 * Drop this header + ORBextractor.cc into introspective_ORB_SLAM/{include,src} in place of the originals and link
 * libivslam_gpu.so.  Needs OpenCV headers to compile, like the file it replaces; in this repo it
 * is compile-checked against the minimal stand-in under tests/fake_opencv/.
 *
 * Differences a maintainer should know (all documented in DESIGN.md):
 *  - there is no CPU fallback: the constructor throws std::runtime_error if no sm_100 GPU is usable;
 *  - mvImagePyramid / mvQualityImagePyramid are filled lazily by SyncPyramidsToHost() (the device keeps the master
 *    copy; the GPU stereo matcher never needs them on the host).  Code that reads the pyramids directly, like the
 *    reference's Frame::ComputeStereoMatches, must call it first — the shim's ComputeStereoMatches does not need to;
 *  - ComputeKeyPointsOctTree / DistributeOctTree (dead code in the reference) are available as an optional mode,
 *    SetKeypointMode(1); the protected member functions themselves and ExtractorNode are not exposed.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>

#include <opencv2/core/core.hpp>

struct ivg_extractor;

namespace ORB_SLAM2 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
               bool enableIntrospection = false);
  ~ORBextractor();
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // Compute the ORB features and descriptors on an image; `mask` is the IV-SLAM cost-map (or empty).
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                  cv::OutputArray descriptors);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<cv::Mat> mvImagePyramid;
  std::vector<cv::Mat> mvQualityImagePyramid;

  // 0 (default): ComputeKeyPointsOld, the path the reference runs; 1: ComputeKeyPointsOctTree + DistributeOctTree, which
  // the reference compiles but never calls (src/ORBextractor.cc:1247-1248).  See ivg_extractor_set_mode.
  void SetKeypointMode(int mode);
  // Copies the device pyramids of the last operator() call into mvImagePyramid / mvQualityImagePyramid.
  void SyncPyramidsToHost();
  // N4 (optional): the steps the reference runs on the CPU before operator() — cv::remap with the CV_32FC1 maps of
  // cv::initUndistortRectifyMap (Examples/Stereo/stereo_kitti.cc:463-464; the cost-map too, :519-521) and
  // Tracking::GrabImageStereo's cvtColor (src/Tracking.cc:278-294) — fused into the upload.  ExtractRaw takes the frame as
  // read from the camera (CV_8UC1 / CV_8UC3 / CV_8UC4; rgb = Tracking::mbRGB) and the un-rectified cost-map (or empty).
  void SetRectifyMaps(const cv::Mat& M1, const cv::Mat& M2);    // empty Mats clear the maps
  void ExtractRaw(const cv::Mat& raw, const cv::Mat& cost, bool rgb, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);
  // The underlying C-ABI handle (used by the GPU Frame::ComputeStereoMatches replacement).
  ivg_extractor* handle() const { return mHandle; }
  // for ExtractStereoGPU (the one-call stereo front-end below): what operator() reads / sets on its own (src/ORBextractor.cc:1231)
  bool IntrospectionEnabled() const { return benableIntrospection; }
  void SetQualityScoresAvailable(bool available) { bqualityScoresAvailable = available; }

  // CUDA device used by extractors constructed AFTER the call (the reference's constructor has no such argument, so it is
  // process-wide state; default: environment variable IVSLAM_DEVICE, else device 0).
  static void SetDevice(int device);
  static int GetDevice();

 protected:
  int nfeatures;
  double scaleFactor;
  int nlevels;
  int iniThFAST;
  int minThFAST;
  bool benableIntrospection = false;
  bool bqualityScoresAvailable = false;

  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;

  ivg_extractor* mHandle = nullptr;
  std::vector<unsigned char> mKeypointStage;   // ivg_keypoint records (same layout as cv::KeyPoint)
};

// Replacement body for Frame::ComputeStereoMatches (introspective_ORB_SLAM/src/Frame.cc:758-932): fills mvuRight /
// mvDepth for the N keypoints the left extractor produced in its last call.  maxD is what the reference computes as
// mbf/mb (it reads mb before assigning it; pass fx).  See shim/Frame_ComputeStereoMatches.cc.
void ComputeStereoMatchesGPU(ORBextractor* left, ORBextractor* right, int N, float mbf, float maxD,
                             std::vector<float>& mvuRight, std::vector<float>& mvDepth);

// Optional: the whole stereo front-end of Frame::Frame in one call from one thread — the two ExtractORB threads
// (src/Frame.cc:115-125) and ComputeStereoMatches (:127).  Both eyes and the matcher are queued back to back on the device and
// the host waits once: no thread creation per frame, no two threads contending for the driver (0.18 -> 0.13 ms per KITTI frame).
// maskLeft: the cost-map Frame::Frame hands to both ExtractORBWeighted threads, or an empty Mat (it weights an eye only if that
// extractor was built with introspection; the reference builds the right one without).  See shim/Frame_ComputeStereoMatches.cc for the lines to change in Frame.cc.
void ExtractStereoGPU(ORBextractor* left, ORBextractor* right, const cv::Mat& imLeft, const cv::Mat& imRight, const cv::Mat& maskLeft,
                      std::vector<cv::KeyPoint>& keysLeft, cv::Mat& descLeft, std::vector<cv::KeyPoint>& keysRight, cv::Mat& descRight,
                      float mbf, float maxD, std::vector<float>& mvuRight, std::vector<float>& mvDepth);

}  // namespace ORB_SLAM2

#endif  // ORBEXTRACTOR_H

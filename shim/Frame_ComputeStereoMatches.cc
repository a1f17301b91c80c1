/* shim/Frame_ComputeStereoMatches.cc — GPU body for Frame::ComputeStereoMatches
 * (introspective_ORB_SLAM/src/Frame.cc:758-932, declared include/Frame.h:163).
 *
 * In Frame.cc the maintainer replaces the body of Frame::ComputeStereoMatches() by
 *
 *     ORB_SLAM2::ComputeStereoMatchesGPU(mpORBextractorLeft, mpORBextractorRight, N, mbf, fx, mvuRight, mvDepth);
 *
 * (fx: the reference derives maxD = mbf/mb with mb still unassigned at that point).  The matcher runs on
 * what the two extractors left on the device — pyramids, keypoints, descriptors — so nothing is uploaded. */
#include <algorithm>
#include <stdexcept>
#include <string>

#include "ORBextractor.h"
#include "ivslam_gpu.h"

namespace ORB_SLAM2 {

void ComputeStereoMatchesGPU(ORBextractor* left, ORBextractor* right, int N, float mbf, float maxD,
                             std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
  const int cap = ivg_max_keypoints(left->handle());
  std::vector<float> u(cap, -1.0f), d(cap, -1.0f);
  const int rc = ivg_stereo_match(left->handle(), right->handle(), mbf, maxD, u.data(), d.data(), cap);
  if (rc != IVG_OK) throw std::runtime_error(std::string("ivg_stereo_match: ") + ivg_strerror(rc) + " " + ivg_last_cuda_error());
  mvuRight.assign(u.begin(), u.begin() + N);      // mvuRight = vector<float>(N,-1.0f)  (src/Frame.cc:760-761)
  mvDepth.assign(d.begin(), d.begin() + N);
}

/* One call for the whole stereo front-end.  In Frame::Frame (src/Frame.cc:115-127) the maintainer replaces
 *
 *     thread threadLeft(&Frame::ExtractORBWeighted, this, 0, imLeft, costImg);   // or ExtractORB
 *     thread threadRight(&Frame::ExtractORB, this, 1, imRight);
 *     threadLeft.join(); threadRight.join();
 *     ...
 *     ComputeStereoMatches();
 *
 * by
 *
 *     ORB_SLAM2::ExtractStereoGPU(mpORBextractorLeft, mpORBextractorRight, imLeft, imRight, costImg /* or cv::Mat() *\/,
 *                                 mvKeys, mDescriptors, mvKeysRight, mDescriptorsRight, mbf, fx, mvuRight, mvDepth);
 *     N = mvKeys.size();
 *
 * and leaves everything between (UndistortKeyPoints, the quality scores) where it is. */
void ExtractStereoGPU(ORBextractor* left, ORBextractor* right, const cv::Mat& imLeft, const cv::Mat& imRight, const cv::Mat& maskLeft,
                      std::vector<cv::KeyPoint>& keysLeft, cv::Mat& descLeft, std::vector<cv::KeyPoint>& keysRight, cv::Mat& descRight,
                      float mbf, float maxD, std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
  if (imLeft.empty() || imRight.empty() || imLeft.cols != imRight.cols || imLeft.rows != imRight.rows || imLeft.step != imRight.step)
    throw std::runtime_error("ExtractStereoGPU: two non-empty images of the same size and row step expected");
  const bool weighted = !maskLeft.empty() && left->IntrospectionEnabled();     // src/ORBextractor.cc:1231
  const bool useMask = !maskLeft.empty() && (left->IntrospectionEnabled() || right->IntrospectionEnabled());
  left->SetQualityScoresAvailable(weighted);
  right->SetQualityScoresAvailable(!maskLeft.empty() && right->IntrospectionEnabled());
  const int cap = std::max(ivg_max_keypoints(left->handle()), ivg_max_keypoints(right->handle()));
  keysLeft.resize(cap); keysRight.resize(cap);
  cv::Mat dL(cap, 32, CV_8U), dR(cap, 32, CV_8U);
  std::vector<float> u(cap, -1.0f), d(cap, -1.0f);
  int nL = 0, nR = 0;
  const int rc = ivg_extract_stereo(left->handle(), right->handle(), imLeft.data, imRight.data, imLeft.cols, imLeft.rows, imLeft.step,
                                    useMask ? maskLeft.data : nullptr, useMask ? (size_t)maskLeft.step : 0,
                                    reinterpret_cast<ivg_keypoint*>(keysLeft.data()), dL.data, &nL,
                                    reinterpret_cast<ivg_keypoint*>(keysRight.data()), dR.data, &nR, mbf, maxD, u.data(), d.data(), cap);
  if (rc != IVG_OK) throw std::runtime_error(std::string("ivg_extract_stereo: ") + ivg_strerror(rc) + " " + ivg_last_cuda_error());
  keysLeft.resize(nL); keysRight.resize(nR);
  if (nL) dL.rowRange(0, nL).copyTo(descLeft); else descLeft.release();
  if (nR) dR.rowRange(0, nR).copyTo(descRight); else descRight.release();
  mvuRight.assign(u.begin(), u.begin() + nL);
  mvDepth.assign(d.begin(), d.begin() + nL);
}

}  // namespace ORB_SLAM2

/* shim/Frame_ComputeStereoMatches.cc — GPU body for Frame::ComputeStereoMatches
 * (introspective_ORB_SLAM/src/Frame.cc:758-932, declared include/Frame.h:163).
 *
 * In Frame.cc the maintainer replaces the body of Frame::ComputeStereoMatches() by
 *
 *     ORB_SLAM2::ComputeStereoMatchesGPU(mpORBextractorLeft, mpORBextractorRight, N, mbf, fx, mvuRight, mvDepth);
 *
 * (fx: the reference derives maxD = mbf/mb with mb still unassigned at that point).  The matcher runs on
 * what the two extractors left on the device — pyramids, keypoints, descriptors — so nothing is uploaded. */
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

}  // namespace ORB_SLAM2

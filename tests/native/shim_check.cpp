// Drives shim/ORBextractor.{h,cc} + shim/Frame_ComputeStereoMatches.cc (compiled against tests/fake_opencv) exactly the
// way Frame::Frame does in the reference (src/Frame.cc:115-125, :193): two extractor objects, operator() on each eye,
// then ComputeStereoMatches.  Built and called by tests/test_shim.py.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>
#include <vector>

#include "ORBextractor.h"

extern "C" int shim_stereo_frame(const unsigned char* left, const unsigned char* right, int w, int h, int nfeatures, int iniTh,
                                 int minTh, float mbf, float maxD, void* kpsL, unsigned char* descL, int* nL, float* uRight,
                                 float* depth, int* levels, float* scale1, unsigned char* pyr1, int* pyr1w, int* pyr1h) {
  try {
    ORB_SLAM2::ORBextractor exL(nfeatures, 1.2f, 8, iniTh, minTh), exR(nfeatures, 1.2f, 8, iniTh, minTh);
    cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)w), imR(h, w, CV_8UC1, (void*)right, (size_t)w), none;
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    exL(imL, none, kL, dL);
    exR(imR, none, kR, dR);
    std::vector<float> u, d;
    ORB_SLAM2::ComputeStereoMatchesGPU(&exL, &exR, (int)kL.size(), mbf, maxD, u, d);
    *nL = (int)kL.size();
    std::memcpy(kpsL, kL.data(), kL.size() * sizeof(cv::KeyPoint));
    for (int i = 0; i < dL.rows; ++i) std::memcpy(descL + 32 * i, dL.data + (size_t)i * dL.step, 32);
    std::memcpy(uRight, u.data(), u.size() * sizeof(float));
    std::memcpy(depth, d.data(), d.size() * sizeof(float));
    // the one-call front-end (ExtractStereoGPU) must return exactly what the two extractor calls + ComputeStereoMatchesGPU returned:
    // first call = explicit matcher launch (links the pair), second and third = the matcher queued behind the right eye's run
    for (int rep = 0; rep < 3; ++rep) {
      std::vector<cv::KeyPoint> k2, kR2;
      cv::Mat d2, dR2;
      std::vector<float> u2, dd2;
      ORB_SLAM2::ExtractStereoGPU(&exL, &exR, imL, imR, none, k2, d2, kR2, dR2, mbf, maxD, u2, dd2);
      if (k2.size() != kL.size() || kR2.size() != kR.size() || u2.size() != u.size()) return -2;
      if (std::memcmp(k2.data(), kL.data(), kL.size() * sizeof(cv::KeyPoint)) || std::memcmp(kR2.data(), kR.data(), kR.size() * sizeof(cv::KeyPoint))) return -3;
      for (int i = 0; i < dL.rows; ++i) if (std::memcmp(d2.data + (size_t)i * d2.step, dL.data + (size_t)i * dL.step, 32)) return -4;
      for (int i = 0; i < dR.rows; ++i) if (std::memcmp(dR2.data + (size_t)i * dR2.step, dR.data + (size_t)i * dR.step, 32)) return -4;
      if (std::memcmp(u2.data(), u.data(), u.size() * sizeof(float)) || std::memcmp(dd2.data(), d.data(), d.size() * sizeof(float))) return -5;
    }
    *levels = exL.GetLevels();
    *scale1 = exL.GetScaleFactors()[1];
    exL.SyncPyramidsToHost();
    *pyr1w = exL.mvImagePyramid[1].cols; *pyr1h = exL.mvImagePyramid[1].rows;
    for (int y = 0; y < *pyr1h; ++y) std::memcpy(pyr1 + (size_t)y * *pyr1w, exL.mvImagePyramid[1].data + (size_t)y * exL.mvImagePyramid[1].step, *pyr1w);
    return 0;
  } catch (const std::exception& e) {
    return -1;
  }
}

extern "C" int shim_construct_only() {
  try { ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7); return 0; }
  catch (const std::exception&) { return -1; }
}

// Drop-in latency: one stereo frame at a time exactly like the reference's Frame constructor — two std::threads run the
// two extractor objects (src/Frame.cc:115-125), join, then ComputeStereoMatches (:193).  Returns mean milliseconds.
#include "ivslam_gpu.h"
// graph bit 2: the whole frame through ExtractStereoGPU (one call from one thread) instead of two threads + ComputeStereoMatches;
// graph bit 0: CUDA-graph replay of the kernel sequence; bit 1: the caller keeps its two images in page-locked memory
// (cv::cuda::HostMem / cudaHostAlloc — a one-line change where the reference allocates imLeft/imRight) so the upload is a
// plain asynchronous DMA instead of the driver's pageable bounce path.
extern "C" double shim_frame_latency_ms(const unsigned char* left, const unsigned char* right, int w, int h, int nfeatures,
                                        int iniTh, int minTh, float mbf, float maxD, int iters, int graph) {
  void* pin[2] = {nullptr, nullptr};
  try {
    ORB_SLAM2::ORBextractor exL(nfeatures, 1.2f, 8, iniTh, minTh), exR(nfeatures, 1.2f, 8, iniTh, minTh);
    ivg_set_graph_mode(exL.handle(), graph & 1);
    ivg_set_graph_mode(exR.handle(), graph & 1);
    if (graph & 2) {
      if (ivg_host_alloc(&pin[0], (size_t)w * h) || ivg_host_alloc(&pin[1], (size_t)w * h)) return -1.0;
      std::memcpy(pin[0], left, (size_t)w * h); std::memcpy(pin[1], right, (size_t)w * h);
      left = (const unsigned char*)pin[0]; right = (const unsigned char*)pin[1];
    }
    cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)w), imR(h, w, CV_8UC1, (void*)right, (size_t)w), none;
    std::vector<cv::KeyPoint> kL, kR;
    cv::Mat dL, dR;
    std::vector<float> u, d;
    auto frame = [&] {
      if (graph & 4) { ORB_SLAM2::ExtractStereoGPU(&exL, &exR, imL, imR, none, kL, dL, kR, dR, mbf, maxD, u, d); return; }
      std::thread tl([&] { exL(imL, none, kL, dL); });
      std::thread tr([&] { exR(imR, none, kR, dR); });
      tl.join(); tr.join();
      ORB_SLAM2::ComputeStereoMatchesGPU(&exL, &exR, (int)kL.size(), mbf, maxD, u, d);
    };
    for (int i = 0; i < 5; ++i) frame();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) frame();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / iters;
    if (std::getenv("IVSLAM_LATENCY_BREAKDOWN")) {      // developer: where the frame time goes on the host side
      double tExt = 0, tSt = 0, tOne = 0;
      for (int i = 0; i < iters; ++i) {
        const auto a = std::chrono::steady_clock::now();
        std::thread tl([&] { exL(imL, none, kL, dL); });
        std::thread tr([&] { exR(imR, none, kR, dR); });
        tl.join(); tr.join();
        const auto b = std::chrono::steady_clock::now();
        ORB_SLAM2::ComputeStereoMatchesGPU(&exL, &exR, (int)kL.size(), mbf, maxD, u, d);
        const auto c = std::chrono::steady_clock::now();
        exL(imL, none, kL, dL);                           // one eye alone on the calling thread: no thread spawn, no contention
        const auto e = std::chrono::steady_clock::now();
        tExt += std::chrono::duration<double, std::micro>(b - a).count();
        tSt += std::chrono::duration<double, std::micro>(c - b).count();
        tOne += std::chrono::duration<double, std::micro>(e - c).count();
      }
      std::fprintf(stderr, "breakdown (us): two extractor threads spawn..join %.1f, ComputeStereoMatches %.1f, one operator() inline %.1f\n",
                   tExt / iters, tSt / iters, tOne / iters);
    }
    for (void* p : pin) if (p) ivg_host_free(p);
    return ms;
  } catch (const std::exception&) {
    return -1.0;
  }
}

// N4: SetRectifyMaps + ExtractRaw on a BGR frame; returns the keypoints/descriptors and the rectified gray level 0.
extern "C" int shim_extract_raw(const unsigned char* bgr, int sw, int sh, const float* mapx, const float* mapy, int w, int h, int nfeatures,
                                void* kps, unsigned char* desc, int* n, unsigned char* level0) {
  try {
    ORB_SLAM2::ORBextractor ex(nfeatures, 1.2f, 8, 20, 7);
    cv::Mat raw(sh, sw, CV_8UC3, (void*)bgr, (size_t)sw * 3), none;
    cv::Mat m1(h, w, CV_32FC1, (void*)mapx, (size_t)w * 4), m2(h, w, CV_32FC1, (void*)mapy, (size_t)w * 4);
    ex.SetRectifyMaps(m1, m2);
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    ex.ExtractRaw(raw, none, false, k, d);
    *n = (int)k.size();
    std::memcpy(kps, k.data(), k.size() * sizeof(cv::KeyPoint));
    for (int i = 0; i < d.rows; ++i) std::memcpy(desc + 32 * i, d.data + (size_t)i * d.step, 32);
    ex.SyncPyramidsToHost();
    for (int y = 0; y < h; ++y) std::memcpy(level0 + (size_t)y * w, ex.mvImagePyramid[0].data + (size_t)y * ex.mvImagePyramid[0].step, w);
    return 0;
  } catch (const std::exception& e) {
    return -1;
  }
}

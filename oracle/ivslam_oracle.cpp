// =====================================================================================
// ivslam_oracle.cpp — CPU restatement of IV-SLAM's stereo front-end.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the CPU baseline for
// bench.py.  Nothing under iv_slam_b200/ (the product) may include, link or call it;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs do.
//
// PARITY PIN: this restatement is checked against the reference itself.  oracle/_ref compiles the UNMODIFIED
// introspective_ORB_SLAM/src/ORBextractor.cc (whole file) and Frame::ComputeStereoMatches (src/Frame.cc:758-932, cut
// out verbatim) over an OpenCV-compat layer (oracle/refbuild/); tests/test_ref_pin.py requires bit-for-bit equality of
// everything this file produces with that build on the BASELINE configurations, fuzz geometries, cost-map extremes
// and the stereo stress case.  The pixel arithmetic of the un-vendored OpenCV (resize, GaussianBlur, FAST,
// fastAtan2 — the reference pins no OpenCV version, introspective_ORB_SLAM/CMakeLists.txt:37-46) is pinned to this
// image's opencv-python-headless 4.13.0 by tests/test_oracle_vs_cv2.py, byte for byte, and those same functions are
// what the compat layer of oracle/_ref forwards to.  tests/golden/ holds vectors generated from cv2 by
// tests/golden/make_golden.py.
//
// All file:line citations are relative to /root/reference/introspective_ORB_SLAM/.
// Build: see oracle/Makefile  (g++ -O3 -march=native -ffp-contract=off, the reference's
// own flags from CMakeLists.txt:16-17 plus contraction off = the canonical float
// semantics).
// =====================================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

// ------------------------------------------------------------------------------------
// cvRound: round-half-to-even (x86 cvtss2si / cvtsd2si under the default MXCSR mode).
// ------------------------------------------------------------------------------------
#if defined(__SSE2__)
inline int cv_round(float v) { return _mm_cvtss_si32(_mm_set_ss(v)); }     // exactly OpenCV's cvRound on SSE2
inline int cv_round(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }
#else
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
#endif

// ------------------------------------------------------------------------------------
// cv::resize(src, dst, INTER_LINEAR) for CV_8UC1 (call site src/ORBextractor.cc:1311,
// :1341).  OpenCV's 8-bit bilinear path: per-axis taps and 11-bit fixed-point
// coefficients are computed in float, rows are filtered horizontally into int32 and
// combined vertically with two >>16 multiplies.
// ------------------------------------------------------------------------------------
struct AxisTab {
  std::vector<int> ofs;        // first tap index
  std::vector<int> ofs1;       // second tap index (clamped)
  std::vector<short> c0, c1;   // coefficients, scale 2048
};

AxisTab make_axis(int S, int D) {
  AxisTab t;
  t.ofs.resize(D); t.ofs1.resize(D); t.c0.resize(D); t.c1.resize(D);
  const double inv_scale = (double)D / S;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < D; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= S - 1) { f = 0.f; s = S - 1; }
    t.ofs[d] = s;
    t.ofs1[d] = std::min(s + 1, S - 1);
    t.c0[d] = (short)cv_round((1.f - f) * 2048.f);
    t.c1[d] = (short)cv_round(f * 2048.f);
  }
  return t;
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                      uint8_t* dst, int dw, int dh, size_t dstride) {
  AxisTab X = make_axis(sw, dw), Y = make_axis(sh, dh);
  std::vector<int> row0(dw), row1(dw);
  int cached0 = -1, cached1 = -1;
  auto hfilter = [&](int sy, std::vector<int>& out) {
    const uint8_t* s = src + (size_t)sy * sstride;
    for (int d = 0; d < dw; ++d) out[d] = s[X.ofs[d]] * X.c0[d] + s[X.ofs1[d]] * X.c1[d];
  };
  for (int y = 0; y < dh; ++y) {
    int sy0 = Y.ofs[y], sy1 = Y.ofs1[y];
    if (cached1 == sy0) { std::swap(row0, row1); std::swap(cached0, cached1); }
    if (cached0 != sy0) { hfilter(sy0, row0); cached0 = sy0; }
    if (sy1 == sy0) { row1 = row0; cached1 = sy1; }
    else if (cached1 != sy1) { hfilter(sy1, row1); cached1 = sy1; }
    const int b0 = Y.c0[y], b1 = Y.c1[y];
    uint8_t* o = dst + (size_t)y * dstride;
    for (int d = 0; d < dw; ++d) {
      int v = (((b0 * (row0[d] >> 4)) >> 16) + ((b1 * (row1[d] >> 4)) >> 16) + 2) >> 2;
      o[d] = (uint8_t)v;
    }
  }
}

// ------------------------------------------------------------------------------------
// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1
// (call site src/ORBextractor.cc:1277).  OpenCV >= 3.4 fixed-point path: Q8 kernel
// {18,34,48,56,48,34,18}, 16-bit row sums, (v + 2^15) >> 16.
// ------------------------------------------------------------------------------------
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
  return p;
}

void gauss7_u8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  thread_local std::vector<uint16_t> tmp;
  thread_local std::vector<uint8_t> pad;
  tmp.resize((size_t)w * h);
  pad.resize((size_t)w + 6);
  for (int y = 0; y < h; ++y) {                       // horizontal pass on a reflect-101 padded copy of the row
    const uint8_t* s = src + (size_t)y * sstride;
    uint8_t* p = pad.data();
    for (int x = -3; x < 0; ++x) p[x + 3] = s[reflect101(x, w)];
    std::memcpy(p + 3, s, w);
    for (int x = w; x < w + 3; ++x) p[x + 3] = s[reflect101(x, w)];
    uint16_t* t = &tmp[(size_t)y * w];
    for (int x = 0; x < w; ++x)
      t[x] = (uint16_t)(18 * (p[x] + p[x + 6]) + 34 * (p[x + 1] + p[x + 5]) + 48 * (p[x + 2] + p[x + 4]) + 56 * p[x + 3]);
  }
  for (int y = 0; y < h; ++y) {
    const uint16_t* r[7];
    for (int k = 0; k < 7; ++k) r[k] = &tmp[(size_t)reflect101(y + k - 3, h) * w];
    uint8_t* o = dst + (size_t)y * dstride;
    const uint16_t *r0 = r[0], *r1 = r[1], *r2 = r[2], *r3 = r[3], *r4 = r[4], *r5 = r[5], *r6 = r[6];
    for (int x = 0; x < w; ++x) {
      const uint32_t acc = 18u * ((uint32_t)r0[x] + r6[x]) + 34u * ((uint32_t)r1[x] + r5[x]) + 48u * ((uint32_t)r2[x] + r4[x]) + 56u * r3[x];
      o[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
  (void)K;
}

// ------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, true) == FAST-9/16 + 3x3 non-max suppression
// (call sites src/ORBextractor.cc:1045,:1051).  SURVEY Appendix A.3.
// Output order is row-major; coordinates are relative to the window passed in.
// ------------------------------------------------------------------------------------
struct Corner { int x, y, score; };

const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Largest t such that the pixel is still a FAST-9 corner at threshold t
// (= max over the 16 arcs of 9 of the arc's minimum |difference| of one sign, minus 1).
inline int corner_score(const int d[16]) {
  int best = INT_MIN;
  for (int k = 0; k < 16; ++k) {
    int mn = INT_MAX, mx = INT_MIN;
    for (int i = 0; i < 9; ++i) { int v = d[(k + i) & 15]; mn = std::min(mn, v); mx = std::max(mx, v); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;
}

struct FastScratch { std::vector<uint8_t> score; std::vector<int> pos; };

void fast9_nms(const uint8_t* img, int w, int h, size_t stride, int th, bool nms,
               std::vector<Corner>& out, FastScratch& sc) {
  out.clear();
  if (w < 7 || h < 7) return;
  th = std::min(std::max(th, 0), 255);
  int off[16];
  for (int k = 0; k < 16; ++k) off[k] = RING_DY[k] * (int)stride + RING_DX[k];
  sc.score.assign((size_t)w * h, 0);
  sc.pos.clear();
  uint8_t* S = sc.score.data();
  auto test_pixel = [&](const uint8_t* p, int x, int y) {
    const int v = p[0], hi = v + th, lo = v - th;
    // every arc of 9 contains one pixel of each opposing pair (k, k+8): cheap rejection
    int br = 0, dk = 0;  // bit k set if ring[k] brighter / darker than the band
    bool alive_b = true, alive_d = true;
    for (int k = 0; k < 8 && (alive_b || alive_d); ++k) {
      int a = p[off[k]], b = p[off[k + 8]];
      if (a > hi) br |= 1 << k;
      if (b > hi) br |= 1 << (k + 8);
      if (a < lo) dk |= 1 << k;
      if (b < lo) dk |= 1 << (k + 8);
      alive_b = alive_b && (a > hi || b > hi);
      alive_d = alive_d && (a < lo || b < lo);
    }
    if (!alive_b && !alive_d) return;
    auto has_arc9 = [](int m) {
      unsigned r = (unsigned)m | ((unsigned)m << 16);
      unsigned t = r & (r >> 1);
      t &= t >> 2;
      t &= t >> 4;
      t &= r >> 8;
      return t != 0;
    };
    if (!((alive_b && has_arc9(br)) || (alive_d && has_arc9(dk)))) return;
    int d[16];
    for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
    int s = corner_score(d);
    if (!nms) out.push_back({x, y, s});
    else { S[(size_t)y * w + x] = (uint8_t)s; sc.pos.push_back(y * w + x); }
  };
  for (int y = 3; y < h - 3; ++y) {
    const uint8_t* row = img + (size_t)y * stride;
    int x = 3;
#if defined(__SSE2__)
    // 16 pixels at a time (the reference's OpenCV FAST is SIMD too): reject blocks where no pixel passes the
    // compass test on diameters 0-8 and 4-12, run the scalar test only on the pixels that do.
    const __m128i delta = _mm_set1_epi8((char)-128), tv = _mm_set1_epi8((char)th);
    for (; x + 16 <= w - 3; x += 16) {
      const uint8_t* p = row + x;
      const __m128i c = _mm_loadu_si128((const __m128i*)p);
      const __m128i v0 = _mm_xor_si128(_mm_adds_epu8(c, tv), delta), v1 = _mm_xor_si128(_mm_subs_epu8(c, tv), delta);
      auto ld = [&](int k) { return _mm_xor_si128(_mm_loadu_si128((const __m128i*)(p + off[k])), delta); };
      const __m128i x0 = ld(0), x8 = ld(8), x4 = ld(4), x12 = ld(12);
      __m128i mb = _mm_and_si128(_mm_or_si128(_mm_cmpgt_epi8(x0, v0), _mm_cmpgt_epi8(x8, v0)),
                                 _mm_or_si128(_mm_cmpgt_epi8(x4, v0), _mm_cmpgt_epi8(x12, v0)));
      __m128i md = _mm_and_si128(_mm_or_si128(_mm_cmpgt_epi8(v1, x0), _mm_cmpgt_epi8(v1, x8)),
                                 _mm_or_si128(_mm_cmpgt_epi8(v1, x4), _mm_cmpgt_epi8(v1, x12)));
      int m = _mm_movemask_epi8(_mm_or_si128(mb, md));
      if (m) {   // two more diameters before falling back to the scalar test
        const __m128i x2 = ld(2), x10 = ld(10), x6 = ld(6), x14 = ld(14);
        mb = _mm_and_si128(mb, _mm_and_si128(_mm_or_si128(_mm_cmpgt_epi8(x2, v0), _mm_cmpgt_epi8(x10, v0)),
                                             _mm_or_si128(_mm_cmpgt_epi8(x6, v0), _mm_cmpgt_epi8(x14, v0))));
        md = _mm_and_si128(md, _mm_and_si128(_mm_or_si128(_mm_cmpgt_epi8(v1, x2), _mm_cmpgt_epi8(v1, x10)),
                                             _mm_or_si128(_mm_cmpgt_epi8(v1, x6), _mm_cmpgt_epi8(v1, x14))));
        m = _mm_movemask_epi8(_mm_or_si128(mb, md));
      }
      while (m) {
        const int i = __builtin_ctz(m);
        m &= m - 1;
        test_pixel(p + i, x + i, y);
      }
    }
#endif
    for (; x < w - 3; ++x) test_pixel(row + x, x, y);
  }
  if (!nms) return;
  for (int idx : sc.pos) {     // detection order is row-major, so is the emission order
    const uint8_t* r = S + idx;
    const int s = r[0];
    if (s == 0) continue;      // stored 0 == not a corner at th; a corner scoring 0 never beats its neighbours
    if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] && s > r[w - 1] && s > r[w] && s > r[w + 1])
      out.push_back({idx % w, idx / w, s});
  }
}

// ------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) in degrees (call site src/ORBextractor.cc:104); SURVEY A.4.
// Plain float multiply/add chain (file is compiled with -ffp-contract=off).
// ------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
  const float sc = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
  const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + 2.2204460492503131e-16f);  // (float)DBL_EPSILON
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + 2.2204460492503131e-16f);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ------------------------------------------------------------------------------------
// KeyPoint with cv::KeyPoint's memory layout (28 bytes).
// ------------------------------------------------------------------------------------
struct KeyPoint {
  float x, y, size, angle, response;
  int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// cv::KeyPointsFilter::retainBest (OpenCV features2d/src/keypoint.cpp) followed by the
// caller's vector::resize(n) (src/ORBextractor.cc:1146-1148, :1164-1165): when the list is
// longer than n the survivors are the first n elements after the real libstdc++
// std::nth_element with the response-greater comparator.
void retain_best_resize(std::vector<KeyPoint>& k, int n) {
  if (n >= 0 && k.size() > (size_t)n) {
    if (n == 0) { k.clear(); return; }
    std::nth_element(k.begin(), k.begin() + n - 1, k.end(),
                     [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    k.resize(n);
  }
}

const int kBriefPattern[256 * 4] = {
#include "../include/ivslam_brief_pattern.inc"
};

const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;


// ------------------------------------------------------------------------------------
// ORBextractor (include/ORBextractor.h:54-128, src/ORBextractor.cc:411-476).
// ------------------------------------------------------------------------------------
struct Level {
  int w = 0, h = 0;
  std::vector<uint8_t> img, blur, qual;
};

struct Extractor {
  int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;
  bool enableIntrospection;
  bool qualityAvailable = false;
  int kp_mode = 0;     // 0: ComputeKeyPointsOld (live in the reference), 1: ComputeKeyPointsOctTree (compiled but dead there)
  int trig_mode = 0;   // 0: glibc cosf/sinf (the reference, src/ORBextractor.cc:114); 1: (float)cos((double)x)
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nPerLevel, umax;
  std::vector<Level> lv;
  std::vector<std::vector<KeyPoint>> levelKeys;   // per level, level coordinates, after orientation
  FastScratch fs;
  long n_fast_calls = 0, n_raw = 0;
  double t_stage[6] = {0, 0, 0, 0, 0, 0};   // seconds: pyramid, FAST, selection, orientation, blur, descriptors

  Extractor(int nf, float sf, int nl, int ini, int mn, bool intro)
      : nfeatures(nf), scaleFactor(sf), nlevels(nl), iniThFAST(ini), minThFAST(mn), enableIntrospection(intro) {
    // src/ORBextractor.cc:417-432: float tables, scaleFactor member is double
    scale.resize(nl); sigma2.resize(nl); invScale.resize(nl); invSigma2.resize(nl);
    scale[0] = 1.f; sigma2[0] = 1.f;
    for (int i = 1; i < nl; ++i) { scale[i] = (float)(scale[i - 1] * scaleFactor); sigma2[i] = scale[i] * scale[i]; }
    for (int i = 0; i < nl; ++i) { invScale[i] = 1.f / scale[i]; invSigma2[i] = 1.f / sigma2[i]; }
    // :437-452 features per level
    nPerLevel.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) { nPerLevel[l] = cv_round(nDesired); sum += nPerLevel[l]; nDesired *= factor; }
    nPerLevel[nl - 1] = std::max(nfeatures - sum, 0);
    // :458-475 circular patch row extents
    umax.assign(HALF_PATCH_SIZE + 1, 0);
    int vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0; ++v0;
    }
    lv.resize(nl);
    levelKeys.resize(nl);
  }
};

// ComputePyramid / ComputeQualityImagePyramid (src/ORBextractor.cc:1298-1357): each level is
// resized from the previous one.  The 19-px reflected frame the reference adds around every
// level is never read on this path and is not materialised.
void compute_pyramid(Extractor& e, const uint8_t* img, int w, int h, size_t stride, bool quality) {
  for (int l = 0; l < e.nlevels; ++l) {
    float sc = e.invScale[l];
    int lw = cv_round((float)w * sc), lh = cv_round((float)h * sc);
    Level& L = e.lv[l];
    L.w = lw; L.h = lh;
    std::vector<uint8_t>& dst = quality ? L.qual : L.img;
    dst.resize((size_t)lw * lh);
    if (l == 0) {
      for (int y = 0; y < h; ++y) std::memcpy(&dst[(size_t)y * w], img + (size_t)y * stride, w);
    } else {
      const Level& P = e.lv[l - 1];
      const std::vector<uint8_t>& src = quality ? P.qual : P.img;
      resize_linear_u8(src.data(), P.w, P.h, P.w, dst.data(), lw, lh, lw);
    }
  }
}

// Grid geometry of one level (src/ORBextractor.cc:884-907).  Returns false for geometries the
// reference itself cannot run (zero grid => division by zero) or where a non-final cell
// window would leave the search area (only possible for tiny levels); both the oracle and
// the CUDA path refuse these identically.
struct Grid { int cols, rows, cellW, cellH, W, H, maxBX, maxBY, nCells; };
bool level_grid(const Extractor& e, int level, Grid& g) {
  float imageRatio = (float)e.lv[0].w / e.lv[0].h;
  int n = e.nPerLevel[level];
  g.cols = (int)std::sqrt((float)n / (5 * imageRatio));
  g.rows = (int)(imageRatio * g.cols);
  g.maxBX = e.lv[level].w - EDGE_THRESHOLD;
  g.maxBY = e.lv[level].h - EDGE_THRESHOLD;
  g.W = g.maxBX - EDGE_THRESHOLD;
  g.H = g.maxBY - EDGE_THRESHOLD;
  if (g.cols < 1 || g.rows < 1 || g.W < 1 || g.H < 1) return false;
  g.cellW = (int)std::ceil((float)g.W / g.cols);
  g.cellH = (int)std::ceil((float)g.H / g.rows);
  g.nCells = g.rows * g.cols;
  if ((g.cols - 1) * g.cellW > g.W || (g.rows - 1) * g.cellH > g.H) return false;
  return true;
}

// IC_Angle (src/ORBextractor.cc:78-105)
float ic_angle(const Level& L, float px, float py, const std::vector<int>& umax) {
  int m01 = 0, m10 = 0;
  const int step = L.w;
  const uint8_t* c = &L.img[(size_t)cv_round(py) * step + cv_round(px)];
  for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m10 += u * c[u];
  for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
    int vsum = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int a = c[u + v * step], b = c[u - v * step];
      vsum += a - b;
      m10 += u * (a + b);
    }
    m01 += v * vsum;
  }
  return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor (src/ORBextractor.cc:108-148) on the blurred level.
void orb_descriptor(const Extractor& e, const Level& L, const KeyPoint& kp, uint8_t* desc) {
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  float angle = kp.angle * factorPI;
  float a, b;
  if (e.trig_mode == 0) { a = cosf(angle); b = sinf(angle); }
  else { a = (float)std::cos((double)angle); b = (float)std::sin((double)angle); }
  const int step = L.w;
  const uint8_t* c = &L.blur[(size_t)cv_round(kp.y) * step + cv_round(kp.x)];
  const int* pat = kBriefPattern;
  for (int i = 0; i < 32; ++i) {
    int val = 0;
    for (int k = 0; k < 8; ++k, pat += 4) {
      float x0 = (float)pat[0], y0 = (float)pat[1], x1 = (float)pat[2], y1 = (float)pat[3];
      int t0 = c[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
      int t1 = c[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
      val |= (t0 < t1) << k;
    }
    desc[i] = (uint8_t)val;
  }
}

// ComputeKeyPointsOld (src/ORBextractor.cc:880-1213) — the live selection path
// (operator() calls it at :1248; the OctTree variant at :1247 is commented out).
int compute_keypoints_old(Extractor& e) {
  std::vector<Corner> corners;
  for (int level = 0; level < e.nlevels; ++level) {
    Level& L = e.lv[level];
    std::vector<KeyPoint>& keypoints = e.levelKeys[level];
    keypoints.clear();
    Grid g;
    if (!level_grid(e, level, g)) return -2;
    const int nDesired = e.nPerLevel[level];
    const int levelRows = g.rows, levelCols = g.cols, cellW = g.cellW, cellH = g.cellH;
    const int minBX = EDGE_THRESHOLD, minBY = EDGE_THRESHOLD, maxBX = g.maxBX, maxBY = g.maxBY;
    const int nCells = g.nCells;
    int nfeaturesCell = (int)std::ceil((float)nDesired / nCells);
    const bool weighted = e.qualityAvailable && e.enableIntrospection;

    std::vector<std::vector<KeyPoint>> cellKP((size_t)nCells);
    std::vector<int> nToRetain(nCells, 0), nTotal(nCells, 0);
    std::vector<char> bNoMore(nCells, 0);
    std::vector<int> iniXCol(levelCols), iniYRow(levelRows);
    int nNoMore = 0, nToDistribute = 0;

    float hY = (float)(cellH + 6);   // :935 — declared once; stale across loops
    std::vector<float> nfeatures_cell(nCells, (float)nfeaturesCell), cell_weights(nCells, 0.f);
    float cell_weights_sum = 0.0f;
    if (weighted) {   // :942-987
      for (int i = 0; i < levelRows; ++i) {
        const float iniY = (float)(minBY + i * cellH - 3);
        iniYRow[i] = (int)iniY;
        if (i == levelRows - 1) { hY = maxBY + 3 - iniY; if (hY <= 0) continue; }
        float hX = (float)(cellW + 6);
        for (int j = 0; j < levelCols; ++j) {
          float iniX;
          if (i == 0) { iniX = (float)(minBX + j * cellW - 3); iniXCol[j] = (int)iniX; }
          else iniX = (float)iniXCol[j];
          if (j == levelCols - 1) { hX = maxBX + 3 - iniX; if (hX <= 0) continue; }
          unsigned long sum = 0;
          for (int y = (int)iniY; y < (int)(iniY + hY); ++y)
            for (int x = (int)iniX; x < (int)(iniX + hX); ++x) sum += L.qual[(size_t)y * L.w + x];
          float cost = static_cast<float>(sum) / static_cast<float>(hX * hY);
          float qual_score = (float)(1.0 / (1.0 + cost / 255));
          float qual_score_norm = 2 * qual_score - 1;
          cell_weights[i * levelCols + j] = qual_score_norm;
          cell_weights_sum += qual_score_norm;
        }
      }
    }

    for (int i = 0; i < levelRows; ++i) {   // :989-1099
      const float iniY = (float)(minBY + i * cellH - 3);
      iniYRow[i] = (int)iniY;
      if (i == levelRows - 1) { hY = maxBY + 3 - iniY; if (hY <= 0) continue; }
      float hX = (float)(cellW + 6);
      for (int j = 0; j < levelCols; ++j) {
        float iniX;
        if (i == 0) { iniX = (float)(minBX + j * cellW - 3); iniXCol[j] = (int)iniX; }
        else iniX = (float)iniXCol[j];
        if (j == levelCols - 1) { hX = maxBX + 3 - iniX; if (hX <= 0) continue; }
        const int c = i * levelCols + j;
        if (weighted)
          nfeatures_cell[c] = std::max(1.0f, std::ceil((float)nDesired * cell_weights[c] / cell_weights_sum));

        const int x0 = (int)iniX, y0 = (int)iniY, ww = (int)(iniX + hX) - x0, wh = (int)(iniY + hY) - y0;
        if (x0 < 0 || y0 < 0 || x0 + ww > L.w || y0 + wh > L.h) return -2;
        const uint8_t* win = &L.img[(size_t)y0 * L.w + x0];
        const auto tf0 = std::chrono::steady_clock::now();
        fast9_nms(win, ww, wh, L.w, e.iniThFAST, true, corners, e.fs); e.n_fast_calls++;
        if (corners.size() <= 3) { fast9_nms(win, ww, wh, L.w, e.minThFAST, true, corners, e.fs); e.n_fast_calls++; }
        e.t_stage[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - tf0).count();
        std::vector<KeyPoint>& kc = cellKP[c];
        kc.resize(corners.size());
        for (size_t k = 0; k < corners.size(); ++k)
          kc[k] = KeyPoint{(float)corners[k].x, (float)corners[k].y, 7.f, -1.f, (float)corners[k].score, 0, -1};
        e.n_raw += (long)corners.size();
        if (weighted) {   // :1058-1080
          for (size_t k = 0; k < kc.size(); ++k) {
            float cost = static_cast<float>(L.qual[(size_t)(y0 + (int)kc[k].y) * L.w + (x0 + (int)kc[k].x)]);
            kc[k].response *= 2 * (1.0f / (1.0f + cost / 255.0f)) - 1;
          }
        }
        const int nKeys = (int)kc.size();
        nTotal[c] = nKeys;
        if (nKeys > nfeatures_cell[c]) { nToRetain[c] = (int)nfeatures_cell[c]; bNoMore[c] = 0; }
        else {
          nToRetain[c] = nKeys;
          nToDistribute = (int)(nToDistribute + (nfeatures_cell[c] - nKeys));
          bNoMore[c] = 1; nNoMore++;
        }
      }
    }

    while (nToDistribute > 0 && nNoMore < nCells) {   // :1103-1133 (runs once)
      int nNew = 0;
      for (int c = 0; c < nCells; ++c) {
        if (!bNoMore[c]) {
          nNew = (int)(nfeatures_cell[c] + std::ceil((float)nToDistribute / (nCells - nNoMore)));
          if (nTotal[c] > nNew) { nToRetain[c] = nNew; bNoMore[c] = 0; }
          else { nToRetain[c] = nTotal[c]; nToDistribute += nNew - nTotal[c]; bNoMore[c] = 1; nNoMore++; }
        }
      }
      nToDistribute = 0;
    }

    const int scaledPatchSize = (int)(PATCH_SIZE * e.scale[level]);
    for (int i = 0; i < levelRows; ++i)   // :1141-1160
      for (int j = 0; j < levelCols; ++j) {
        std::vector<KeyPoint>& kc = cellKP[i * levelCols + j];
        retain_best_resize(kc, nToRetain[i * levelCols + j]);
        for (size_t k = 0; k < kc.size(); ++k) {
          kc[k].x += iniXCol[j]; kc[k].y += iniYRow[i];
          kc[k].octave = level; kc[k].size = (float)scaledPatchSize;
          keypoints.push_back(kc[k]);
        }
      }
    if ((int)keypoints.size() > nDesired) retain_best_resize(keypoints, nDesired);   // :1162-1166
  }
  const auto ta0 = std::chrono::steady_clock::now();
  for (int level = 0; level < e.nlevels; ++level)   // :1208-1210
    for (KeyPoint& kp : e.levelKeys[level]) kp.angle = ic_angle(e.lv[level], kp.x, kp.y, e.umax);
  e.t_stage[3] += std::chrono::duration<double>(std::chrono::steady_clock::now() - ta0).count();
  return 0;
}

// ------------------------------------------------------------------------------------
// ComputeKeyPointsOctTree + DistributeOctTree + ExtractorNode::DivideNode
// (src/ORBextractor.cc:771-878, :545-769, :487-543).  This path is compiled but DEAD in the reference (operator() calls
// ComputeKeyPointsOld, :1247-1248); it is provided as an optional mode because the north star names it.
// One deliberate difference: the reference sorts (size, ExtractorNode*) pairs (:690), so ties between nodes of equal
// size are broken by heap addresses — nondeterministic across runs.  Here a tie is broken by creation
// order (the node created later counts as the larger one), which makes the result a pure function of the image.
// The per-cell quality score the reference parks in KeyPoint::size (:826-849) is overwritten before anyone reads it
// (:861) and is not computed.
// ------------------------------------------------------------------------------------
struct OctNode {
  int ulx, uly, brx, bry;          // UL and BR corners (UR.x == BR.x, BL.y == BR.y)
  std::vector<int> keys;           // indices into the level's key list, in insertion order
  bool noMore = false;
  long seq = 0;                    // creation order (stands in for the pointer in the reference's sort)
};

void divide_node(const OctNode& n, const std::vector<KeyPoint>& K, OctNode out[4]) {
  const int halfX = (int)std::ceil(static_cast<float>(n.brx - n.ulx) / 2);
  const int halfY = (int)std::ceil(static_cast<float>(n.bry - n.uly) / 2);
  const int mx = n.ulx + halfX, my = n.uly + halfY;
  out[0].ulx = n.ulx; out[0].uly = n.uly; out[0].brx = mx;    out[0].bry = my;
  out[1].ulx = mx;    out[1].uly = n.uly; out[1].brx = n.brx; out[1].bry = my;
  out[2].ulx = n.ulx; out[2].uly = my;    out[2].brx = mx;    out[2].bry = n.bry;
  out[3].ulx = mx;    out[3].uly = my;    out[3].brx = n.brx; out[3].bry = n.bry;
  for (int k : n.keys) {
    const KeyPoint& kp = K[k];
    if (kp.x < mx) { if (kp.y < my) out[0].keys.push_back(k); else out[2].keys.push_back(k); }
    else if (kp.y < my) out[1].keys.push_back(k);
    else out[3].keys.push_back(k);
  }
  for (int c = 0; c < 4; ++c) out[c].noMore = out[c].keys.size() == 1;
}

std::vector<KeyPoint> distribute_octree(const std::vector<KeyPoint>& K, int minX, int maxX, int minY, int maxY, int N) {
  std::vector<KeyPoint> result;
  const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
  if (nIni < 1) return result;     // the reference divides by zero here (taller-than-wide levels); defined as "no keypoints"
  const float hX = static_cast<float>(maxX - minX) / nIni;
  std::list<OctNode> nodes;
  long seq = 0;
  std::vector<OctNode*> ini(nIni);
  for (int i = 0; i < nIni; ++i) {
    OctNode n;
    n.ulx = (int)(hX * static_cast<float>(i)); n.uly = 0;
    n.brx = (int)(hX * static_cast<float>(i + 1)); n.bry = maxY - minY;
    n.seq = seq++;
    nodes.push_back(n);
    ini[i] = &nodes.back();
  }
  for (size_t i = 0; i < K.size(); ++i) {
    const int b = (int)(K[i].x / hX);
    if (b < 0 || b >= nIni) continue;   // cannot happen for coordinates inside the level
    ini[b]->keys.push_back((int)i);
  }
  for (auto it = nodes.begin(); it != nodes.end();) {
    if (it->keys.size() == 1) { it->noMore = true; ++it; }
    else if (it->keys.empty()) it = nodes.erase(it);
    else ++it;
  }
  typedef std::list<OctNode>::iterator It;
  std::vector<std::pair<int, It>> expand;   // (size, node); order key = (size, seq)
  auto add_children = [&](OctNode ch[4], int* nToExpand) {
    for (int c = 0; c < 4; ++c)
      if (!ch[c].keys.empty()) {
        ch[c].seq = seq++;
        nodes.push_front(ch[c]);
        if (ch[c].keys.size() > 1) { if (nToExpand) ++*nToExpand; expand.push_back(std::make_pair((int)ch[c].keys.size(), nodes.begin())); }
      }
  };
  bool finish = false;
  while (!finish) {
    const int prevSize = (int)nodes.size();
    int nToExpand = 0;
    expand.clear();
    for (auto it = nodes.begin(); it != nodes.end();) {
      if (it->noMore) { ++it; continue; }
      OctNode ch[4];
      divide_node(*it, K, ch);
      add_children(ch, &nToExpand);
      it = nodes.erase(it);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
    else if ((int)nodes.size() + nToExpand * 3 > N) {
      while (!finish) {
        const int prev2 = (int)nodes.size();
        std::vector<std::pair<int, It>> prev = expand;
        expand.clear();
        std::sort(prev.begin(), prev.end(), [](const std::pair<int, It>& a, const std::pair<int, It>& b) {
          return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq;
        });
        for (int j = (int)prev.size() - 1; j >= 0; --j) {
          OctNode ch[4];
          divide_node(*prev[j].second, K, ch);
          add_children(ch, nullptr);
          nodes.erase(prev[j].second);
          if ((int)nodes.size() >= N) break;
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev2) finish = true;
      }
    }
  }
  result.reserve(nodes.size());
  for (const OctNode& n : nodes) {
    int best = n.keys[0];
    float mr = K[best].response;
    for (size_t k = 1; k < n.keys.size(); ++k)
      if (K[n.keys[k]].response > mr) { best = n.keys[k]; mr = K[best].response; }
    result.push_back(K[best]);
  }
  return result;
}

int compute_keypoints_octree(Extractor& e) {
  std::vector<Corner> corners;
  const float Wc = 30;
  for (int level = 0; level < e.nlevels; ++level) {
    Level& L = e.lv[level];
    const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
    const int maxBX = L.w - EDGE_THRESHOLD + 3, maxBY = L.h - EDGE_THRESHOLD + 3;
    std::vector<KeyPoint> toDistribute;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / Wc), nRows = (int)(height / Wc);
    if (nCols < 1 || nRows < 1) return -2;      // the reference divides by zero
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    for (int i = 0; i < nRows; ++i) {
      const float iniY = (float)(minBY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBY - 3) continue;
      if (maxY > maxBY) maxY = (float)maxBY;
      for (int j = 0; j < nCols; ++j) {
        const float iniX = (float)(minBX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBX - 6) continue;
        if (maxX > maxBX) maxX = (float)maxBX;
        const int x0 = (int)iniX, y0 = (int)iniY, ww = (int)maxX - x0, wh = (int)maxY - y0;
        const uint8_t* win = &L.img[(size_t)y0 * L.w + x0];
        fast9_nms(win, ww, wh, L.w, e.iniThFAST, true, corners, e.fs); e.n_fast_calls++;
        if (corners.empty()) { fast9_nms(win, ww, wh, L.w, e.minThFAST, true, corners, e.fs); e.n_fast_calls++; }
        for (const Corner& c : corners)
          toDistribute.push_back(KeyPoint{(float)c.x + j * wCell, (float)c.y + i * hCell, 7.f, -1.f, (float)c.score, 0, -1});
        e.n_raw += (long)corners.size();
      }
    }
    std::vector<KeyPoint>& keypoints = e.levelKeys[level];
    keypoints = distribute_octree(toDistribute, minBX, maxBX, minBY, maxBY, e.nPerLevel[level]);
    const int scaledPatchSize = (int)(PATCH_SIZE * e.scale[level]);
    for (KeyPoint& kp : keypoints) { kp.x += minBX; kp.y += minBY; kp.octave = level; kp.size = (float)scaledPatchSize; }
  }
  for (int level = 0; level < e.nlevels; ++level)
    for (KeyPoint& kp : e.levelKeys[level]) kp.angle = ic_angle(e.lv[level], kp.x, kp.y, e.umax);
  return 0;
}

// ORBextractor::operator() (src/ORBextractor.cc:1224-1296)
int extract(Extractor& e, const uint8_t* img, int w, int h, size_t stride, const uint8_t* cost,
            size_t cost_stride, KeyPoint* kps, uint8_t* desc, int cap, int* n_out) {
  *n_out = 0;
  if (!img || w <= 0 || h <= 0) return 0;   // :1227-1228 empty image => silent return
  if (cost && e.enableIntrospection) { e.qualityAvailable = true; compute_pyramid(e, cost, w, h, cost_stride, true); }
  else e.qualityAvailable = false;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  const auto t0 = now();
  compute_pyramid(e, img, w, h, stride, false);
  const auto t1 = now();
  const double fast_before = e.t_stage[1], ang_before = e.t_stage[3];
  int rc = e.kp_mode == 1 ? compute_keypoints_octree(e) : compute_keypoints_old(e);
  if (rc) return rc;
  const auto t2 = now();
  e.t_stage[0] += secs(t0, t1);
  e.t_stage[2] += secs(t1, t2) - (e.t_stage[1] - fast_before) - (e.t_stage[3] - ang_before);
  int n = 0;
  for (int l = 0; l < e.nlevels; ++l) n += (int)e.levelKeys[l].size();
  if (n > cap) return -3;
  int offset = 0;
  for (int l = 0; l < e.nlevels; ++l) {
    std::vector<KeyPoint>& keys = e.levelKeys[l];
    Level& L = e.lv[l];
    if (keys.empty()) { L.blur.clear(); continue; }
    L.blur.resize((size_t)L.w * L.h);
    const auto tb0 = now();
    gauss7_u8(L.img.data(), L.w, L.h, L.w, L.blur.data(), L.w);
    const auto tb1 = now();
    e.t_stage[4] += secs(tb0, tb1);
    for (size_t i = 0; i < keys.size(); ++i) {
      orb_descriptor(e, L, keys[i], desc + (size_t)(offset + i) * 32);
      KeyPoint k = keys[i];
      if (l != 0) { float sc = e.scale[l]; k.x *= sc; k.y *= sc; }
      kps[offset + i] = k;
    }
    offset += (int)keys.size();
    e.t_stage[5] += secs(tb1, now());
  }
  *n_out = n;
  return 0;
}

// ------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches (src/Frame.cc:758-932); DescriptorDistance
// (src/ORBmatcher.cc:1700-1716), TH_HIGH / TH_LOW (src/ORBmatcher.cc:37-38).
// maxD is an argument: the reference reads mbf/mb with mb not yet assigned.
// An empty match list is a no-op instead of the reference's out-of-range read (Q8).
// ------------------------------------------------------------------------------------
const int TH_HIGH = 100, TH_LOW = 50;

inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y; std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
    unsigned v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

int stereo_match(const Extractor& eL, const Extractor& eR, const KeyPoint* kL, int N, const uint8_t* dL,
                 const KeyPoint* kR, int Nr, const uint8_t* dR, float mbf, float maxD,
                 float* uRight, float* depth, int* bestDistOut, int* sadOut) {
  for (int i = 0; i < N; ++i) { uRight[i] = -1.f; depth[i] = -1.f; if (bestDistOut) bestDistOut[i] = -1; if (sadOut) sadOut[i] = -1; }
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = eL.lv[0].h;
  std::vector<std::vector<size_t>> vRowIndices(nRows);
  for (int iR = 0; iR < Nr; ++iR) {
    const float kpY = kR[iR].y;
    const float r = 2.0f * eL.scale[kR[iR].octave];
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; ++yi)
      if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);   // reference indexes unchecked
  }
  const float minD = 0;
  std::vector<std::pair<int, int>> vDistIdx;
  vDistIdx.reserve(N);
  for (int iL = 0; iL < N; ++iL) {
    const KeyPoint& kpL = kL[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<size_t>& cand = vRowIndices[row];
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH; size_t bestIdxR = 0;
    for (size_t iC = 0; iC < cand.size(); ++iC) {
      const size_t iR = cand[iC];
      const KeyPoint& kpR = kR[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = descriptor_distance(dL + 32 * (size_t)iL, dR + 32 * iR);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (bestDistOut) bestDistOut[iL] = bestDist;
    if (bestDist < thOrbDist) {
      const float uR0 = kR[bestIdxR].x;
      const float sf = eL.invScale[kpL.octave];
      const float scaleduL = std::round(kpL.x * sf), scaledvL = std::round(kpL.y * sf), scaleduR0 = std::round(uR0 * sf);
      const int w = 5, L = 5;
      const Level& PL = eL.lv[kpL.octave];
      const Level& PR = eR.lv[kpL.octave];
      const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)scaleduR0;
      const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
      if (iniu < 0 || endu >= PR.w) continue;
      // windows the reference would take with rowRange/colRange; outside the level it throws
      if (cy - w < 0 || cy + w + 1 > PL.h || cxL - w < 0 || cxL + w + 1 > PL.w || cxR - L - w < 0) continue;
      const int cL = PL.img[(size_t)cy * PL.w + cxL];
      int bestSAD = INT_MAX, bestincR = 0;
      float vDists[2 * 5 + 1];
      for (int incR = -L; incR <= L; ++incR) {
        const int cR = PR.img[(size_t)cy * PR.w + cxR + incR];
        float dist = 0;   // exact: integer-valued float sums < 2^24
        for (int dy = -w; dy <= w; ++dy)
          for (int dx = -w; dx <= w; ++dx) {
            float a = (float)PL.img[(size_t)(cy + dy) * PL.w + cxL + dx] - (float)cL;
            float b = (float)PR.img[(size_t)(cy + dy) * PR.w + cxR + incR + dx] - (float)cR;
            dist += std::fabs(a - b);
          }
        if (dist < bestSAD) { bestSAD = (int)dist; bestincR = incR; }
        vDists[L + incR] = dist;
      }
      if (bestincR == -L || bestincR == L) continue;
      const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
      const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
      if (deltaR < -1 || deltaR > 1) continue;
      float bestuR = eL.scale[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
      float disparity = (uL - bestuR);
      if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
        depth[iL] = mbf / disparity;
        uRight[iL] = bestuR;
        if (sadOut) sadOut[iL] = bestSAD;
        vDistIdx.push_back(std::pair<int, int>(bestSAD, iL));
      }
    }
  }
  if (vDistIdx.empty()) return 0;
  std::sort(vDistIdx.begin(), vDistIdx.end());
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
    if (vDistIdx[i].first < thDist) break;
    uRight[vDistIdx[i].second] = -1; depth[vDistIdx[i].second] = -1;
  }
  return 0;
}

}  // namespace

// =====================================================================================
// C interface for ctypes (tests / bench only).
// =====================================================================================
// ------------------------------------------------------------------------------------------------------------------
// N4 input prologue: cv::remap(INTER_LINEAR, CV_32FC1 maps, BORDER_CONSTANT 0) and cvtColor(*2GRAY).
// Call sites in the reference: Examples/Stereo/stereo_kitti.cc:463-464,520, stereo_euroc.cc:369-370,397 (remap);
// src/Tracking.cc:278-294 (cvtColor).  The arithmetic is OpenCV's (un-vendored; pinned to the 4.13.0 wheel of this image):
//   imgproc/src/imgwarp.cpp  RemapInvoker (sx = cvRound(mapx*INTER_TAB_SIZE), INTER_BITS = 5), initInterTab2D
//     (INTER_REMAP_COEF_BITS = 15; the (0,0) entry is saturate_cast<short>(32768) = 32767 and its correction lands on an
//     element that is recomputed afterwards), remapBilinear<FixedPtCast<int, uchar, 15>> (taps outside the source = 0);
//   imgproc/src/color_rgb.simd.hpp  RGB2Gray<uchar>: (B*3735 + G*19235 + R*9798 + 2^14) >> 15.
// Checked bit-for-bit against cv2.remap / cv2.cvtColor in tests/test_oracle_vs_cv2.py.
static inline int cv_round_x32(float m) {
  const float v = m * 32.0f;
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return (int)0x80000000;   // cvtss2si "integer indefinite"
  return (int)lrintf(v);                                                     // round half to even (default rounding mode)
}

static void prologue_frame(const uint8_t* src, int sw, int sh, size_t sstride, int cn, int rgb, const float* mapx, const float* mapy,
                           size_t mstride, uint8_t* dst, int W, int H, size_t dstride) {
  const int nc = cn == 1 ? 1 : 3;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int v[3] = {0, 0, 0};
      if (!mapx) {
        for (int c = 0; c < nc; ++c) v[c] = src[(size_t)y * sstride + (size_t)x * cn + c];
      } else {
        const int sx = cv_round_x32(mapx[(size_t)y * mstride + x]), sy = cv_round_x32(mapy[(size_t)y * mstride + x]);
        const int fx = sx & 31, fy = sy & 31;
        const int ix = std::min(std::max(sx >> 5, -32768), 32767), iy = std::min(std::max(sy >> 5, -32768), 32767);
        int w[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
        if (fx == 0 && fy == 0) w[0] = 32767;
        for (int c = 0; c < nc; ++c) {
          int acc = 0;
          for (int t = 0; t < 4; ++t) {
            const int tx = ix + (t & 1), ty = iy + (t >> 1);
            const int pix = (tx >= 0 && tx < sw && ty >= 0 && ty < sh) ? src[(size_t)ty * sstride + (size_t)tx * cn + c] : 0;
            acc += pix * w[t];
          }
          v[c] = (acc + (1 << 14)) >> 15;
        }
      }
      int g = v[0];
      if (cn != 1) {
        const int b = rgb ? v[2] : v[0], r = rgb ? v[0] : v[2];
        g = (b * 3735 + v[1] * 19235 + r * 9798 + (1 << 14)) >> 15;
      }
      dst[(size_t)y * dstride + x] = (uint8_t)g;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// N2: the per-frame Hamming-search consumers of the extractor's output in Track():
//   ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)   src/ORBmatcher.cc:1372-1519
//   ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th)                   src/ORBmatcher.cc:45-133
//   Frame::GetFeaturesInArea (src/Frame.cc:620-668), ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1650-1695),
//   RadiusByViewingCos (:135-141), DescriptorDistance (:1700-1716).
// MapPoint / Frame objects are flattened into arrays by the caller; flags bit0 = the point takes part (LastFrame:
// mvpMapPoints[i] && !mvbOutlier[i]; local map: mbTrackInView && !isBad()), bit1 = pMP->Observations() > 0 (such a point
// blocks the keypoint it is assigned to for all later points — the loops are sequential and order dependent).
// Float pin: `Rcw*x3Dw+tcw` is cv::gemm's small-matrix path: float products summed left to right in float, then
// (float)(sum*1.0 + t*1.0) in double (checked against cv2.gemm in tests); the remaining expressions are evaluated
// without FMA contraction (-ffp-contract=off), like the rest of this file.
struct ProjFrame {
  const KeyPoint* kps; int N; const uint8_t* desc; const float* uRight;
  const int* gridStart; const int* gridIdx;                      // 64 x 48 CSR of orc_frame_post (cell = col*48 + row)
  float minX, minY, invW, invH;
  const float* scale;
  int nLevels;                                                   // entries in scale[]: points whose octave is outside are skipped (as the CUDA path does)
};

static void features_in_area(const ProjFrame& F, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) {
  out.clear();
  const int COLS = 64, ROWS = 48;
  const int nMinCellX = std::max(0, (int)std::floor((x - F.minX - r) * F.invW));
  if (nMinCellX >= COLS) return;
  const int nMaxCellX = std::min(COLS - 1, (int)std::ceil((x - F.minX + r) * F.invW));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)std::floor((y - F.minY - r) * F.invH));
  if (nMinCellY >= ROWS) return;
  const int nMaxCellY = std::min(ROWS - 1, (int)std::ceil((y - F.minY + r) * F.invH));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
    for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
      for (int j = F.gridStart[ix * ROWS + iy]; j < F.gridStart[ix * ROWS + iy + 1]; ++j) {
        const int idx = F.gridIdx[j];
        const KeyPoint& kp = F.kps[idx];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float distx = kp.x - x, disty = kp.y - y;
        if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
      }
}

static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; ++i) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// cv::Mat x3Dc = Rcw*x3Dw+tcw  (src/ORBmatcher.cc:1408): one cv::gemm call; for 3x3 * 3x1 CV_32F OpenCV takes its small-matrix
// path: float products summed left to right in float, then (float)(sum*alpha + c*beta) in double.
static inline void transform_point(const float* R, const float* t, const float* X, float* out) {
  for (int r = 0; r < 3; ++r) {
    const float t0 = R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2];
    out[r] = (float)((double)t0 * 1.0 + (double)t[r] * 1.0);
  }
}

static int search_by_projection_last(const ProjFrame& F, int n, const float* world, const uint8_t* desc, const int* octave, const float* angle,
                                     const uint8_t* flags, const float* Rcw, const float* tcw, float fx, float fy, float cx, float cy, float mbf,
                                     float maxX, float maxY, int mode, float th, int checkOri, int* match) {
  const int HISTO_LENGTH = 30;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  std::vector<uint8_t> blocked(F.N, 0);        // CurrentFrame.mvpMapPoints[i2] && ->Observations() > 0
  for (int i = 0; i < F.N; ++i) match[i] = -1;
  int nmatches = 0;
  std::vector<int> cand;
  for (int i = 0; i < n; ++i) {
    if (!(flags[i] & 1)) continue;
    float c3[3];
    transform_point(Rcw, tcw, world + 3 * i, c3);
    const float xc = c3[0], yc = c3[1];
    const float invzc = 1.0 / c3[2];
    if (invzc < 0) continue;
    float u = fx * xc * invzc + cx;
    float v = fy * yc * invzc + cy;
    if (!(std::isfinite(u) && std::isfinite(v))) continue;       // zc == 0: the reference goes on with inf/NaN (undefined casts)
    if (u < F.minX || u > maxX) continue;
    if (v < F.minY || v > maxY) continue;
    const int nLastOctave = octave[i];
    if (nLastOctave < 0 || nLastOctave >= F.nLevels) continue;   // the reference would index mvScaleFactors out of range
    const float radius = th * F.scale[nLastOctave];
    if (mode == 1) features_in_area(F, u, v, radius, nLastOctave, -1, cand);
    else if (mode == 2) features_in_area(F, u, v, radius, 0, nLastOctave, cand);
    else features_in_area(F, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (blocked[i2]) continue;
      if (F.uRight[i2] > 0) {
        const float ur = u - mbf * invzc;
        const float er = std::fabs(ur - F.uRight[i2]);
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(desc + 32 * (size_t)i, F.desc + 32 * (size_t)i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      match[bestIdx2] = i;
      blocked[bestIdx2] = (flags[i] & 2) ? 1 : 0;
      nmatches++;
      if (checkOri) {
        float rot = angle[i] - F.kps[bestIdx2].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin = round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(bestIdx2);
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { match[idx] = -1; nmatches--; }
  }
  return nmatches;
}

static int search_by_projection_map(const ProjFrame& F, int n, const float* proj, const float* viewCos, const int* level, const uint8_t* desc,
                                    const uint8_t* flags, const uint8_t* curBlocked, float th, float nnratio, int* match) {
  std::vector<uint8_t> blocked(F.N, 0);
  for (int i = 0; i < F.N; ++i) { match[i] = -1; if (curBlocked) blocked[i] = curBlocked[i]; }
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  std::vector<int> cand;
  for (int iMP = 0; iMP < n; ++iMP) {
    if (!(flags[iMP] & 1)) continue;
    const int nPredictedLevel = level[iMP];
    if (nPredictedLevel < 0 || nPredictedLevel >= F.nLevels) continue;
    float r = viewCos[iMP] > 0.998 ? 2.5 : 4.0;
    if (bFactor) r *= th;
    features_in_area(F, proj[3 * iMP], proj[3 * iMP + 1], r * F.scale[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (blocked[idx]) continue;
      if (F.uRight[idx] > 0) {
        const float er = std::fabs(proj[3 * iMP + 2] - F.uRight[idx]);
        if (er > r * F.scale[nPredictedLevel]) continue;
      }
      const int dist = descriptor_distance(desc + 32 * (size_t)iMP, F.desc + 32 * (size_t)idx);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.kps[idx].octave; bestIdx = idx; }
      else if (dist < bestDist2) { bestLevel2 = F.kps[idx].octave; bestDist2 = dist; }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      match[bestIdx] = iMP;
      blocked[bestIdx] = (flags[iMP] & 2) ? 1 : 0;
      nmatches++;
    }
  }
  return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)  src/ORBmatcher.cc:165-294, flattened: the points are the
// key-frame keypoints of the shared vocabulary nodes in the reference's traversal order, nodeSlot[i] selects F's list.
static int search_by_bow(const KeyPoint* kpsF, int N, const uint8_t* descF, int n, const uint8_t* desc, const float* angle, const uint8_t* flags,
                         const int* nodeSlot, int nNodes, const int* nodeStart, const int* nodeIdx, float nnratio, int checkOri, int* match) {
  const int HISTO_LENGTH = 30;
  std::vector<int> rotHist[HISTO_LENGTH];
  const float factor = 1.0f / HISTO_LENGTH;
  for (int i = 0; i < N; ++i) match[i] = -1;
  int nmatches = 0;
  for (int i = 0; i < n; ++i) {
    if (!(flags[i] & 1)) continue;                       // !pMP || pMP->isBad()
    const int slot = nodeSlot[i];
    if (slot < 0 || slot >= nNodes) continue;
    int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
    for (int j = nodeStart[slot]; j < nodeStart[slot + 1]; ++j) {
      const int realIdxF = nodeIdx[j];
      if (realIdxF < 0 || realIdxF >= N) continue;
      if (match[realIdxF] >= 0) continue;               // vpMapPointMatches[realIdxF] already set
      const int dist = descriptor_distance(desc + 32 * (size_t)i, descF + 32 * (size_t)realIdxF);
      if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist1 <= TH_LOW) {
      if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
        match[bestIdxF] = i;
        if (checkOri) {
          float rot = angle[i] - kpsF[bestIdxF].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(bestIdxF);
        }
        nmatches++;
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { match[idx] = -1; nmatches--; }
    }
  }
  return nmatches;
}

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride, uint8_t* dst, int dw, int dh, size_t dstride) {
  resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_gauss7_u8(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride) {
  gauss7_u8(src, w, h, sstride, dst, dstride);
}
int orc_fast9(const uint8_t* img, int w, int h, size_t stride, int th, int nms, int* xs, int* ys, int* scores, int cap) {
  std::vector<Corner> out; FastScratch fs;
  fast9_nms(img, w, h, stride, th, nms != 0, out, fs);
  int n = (int)out.size();
  for (int i = 0; i < n && i < cap; ++i) { xs[i] = out[i].x; ys[i] = out[i].y; scores[i] = out[i].score; }
  return n;
}
// Same detector with thread-local scratch, for callers that run it on many small windows (the OpenCV-compat layer of
// oracle/refbuild/, where the unmodified reference calls cv::FAST per cell).  Returns n triples (x, y, score).
const int* orc_fast9_tl(const uint8_t* img, int w, int h, size_t stride, int th, int nms, int* n) {
  thread_local std::vector<Corner> out; thread_local FastScratch fs;
  fast9_nms(img, w, h, stride, th, nms != 0, out, fs);
  *n = (int)out.size();
  static_assert(sizeof(Corner) == 3 * sizeof(int), "Corner is three ints");
  return out.empty() ? nullptr : &out[0].x;
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_fast_atan2_array(const float* y, const float* x, float* out, int n) { for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]); }

// retainBest + resize(n) on (response, id) pairs; returns the new length, reorders in place.
int orc_retain_best(float* resp, int* ids, int len, int n) {
  std::vector<KeyPoint> k(len);
  for (int i = 0; i < len; ++i) { k[i] = KeyPoint{0, 0, 0, 0, resp[i], 0, ids[i]}; }
  retain_best_resize(k, n);
  for (size_t i = 0; i < k.size(); ++i) { resp[i] = k[i].response; ids[i] = k[i].class_id; }
  return (int)k.size();
}

void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int enableIntrospection) {
  if (nlevels < 1 || nlevels > 16 || nfeatures < 1) return nullptr;
  return new Extractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, enableIntrospection != 0);
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }
void orc_set_trig_mode(void* h, int mode) { ((Extractor*)h)->trig_mode = mode; }
void orc_set_keypoint_mode(void* h, int mode) { ((Extractor*)h)->kp_mode = mode; }
int orc_features_per_level(void* h, int* out) { Extractor* e = (Extractor*)h; for (int l = 0; l < e->nlevels; ++l) out[l] = e->nPerLevel[l]; return e->nlevels; }
int orc_scale_factors(void* h, float* out) { Extractor* e = (Extractor*)h; for (int l = 0; l < e->nlevels; ++l) out[l] = e->scale[l]; return e->nlevels; }
int orc_umax(void* h, int* out) { Extractor* e = (Extractor*)h; for (int v = 0; v <= 15; ++v) out[v] = e->umax[v]; return 16; }

int orc_extract(void* h, const uint8_t* img, int w, int hgt, size_t stride, const uint8_t* cost, size_t cost_stride,
                void* kps, uint8_t* desc, int cap, int* n_out) {
  return extract(*(Extractor*)h, img, w, hgt, stride, cost, cost_stride, (KeyPoint*)kps, desc, cap, n_out);
}
// pyramid only (used to stage levels for stereo-only tests)
int orc_compute_pyramid(void* h, const uint8_t* img, int w, int hgt, size_t stride) {
  compute_pyramid(*(Extractor*)h, img, w, hgt, stride, false); return 0;
}
int orc_level_size(void* h, int level, int* w, int* hgt) { Extractor* e = (Extractor*)h; *w = e->lv[level].w; *hgt = e->lv[level].h; return 0; }
// which: 0 image pyramid, 1 blurred level, 2 quality (cost-map) pyramid; returns -1 if that plane is absent
int orc_get_level(void* h, int level, int which, uint8_t* dst, size_t dstride) {
  Extractor* e = (Extractor*)h; const Level& L = e->lv[level];
  const std::vector<uint8_t>& s = which == 0 ? L.img : which == 1 ? L.blur : L.qual;
  if (s.size() != (size_t)L.w * L.h) return -1;
  for (int y = 0; y < L.h; ++y) std::memcpy(dst + (size_t)y * dstride, &s[(size_t)y * L.w], L.w);
  return 0;
}
// per-level keypoints in level coordinates (after selection + orientation); returns the count
int orc_get_level_keypoints(void* h, int level, void* kps, int cap) {
  Extractor* e = (Extractor*)h; const std::vector<KeyPoint>& k = e->levelKeys[level];
  for (size_t i = 0; i < k.size() && (int)i < cap; ++i) ((KeyPoint*)kps)[i] = k[i];
  return (int)k.size();
}
int orc_level_grid(void* h, int level, int* out /*cols,rows,cellW,cellH*/) {
  Grid g; if (!level_grid(*(Extractor*)h, level, g)) return -2;
  out[0] = g.cols; out[1] = g.rows; out[2] = g.cellW; out[3] = g.cellH; return 0;
}
void orc_stage_seconds(void* h, double* out6) { Extractor* e = (Extractor*)h; for (int i = 0; i < 6; ++i) out6[i] = e->t_stage[i]; }
void orc_stats(void* h, long* n_fast_calls, long* n_raw) { Extractor* e = (Extractor*)h; *n_fast_calls = e->n_fast_calls; *n_raw = e->n_raw; }

int orc_stereo_match(void* left, void* right, const void* kL, int N, const uint8_t* dL, const void* kR, int Nr,
                     const uint8_t* dR, float mbf, float maxD, float* uRight, float* depth, int* bestDist, int* sad) {
  return stereo_match(*(Extractor*)left, *(Extractor*)right, (const KeyPoint*)kL, N, dL, (const KeyPoint*)kR, Nr, dR,
                      mbf, maxD, uRight, depth, bestDist, sad);
}

// N1: mvKeyQualScore (src/Frame.cc:128-143), UndistortKeyPoints for k1 == 0 (:696-700), AssignFeaturesToGrid + PosInGrid
// (:415-430, :670-680).  cost may be null (=> scores 1.0).  Grid as CSR: cell = col*48 + row, indices ascending.
void orc_frame_post(const void* kps_, int N, const uint8_t* cost, size_t cost_stride, float minX, float maxX, float minY, float maxY,
                    float* qual, int* gridStart, int* gridIdx) {
  const KeyPoint* k = (const KeyPoint*)kps_;
  const int COLS = 64, ROWS = 48;
  for (int i = 0; i < N; ++i) {
    if (cost) {
      int px = static_cast<int>(std::round(k[i].x)), py = static_cast<int>(std::round(k[i].y));
      float c = static_cast<float>(cost[(size_t)py * cost_stride + px]);
      float qual_score = 1.0 / (1.0 + c / 256);
      float qual_score_norm = 2 * qual_score - 1;
      qual[i] = qual_score_norm;
    } else qual[i] = 1.0f;
  }
  const float invW = static_cast<float>(COLS) / (maxX - minX), invH = static_cast<float>(ROWS) / (maxY - minY);
  std::vector<std::vector<int>> grid(COLS * ROWS);
  for (int i = 0; i < N; ++i) {
    int posX = (int)std::round((k[i].x - minX) * invW), posY = (int)std::round((k[i].y - minY) * invH);
    if (posX < 0 || posX >= COLS || posY < 0 || posY >= ROWS) continue;
    grid[posX * ROWS + posY].push_back(i);
  }
  int run = 0;
  for (int c = 0; c < COLS * ROWS; ++c) { gridStart[c] = run; for (int i : grid[c]) gridIdx[run++] = i; }
  gridStart[COLS * ROWS] = run;
}

// One stereo frame with the reference's threading: two extraction threads, then matching on
// the caller (src/Frame.cc:115-125, :193).  cost applies to the left eye only (the right
// extractor is built without introspection, src/Tracking.cc:182-183).
int orc_stereo_frame(void* left, void* right, const uint8_t* imgL, const uint8_t* imgR, int w, int hgt, size_t stride,
                     const uint8_t* cost, size_t cost_stride, float mbf, float maxD, int cap,
                     void* kL, uint8_t* dL, int* nL, void* kR, uint8_t* dR, int* nR, float* uRight, float* depth, int threads) {
  int rcL = 0, rcR = 0;
  auto runL = [&] { rcL = extract(*(Extractor*)left, imgL, w, hgt, stride, cost, cost_stride, (KeyPoint*)kL, dL, cap, nL); };
  auto runR = [&] { rcR = extract(*(Extractor*)right, imgR, w, hgt, stride, cost, cost_stride, (KeyPoint*)kR, dR, cap, nR); };
  if (threads >= 2) { std::thread tl(runL), tr(runR); tl.join(); tr.join(); }
  else { runL(); runR(); }
  if (rcL) return rcL;
  if (rcR) return rcR;
  return stereo_match(*(Extractor*)left, *(Extractor*)right, (const KeyPoint*)kL, *nL, dL, (const KeyPoint*)kR, *nR, dR,
                      mbf, maxD, uRight, depth, nullptr, nullptr);
}

// Frame-parallel batch for the CPU baseline: `workers` threads, each owning an extractor
// pair and pulling frames from a shared counter; every frame is processed as in
// orc_stereo_frame with threads=1.  Images are contiguous (n frames of hgt*stride bytes).
// Outputs: counts per frame and a checksum so the work cannot be optimised away.
int orc_stereo_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int n,
                     const uint8_t* imgsL, const uint8_t* imgsR, int w, int hgt, size_t stride,
                     float mbf, float maxD, int workers, int* nL_out, int* nMatched_out) {
  std::atomic<int> next(0);
  std::atomic<int> err(0);
  auto work = [&] {
    Extractor eL(nfeatures, scaleFactor, nlevels, iniTh, minTh, false), eR(nfeatures, scaleFactor, nlevels, iniTh, minTh, false);
    const int cap = nfeatures + 64;
    std::vector<KeyPoint> kL(cap), kR(cap);
    std::vector<uint8_t> dL((size_t)cap * 32), dR((size_t)cap * 32);
    std::vector<float> uR(cap), dep(cap);
    for (;;) {
      int f = next.fetch_add(1);
      if (f >= n) break;
      int nl = 0, nr = 0;
      int rc = orc_stereo_frame(&eL, &eR, imgsL + (size_t)f * hgt * stride, imgsR + (size_t)f * hgt * stride, w, hgt, stride,
                                nullptr, 0, mbf, maxD, cap, kL.data(), dL.data(), &nl, kR.data(), dR.data(), &nr, uR.data(), dep.data(), 1);
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

void orc_prologue(const uint8_t* src, int sw, int sh, size_t sstride, int cn, int rgb, const float* mapx, const float* mapy,
                  size_t mstride, uint8_t* dst, int W, int H, size_t dstride) {
  prologue_frame(src, sw, sh, sstride, cn, rgb, mapx, mapy, mstride, dst, W, H, dstride);
}

void orc_transform_point(const float* R, const float* t, const float* X, float* out) { transform_point(R, t, X, out); }

int orc_search_by_projection_last(const void* kps, int N, const uint8_t* descCur, const float* uRight, const int* gridStart, const int* gridIdx,
                                  const float* scale, int nlevels, float minX, float maxX, float minY, float maxY,
                                  int n, const float* world, const uint8_t* desc, const int* octave, const float* angle, const uint8_t* flags,
                                  const float* Rcw, const float* tcw, float fx, float fy, float cx, float cy, float mbf, int mode, float th,
                                  int checkOri, int* match) {
  ProjFrame F{(const KeyPoint*)kps, N, descCur, uRight, gridStart, gridIdx, minX, minY, 64.0f / (maxX - minX), 48.0f / (maxY - minY), scale, nlevels};
  return search_by_projection_last(F, n, world, desc, octave, angle, flags, Rcw, tcw, fx, fy, cx, cy, mbf, maxX, maxY, mode, th, checkOri, match);
}

int orc_search_by_projection_map(const void* kps, int N, const uint8_t* descCur, const float* uRight, const int* gridStart, const int* gridIdx,
                                 const float* scale, int nlevels, float minX, float maxX, float minY, float maxY,
                                 int n, const float* proj, const float* viewCos, const int* level, const uint8_t* desc, const uint8_t* flags,
                                 const uint8_t* curBlocked, float th, float nnratio, int* match) {
  ProjFrame F{(const KeyPoint*)kps, N, descCur, uRight, gridStart, gridIdx, minX, minY, 64.0f / (maxX - minX), 48.0f / (maxY - minY), scale, nlevels};
  return search_by_projection_map(F, n, proj, viewCos, level, desc, flags, curBlocked, th, nnratio, match);
}

int orc_search_by_bow(const void* kpsF, int N, const uint8_t* descF, int n, const uint8_t* desc, const float* angle, const uint8_t* flags,
                      const int* nodeSlot, int nNodes, const int* nodeStart, const int* nodeIdx, float nnratio, int checkOri, int* match) {
  return search_by_bow((const KeyPoint*)kpsF, N, descF, n, desc, angle, flags, nodeSlot, nNodes, nodeStart, nodeIdx, nnratio, checkOri, match);
}

}  // extern "C"


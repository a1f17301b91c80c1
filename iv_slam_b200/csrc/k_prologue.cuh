// k_prologue.cuh — N4: the input prologue that runs before ORBextractor::operator() in the reference,
// fused into the level-0 ingest:
//   * cv::remap(src, dst, M1, M2, INTER_LINEAR) with CV_32FC1 maps and the default BORDER_CONSTANT(0)
//     (introspective_ORB_SLAM/Examples/Stereo/stereo_kitti.cc:463-464, :520; stereo_euroc.cc:369-370, :397), on the raw
//     1/3/4-channel frame, and
//   * cvtColor(..., CV_{BGR,RGB,BGRA,RGBA}2GRAY) (src/Tracking.cc:278-294).
// Both are OpenCV (un-vendored, pinned to 4.13 like the rest of the path) fixed-point routines:
//   remap   sx = cvRound(mapx*32), sy likewise; integer part = s>>5 (saturated to short), fraction = s&31; the four taps
//           are weighted by the 15-bit table (32-fy)(32-fx)*32, ... (entry (0,0) holds 32767: saturate_cast<short>(32768)),
//           taps outside the source read 0, result = (sum + 2^14) >> 15;
//   gray    (B*3735 + G*19235 + R*9798 + 2^14) >> 15.
// One thread produces four level-0 pixels and stores them as one aligned word of the row-pitched plane.
#pragma once
#include "common.cuh"

namespace ivg {

struct PrologueArgs {
  const uint8_t* src; size_t srcFrameBytes;    // raw frames, contiguous rows of sw*cn bytes
  int sw, sh, cn, rgb;                         // source size, channels (1, 3, 4), 1 = R first
  const float* mapx; const float* mapy;        // [H][W] or null (no remap: sw == W, sh == H)
  uint8_t* plane; size_t planeBytes;           // destination: level 0 of each frame's pyramid buffer
  int W, H, pitch;
};

__device__ __forceinline__ int cv_round_x32(float m) {
  const float v = __fmul_rn(m, 32.0f);
  // cvRound = cvtss2si: NaN and out-of-range give the "integer indefinite" value
  if (!(v > -2147483648.0f && v < 2147483648.0f)) return (int)0x80000000;
  return __float2int_rn(v);
}

__device__ __forceinline__ int to_gray(int c0, int c1, int c2, int rgb) {
  const int b = rgb ? c2 : c0, r = rgb ? c0 : c2;
  return (b * 3735 + c1 * 19235 + r * 9798 + (1 << 14)) >> 15;
}

__global__ void __launch_bounds__(256) k_prologue(PrologueArgs P) {
  const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x4 >= P.W) return;
  const uint8_t* S = P.src + (size_t)blockIdx.z * P.srcFrameBytes;
  const int cn = P.cn, rowB = P.sw * cn;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x4 + i;
    if (x >= P.W) break;
    int v[3] = {0, 0, 0};
    const int nc = cn == 1 ? 1 : 3;
    if (!P.mapx) {
      const uint8_t* p = S + (size_t)y * rowB + (size_t)x * cn;
      for (int c = 0; c < nc; ++c) v[c] = p[c];
    } else {
      const int sx = cv_round_x32(__ldg(P.mapx + (size_t)y * P.W + x)), sy = cv_round_x32(__ldg(P.mapy + (size_t)y * P.W + x));
      const int fx = sx & 31, fy = sy & 31;
      const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
      int w00 = (32 - fy) * (32 - fx) * 32;
      const int w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
      if ((fx | fy) == 0) w00 = 32767;
      const bool x0in = ix >= 0 && ix < P.sw, x1in = ix + 1 >= 0 && ix + 1 < P.sw;
      const bool y0in = iy >= 0 && iy < P.sh, y1in = iy + 1 >= 0 && iy + 1 < P.sh;
      const uint8_t* r0 = S + (size_t)iy * rowB + (size_t)ix * cn;       // only dereferenced where the tap is inside
      const uint8_t* r1 = r0 + rowB;
      for (int c = 0; c < nc; ++c) {
        const int t00 = (x0in && y0in) ? r0[c] : 0, t01 = (x1in && y0in) ? r0[cn + c] : 0;
        const int t10 = (x0in && y1in) ? r1[c] : 0, t11 = (x1in && y1in) ? r1[cn + c] : 0;
        v[c] = (t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11 + (1 << 14)) >> 15;
      }
    }
    const int g = cn == 1 ? v[0] : to_gray(v[0], v[1], v[2], P.rgb);
    out |= (uint32_t)g << (8 * i);
  }
  *reinterpret_cast<uint32_t*>(P.plane + (size_t)blockIdx.z * P.planeBytes + (size_t)y * P.pitch + x4) = out;   // bytes past W land in the row padding
}

}  // namespace ivg

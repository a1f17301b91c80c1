// k_pyramid.cuh — K1: one pyramid level from the previous one.
// Replaces cv::resize(INTER_LINEAR) inside ORBextractor::ComputePyramid / ComputeQualityImagePyramid
// (introspective_ORB_SLAM/src/ORBextractor.cc:1298-1357, resize calls at :1311 and :1341).
// Arithmetic = OpenCV's 8-bit fixed-point bilinear path: taps and Q11 coefficients are
// precomputed on the host exactly as OpenCV derives them; the kernel does the integer part.
// HBM-bound stencil: each thread produces 4 adjacent output pixels (one 32-bit store), source rows are read
// through the read-only path (L1/L2 resident: a level is at most a few MB).
#pragma once
#include "common.cuh"

namespace ivg {

__global__ void __launch_bounds__(256) k_resize_level(FrameSet fs, int level, int which /*0 image, 1 cost-map*/) {
  const LevelDev& D = fs.lv[level];
  const LevelDev& S = fs.lv[level - 1];
  const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x4 >= D.w || y >= D.h) return;
  uint8_t* plane = (which ? fs.qual : fs.pyr) + (size_t)blockIdx.z * fs.planeBytes;
  const uint8_t* src = plane + S.planeOff;
  uint8_t* dst = plane + D.planeOff;
  const ResizeTap* tx = fs.rtab + D.rtabX;
  const ResizeTap ty = fs.rtab[D.rtabY + y];
  const uint8_t* r0 = src + (size_t)ty.s0 * S.pitch;
  const uint8_t* r1 = src + (size_t)ty.s1 * S.pitch;
  const int b0 = ty.c0, b1 = ty.c1;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = min(x4 + i, D.w - 1);
    const ResizeTap t = tx[x];
    const int t0 = __ldg(r0 + t.s0) * t.c0 + __ldg(r0 + t.s1) * t.c1;
    const int t1 = __ldg(r1 + t.s0) * t.c0 + __ldg(r1 + t.s1) * t.c1;
    const int v = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
    out |= (uint32_t)(v & 0xFF) << (8 * i);
  }
  *reinterpret_cast<uint32_t*>(dst + (size_t)y * D.pitch + x4) = out;
}

}  // namespace ivg

// k_blur.cuh — K5: 7x7 Gaussian blur (sigma 2, BORDER_REFLECT_101) of every pyramid level.
// Replaces cv::GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on the clone of each level
// (introspective_ORB_SLAM/src/ORBextractor.cc:1276-1277).  Arithmetic = OpenCV's 8-bit fixed-point path
// (SURVEY Appendix A.2): Q8 kernel {18,34,48,56,48,34,18}, 16-bit horizontal sums, (v + 2^15) >> 16.
// Separable, one CTA per 64x32 tile (tile table spans all levels: one launch per batch); the tile + 3 px halo is
// staged in shared memory with the reflection applied while loading, so the inner loops are branch-free.
#pragma once
#include "common.cuh"

namespace ivg {

constexpr int BT_PW = BT_W + 8;          // staged row pitch (70 used)
constexpr int BT_PH = BT_H + 6;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_gauss7(FrameSet fs) {
  __shared__ uint8_t spix[BT_PH * BT_PW];
  __shared__ __align__(8) uint16_t shs[BT_PH * BT_W];

  int level = 0;
#pragma unroll 1
  for (int l = 1; l < fs.nlevels; ++l)
    if ((int)blockIdx.x >= fs.lv[l].btBase) level = l;
  const LevelDev& L = fs.lv[level];
  const int t = blockIdx.x - L.btBase;
  const int x0 = (t % L.btX) * BT_W, y0 = (t / L.btX) * BT_H;
  const size_t frameOff = (size_t)blockIdx.y * fs.planeBytes + L.planeOff;
  const uint8_t* img = fs.pyr + frameOff;
  const int tid = threadIdx.x;

  for (int i = tid; i < BT_PH * (BT_W + 6); i += 256) {
    const int r = i / (BT_W + 6), c = i - r * (BT_W + 6);
    const int gy = reflect101(min(y0 - 3 + r, L.h + 2), L.h), gx = reflect101(min(x0 - 3 + c, L.w + 2), L.w);
    spix[r * BT_PW + c] = __ldg(img + (size_t)gy * L.pitch + gx);
  }
  __syncthreads();
  for (int i = tid; i < BT_PH * BT_W; i += 256) {
    const int r = i / BT_W, c = i - r * BT_W;
    const uint8_t* p = spix + r * BT_PW + c;
    shs[i] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
  }
  __syncthreads();
  uint8_t* dst = fs.blur + frameOff;
  for (int g = tid; g < (BT_W / 4) * BT_H; g += 256) {
    const int ry = g / (BT_W / 4), rx4 = (g % (BT_W / 4)) * 4;
    const int gy = y0 + ry, gx = x0 + rx4;
    if (gy >= L.h || gx >= L.pitch) continue;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint16_t* h = shs + ry * BT_W + rx4 + i;
      const uint32_t v = 18u * (h[0] + h[6 * BT_W]) + 34u * (h[BT_W] + h[5 * BT_W]) + 48u * (h[2 * BT_W] + h[4 * BT_W]) + 56u * h[3 * BT_W];
      out |= ((v + 32768u) >> 16) << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)gy * L.pitch + gx) = out;
  }
}

}  // namespace ivg

// k_blur.cuh — K5: 7x7 Gaussian blur (sigma 2, BORDER_REFLECT_101) of every pyramid level.
// Replaces cv::GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on the clone of each level
// (introspective_ORB_SLAM/src/ORBextractor.cc:1276-1277).  Arithmetic = OpenCV's 8-bit fixed-point path
// (SURVEY Appendix A.2): Q8 kernel {18,34,48,56,48,34,18}, 16-bit horizontal sums, (v + 2^15) >> 16.
//
// Separable, one CTA per 128x32 tile (tile table spans all levels and frames: one launch per batch).
//   stage       tile + 3 px halo into shared memory with aligned 32-bit loads (reflection applied while loading; only
//               words that straddle the image edge take the per-byte path);
//   horizontal  4 pixels per thread in packed 16-bit lanes: the row sums are < 2^16, so one IMAD on a register holding
//               two pixels (x, x+2) is two exact multiply-adds — 8 masked funnel-shifted windows feed both the even
//               and the odd pixel pair (28 instructions per 4 pixels);
//   vertical    4 pixels x 4 rows per thread from the packed 16-bit plane, one aligned 32-bit store per row.
#pragma once
#include "common.cuh"

namespace ivg {

constexpr int BL_W = 128, BL_H = 32;
constexpr int BL_PW = (BL_W + 8) / 4;     // staged words per row: x0-4 .. x0+131
constexpr int BL_PH = BL_H + 6;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_gauss7(FrameSet fs) {
  __shared__ __align__(16) uint32_t spx[BL_PH * BL_PW];
  __shared__ __align__(16) uint2 shs[BL_PH * (BL_W / 4)];

  int level = 0;
#pragma unroll 1
  for (int l = 1; l < fs.nlevels; ++l)
    if ((int)blockIdx.x >= fs.lv[l].btBase) level = l;
  const LevelDev& L = fs.lv[level];
  const int t = blockIdx.x - L.btBase;
  const int x0 = (t % L.btX) * BL_W, y0 = (t / L.btX) * BL_H;
  const size_t frameOff = (size_t)blockIdx.y * fs.planeBytes + L.planeOff;
  const uint8_t* img = fs.pyr + frameOff;
  const int tid = threadIdx.x;

  {
    const int w = L.w, h = L.h, pitch = L.pitch;
    for (int i = tid; i < BL_PH * BL_PW; i += 256) {
      const int r = i / BL_PW, g = i - r * BL_PW;
      // reflect-101 of a row index in [-3, h+2]: one reflection suffices (h >= 4)
      int gy = min(y0 - 3 + r, h + 2);
      gy = abs(gy);
      gy = min(gy, 2 * h - 2 - gy);
      const int gx = x0 - 4 + 4 * g;
      const uint8_t* row = img + (size_t)gy * pitch;
      uint32_t v = 0;
      if (gx >= 0 && gx + 3 < w) v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
      else if (gx <= w + 2) {                        // straddles an image edge; columns past w+2 are never used
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int x = abs(min(gx + k, w + 2));
          x = min(x, 2 * w - 2 - x);
          v |= (uint32_t)__ldg(row + x) << (8 * k);
        }
      }
      spx[i] = v;
    }
  }
  __syncthreads();

  const uint32_t M = 0x00FF00FFu;
  for (int i = tid; i < BL_PH * (BL_W / 4); i += 256) {
    const int r = i / (BL_W / 4), g = i - r * (BL_W / 4);
    const uint32_t* w = spx + r * BL_PW + g;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];        // pixels x-4..x-1, x..x+3, x+4..x+7
    const uint32_t G0 = __funnelshift_r(w0, w1, 8) & M, G1 = __funnelshift_r(w0, w1, 16) & M, G2 = __funnelshift_r(w0, w1, 24) & M;
    const uint32_t G3 = w1 & M;
    const uint32_t G4 = __funnelshift_r(w1, w2, 8) & M, G5 = __funnelshift_r(w1, w2, 16) & M, G6 = __funnelshift_r(w1, w2, 24) & M;
    const uint32_t G7 = w2 & M;
    // lanes of G_j: pixels (x-3+j, x-1+j)
    const uint32_t hE = 18u * (G0 + G6) + 34u * (G1 + G5) + 48u * (G2 + G4) + 56u * G3;   // h(x), h(x+2)
    const uint32_t hO = 18u * (G1 + G7) + 34u * (G2 + G6) + 48u * (G3 + G5) + 56u * G4;   // h(x+1), h(x+3)
    shs[i] = make_uint2(hE, hO);
  }
  __syncthreads();

  uint8_t* dst = fs.blur + frameOff;
  {
    const int g = tid & 31, seg = tid >> 5;                 // 32 column groups x 8 row segments of 4 rows
    const int gx = x0 + 4 * g;
    if (gx < L.pitch) {
      uint32_t e[10], o[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) { const uint2 v = shs[(seg * 4 + j) * (BL_W / 4) + g]; e[j] = v.x; o[j] = v.y; }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int gy = y0 + seg * 4 + rr;
        if (gy >= L.h) break;
        uint32_t res[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t hv[7];
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            const uint32_t wv = (q & 1) ? o[rr + k] : e[rr + k];
            hv[k] = (q & 2) ? (wv >> 16) : (wv & 0xFFFFu);
          }
          res[q] = (18u * (hv[0] + hv[6]) + 34u * (hv[1] + hv[5]) + 48u * (hv[2] + hv[4]) + 56u * hv[3] + 32768u) >> 16;
        }
        // pixel order x, x+1, x+2, x+3 = (E lo, O lo, E hi, O hi) = q 0, 1, 2, 3
        *reinterpret_cast<uint32_t*>(dst + (size_t)gy * L.pitch + gx) = res[0] | (res[1] << 8) | (res[2] << 16) | (res[3] << 24);
      }
    }
  }
}

}  // namespace ivg

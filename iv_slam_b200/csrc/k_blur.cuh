// k_blur.cuh — K5: 7x7 Gaussian blur (sigma 2, BORDER_REFLECT_101) of every pyramid level.
// Replaces cv::GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on the clone of each level
// (introspective_ORB_SLAM/src/ORBextractor.cc:1276-1277).  Arithmetic = OpenCV's 8-bit fixed-point path
// (SURVEY Appendix A.2): Q8 kernel {18,34,48,56,48,34,18}, 16-bit horizontal sums, (v + 2^15) >> 16.
//
// Separable, one CTA per 128x64 tile (tile table spans all levels and frames: one launch per batch).
//   stage       tile + halo (160 x 70 bytes) lands in shared memory through ONE TMA box load (cp.async.bulk.tensor.3d over
//               x, y, frame; out-of-image bytes arrive as zeros) signalled on an mbarrier; only tiles that touch an image
//               edge then patch their halo by reflection from the bytes already in shared memory;
//   horizontal  4 pixels per thread with DP4A on byte windows: the Q8 taps fit a byte, so h(x) is two 4-byte dot products
//               ({18,34,48,56} and {48,34,18,0}) over windows cut out of three loaded words by funnel shifts
//               (6 SHF + 8 IDP.4A per 4 pixels); the row sums are < 2^16 and are stored two per word;
//   vertical    4 pixels x 8 rows per thread: the row sums of two vertically adjacent rows share a word, so IDP.2A (two
//               16-bit x 8-bit products) takes a 7-tap column sum in four instructions; one aligned 32-bit store per row.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace ivg {

constexpr int BL_W = 128, BL_H = 64;
constexpr int BL_RPS = BL_H / 8;          // output rows per 32-thread row segment in the vertical pass
constexpr int BL_BOXW = 160;              // TMA box width in bytes: x0-16 .. x0+143 (the box must start 16-byte aligned)
constexpr int BL_X0 = 16;                 // staged byte index of tile column 0
constexpr int BL_PW = BL_BOXW / 4;        // staged words per row
constexpr int BL_PH = BL_H + 6;

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_gauss7(FrameSet fs, const __grid_constant__ TmaMaps maps) {
  __shared__ __align__(128) uint32_t spx[BL_PH * BL_PW];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) uint32_t shs[(BL_PH / 2) * BL_W];      // row sums, two vertically adjacent rows per word

  const uint32_t tt = __ldg(fs.blurTiles + blockIdx.x);       // host-built tile table: level | tile x | tile y
  const int level = tt >> 28;
  const LevelDev& L = fs.lv[level];
  const int x0 = (int)((tt >> 14) & 0x3FFF) * BL_W, y0 = (int)(tt & 0x3FFF) * BL_H;
  const size_t frameOff = (size_t)blockIdx.y * fs.planeBytes + L.planeOff;
  const int tid = threadIdx.x;

  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, BL_PH * BL_BOXW);
    tma_load_3d(spx, &maps.m[level], &bar, x0 - BL_X0, y0 - 3, (int)blockIdx.y);
  }
  mbar_wait(&bar, 0);
  {
    // reflect-101 patches, only where the box left the image (block-uniform conditions): rows first, then columns
    const int w = L.w, h = L.h;
    uint8_t* sb = reinterpret_cast<uint8_t*>(spx);
    const bool top = y0 == 0, bottom = y0 + BL_H + 3 > h;
    if (top || bottom) {
      for (int i = tid; i < 6 * BL_PW; i += 256) {
        const int k = i / BL_PW, g = i - k * BL_PW;
        if (k < 3) {                       // rows -3..-1 <- rows 3..1
          if (top) spx[k * BL_PW + g] = spx[(6 - k) * BL_PW + g];
        } else if (bottom) {               // rows h..h+2 <- rows h-2..h-4
          const int gy = h + (k - 3), r = gy - (y0 - 3);
          if (r < BL_PH) spx[r * BL_PW + g] = spx[(2 * h - 2 - gy - (y0 - 3)) * BL_PW + g];
        }
      }
      __syncthreads();
    }
    const bool left = x0 == 0, right = x0 + BL_W + 3 > w;
    if (left || right) {
      for (int i = tid; i < 6 * BL_PH; i += 256) {
        const int r = i / 6, k = i - r * 6;
        uint8_t* row = sb + r * BL_BOXW;
        if (k < 3) {                       // columns -3..-1 <- columns 3..1   (staged index = column - x0 + BL_X0)
          if (left) row[BL_X0 - 3 + k] = row[BL_X0 + 3 - k];
        } else if (right) {                // columns w..w+2 <- columns w-2..w-4
          const int gx = w + (k - 3), c = gx - x0 + BL_X0;
          if (c < BL_BOXW) row[c] = row[2 * w - 2 - gx - x0 + BL_X0];
        }
      }
    }
  }
  __syncthreads();

  // horizontal: one item = two staged rows (2p, 2p+1) x 4 columns; the two rows' sums go into one word per column
  // (row 2p in the low half), which is the operand layout of IDP.2A in the vertical pass
  for (int i = tid; i < (BL_PH / 2) * (BL_W / 4); i += 256) {
    const int p = i / (BL_W / 4), g = i - p * (BL_W / 4);
    const uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24), KB = 48u | (34u << 8) | (18u << 16);
    uint32_t h[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const uint32_t* w = spx + (2 * p + rr) * BL_PW + (BL_X0 / 4 - 1) + g;
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];        // pixels x-4..x-1, x..x+3, x+4..x+7
      // DP4A on byte windows: h(x+j) = {18,34,48,56} . p[x+j-3 .. x+j] + {48,34,18,0} . p[x+j+1 .. x+j+4]
      h[rr][0] = __dp4a(__funnelshift_r(w1, w2, 8), KB, __dp4a(__funnelshift_r(w0, w1, 8), KA, 0u));
      h[rr][1] = __dp4a(__funnelshift_r(w1, w2, 16), KB, __dp4a(__funnelshift_r(w0, w1, 16), KA, 0u));
      h[rr][2] = __dp4a(__funnelshift_r(w1, w2, 24), KB, __dp4a(__funnelshift_r(w0, w1, 24), KA, 0u));
      h[rr][3] = __dp4a(w2, KB, __dp4a(w1, KA, 0u));
    }
    reinterpret_cast<uint4*>(shs)[i] = make_uint4(h[0][0] | (h[1][0] << 16), h[0][1] | (h[1][1] << 16), h[0][2] | (h[1][2] << 16), h[0][3] | (h[1][3] << 16));
  }
  __syncthreads();

  uint8_t* dst = fs.blur + frameOff;
  {
    const int g = tid & 31, seg = tid >> 5;                 // 32 column groups x 8 row segments of 8 rows
    const int gx = x0 + 4 * g;
    if (gx < L.pitch) {
      // vertical: output row rr needs staged rows rr..rr+6 = four row pairs; IDP.2A multiplies the two 16-bit halves of a
      // pair word by two byte weights, so a 7-tap column sum is four instructions whatever the parity of rr
      const uint32_t WE = 18u | (34u << 8) | (48u << 16) | (56u << 24), WE2 = 48u | (34u << 8) | (18u << 16);
      const uint32_t WO = (18u << 8) | (34u << 16) | (48u << 24), WO2 = 56u | (48u << 8) | (34u << 16) | (18u << 24);
      uint4 P[BL_RPS / 2 + 3];
#pragma unroll
      for (int j = 0; j < BL_RPS / 2 + 3; ++j) P[j] = reinterpret_cast<const uint4*>(shs)[(seg * (BL_RPS / 2) + j) * (BL_W / 4) + g];
#pragma unroll
      for (int rr = 0; rr < BL_RPS; ++rr) {
        const int gy = y0 + seg * BL_RPS + rr;
        if (gy >= L.h) break;
        const int m = rr >> 1;
        const uint32_t wa = (rr & 1) ? WO : WE, wb = (rr & 1) ? WO2 : WE2;
        uint32_t res[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t p0 = c == 0 ? P[m].x : c == 1 ? P[m].y : c == 2 ? P[m].z : P[m].w;
          const uint32_t p1 = c == 0 ? P[m + 1].x : c == 1 ? P[m + 1].y : c == 2 ? P[m + 1].z : P[m + 1].w;
          const uint32_t p2 = c == 0 ? P[m + 2].x : c == 1 ? P[m + 2].y : c == 2 ? P[m + 2].z : P[m + 2].w;
          const uint32_t p3 = c == 0 ? P[m + 3].x : c == 1 ? P[m + 3].y : c == 2 ? P[m + 3].z : P[m + 3].w;
          res[c] = __dp2a_hi(p3, wb, __dp2a_lo(p2, wb, __dp2a_hi(p1, wa, __dp2a_lo(p0, wa, 32768u)))) >> 16;
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)gy * L.pitch + gx) = res[0] | (res[1] << 8) | (res[2] << 16) | (res[3] << 24);
      }
    }
  }
}

}  // namespace ivg

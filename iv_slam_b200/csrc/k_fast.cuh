// k_fast.cuh — K2: FAST-9/16 corner score + cell-masked 3x3 non-max suppression over every pyramid level.
//
// Replaces the per-cell cv::FAST(cellImage, kps, th, true) calls of ComputeKeyPointsOld
// (introspective_ORB_SLAM/src/ORBextractor.cc:1045 and :1051).  SURVEY Appendix A.3:
//   * the corner score S (largest threshold at which the pixel is still a 9-arc corner) does not depend on the
//     threshold, and "corner at th" <=> S >= th, so ONE score pass serves both iniThFAST and minThFAST;
//   * the reference runs FAST per cell window, so NMS at the edge of a cell's detect range sees zeros for pixels
//     that belong to the neighbouring cell: neighbours outside the centre's own detect range are skipped, using
//     per-level x/y flag tables (FLAG_FIRST / FLAG_LAST mark where a range begins / ends);
//   * pixels outside every detect range (including the rows IV-SLAM's stale-hY quirk never searches)
//     are not candidates.
// Output: candidate map, one byte per pixel = S if the pixel survives NMS (S >= minThFAST), else 0.
//
// One CTA = one 64x32 tile of one level of one frame (tile table spans all levels: one launch per batch).
// Integer stencil work: the pixel tile (+4 px halo) is staged in shared memory with aligned 32-bit loads, the
// arc test runs for every pixel, the exact score only for the compacted list of pixels that pass (dense warps).
#pragma once
#include "common.cuh"

namespace ivg {

constexpr int FT_PITCH = 80;                 // bytes per staged row: x0-4 .. x0+75
constexpr int FT_ROWS = FT_H + 8;            // y0-4 .. y0+35
constexpr int FT_TESTW = FT_W + 2, FT_TESTH = FT_H + 2;

__device__ __forceinline__ bool arc9(uint32_t m) {   // 9 contiguous set bits in a circular 16-bit mask
  m |= m << 16;
  uint32_t t = m & (m >> 1);
  t &= t >> 2;
  t &= t >> 4;
  t &= m >> 8;
  return t != 0;
}

// exact score from the 16 ring differences d[k] = centre - ring[k]
__device__ __forceinline__ int fast_score16(const int (&d)[16]) {
  int lo3[16], hi3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo3[k] = min(min(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
    hi3[k] = max(max(d[k], d[(k + 1) & 15]), d[(k + 2) & 15]);
  }
  int best = -1024, worst = 1024;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int mn = min(min(lo3[k], lo3[(k + 3) & 15]), lo3[(k + 6) & 15]);   // min over the arc k..k+8
    const int mx = max(max(hi3[k], hi3[(k + 3) & 15]), hi3[(k + 6) & 15]);
    best = max(best, mn);
    worst = min(worst, mx);
  }
  return max(best, -worst) - 1;
}

__global__ void __launch_bounds__(256) k_fast_nms(FrameSet fs) {
  __shared__ __align__(16) uint8_t spix[FT_ROWS * FT_PITCH];
  __shared__ __align__(16) uint8_t sscore[FT_ROWS * FT_PITCH];
  __shared__ uint8_t sxf[FT_PITCH], syf[FT_ROWS];
  __shared__ uint16_t slist[FT_TESTW * FT_TESTH];
  __shared__ int scount;

  // locate the level of this tile
  int level = 0;
#pragma unroll 1
  for (int l = 1; l < fs.nlevels; ++l)
    if ((int)blockIdx.x >= fs.lv[l].ftBase) level = l;
  const LevelDev& L = fs.lv[level];
  const int t = blockIdx.x - L.ftBase;
  const int x0 = FT_ORG + (t % L.ftX) * FT_W, y0 = FT_ORG + (t / L.ftX) * FT_H;
  const size_t frameOff = (size_t)blockIdx.y * fs.planeBytes + L.planeOff;
  const uint8_t* img = fs.pyr + frameOff;
  const int tid = threadIdx.x;

  if (tid == 0) scount = 0;
  // stage pixels: rows y0-4.., 20 aligned words per row starting at x0-4 (x0 is a multiple of 16)
  for (int i = tid; i < FT_ROWS * (FT_PITCH / 4); i += 256) {
    const int r = i / (FT_PITCH / 4), c = i % (FT_PITCH / 4);
    const int gy = y0 - 4 + r, gx = x0 - 4 + 4 * c;
    uint32_t v = 0;
    if (gy < L.h && gx < L.pitch) v = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)gy * L.pitch + gx));
    reinterpret_cast<uint32_t*>(spix)[i] = v;
    reinterpret_cast<uint32_t*>(sscore)[i] = 0;
  }
  if (tid < FT_PITCH) { const int gx = x0 - 4 + tid; sxf[tid] = gx < L.w ? fs.xflags[L.flagX + gx] : 0; }
  if (tid >= 128 && tid < 128 + FT_ROWS) { const int gy = y0 - 4 + (tid - 128); syf[tid - 128] = gy < L.h ? fs.yflags[L.flagY + gy] : 0; }
  __syncthreads();

  const int th = fs.scoreTh;   // min(iniThFAST, minThFAST)
  constexpr int DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  constexpr int DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

  // phase 1: 9-arc test at minThFAST for every in-range pixel of the tile + 1 px ring
  for (int i = tid; i < FT_TESTW * FT_TESTH; i += 256) {
    const int ry = i / FT_TESTW, rx = i - ry * FT_TESTW;
    const int sy = ry + 3, sx = rx + 3;                 // staged coordinates of pixel (x0-1+rx, y0-1+ry)
    if (!((sxf[sx] & FLAG_IN) && (syf[sy] & FLAG_IN))) continue;
    const uint8_t* p = spix + sy * FT_PITCH + sx;
    const int v = p[0], hi = v + th, lo = v - th;
    uint32_t br = 0, dk = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int r = p[DY[k] * FT_PITCH + DX[k]];
      br |= (r > hi ? 1u : 0u) << k;
      dk |= (r < lo ? 1u : 0u) << k;
    }
    if (arc9(br) || arc9(dk)) slist[atomicAdd(&scount, 1)] = (uint16_t)(sy * FT_PITCH + sx);
  }
  __syncthreads();

  // phase 2: exact score for the pixels that passed (dense)
  const int n = scount;
  for (int j = tid; j < n; j += 256) {
    const int idx = slist[j];
    const uint8_t* p = spix + idx;
    const int v = p[0];
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - (int)p[DY[k] * FT_PITCH + DX[k]];
    sscore[idx] = (uint8_t)fast_score16(d);
  }
  __syncthreads();

  // phase 3: NMS inside the centre's own detect range, 4 pixels per thread, one aligned 32-bit store
  uint8_t* cand = fs.cand + frameOff;
  for (int g = tid; g < (FT_W / 4) * FT_H; g += 256) {
    const int ry = g / (FT_W / 4), rx4 = (g % (FT_W / 4)) * 4;
    const int gy = y0 + ry, gx = x0 + rx4;
    if (gy >= L.h || gx >= L.pitch) continue;
    const int sy = ry + 4;
    const uint8_t yf = syf[sy];
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int sx = rx4 + 4 + i;
      const uint8_t* s = sscore + sy * FT_PITCH + sx;
      const int c = s[0];
      if (c == 0) continue;
      const uint8_t xf = sxf[sx];
      const bool l = !(xf & FLAG_FIRST), r = !(xf & FLAG_LAST), u = !(yf & FLAG_FIRST), d = !(yf & FLAG_LAST);
      bool keep = true;
      if (l) keep = keep && c > s[-1];
      if (r) keep = keep && c > s[1];
      if (u) {
        keep = keep && c > s[-FT_PITCH];
        if (l) keep = keep && c > s[-FT_PITCH - 1];
        if (r) keep = keep && c > s[-FT_PITCH + 1];
      }
      if (d) {
        keep = keep && c > s[FT_PITCH];
        if (l) keep = keep && c > s[FT_PITCH - 1];
        if (r) keep = keep && c > s[FT_PITCH + 1];
      }
      if (keep) out |= (uint32_t)c << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(cand + (size_t)gy * L.pitch + gx) = out;
  }
}

}  // namespace ivg

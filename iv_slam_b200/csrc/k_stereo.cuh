// k_stereo.cuh — K7-K10: Frame::ComputeStereoMatches (introspective_ORB_SLAM/src/Frame.cc:758-932).
//
//   k_stereo_index   one CTA per stereo pair: counting sort of the right keypoints by image row (int)y into a CSR
//       table (the reference's vRowIndices build, :768-785, stores every keypoint in all rows of its band; here each
//       keypoint is stored once and the band test is applied by the searcher).
//   k_stereo_match   one warp per left keypoint.
//     * candidate search (:787-841): the reference scans the list of row (int)vL, i.e. the right keypoints whose
//       band [floor(y-r), ceil(y+r)], r = 2*scale[octave], covers that row.  Candidate membership is a pure predicate
//       of (left kp, right kp), and the winner is the minimum of (Hamming distance, iR) because the list is in
//       ascending iR and the update is a strict '<' starting from TH_HIGH=100 (src/ORBmatcher.cc:37) — so lanes test
//       the right keypoints of the rows within the widest possible band in parallel (any order) and a warp arg-min
//       over the packed (dist<<16 | iR) reproduces the result exactly.  Distance = popcount of 8 xor-ed words
//       (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1700-1716).
//     * SAD refinement (:843-915): 11x11 window around the rounded level coordinates in the left keypoint's octave of
//       the UNBLURRED pyramids, 11 horizontal shifts, each patch minus its own centre, L1 norm (integers, exact),
//       parabola fit in float, disparity / depth.
//   k_stereo_median  one CTA per stereo pair: median of the surviving SADs by two-pass radix select (the sort at :918
//     only feeds the median), then invalidates matches with SAD >= 1.5*1.4*median (:919-931).
//     No surviving match = no-op.
// maxD is an argument.
#pragma once
#include "common.cuh"

namespace ivg {

struct StereoArgs {
  const uint8_t* kpL; const uint8_t* descL; const int* nL;   // [nPairs][cap] records (28 B) / 32 B; counts
  const uint8_t* kpR; const uint8_t* descR; const int* nR;
  const uint8_t* pyrL; const uint8_t* pyrR;                  // image pyramids, [nPairs][planeBytes]
  size_t planeBytes;
  int cap;
  int nRows;                                                 // rows of level 0
  float mbf, maxD;
  float* uRight; float* depth; int* sad;                     // [nPairs][cap]
  int* bestDist;                                             // optional debug [nPairs][cap] or null
  uint4* sorted;                                             // [nPairs][cap] right keypoints bucketed by row: (x bits, y bits, octave, iR)
  int* rowStart;                                             // [nPairs][nRows + 1]
  int bandMargin;                                            // rows to scan either side of (int)vL: ceil(2*scale[last]) + 2
};

constexpr int TH_HIGH = 100, TH_LOW = 50;

__global__ void __launch_bounds__(256) k_stereo_index(StereoArgs A) {
  extern __shared__ int ssh[];               // nRows + 1 counters, then the scatter cursors in place
  __shared__ int wsum[8];
  const size_t pair = blockIdx.x;
  const int Nr = A.nR[pair], nRows = A.nRows, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t* kr0 = A.kpR + pair * A.cap * 28;
  for (int i = tid; i <= nRows; i += 256) ssh[i] = 0;
  __syncthreads();
  for (int iR = tid; iR < Nr; iR += 256) {
    const float y = reinterpret_cast<const float*>(kr0 + (size_t)iR * 28)[1];
    atomicAdd(&ssh[min(max((int)y, 0), nRows - 1)], 1);
  }
  __syncthreads();
  // exclusive scan over nRows+1 entries: each thread owns a contiguous chunk
  const int per = (nRows + 1 + 255) / 256, b0 = tid * per, b1 = min(b0 + per, nRows + 1);
  int local = 0;
  for (int i = b0; i < b1; ++i) local += ssh[i];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int base = incl - local;
  for (int w = 0; w < warp; ++w) base += wsum[w];
  int* rs = A.rowStart + pair * (size_t)(nRows + 1);
  for (int i = b0; i < b1; ++i) { const int c = ssh[i]; ssh[i] = base; rs[i] = base; base += c; }
  __syncthreads();
  uint4* out = A.sorted + pair * A.cap;
  for (int iR = tid; iR < Nr; iR += 256) {
    const float* kr = reinterpret_cast<const float*>(kr0 + (size_t)iR * 28);
    const float x = kr[0], y = kr[1];
    const int oct = reinterpret_cast<const int*>(kr)[5];
    const int pos = atomicAdd(&ssh[min(max((int)y, 0), nRows - 1)], 1);   // order inside a row is irrelevant (arg-min is order-free)
    out[pos] = make_uint4(__float_as_uint(x), __float_as_uint(y), (unsigned)oct, (unsigned)iR);
  }
}

__global__ void __launch_bounds__(256) k_stereo_match(FrameSet fs, StereoArgs A) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pair = blockIdx.y;
  const int iL = blockIdx.x * 8 + warp;
  const int N = A.nL[pair], Nr = A.nR[pair];
  if (iL >= A.cap) return;
  const size_t o = pair * A.cap + iL;
  if (iL >= N) {   // slots past the keypoint count read as "no match"
    if (lane == 0) { A.uRight[o] = -1.f; A.depth[o] = -1.f; A.sad[o] = -1; }
    return;
  }
  const float* kl = reinterpret_cast<const float*>(A.kpL + o * 28);
  const float uL = kl[0], vL = kl[1];
  const int levelL = reinterpret_cast<const int*>(kl)[5];
  const int row = (int)vL;
  const float minU = __fsub_rn(uL, A.maxD), maxU = uL;   // minD = 0
  unsigned best = ((unsigned)TH_HIGH << 16) | 0xFFFFu;
  uint32_t dl[8];
  {
    const uint32_t* d = reinterpret_cast<const uint32_t*>(A.descL + o * 32);
#pragma unroll
    for (int k = 0; k < 8; ++k) dl[k] = __ldg(d + k);
  }
  if (row >= 0 && row < A.nRows && !(maxU < 0) && Nr > 0) {
    const uint8_t* dr0 = A.descR + pair * A.cap * 32;
    const int* rs = A.rowStart + pair * (size_t)(A.nRows + 1);
    const uint4* srt = A.sorted + pair * A.cap;
    const int jb = __ldg(rs + max(row - A.bandMargin, 0)), je = __ldg(rs + min(row + A.bandMargin + 1, A.nRows));
    for (int j = jb + lane; j < je; j += 32) {
      const uint4 e = __ldg(srt + j);
      const int octR = (int)e.z;
      if (octR < levelL - 1 || octR > levelL + 1) continue;
      const float uR = __uint_as_float(e.x), kpY = __uint_as_float(e.y);
      if (!(uR >= minU && uR <= maxU)) continue;
      const float r = __fmul_rn(2.0f, fs.lv[octR].scale);
      const int maxr = (int)ceilf(__fadd_rn(kpY, r)), minr = (int)floorf(__fsub_rn(kpY, r));
      if (row < minr || row > maxr) continue;
      const uint4* d = reinterpret_cast<const uint4*>(dr0 + (size_t)e.w * 32);
      const uint4 da = __ldg(d), db = __ldg(d + 1);
      const int dist = __popc(dl[0] ^ da.x) + __popc(dl[1] ^ da.y) + __popc(dl[2] ^ da.z) + __popc(dl[3] ^ da.w) +
                       __popc(dl[4] ^ db.x) + __popc(dl[5] ^ db.y) + __popc(dl[6] ^ db.z) + __popc(dl[7] ^ db.w);
      if (dist < TH_HIGH) best = min(best, ((unsigned)dist << 16) | e.w);
    }
  }
#pragma unroll
  for (int s = 16; s; s >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, s));
  const int bestDist = best >> 16;
  const int bestIdxR = best & 0xFFFF;
  float outU = -1.f, outD = -1.f;
  int outSad = -1;
  if (A.bestDist && lane == 0) A.bestDist[o] = bestDist;
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  if (bestDist < thOrbDist) {
    const float uR0 = reinterpret_cast<const float*>(A.kpR + (pair * A.cap + bestIdxR) * 28)[0];
    const LevelDev& L = fs.lv[levelL];
    const float sf = L.invScale;
    const float scaleduL = roundf(__fmul_rn(uL, sf)), scaledvL = roundf(__fmul_rn(vL, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int w = 5, Ls = 5;
    const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)scaleduR0;
    const float iniu = scaleduR0 + (float)(Ls - w), endu = scaleduR0 + (float)(Ls + w + 1);
    bool ok = !(iniu < 0 || endu >= (float)L.w);
    ok = ok && !(cy - w < 0 || cy + w + 1 > L.h || cxL - w < 0 || cxL + w + 1 > L.w || cxR - Ls - w < 0);
    if (ok) {
      const uint8_t* PL = A.pyrL + pair * A.planeBytes + L.planeOff;
      const uint8_t* PR = A.pyrR + pair * A.planeBytes + L.planeOff;
      const int cL = __ldg(PL + (size_t)cy * L.pitch + cxL);
      int cR[11];
#pragma unroll
      for (int s = 0; s < 11; ++s) cR[s] = __ldg(PR + (size_t)cy * L.pitch + cxR + s - Ls);
      int acc[11];
#pragma unroll
      for (int s = 0; s < 11; ++s) acc[s] = 0;
      for (int p = lane; p < 121; p += 32) {
        const int dy = p / 11 - w, dx = p % 11 - w;
        const int lv = (int)__ldg(PL + (size_t)(cy + dy) * L.pitch + cxL + dx) - cL;
        const uint8_t* rr = PR + (size_t)(cy + dy) * L.pitch + cxR + dx - Ls;
#pragma unroll
        for (int s = 0; s < 11; ++s) acc[s] += abs(lv - ((int)__ldg(rr + s) - cR[s]));
      }
#pragma unroll
      for (int s = 0; s < 11; ++s) {
#pragma unroll
        for (int q = 16; q; q >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], q);
      }
      int bestSad = 0x7fffffff, bestinc = 0;
#pragma unroll
      for (int s = 0; s < 11; ++s)
        if (acc[s] < bestSad) { bestSad = acc[s]; bestinc = s - Ls; }
      if (!(bestinc == -Ls || bestinc == Ls)) {
        float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int s = 1; s < 10; ++s)
          if (s - Ls == bestinc) { d1 = (float)acc[s - 1]; d2 = (float)acc[s]; d3 = (float)acc[s + 1]; }
        const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
        if (!(deltaR < -1.f || deltaR > 1.f)) {
          float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
          float disparity = __fsub_rn(uL, bestuR);
          if (disparity >= 0.f && disparity < A.maxD) {
            if (disparity <= 0.f) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
            outD = __fdiv_rn(A.mbf, disparity);
            outU = bestuR;
            outSad = bestSad;
          }
        }
      }
    }
  }
  if (lane == 0) { A.uRight[o] = outU; A.depth[o] = outD; A.sad[o] = outSad; }
}

__global__ void __launch_bounds__(256) k_stereo_median(StereoArgs A) {
  __shared__ int hist[256];
  __shared__ int sel[3];   // [0] chosen high byte, [1] rank inside it, [2] median value
  const size_t pair = blockIdx.x;
  const int N = A.nL[pair];
  const int* sad = A.sad + pair * A.cap;
  const int tid = threadIdx.x;
  hist[tid] = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < N; i += 256) {
    const int s = sad[i];
    if (s >= 0) { atomicAdd(&hist[(s >> 8) & 0xFF], 1); ++cnt; }
  }
  __syncthreads();
  if (tid == 0) {
    int total = 0;
    for (int b = 0; b < 256; ++b) total += hist[b];
    sel[0] = -1;
    if (total > 0) {
      int k = total / 2, b = 0;                       // vDistIdx[size/2] of the ascending sort
      while (k >= hist[b]) { k -= hist[b]; ++b; }
      sel[0] = b; sel[1] = k;
    }
  }
  __syncthreads();
  const int hb = sel[0];
  if (hb < 0) return;                                  // no matches: nothing to filter
  hist[tid] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += 256) {
    const int s = sad[i];
    if (s >= 0 && ((s >> 8) & 0xFF) == hb) atomicAdd(&hist[s & 0xFF], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int k = sel[1], b = 0;
    while (k >= hist[b]) { k -= hist[b]; ++b; }
    sel[2] = (hb << 8) | b;
  }
  __syncthreads();
  const float median = (float)sel[2];
  const float thDist = __fmul_rn(1.5f * 1.4f, median);
  for (int i = tid; i < N; i += 256) {
    const int s = sad[i];
    if (s >= 0 && !((float)s < thDist)) { A.uRight[pair * A.cap + i] = -1.f; A.depth[pair * A.cap + i] = -1.f; }
  }
}

}  // namespace ivg

// k_stereo.cuh — K7-K10: Frame::ComputeStereoMatches (introspective_ORB_SLAM/src/Frame.cc:758-932).
//
//   k_stereo_index   one CTA per stereo pair: counting sort of the right keypoints by (octave, image row (int)y) into a
//       CSR table (the reference's vRowIndices build, :768-785, stores every keypoint in all rows of its band; here
//       each keypoint is stored once and the band test is applied by the searcher, which only visits the three octaves
//       levelL-1..levelL+1 the reference accepts (:815) and, per octave, the rows its band radius can reach).
//   k_stereo_match   one warp per 8 left keypoints: the candidate search runs one keypoint at a time on all 32 lanes,
//     the SAD refinement of the 8 keypoints runs together, 4 lanes per keypoint.
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
  int* rowStart;                                             // [nPairs][nLevels * nRows + 1], bin = octave * nRows + row
  int nLevels;
  float* hostU; float* hostD;                                // optional: pinned host copies of uRight / depth, written by the last kernel of the matcher
};

constexpr int TH_HIGH = 100, TH_LOW = 50;

__device__ __forceinline__ int stereo_bin(float y, int oct, int nRows, int nLevels) {
  if (oct < 0 || oct >= nLevels) return -1;      // the reference would index mvScaleFactors out of range; such keypoints never match here
  return oct * nRows + min(max((int)y, 0), nRows - 1);
}

// 256 threads per pair in batches, 1024 when a handful of frames is all there is (the only CTA of the launch: the loops over
// bins and keypoints get four times shorter)
__global__ void __launch_bounds__(1024) k_stereo_index(StereoArgs A) {
  extern __shared__ int ssh[];               // nBins + 1 counters, then the scatter cursors in place
  __shared__ int wsum[32];
  const size_t pair = blockIdx.x;
  const int Nr = A.nR[pair], nRows = A.nRows, nBins = A.nRows * A.nLevels, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x;
  const uint8_t* kr0 = A.kpR + pair * A.cap * 28;
  for (int i = tid; i <= nBins; i += nthr) ssh[i] = 0;
  __syncthreads();
  for (int iR = tid; iR < Nr; iR += nthr) {
    const float* kr = reinterpret_cast<const float*>(kr0 + (size_t)iR * 28);
    const int bin = stereo_bin(kr[1], reinterpret_cast<const int*>(kr)[5], nRows, A.nLevels);
    if (bin >= 0) atomicAdd(&ssh[bin], 1);
  }
  __syncthreads();
  // exclusive scan over nBins+1 entries: each thread owns a contiguous chunk
  const int per = (nBins + nthr) / nthr, b0 = min(tid * per, nBins + 1), b1 = min(b0 + per, nBins + 1);
  int local = 0;
  for (int i = b0; i < b1; ++i) local += ssh[i];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int base = incl - local;
  {
    // sum of the warps before this one: one value per lane and a shuffle reduction
    int ws = lane < warp ? wsum[lane] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
    base += ws;
  }
  int* rs = A.rowStart + pair * (size_t)(nBins + 1);
  for (int i = b0; i < b1; ++i) { const int c = ssh[i]; ssh[i] = base; rs[i] = base; base += c; }
  __syncthreads();
  uint4* out = A.sorted + pair * A.cap;
  for (int iR = tid; iR < Nr; iR += nthr) {
    const float* kr = reinterpret_cast<const float*>(kr0 + (size_t)iR * 28);
    const float x = kr[0], y = kr[1];
    const int oct = reinterpret_cast<const int*>(kr)[5];
    const int bin = stereo_bin(y, oct, nRows, A.nLevels);
    if (bin < 0) continue;
    const int pos = atomicAdd(&ssh[bin], 1);   // order inside a bin is irrelevant (arg-min is order-free)
    out[pos] = make_uint4(__float_as_uint(x), __float_as_uint(y), (unsigned)oct, (unsigned)iR);
  }
}

constexpr int SM_KP = 8;                   // left keypoints per warp
#ifndef IVG_SM_KP_LAT
#define IVG_SM_KP_LAT 2
#endif
constexpr int SM_KP_LAT = IVG_SM_KP_LAT;   // the same in the one-frame-at-a-time configuration (1 and 4 measured: see profiles/README.md)
constexpr int SM_ROWB = 48;                // staged bytes per patch row: right strip at [0,21), left patch at [32,43)
constexpr int SM_SLOT = 11 * SM_ROWB;      // one keypoint's 11 rows
#ifndef IVG_SM_WARPS
#define IVG_SM_WARPS 4
#endif
constexpr int SM_WARPS = IVG_SM_WARPS;

// [b,0,b,0] with b = byte k of w: one 8-bit value in both 16-bit lanes
__device__ __forceinline__ unsigned dup16(unsigned w, int k) { return __byte_perm(w, 0u, 0x4040u | (unsigned)k | ((unsigned)k << 8)); }

#ifndef IVG_SM_MINB
#define IVG_SM_MINB 10
#endif
// KP = left keypoints per warp: SM_KP for batches (the per-keypoint chains of ten resident CTAs hide each other's latency),
// SM_KP_LAT when one frame's 2000 keypoints are all there is: four times the warps, a quarter of the dependent loads each.
template <int KP>
__global__ void __launch_bounds__(32 * SM_WARPS, KP == SM_KP ? IVG_SM_MINB : 1) k_stereo_match(FrameSet fs, StereoArgs A) {
  __shared__ __align__(16) uint8_t patch[SM_WARPS][KP][SM_SLOT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t pair = blockIdx.y;
  const int iL0 = (blockIdx.x * SM_WARPS + warp) * KP;
  const int N = A.nL[pair], Nr = A.nR[pair];
  if (iL0 >= A.cap) return;
  {
    uint4* z = reinterpret_cast<uint4*>(&patch[warp][0][0]);   // pad bytes must read as zero
    for (int i = lane; i < KP * SM_SLOT / 16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();

  // ---- phase 1: one left keypoint at a time, the whole warp searches its candidates and stages the SAD patches.
  // Up front lane k (< 8) loads keypoint k's record and lanes 3k..3k+2 look up its three CSR segments (octaves
  // levelL-1, levelL, levelL+1), so the per-keypoint dependent chain is: candidates -> descriptors -> patch rows.
  // Lanes 4k..4k+3 remember keypoint k's parameters for phase 2.
  bool kHit = false;          // lane k (< 8): keypoint k found a match with distance < thOrbDist
  float kUR0 = 0.f;           //               x of that right keypoint
  const uint8_t* dr0 = A.descR + pair * A.cap * 32;
  const int nBins = A.nRows * A.nLevels;
  const int* rs = A.rowStart + pair * (size_t)(nBins + 1);
  const uint4* srt = A.sorted + pair * A.cap;
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  float myU = 0.f, myV = 0.f;
  int myLev = -1;
  if (lane < KP && iL0 + lane < N) {
    const float* kl = reinterpret_cast<const float*>(A.kpL + (pair * A.cap + iL0 + lane) * 28);
    myU = kl[0]; myV = kl[1]; myLev = reinterpret_cast<const int*>(kl)[5];
  }
  int segB = 0, segN = 0;
  {
    const int k3 = lane / 3, t = lane - 3 * k3, src = min(k3, KP - 1);
    const float u = __shfl_sync(0xffffffffu, myU, src), v = __shfl_sync(0xffffffffu, myV, src);
    const int lev = __shfl_sync(0xffffffffu, myLev, src);
    const int oct = lev - 1 + t, row = (int)v;
    if (k3 < KP && lev >= 0 && lev < A.nLevels && oct >= 0 && oct < A.nLevels && row >= 0 && row < A.nRows && !(u < 0) && Nr > 0) {
      // rows whose bucket can hold a keypoint with row in [floor(y - r), ceil(y + r)], r = 2*scale[oct]
      const int m = (int)ceilf(__fmul_rn(2.0f, fs.lv[oct].scale)) + 1;
      segB = __ldg(rs + oct * A.nRows + max(row - m, 0));
      segN = __ldg(rs + oct * A.nRows + min(row + m + 1, A.nRows)) - segB;
    }
  }

  for (int k = 0; k < KP; ++k) {
    const int iL = iL0 + k;
    if (iL >= A.cap) break;
    const size_t o = pair * A.cap + iL;
    if (iL >= N) {   // slots past the keypoint count read as "no match"
      if (lane == 0) { A.uRight[o] = -1.f; A.depth[o] = -1.f; A.sad[o] = -1; }
      continue;
    }
    const float uL = __shfl_sync(0xffffffffu, myU, k), vL = __shfl_sync(0xffffffffu, myV, k);
    const int levelL = __shfl_sync(0xffffffffu, myLev, k);
    const int b0 = __shfl_sync(0xffffffffu, segB, 3 * k), b1 = __shfl_sync(0xffffffffu, segB, 3 * k + 1), b2 = __shfl_sync(0xffffffffu, segB, 3 * k + 2);
    const int n0 = __shfl_sync(0xffffffffu, segN, 3 * k), n1 = __shfl_sync(0xffffffffu, segN, 3 * k + 1), n2 = __shfl_sync(0xffffffffu, segN, 3 * k + 2);
    const int row = (int)vL;
    const float minU = __fsub_rn(uL, A.maxD), maxU = uL;   // minD = 0
    unsigned best = ((unsigned)TH_HIGH << 16) | 0xFFFFu;
    float bestU = 0.f;
    const int total = n0 + n1 + n2;
    if (total > 0) {
      uint32_t dl[8];
      {
        const uint4* d = reinterpret_cast<const uint4*>(A.descL + o * 32);
        const uint4 a = __ldg(d), b = __ldg(d + 1);
        dl[0] = a.x; dl[1] = a.y; dl[2] = a.z; dl[3] = a.w; dl[4] = b.x; dl[5] = b.y; dl[6] = b.z; dl[7] = b.w;
      }
      for (int j = lane; j < total; j += 32) {
        const int idx = j < n0 ? b0 + j : (j < n0 + n1 ? b1 + (j - n0) : b2 + (j - n0 - n1));
        const uint4 e = __ldg(srt + idx);
        const float uR = __uint_as_float(e.x), kpY = __uint_as_float(e.y);
        if (!(uR >= minU && uR <= maxU)) continue;
        const float r = __fmul_rn(2.0f, fs.lv[(int)e.z].scale);
        const int maxr = (int)ceilf(__fadd_rn(kpY, r)), minr = (int)floorf(__fsub_rn(kpY, r));
        if (row < minr || row > maxr) continue;
        const uint4* d = reinterpret_cast<const uint4*>(dr0 + (size_t)e.w * 32);
        const uint4 da = __ldg(d), db = __ldg(d + 1);
        const int dist = __popc(dl[0] ^ da.x) + __popc(dl[1] ^ da.y) + __popc(dl[2] ^ da.z) + __popc(dl[3] ^ da.w) +
                         __popc(dl[4] ^ db.x) + __popc(dl[5] ^ db.y) + __popc(dl[6] ^ db.z) + __popc(dl[7] ^ db.w);
        const unsigned cand = ((unsigned)dist << 16) | e.w;
        if (dist < TH_HIGH && cand < best) { best = cand; bestU = uR; }
      }
    }
    const unsigned mine = best;
#pragma unroll
    for (int s = 16; s; s >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, s));
    const int bestDist = best >> 16;
    if (A.bestDist && lane == 0) A.bestDist[o] = bestDist;
    if (bestDist < thOrbDist) {
      // x of the winning right keypoint: held by the lane that found it (iR is unique, so is the winner)
      const float uR0 = __shfl_sync(0xffffffffu, bestU, __ffs(__ballot_sync(0xffffffffu, mine == best)) - 1);
      if (lane == k) { kHit = true; kUR0 = uR0; }
    }
  }

  // ---- phase 1b: lane k (< 8) derives keypoint k's SAD window from its own record: 8 keypoints in parallel
  bool kDo = false;
  float kSUR0 = 0.f, kScale = 1.f;
  int kOffR = 0, kOffL = 0, kPitch = 0;
  if (lane < KP && iL0 + lane < A.cap) {
    if (kHit) {
      const LevelDev& L = fs.lv[myLev];
      const float sf = L.invScale;
      kScale = L.scale;
      const float scaleduL = roundf(__fmul_rn(myU, sf)), scaledvL = roundf(__fmul_rn(myV, sf));
      kSUR0 = roundf(__fmul_rn(kUR0, sf));
      const int w = 5, Ls = 5;
      const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)kSUR0;
      const float iniu = kSUR0 + (float)(Ls - w), endu = kSUR0 + (float)(Ls + w + 1);
      bool ok = !(iniu < 0 || endu >= (float)L.w);
      ok = ok && !(cy - w < 0 || cy + w + 1 > L.h || cxL - w < 0 || cxL + w + 1 > L.w || cxR - Ls - w < 0);
      if (ok) {
        kDo = true;
        kPitch = L.pitch;
        kOffR = (int)L.planeOff + (cy - w) * L.pitch + (cxR - Ls - w);
        kOffL = (int)L.planeOff + (cy - w) * L.pitch + (cxL - w);
      }
    }
    if (!kDo && iL0 + lane < N) { const size_t o = pair * A.cap + iL0 + lane; A.uRight[o] = -1.f; A.depth[o] = -1.f; A.sad[o] = -1; }
  }

  // ---- phase 1c: stage the 11 patch rows of every keypoint that goes on: lanes 0..20 the right strip (columns
  // cxR-10..cxR+10), lanes 21..31 the left patch (columns cxL-5..cxL+5)
  const unsigned doMask = __ballot_sync(0xffffffffu, kDo);
  for (unsigned rest = doMask; rest;) {
    const int k = __ffs(rest) - 1;
    rest &= rest - 1;
    const int offR = __shfl_sync(0xffffffffu, kOffR, k), offL = __shfl_sync(0xffffffffu, kOffL, k), pitch = __shfl_sync(0xffffffffu, kPitch, k);
    const uint8_t* src = lane < 21 ? A.pyrR + pair * A.planeBytes + offR + lane : A.pyrL + pair * A.planeBytes + offL + (lane - 21);
    uint8_t* dst = &patch[warp][k][lane < 21 ? lane : lane + 11];
    uint8_t v[11];
#pragma unroll
    for (int dy = 0; dy < 11; ++dy) v[dy] = __ldg(src + (size_t)dy * pitch);
#pragma unroll
    for (int dy = 0; dy < 11; ++dy) dst[dy * SM_ROWB] = v[dy];
  }
  // lanes 4k..4k+3 take over keypoint k's parameters for phase 2
  const bool gDo = __shfl_sync(0xffffffffu, (int)kDo, lane >> 2) != 0;
  const float gUL = __shfl_sync(0xffffffffu, myU, lane >> 2), gUR0 = __shfl_sync(0xffffffffu, kSUR0, lane >> 2);
  const float gScale = __shfl_sync(0xffffffffu, kScale, lane >> 2);
  if (!__any_sync(0xffffffffu, gDo)) return;
  __syncwarp();

  // ---- phase 2: SAD refinement, 4 lanes per keypoint (lane j takes patch rows j, j+4, j+8), two pixels per
  // instruction in packed 16-bit lanes.  With e = L - cL, d' = (R - cR_s) - e for shift s:
  //   |d'| = 2*max(d',0) - d',  max(d',0) = max(R + (cL - L), cR_s) - cR_s   (one VIADDMNMX.S16x2 per pixel pair)
  // and the sum of d' over the window comes from window sums of R and the sum of L.  All integers: exact.
  // The 11th pixel of a row is paired with a dummy whose max() is exactly cR_s.
  const int j = lane & 3, kk = min(lane >> 2, KP - 1);      // lanes past 4*KP idle along (gDo is false for them)
  const uint8_t* S = &patch[warp][kk][0];
  unsigned cR2[11];
  int cL;
  {
    const uint4 r5 = *reinterpret_cast<const uint4*>(S + 5 * SM_ROWB);           // right bytes 0..15 of the centre row
    cR2[0] = dup16(r5.y, 1); cR2[1] = dup16(r5.y, 2); cR2[2] = dup16(r5.y, 3);     // cR_s = right byte 5 + s
    cR2[3] = dup16(r5.z, 0); cR2[4] = dup16(r5.z, 1); cR2[5] = dup16(r5.z, 2); cR2[6] = dup16(r5.z, 3);
    cR2[7] = dup16(r5.w, 0); cR2[8] = dup16(r5.w, 1); cR2[9] = dup16(r5.w, 2); cR2[10] = dup16(r5.w, 3);
    cL = S[5 * SM_ROWB + 32 + 5];
  }
  const unsigned cL2p1 = (unsigned)(cL + 1) * 0x00010001u;
  unsigned acc[11], accW[11], sumL2 = 0;
#pragma unroll
  for (int s = 0; s < 11; ++s) { acc[s] = 0; accW[s] = 0; }
#pragma unroll 1
  for (int dy = j; dy < 11; dy += 4) {
    const uint8_t* rp = S + dy * SM_ROWB;
    const uint4 ra = *reinterpret_cast<const uint4*>(rp);
    const uint2 rb = *reinterpret_cast<const uint2*>(rp + 16);
    const uint4 la = *reinterpret_cast<const uint4*>(rp + 32);
    const unsigned w[6] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y};
    unsigned R2[21];                       // R2[c] = right bytes (c, c+1) in 16-bit lanes
#pragma unroll
    for (int c = 0; c < 21; ++c) {
      const int i = c >> 2, m = c & 3;
      R2[c] = m == 0 ? __byte_perm(w[i], 0u, 0x4140) : m == 1 ? __byte_perm(w[i], 0u, 0x4241) : m == 2 ? __byte_perm(w[i], 0u, 0x4342)
                     : __byte_perm(__funnelshift_r(w[i], w[i + 1 < 6 ? i + 1 : 5], 24), 0u, 0x4140);
    }
    unsigned L2[6] = {__byte_perm(la.x, 0u, 0x4140), __byte_perm(la.x, 0u, 0x4342), __byte_perm(la.y, 0u, 0x4140),
                      __byte_perm(la.y, 0u, 0x4342), __byte_perm(la.z, 0u, 0x4140), __byte_perm(la.z, 0u, 0x4342)};   // byte 11 is a zero pad
    sumL2 += L2[0] + L2[1] + L2[2] + L2[3] + L2[4] + L2[5];
    unsigned ne2[6];                       // (cL - L) per 16-bit lane
#pragma unroll
    for (int i = 0; i < 6; ++i) ne2[i] = __vadd2(cL2p1, ~L2[i]);
    ne2[5] = (ne2[5] & 0xFFFFu) | 0xC0000000u;                    // dummy partner: -16384 => its max() is cR_s
    // window sums of the right strip: V = 5 full pairs (sliding, stride 2), plus the single 11th pixel
    unsigned V0 = R2[0] + R2[2] + R2[4] + R2[6] + R2[8], V1 = R2[1] + R2[3] + R2[5] + R2[7] + R2[9];
#pragma unroll
    for (int s = 0; s < 11; ++s) {
      unsigned& V = (s & 1) ? V1 : V0;
      accW[s] += V + (R2[s + 10] & 0xFFFFu);
      if (s + 2 < 11) V = V - R2[s] + R2[s + 10];
      const unsigned m0 = __viaddmax_s16x2(R2[s], ne2[0], cR2[s]), m1 = __viaddmax_s16x2(R2[s + 2], ne2[1], cR2[s]);
      const unsigned m2 = __viaddmax_s16x2(R2[s + 4], ne2[2], cR2[s]), m3 = __viaddmax_s16x2(R2[s + 6], ne2[3], cR2[s]);
      const unsigned m4 = __viaddmax_s16x2(R2[s + 8], ne2[4], cR2[s]), m5 = __viaddmax_s16x2(R2[s + 10], ne2[5], cR2[s]);
      acc[s] += m0 + m1 + m2 + m3 + m4 + m5;                       // every lane value is in [0, 510]: no carry between lanes
    }
  }
  int T[11];
#pragma unroll
  for (int s = 0; s < 11; ++s) T[s] = 2 * (int)((acc[s] & 0xFFFFu) + (acc[s] >> 16)) - (int)((accW[s] & 0xFFFFu) + (accW[s] >> 16));
  int SL = (int)((sumL2 & 0xFFFFu) + (sumL2 >> 16));
#pragma unroll
  for (int q = 1; q <= 2; q <<= 1) {
#pragma unroll
    for (int s = 0; s < 11; ++s) T[s] += __shfl_xor_sync(0xffffffffu, T[s], q);
    SL += __shfl_xor_sync(0xffffffffu, SL, q);
  }
  if (!gDo || j != 0) return;
  {
    const int Ls = 5;
    const size_t o = pair * A.cap + iL0 + kk;
    float outU = -1.f, outD = -1.f;
    int outSad = -1;
    int sadv[11];
#pragma unroll
    for (int s = 0; s < 11; ++s) sadv[s] = T[s] - 143 * (int)(cR2[s] & 0xFFFFu) - 121 * cL + SL;
    int bestSad = 0x7fffffff, bestinc = 0;
#pragma unroll
    for (int s = 0; s < 11; ++s)
      if (sadv[s] < bestSad) { bestSad = sadv[s]; bestinc = s - Ls; }
    if (!(bestinc == -Ls || bestinc == Ls)) {
      float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int s = 1; s < 10; ++s)
        if (s - Ls == bestinc) { d1 = (float)sadv[s - 1]; d2 = (float)sadv[s]; d3 = (float)sadv[s + 1]; }
      const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
      if (!(deltaR < -1.f || deltaR > 1.f)) {
        float bestuR = __fmul_rn(gScale, __fadd_rn(__fadd_rn(gUR0, (float)bestinc), deltaR));
        float disparity = __fsub_rn(gUL, bestuR);
        if (disparity >= 0.f && disparity < A.maxD) {
          if (disparity <= 0.f) { disparity = 0.01f; bestuR = (float)((double)gUL - 0.01); }
          outD = __fdiv_rn(A.mbf, disparity);
          outU = bestuR;
          outSad = bestSad;
        }
      }
    }
    A.uRight[o] = outU; A.depth[o] = outD; A.sad[o] = outSad;
  }
}

// first bin b with hist[0] + ... + hist[b] > k, and k minus the counts before it; executed by one warp (8 bins per lane)
__device__ __forceinline__ void median_pick_bin(const int* hist, int k, int lane, int& bin, int& rest) {
  int c[8], local = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i] = hist[8 * lane + i]; local += c[i]; }
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const int owner = __ffs(__ballot_sync(0xffffffffu, incl > k)) - 1;      // callers guarantee k < total
  int b = 0, r = k - (incl - local);
  if (lane == owner) {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (b == i && r >= c[i]) { r -= c[i]; ++b; }
    b += 8 * lane;
  }
  bin = __shfl_sync(0xffffffffu, b, owner);
  rest = __shfl_sync(0xffffffffu, r, owner);
}

constexpr int MED_STAGE = 4096;      // SAD values kept in shared memory between the three passes (longer lists are re-read)
__global__ void __launch_bounds__(1024) k_stereo_median(StereoArgs A) {      // 256 threads per pair in batches, 1024 for a handful of frames
  __shared__ int hist[256];
  __shared__ int ssad[MED_STAGE];
  __shared__ int sel[3];   // [0] chosen high byte, [1] rank inside it, [2] median value
  const size_t pair = blockIdx.x;
  const int N = A.nL[pair];
  const int* gsad = A.sad + pair * A.cap;
  const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
  const bool staged = N <= MED_STAGE;
  if (tid < 256) hist[tid] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += nthr) {
    const int s = gsad[i];
    if (staged) ssad[i] = s;
    if (s >= 0) atomicAdd(&hist[(s >> 8) & 0xFF], 1);
  }
  __syncthreads();
  const int* sad = staged ? ssad : gsad;
  if (tid < 32) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += hist[8 * lane + i];
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    int b = -1, k = 0;
    if (t > 0) median_pick_bin(hist, t / 2, lane, b, k);    // vDistIdx[size/2] of the ascending sort
    if (lane == 0) { sel[0] = b; sel[1] = k; }
  }
  __syncthreads();
  const int hb = sel[0];
  float thDist = 3.4e38f;                              // no matches: nothing to filter
  if (hb >= 0) {                                       // block-uniform
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += nthr) {
      const int s = sad[i];
      if (s >= 0 && ((s >> 8) & 0xFF) == hb) atomicAdd(&hist[s & 0xFF], 1);
    }
    __syncthreads();
    if (tid < 32) {
      int b, k;
      median_pick_bin(hist, sel[1], lane, b, k);
      if (lane == 0) sel[2] = (hb << 8) | b;
    }
    __syncthreads();
    const float median = (float)sel[2];
    thDist = __fmul_rn(1.5f * 1.4f, median);
  }
  // outliers out; with host pointers every slot's final value also goes straight into the caller-visible pinned staging (one frame
  // at a time: no device-to-host copy, no second stream to hand over to)
  const int last = A.hostU ? A.cap : N;
  for (int i = tid; i < last; i += nthr) {
    const size_t o = pair * A.cap + i;
    const bool out = i < N && sad[i] >= 0 && !((float)sad[i] < thDist);
    if (out) { A.uRight[o] = -1.f; A.depth[o] = -1.f; }
    if (A.hostU) { A.hostU[o] = out ? -1.f : A.uRight[o]; A.hostD[o] = out ? -1.f : A.depth[o]; }
  }
}

}  // namespace ivg

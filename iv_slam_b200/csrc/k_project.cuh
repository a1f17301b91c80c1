// k_project.cuh — N2: the per-frame Hamming-search consumers of the extractor's output in Track():
//   ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)   src/ORBmatcher.cc:1372-1519
//   ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th)                   src/ORBmatcher.cc:45-133
// with Frame::GetFeaturesInArea (src/Frame.cc:620-668), ComputeThreeMaxima (src/ORBmatcher.cc:1650-1695) and
// DescriptorDistance (:1700-1716), on the keypoints / descriptors / uRight / 64x48 grid of the current frame that are
// already on the device (extraction, stereo matching, N1).
//
// The reference loops are sequential and order dependent: a point whose MapPoint has observations blocks the keypoint
// it takes for every later point.  Split here into
//   k_proj_candidates  warp per point, fully parallel: projection, grid window in the reference's enumeration order
//       (the grid is column-major CSR, so one grid column of the window is one contiguous index range), level / window /
//       stereo tests, Hamming distances; writes the candidate list and the two best (distance, rank) keys — rank in
//       enumeration order reproduces the strict '<' first-wins tie-break and the reference's second-best bookkeeping
//       (second best = second smallest (distance, rank));
//   k_proj_resolve     one warp walks the points in order with the keypoint "blocked" bitmap in shared memory: a
//       point whose two best candidates are still free is decided from registers; only the rare conflicts rescan
//       their list.  Then the rotation histogram vote (round(rot/30), three maxima, 10 % rule).
// Float pin: Rcw*x3Dw+tcw follows cv::gemm's small-matrix path (float products summed left to right, final add in
// double); everything else is evaluated without FMA.
#pragma once
#include "common.cuh"
#include "k_frame.cuh"

namespace ivg {

constexpr int PJ_TH_HIGH = 100, PJ_HISTO = 30;
constexpr unsigned PJ_NONE = (256u << 16) | 0xFFFFu;

struct ProjArgs {
  // current frame (device)
  const uint8_t* kp; const uint8_t* desc; const float* uRight /* may be null */; const int* nPtr; int index, cap;
  const int* gridStart; const int* gridIdx;
  float minX, minY, maxX, maxY, invW, invH;
  float scale[MAX_LEVELS]; int nLevels;
  // points (device copies of the caller's arrays)
  int n;
  const float* world; const float* proj; const float* viewCos; const uint8_t* pdesc; const int* octave; const float* angle;
  const uint8_t* flags; const uint8_t* curBlocked;
  float Rcw[9], tcw[3];
  float fx, fy, cx, cy, mbf;
  int mode;                       // 0: levels nLast-1..nLast+1, 1: forward (>= nLast), 2: backward (<= nLast), 3: local-map variant
  float th, nnratio; int checkOri;
  // work / results
  uint32_t* cand; int candStride; int* candCount;   // [n][candStride]: dist << 16 | keypoint index, in enumeration order
  uint4* tent;                                       // per point: best key, second key, i2a | i2b << 16, levA | levB << 8 | bin << 16 | flags << 24
  int* match; int* nmatches;                         // [cap], scalar
  int8_t* accBin; int* accIdx;                       // [n]
};

__device__ __forceinline__ int pj_hamming(const uint32_t (&a)[8], const uint8_t* b) {
  const uint4 x = __ldg(reinterpret_cast<const uint4*>(b)), y = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  return __popc(a[0] ^ x.x) + __popc(a[1] ^ x.y) + __popc(a[2] ^ x.z) + __popc(a[3] ^ x.w) +
         __popc(a[4] ^ y.x) + __popc(a[5] ^ y.y) + __popc(a[6] ^ y.z) + __popc(a[7] ^ y.w);
}

__device__ __forceinline__ void pj_top2_insert(unsigned& k1, unsigned& k2, unsigned key) {
  if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
}
__device__ __forceinline__ void pj_top2_reduce(unsigned& k1, unsigned& k2) {
#pragma unroll
  for (int s = 16; s; s >>= 1) {
    const unsigned o1 = __shfl_xor_sync(0xffffffffu, k1, s), o2 = __shfl_xor_sync(0xffffffffu, k2, s);
    const unsigned lo = min(k1, o1), hi = max(k1, o1);
    k1 = lo; k2 = min(hi, min(k2, o2));
  }
}

__device__ __forceinline__ int pj_rot_bin(float angLast, float angCur) {
  float rot = __fsub_rn(angLast, angCur);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, 1.0f / PJ_HISTO));
  if (bin == PJ_HISTO) bin = 0;
  return bin;
}

__global__ void __launch_bounds__(256) k_proj_candidates(ProjArgs A) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  if (i >= A.n) return;
  const int N = A.nPtr[A.index];
  const uint8_t fl = A.flags[i];
  unsigned k1 = PJ_NONE, k2 = PJ_NONE;
  int cnt = 0;
  uint32_t* list = A.cand + (size_t)i * A.candStride;
  bool go = (fl & 1) != 0;
  float u = 0.f, v = 0.f, radius = 0.f, xr = 0.f;
  int minLevel = 0, maxLevel = -1;
  if (go) {
    if (A.mode < 3) {
      const float X = A.world[3 * i], Y = A.world[3 * i + 1], Z = A.world[3 * i + 2];
      float c3[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(A.Rcw[3 * r], X), __fmul_rn(A.Rcw[3 * r + 1], Y)), __fmul_rn(A.Rcw[3 * r + 2], Z));
        c3[r] = (float)((double)t0 + (double)A.tcw[r]);
      }
      const float invzc = (float)(1.0 / (double)c3[2]);
      if (invzc < 0) go = false;
      u = __fadd_rn(__fmul_rn(__fmul_rn(A.fx, c3[0]), invzc), A.cx);
      v = __fadd_rn(__fmul_rn(__fmul_rn(A.fy, c3[1]), invzc), A.cy);
      if (!(isfinite(u) && isfinite(v))) go = false;
      if (go && (u < A.minX || u > A.maxX || v < A.minY || v > A.maxY)) go = false;
      const int oct = A.octave[i];
      if (oct < 0 || oct >= A.nLevels) go = false;
      if (go) {
        radius = __fmul_rn(A.th, A.scale[oct]);
        xr = __fsub_rn(u, __fmul_rn(A.mbf, invzc));
        if (A.mode == 1) { minLevel = oct; maxLevel = -1; }
        else if (A.mode == 2) { minLevel = 0; maxLevel = oct; }
        else { minLevel = oct - 1; maxLevel = oct + 1; }
      }
    } else {
      u = A.proj[3 * i]; v = A.proj[3 * i + 1]; xr = A.proj[3 * i + 2];
      const int lev = A.octave[i];
      if (lev < 0 || lev >= A.nLevels) go = false;
      if (go) {
        float r = (double)A.viewCos[i] > 0.998 ? 2.5f : 4.0f;
        if (A.th != 1.0f) r = __fmul_rn(r, A.th);
        radius = __fmul_rn(r, A.scale[lev]);
        minLevel = lev - 1; maxLevel = lev;
      }
    }
  }
  if (go) {
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(u, A.minX), radius), A.invW)));
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(u, A.minX), radius), A.invW)));
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(v, A.minY), radius), A.invH)));
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(v, A.minY), radius), A.invH)));
    if (nMinCellX >= GRID_COLS || nMaxCellX < 0 || nMinCellY >= GRID_ROWS || nMaxCellY < 0) go = false;
    if (go) {
      const bool checkLevels = minLevel > 0 || maxLevel >= 0;
      uint32_t dl[8];
      {
        const uint4* d = reinterpret_cast<const uint4*>(A.pdesc + (size_t)i * 32);
        const uint4 a = __ldg(d), b = __ldg(d + 1);
        dl[0] = a.x; dl[1] = a.y; dl[2] = a.z; dl[3] = a.w; dl[4] = b.x; dl[5] = b.y; dl[6] = b.z; dl[7] = b.w;
      }
      const uint8_t* kp0 = A.kp + (size_t)A.index * A.cap * 28;
      const uint8_t* dc0 = A.desc + (size_t)A.index * A.cap * 32;
      const float* ur0 = A.uRight ? A.uRight + (size_t)A.index * A.cap : nullptr;
      const int* gs = A.gridStart + (size_t)A.index * (GRID_COLS * GRID_ROWS + 1);
      const int* gi = A.gridIdx + (size_t)A.index * A.cap;
      const unsigned ltmask = (1u << lane) - 1u;
      for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
        const int a = __ldg(gs + ix * GRID_ROWS + nMinCellY), b = __ldg(gs + ix * GRID_ROWS + nMaxCellY + 1);
        for (int j0 = a; j0 < b; j0 += 32) {
          const int j = j0 + lane;
          bool pass = false;
          int idx = 0;
          if (j < b) {
            idx = __ldg(gi + j);
            const float* k = reinterpret_cast<const float*>(kp0 + (size_t)idx * 28);
            const int oct = reinterpret_cast<const int*>(k)[5];
            pass = idx < N;
            if (checkLevels && (oct < minLevel || (maxLevel >= 0 && oct > maxLevel))) pass = false;
            const float distx = __fsub_rn(k[0], u), disty = __fsub_rn(k[1], v);
            if (!(fabsf(distx) < radius && fabsf(disty) < radius)) pass = false;
            if (pass && A.curBlocked && A.curBlocked[idx]) pass = false;
            if (pass && ur0) {
              const float uR = ur0[idx];
              if (uR > 0 && fabsf(__fsub_rn(xr, uR)) > radius) pass = false;
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, pass);
          if (pass) {
            const int pos = cnt + __popc(m & ltmask);
            const int dist = pj_hamming(dl, dc0 + (size_t)idx * 32);
            list[pos] = ((unsigned)dist << 16) | (unsigned)idx;
            pj_top2_insert(k1, k2, ((unsigned)dist << 16) | (unsigned)pos);
          }
          cnt += __popc(m);
        }
      }
    }
  }
  pj_top2_reduce(k1, k2);
  __syncwarp();
  if (lane == 0) {
    unsigned ia = 0xFFFFu, ib = 0xFFFFu, levA = 0xFF, levB = 0xFF, bin = 0xFF;
    const uint8_t* kp0 = A.kp + (size_t)A.index * A.cap * 28;
    if (k1 != PJ_NONE) {
      ia = list[k1 & 0xFFFFu] & 0xFFFFu;
      const float* k = reinterpret_cast<const float*>(kp0 + (size_t)ia * 28);
      levA = (unsigned)reinterpret_cast<const int*>(k)[5] & 0xFFu;
      if (A.mode < 3 && A.checkOri) bin = (unsigned)pj_rot_bin(A.angle[i], k[3]);
    }
    if (k2 != PJ_NONE) {
      ib = list[k2 & 0xFFFFu] & 0xFFFFu;
      levB = (unsigned)reinterpret_cast<const int*>(kp0 + (size_t)ib * 28)[5] & 0xFFu;
    }
    A.candCount[i] = cnt;
    A.tent[i] = make_uint4(k1, k2, ia | (ib << 16), levA | (levB << 8) | (bin << 16) | ((unsigned)fl << 24));
  }
}

__global__ void __launch_bounds__(32) k_proj_resolve(ProjArgs A) {
  __shared__ uint32_t blocked[2048];          // one bit per current keypoint (cap <= 65535)
  __shared__ int hist[PJ_HISTO];
  const int lane = threadIdx.x;
  const int N = A.nPtr[A.index];
  for (int w = lane; w < 2048; w += 32) blocked[w] = 0;
  if (lane < PJ_HISTO) hist[lane] = 0;
  for (int c = lane; c < A.cap; c += 32) A.match[c] = -1;
  __syncwarp();
  const uint8_t* kp0 = A.kp + (size_t)A.index * A.cap * 28;
  int nm = 0;
  for (int base = 0; base < A.n; base += 32) {
    const int me = base + lane;
    uint4 t = make_uint4(PJ_NONE, PJ_NONE, 0xFFFFFFFFu, 0xFFFFFFFFu);
    if (me < A.n) t = A.tent[me];
    if (me < A.n) A.accBin[me] = -1;
    unsigned todo = __ballot_sync(0xffffffffu, t.x != PJ_NONE);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int i = base + l;
      unsigned k1 = __shfl_sync(0xffffffffu, t.x, l), k2 = __shfl_sync(0xffffffffu, t.y, l);
      const unsigned ii = __shfl_sync(0xffffffffu, t.z, l), meta = __shfl_sync(0xffffffffu, t.w, l);
      unsigned ia = ii & 0xFFFFu, ib = ii >> 16;
      unsigned levA = meta & 0xFFu, levB = (meta >> 8) & 0xFFu, bin = (meta >> 16) & 0xFFu;
      const unsigned fl = meta >> 24;
      const bool dirty = ((blocked[ia >> 5] >> (ia & 31)) & 1u) || (k2 != PJ_NONE && ((blocked[ib >> 5] >> (ib & 31)) & 1u));
      if (dirty) {   // rare: a keypoint this point wanted was taken by an earlier point with observations — rescan its list
        const uint32_t* list = A.cand + (size_t)i * A.candStride;
        const int cnt = A.candCount[i];
        k1 = k2 = PJ_NONE;
        for (int j = lane; j < cnt; j += 32) {
          const unsigned e = list[j], idx = e & 0xFFFFu;
          if (!((blocked[idx >> 5] >> (idx & 31)) & 1u)) pj_top2_insert(k1, k2, (e & 0xFFFF0000u) | (unsigned)j);
        }
        pj_top2_reduce(k1, k2);
        ia = ib = 0xFFFFu; levA = levB = 0xFF; bin = 0xFF;
        if (k1 != PJ_NONE) {
          ia = list[k1 & 0xFFFFu] & 0xFFFFu;
          const float* k = reinterpret_cast<const float*>(kp0 + (size_t)ia * 28);
          levA = (unsigned)reinterpret_cast<const int*>(k)[5] & 0xFFu;
          if (A.mode < 3 && A.checkOri) bin = (unsigned)pj_rot_bin(A.angle[i], k[3]);
        }
        if (k2 != PJ_NONE) {
          ib = list[k2 & 0xFFFFu] & 0xFFFFu;
          levB = (unsigned)reinterpret_cast<const int*>(kp0 + (size_t)ib * 28)[5] & 0xFFu;
        }
      }
      const int bestDist = (int)(k1 >> 16), bestDist2 = (int)(k2 >> 16);
      bool accept = k1 != PJ_NONE && bestDist <= PJ_TH_HIGH;
      if (accept && A.mode == 3) {
        // bestLevel == bestLevel2 (-1 == -1 never happens here: a best exists) && bestDist > mfNNratio * bestDist2
        const bool sameLevel = k2 != PJ_NONE && levA == levB;
        if (sameLevel && (float)bestDist > __fmul_rn(A.nnratio, (float)bestDist2)) accept = false;
      }
      if (accept && lane == 0) {
        A.match[ia] = i;
        if (fl & 2) blocked[ia >> 5] |= 1u << (ia & 31);
        if (A.mode < 3 && A.checkOri) { A.accBin[i] = (int8_t)bin; A.accIdx[i] = (int)ia; hist[bin]++; }
      }
      nm += accept ? 1 : 0;
      __syncwarp();
    }
  }
  __syncwarp();
  if (A.mode < 3 && A.checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    {   // ComputeThreeMaxima, evaluated redundantly by every lane
      int max1 = 0, max2 = 0, max3 = 0;
      for (int b = 0; b < PJ_HISTO; ++b) {
        const int s = hist[b];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
        else if (s > max3) { max3 = s; ind3 = b; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
    }
    int removed = 0;
    for (int i = lane; i < A.n; i += 32) {
      const int b = A.accBin[i];
      if (b >= 0 && b != ind1 && b != ind2 && b != ind3) { A.match[A.accIdx[i]] = -1; ++removed; }
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, s);
    nm -= removed;
  }
  if (lane == 0) *A.nmatches = nm;
  (void)N;
}

}  // namespace ivg

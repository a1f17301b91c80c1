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
constexpr int PJ_TILE = 64;

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
  uint2* cand; int candStride; int* candCount;      // [n][candStride]: (dist << 16 | keypoint index, octave | rotation bin << 8), in enumeration order
  uint4* tent;                                       // per point, two records: {key1, key2, key3, count | flags << 24} and
                                                     // {idx1 | idx2 << 16, idx3 | lev1 << 16 | lev2 << 24, lev3 | bin1 << 8 | bin2 << 16 | bin3 << 24, 0}
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
__device__ __forceinline__ void pj_top3_insert(unsigned& k1, unsigned& k2, unsigned& k3, unsigned key) {
  if (key < k1) { k3 = k2; k2 = k1; k1 = key; } else if (key < k2) { k3 = k2; k2 = key; } else if (key < k3) k3 = key;
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
  unsigned k1 = PJ_NONE, k2 = PJ_NONE, k3 = PJ_NONE;
  int cnt = 0;
  uint2* list = A.cand + (size_t)i * A.candStride;
  const bool ori = A.mode < 3 && A.checkOri;
  const float angI = ori ? A.angle[i] : 0.f;
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
          unsigned meta = 0xFF00u;
          if (j < b) {
            idx = __ldg(gi + j);
            const float* k = reinterpret_cast<const float*>(kp0 + (size_t)idx * 28);
            const int oct = reinterpret_cast<const int*>(k)[5];
            pass = idx < N;
            meta = ((unsigned)oct & 0xFFu) | ((ori ? (unsigned)pj_rot_bin(angI, k[3]) : 0xFFu) << 8);
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
            list[pos] = make_uint2(((unsigned)dist << 16) | (unsigned)idx, meta);
            pj_top3_insert(k1, k2, k3, ((unsigned)dist << 16) | (unsigned)pos);
          }
          cnt += __popc(m);
        }
      }
    }
  }
  // the three smallest keys of the warp: three rounds of "global minimum, its owner pops"
  unsigned K[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    unsigned m = k1;
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, sft));
    K[r] = m;
    if (m != PJ_NONE && k1 == m) { k1 = k2; k2 = k3; k3 = PJ_NONE; }      // keys are unique: exactly one owner
  }
  __syncwarp();
  if (lane == 0) {
    unsigned idx[3] = {0xFFFFu, 0xFFFFu, 0xFFFFu}, lev[3] = {0xFF, 0xFF, 0xFF}, bin[3] = {0xFF, 0xFF, 0xFF};
#pragma unroll
    for (int r = 0; r < 3; ++r)
      if (K[r] != PJ_NONE) { const uint2 e = list[K[r] & 0xFFFFu]; idx[r] = e.x & 0xFFFFu; lev[r] = e.y & 0xFFu; bin[r] = (e.y >> 8) & 0xFFu; }
    A.candCount[i] = cnt;
    A.tent[2 * i] = make_uint4(K[0], K[1], K[2], (unsigned)min(cnt, 0xFFFF) | ((unsigned)fl << 24));
    A.tent[2 * i + 1] = make_uint4(idx[0] | (idx[1] << 16), idx[2] | (lev[0] << 16) | (lev[1] << 24), lev[2] | (bin[0] << 8) | (bin[1] << 16) | (bin[2] << 24), 0u);
  }
}

__global__ void __launch_bounds__(32) k_proj_resolve(ProjArgs A) {
  __shared__ uint32_t blocked[2048];          // one bit per current keypoint (cap <= 65535)
  __shared__ uint32_t mark[2048];             // scratch: keypoints taken by blocking points of the current chunk
  __shared__ uint2 tile[32][PJ_TILE];         // the first PJ_TILE candidates of every point of the chunk that has to be replayed
  __shared__ int hist[PJ_HISTO];
  const int lane = threadIdx.x;
  for (int w = lane; w < 2048; w += 32) { blocked[w] = 0; mark[w] = 0; }
  if (lane < PJ_HISTO) hist[lane] = 0;
  for (int c = lane; c < A.cap; c += 32) A.match[c] = -1;
  __syncwarp();
  const bool ori = A.mode < 3 && A.checkOri;
  int nm = 0;
  auto is_blocked = [&](unsigned idx) { return ((blocked[idx >> 5] >> (idx & 31)) & 1u) != 0u; };
  auto decide = [&](unsigned k1, unsigned k2, unsigned levA, unsigned levB) {
    bool acc = k1 != PJ_NONE && (int)(k1 >> 16) <= PJ_TH_HIGH;
    // local map: bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2  => no match
    if (acc && A.mode == 3 && k2 != PJ_NONE && levA == levB && (float)(int)(k1 >> 16) > __fmul_rn(A.nnratio, (float)(int)(k2 >> 16))) acc = false;
    return acc;
  };

  const uint4 kNoneA = make_uint4(PJ_NONE, PJ_NONE, PJ_NONE, 0u), kNoneB = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);
  uint4 taNext = kNoneA, tbNext = kNoneB;
  if (lane < A.n) { taNext = A.tent[2 * lane]; tbNext = A.tent[2 * lane + 1]; }
  for (int base = 0; base < A.n; base += 32) {
    const int me = base + lane;
    const uint4 ta = taNext, tb = tbNext;       // the next chunk's records are fetched while this one is resolved
    taNext = kNoneA; tbNext = kNoneB;
    if (me + 32 < A.n) { taNext = A.tent[2 * (me + 32)]; tbNext = A.tent[2 * (me + 32) + 1]; }
    if (me < A.n) A.accBin[me] = -1;
    const bool valid = ta.x != PJ_NONE;
    const unsigned fl = ta.w >> 24;
    const int myCnt = (int)(ta.w & 0xFFFFu) == 0xFFFF ? A.candCount[min(me, A.n - 1)] : (int)(ta.w & 0xFFFFu);
    // The point's three best candidates against the keypoints blocked so far: the first free one is its best, the
    // second free one its second best (needed by the local-map ratio test only).  That is exact as long as enough of
    // the three are free, or the three are the whole list; otherwise the point is replayed from its full list.
    const unsigned key[3] = {ta.x, ta.y, ta.z};
    const unsigned idx[3] = {tb.x & 0xFFFFu, tb.x >> 16, tb.y & 0xFFFFu};
    const unsigned lev[3] = {(tb.y >> 16) & 0xFFu, tb.y >> 24, tb.z & 0xFFu};
    const unsigned bins[3] = {(tb.z >> 8) & 0xFFu, (tb.z >> 16) & 0xFFu, tb.z >> 24};
    unsigned kb = PJ_NONE, ks = PJ_NONE, ia = 0xFFFFu, ib = 0xFFFFu, levb = 0xFF, levs = 0xFF, binb = 0xFF;
    int nfree = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (key[r] != PJ_NONE && !is_blocked(idx[r])) {
        if (nfree == 0) { kb = key[r]; ia = idx[r]; levb = lev[r]; binb = bins[r]; }
        else if (nfree == 1) { ks = key[r]; ib = idx[r]; levs = lev[r]; }
        ++nfree;
      }
    }
    const bool exhausted = myCnt <= 3;
    const bool known = exhausted || nfree >= (A.mode == 3 ? 2 : 1);
    const bool hasB = A.mode == 3 && ks != PJ_NONE;     // the second best only matters for the local-map ratio test
    const bool acc = valid && known && decide(kb, ks, levb, levs);
    const bool blocking = acc && (fl & 2);
    const bool hasA = valid && known && kb != PJ_NONE;
    // A point is "clean" when its decision cannot depend on the other points of the chunk: nobody else in the chunk has
    // the same best keypoint and (local map) its second best is not taken by a blocking point of the chunk.  Clean
    // points commit in parallel; the others are replayed one at a time, in order.
    bool slow = valid && !known;
    const unsigned peers = __match_any_sync(0xffffffffu, hasA ? ia : (0x10000u | (unsigned)lane));
    if (hasA && (peers & ~(1u << lane))) slow = true;
    if (A.mode == 3) {
      if (blocking) atomicOr(&mark[ia >> 5], 1u << (ia & 31));
      __syncwarp();
      if (hasA && hasB && ((mark[ib >> 5] >> (ib & 31)) & 1u)) slow = true;
      __syncwarp();
      if (blocking) mark[ia >> 5] = 0;
      __syncwarp();
    }
    bool done = !valid;
    // stage the candidate lists of the points that will be replayed: all loads are issued before the first store so the
    // chunk pays one global-memory latency, not one per replayed point
    const unsigned pref = __ballot_sync(0xffffffffu, slow);
    for (unsigned rest = pref; rest;) {
      int sl[4], sc[4];
      uint2 ra[4], rb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        sl[q] = rest ? __ffs(rest) - 1 : -1;
        if (rest) rest &= rest - 1;
        sc[q] = sl[q] >= 0 ? __shfl_sync(0xffffffffu, myCnt, sl[q]) : 0;
        const uint2* L = A.cand + (size_t)(base + max(sl[q], 0)) * A.candStride;
        ra[q] = lane < sc[q] ? L[lane] : make_uint2(0, 0);
        rb[q] = lane + 32 < sc[q] ? L[lane + 32] : make_uint2(0, 0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (sl[q] >= 0) { tile[sl[q]][lane] = ra[q]; tile[sl[q]][lane + 32] = rb[q]; }
    }
    __syncwarp();
    for (;;) {
      const unsigned sm = __ballot_sync(0xffffffffu, slow && !done);
      const int first = sm ? __ffs(sm) - 1 : 32;
      if (!slow && !done && lane < first) {           // commit the clean points that precede the next replayed one
        if (acc) {
          A.match[ia] = me;
          if (blocking) atomicOr(&blocked[ia >> 5], 1u << (ia & 31));
          if (ori) { A.accBin[me] = (int8_t)binb; A.accIdx[me] = (int)ia; atomicAdd(&hist[binb], 1); }
        }
        done = true;
        nm += acc ? 1 : 0;
      }
      __syncwarp();
      if (!sm) break;
      // replay point base + first against the current blocked set
      const int i = base + first;
      const int cnt = __shfl_sync(0xffffffffu, myCnt, first);
      const unsigned pfl = __shfl_sync(0xffffffffu, fl, first);
      const uint2* list = A.cand + (size_t)i * A.candStride;
      const bool staged = (pref >> first) & 1u;
      unsigned k1 = PJ_NONE, k2 = PJ_NONE;
      uint2 e1 = make_uint2(0, 0), e2 = make_uint2(0, 0);
      for (int j = lane; j < cnt; j += 32) {
        const uint2 e = (staged && j < PJ_TILE) ? tile[first][j] : list[j];
        if (!is_blocked(e.x & 0xFFFFu)) {
          const unsigned key = (e.x & 0xFFFF0000u) | (unsigned)j;
          if (key < k1) { k2 = k1; e2 = e1; k1 = key; e1 = e; } else if (key < k2) { k2 = key; e2 = e; }
        }
      }
      unsigned K1 = k1, K2 = k2;
      pj_top2_reduce(K1, K2);
      unsigned nia = 0xFFFFu, nlevA = 0xFF, nlevB = 0xFF, nbin = 0xFF;
      if (K1 != PJ_NONE) {
        const int o = __ffs(__ballot_sync(0xffffffffu, k1 == K1)) - 1;      // the global best is some lane's own best
        const unsigned ex = __shfl_sync(0xffffffffu, e1.x, o), ey = __shfl_sync(0xffffffffu, e1.y, o);
        nia = ex & 0xFFFFu; nlevA = ey & 0xFFu; nbin = (ey >> 8) & 0xFFu;
      }
      if (K2 != PJ_NONE) {
        const unsigned m1 = __ballot_sync(0xffffffffu, k1 == K2), m2 = __ballot_sync(0xffffffffu, k2 == K2);
        if (m1) nlevB = __shfl_sync(0xffffffffu, e1.y, __ffs(m1) - 1) & 0xFFu;
        else nlevB = __shfl_sync(0xffffffffu, e2.y, __ffs(m2) - 1) & 0xFFu;
      }
      const bool accept = decide(K1, K2, nlevA, nlevB);
      const bool blk = accept && (pfl & 2);
      if (lane == first) {
        if (accept) {
          A.match[nia] = i;
          if (blk) blocked[nia >> 5] |= 1u << (nia & 31);
          if (ori) { A.accBin[i] = (int8_t)nbin; A.accIdx[i] = (int)nia; hist[nbin]++; }
        }
        done = true;
      }
      nm += (accept && lane == first) ? 1 : 0;
      // a keypoint newly blocked here may be what a later, so far clean, point of the chunk wanted
      if (blk && !done && lane > first && hasA && (ia == nia || (hasB && ib == nia))) slow = true;
      __syncwarp();
    }
  }
  __syncwarp();
#pragma unroll
  for (int s = 16; s; s >>= 1) nm += __shfl_xor_sync(0xffffffffu, nm, s);
  if (ori) {
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
}

}  // namespace ivg

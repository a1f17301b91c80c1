// k_bow.cuh — N2: ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
// (introspective_ORB_SLAM/src/ORBmatcher.cc:165-294) on the keypoints / descriptors of frame F that are on the device.
//
// The reference walks the shared vocabulary nodes; inside a node every key-frame keypoint takes the best still-free
// keypoint of F's list for that node (bestDist1 <= TH_LOW and bestDist1 < mfNNratio * bestDist2), in order.  The loops are
// sequential, but a keypoint of F belongs to exactly one node, so NODES ARE INDEPENDENT: one warp per node replays its
// points in order, lanes = the node's candidate keypoints (strict '<' first-wins ties = minimum of (distance, list
// position)); then the rotation-histogram vote over all nodes (ComputeThreeMaxima, :1650-1695).
#pragma once
#include "common.cuh"
#include "k_project.cuh"

namespace ivg {

constexpr int PJ_TH_LOW = 50;      // ORBmatcher::TH_LOW, src/ORBmatcher.cc:38

struct BowArgs {
  const uint8_t* kp; const uint8_t* desc; const int* nPtr; int index, cap;      // frame F (device results of a handle)
  int n; const uint8_t* pdesc; const float* angle; const uint8_t* flags;        // key-frame points, caller's order
  int nNodes; const int* nodeStart; const int* nodeIdx;                         // F's node lists (CSR)
  const int* ptStart; const int* ptIdx;                                         // points per node slot, in the caller's order (CSR)
  float nnratio; int checkOri;
  int* match; int* nmatches; int* hist;                                         // [cap], scalar, [PJ_HISTO]
  int8_t* accBin; int* accIdx;                                                  // [n]
};

constexpr int BOW_CACHE = 128;     // candidates per node staged in shared memory

__global__ void __launch_bounds__(256) k_bow_match(BowArgs A) {
  __shared__ uint4 sDesc[8][BOW_CACHE][2];
  __shared__ int sIdx[8][BOW_CACHE];
  __shared__ uint32_t sTaken[8][BOW_CACHE / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * 8 + warp;
  if (s >= A.nNodes) return;
  const int N = A.nPtr[A.index];
  const int ca = A.nodeStart[s], cb = A.nodeStart[s + 1];
  const int pa = A.ptStart[s], pb = A.ptStart[s + 1];
  const uint8_t* kp0 = A.kp + (size_t)A.index * A.cap * 28;
  const uint8_t* dc0 = A.desc + (size_t)A.index * A.cap * 32;
  volatile int* match = A.match;
  int nm = 0;
  // Lists of up to BOW_CACHE keypoints are staged once per node in shared memory (index + descriptor) and "taken" is a
  // per-warp bitmask: the per-point loop then touches no global memory except the point's own record.  Longer lists
  // are re-read from global memory tile by tile.
  const int cnt = cb - ca;
  const bool cached = cnt <= BOW_CACHE;
  if (cached) {
    for (int j = lane; j < cnt; j += 32) {
      const int idx = __ldg(A.nodeIdx + ca + j);
      const bool ok = idx >= 0 && idx < N;
      sIdx[warp][j] = ok ? idx : -1;
      if (ok) {
        const uint4* d = reinterpret_cast<const uint4*>(dc0 + (size_t)idx * 32);
        sDesc[warp][j][0] = __ldg(d); sDesc[warp][j][1] = __ldg(d + 1);
      }
    }
    if (lane < BOW_CACHE / 32) sTaken[warp][lane] = 0u;
  }
  __syncwarp();
  // the next point's record is fetched while the current one is matched
  int ni = -1;
  unsigned nfl = 0;
  uint4 nda = make_uint4(0, 0, 0, 0), ndb = nda;
  auto fetch = [&](int p) {
    if (p < pb) {
      ni = A.ptIdx[p]; nfl = A.flags[ni];
      const uint4* d = reinterpret_cast<const uint4*>(A.pdesc + (size_t)ni * 32);
      nda = __ldg(d); ndb = __ldg(d + 1);
    }
  };
  fetch(pa);
  for (int p = pa; p < pb; ++p) {
    const int i = ni;
    const unsigned fl = nfl;
    const uint32_t dl[8] = {nda.x, nda.y, nda.z, nda.w, ndb.x, ndb.y, ndb.z, ndb.w};
    fetch(p + 1);
    if (!(fl & 1)) continue;                            // !pMP || pMP->isBad()
    unsigned k1 = PJ_NONE, k2 = PJ_NONE;
    int myIdx = -1;                                      // keypoint behind this lane's own best key
    if (cached) {
      for (int j0 = 0; j0 < cnt; j0 += 32) {
        const int j = j0 + lane;
        if (j < cnt) {
          const int idx = sIdx[warp][j];
          if (idx >= 0 && !((sTaken[warp][j0 >> 5] >> lane) & 1u)) {
            const uint4 x = sDesc[warp][j][0], y = sDesc[warp][j][1];
            const int dist = __popc(dl[0] ^ x.x) + __popc(dl[1] ^ x.y) + __popc(dl[2] ^ x.z) + __popc(dl[3] ^ x.w) +
                             __popc(dl[4] ^ y.x) + __popc(dl[5] ^ y.y) + __popc(dl[6] ^ y.z) + __popc(dl[7] ^ y.w);
            const unsigned key = ((unsigned)dist << 16) | (unsigned)j;
            if (key < k1) { k2 = k1; k1 = key; myIdx = idx; } else if (key < k2) k2 = key;
          }
        }
      }
    } else {
      for (int j0 = ca; j0 < cb; j0 += 32) {
        const int j = j0 + lane;
        if (j < cb) {
          const int idx = __ldg(A.nodeIdx + j);
          if (idx >= 0 && idx < N && match[idx] < 0) {   // vpMapPointMatches[realIdxF] still NULL (only this lane ever sets it)
            const unsigned key = ((unsigned)pj_hamming(dl, dc0 + (size_t)idx * 32) << 16) | (unsigned)(j - ca);
            if (key < k1) { k2 = k1; k1 = key; myIdx = idx; } else if (key < k2) k2 = key;
          }
        }
      }
    }
    unsigned K1 = k1, K2 = k2;
    pj_top2_reduce(K1, K2);
    const int d1 = (int)(K1 >> 16), d2 = (int)(K2 >> 16);   // 256 when there is no (second) candidate
    if (K1 != PJ_NONE && d1 <= PJ_TH_LOW && (float)d1 < __fmul_rn(A.nnratio, (float)d2)) {
      if (k1 == K1) {                                    // keys are unique: exactly one lane owns the winner
        match[myIdx] = i;
        if (cached) { const int j = (int)(K1 & 0xFFFFu); sTaken[warp][j >> 5] |= 1u << (j & 31); }   // j & 31 == lane: only this lane touches the bit
        if (A.checkOri) {
          const int bin = pj_rot_bin(A.angle[i], reinterpret_cast<const float*>(kp0 + (size_t)myIdx * 28)[3]);
          A.accBin[i] = (int8_t)bin; A.accIdx[i] = myIdx;
          atomicAdd(A.hist + bin, 1);
        }
      }
      ++nm;
    }
    __syncwarp();
  }
  if (lane == 0 && nm) atomicAdd(A.nmatches, nm);
}

// rotation consistency (:266-291): keep the three fullest bins (10 % rule), drop the matches of the others
__global__ void __launch_bounds__(256) k_bow_finish(BowArgs A) {
  __shared__ int removed;
  if (threadIdx.x == 0) removed = 0;
  __syncthreads();
  if (A.checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    int max1 = 0, max2 = 0, max3 = 0;
    for (int b = 0; b < PJ_HISTO; ++b) {
      const int s = A.hist[b];
      if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
      else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
      else if (s > max3) { max3 = s; ind3 = b; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
    int mine = 0;
    for (int i = threadIdx.x; i < A.n; i += 256) {
      const int b = A.accBin[i];
      if (b >= 0 && b != ind1 && b != ind2 && b != ind3) { A.match[A.accIdx[i]] = -1; ++mine; }
    }
    if (mine) atomicAdd(&removed, mine);
  }
  __syncthreads();
  if (threadIdx.x == 0) *A.nmatches -= removed;
}

}  // namespace ivg

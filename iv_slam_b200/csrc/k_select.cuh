// k_select.cuh — K3: keypoint selection of ComputeKeyPointsOld, the LIVE selection path of the reference
// (introspective_ORB_SLAM/src/ORBextractor.cc:880-1213; operator() calls it at :1248, the OctTree variant at :1247
// is commented out — SURVEY F1).
//
//   (k_fast_cells in k_fast.cuh produced, per cell: the row-major corner list, the corner counts at iniThFAST /
//    minThFAST for the "<= 3 keypoints => retry with minThFAST" rule (:1045-1052, post-NMS counts) and the cost-map
//    window sum for IV-SLAM's introspection weighting (:976-984).)
//   k_level_select one CTA per (level, frame): thread 0 replays the sequential budget logic (cell weights :942-987,
//                 per-cell budgets :1028-1031/:1085-1096, the one-shot redistribution loop :1101-1133);
//                 each warp then trims cells with the replayed std::nth_element (retainBest + resize, :1146-1148),
//                 and thread 0 trims the level (:1162-1166).  Output order = reference order.
// Deterministic: no atomics decide any order.
#pragma once
#include "common.cuh"
#include "introselect.h"

namespace ivg {

constexpr int SEL_MAX_CELLS = 1024;      // cells per level the select kernel supports (host check)
constexpr int SEL_WARPS = 8;
// Shared memory is sized per handle (dynamic): FrameSet::selLevelCap level-list entries, selCellCap entries per warp for
// a cell list, selCells per-cell scalars.  Lists longer than the caps are processed in global memory (same code).

// response weight of IV-SLAM's introspection: 2 * (1/(1 + cost/255)) - 1, all float (src/ORBextractor.cc:1070-1071)
__device__ __forceinline__ float introspection_weight(float cost) {
  const float q = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fdiv_rn(cost, 255.0f)));
  return __fsub_rn(__fmul_rn(2.0f, q), 1.0f);
}

struct SelShared {           // views into the dynamic shared memory block
  SelItem* levelBuf;
  SelItem* cellBuf;          // [SEL_WARPS][selCellCap]
  int* nTotal; int* nStored; int* nRetain; int* prefix;
  float* nfc;
  unsigned char* thr; unsigned char* noMore;
};

__host__ __device__ inline size_t sel_smem_bytes(int levelCap, int cellCap, int cells) {
  return (size_t)8 * levelCap + (size_t)8 * SEL_WARPS * cellCap + (size_t)cells * (5 * 4 + 2) + 16;
}

__global__ void __launch_bounds__(SEL_WARPS * 32) k_level_select(FrameSet fs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sTotal, sCount;
  const int SEL_LEVEL_CAP = fs.selLevelCap, SEL_CELL_CAP = fs.selCellCap;
  SelShared S;
  {
    unsigned char* p = smem_raw;
    S.levelBuf = reinterpret_cast<SelItem*>(p); p += (size_t)8 * SEL_LEVEL_CAP;
    S.cellBuf = reinterpret_cast<SelItem*>(p); p += (size_t)8 * SEL_WARPS * SEL_CELL_CAP;
    const int nc = fs.selCells;
    S.nTotal = reinterpret_cast<int*>(p); p += 4 * nc;
    S.nStored = reinterpret_cast<int*>(p); p += 4 * nc;
    S.nRetain = reinterpret_cast<int*>(p); p += 4 * nc;
    S.prefix = reinterpret_cast<int*>(p); p += 4 * nc;
    S.nfc = reinterpret_cast<float*>(p); p += 4 * nc;
    S.thr = p; p += nc;
    S.noMore = p;
  }
  const int level = blockIdx.x;
  const size_t img = blockIdx.y;
  const LevelDev& L = fs.lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nCells = L.nCells;
  const int2* cc = fs.cellCount + img * fs.nCellsTotal + L.cellBase;
  const CellDev* cells = fs.cells + L.cellBase;

  for (int c = tid; c < nCells; c += blockDim.x) {
    const int2 v = cc[c];                         // x: corners at minTh, y: corners at iniTh
    const bool retry = v.y <= 3;                  // cellKeyPoints.size() <= 3 after FAST(iniTh)  (:1047)
    S.thr[c] = (unsigned char)(retry ? fs.minTh : fs.iniTh);
    S.nTotal[c] = retry ? v.x : v.y;
    S.nStored[c] = fs.iniTh < fs.minTh ? v.y : v.x;   // entries in the stored list (corners at scoreTh)
    S.nfc[c] = (float)L.nfeaturesCell;
  }
  __syncthreads();

  if (tid == 0) {
    const int nDesired = L.nDesired;
    if (fs.weighted) {
      // cell weights, row-major, float accumulation order as in the reference (:942-987)
      const uint32_t* cost = fs.cellCost + img * fs.nCellsTotal + L.cellBase;
      float wsum = 0.0f;
      for (int c = 0; c < nCells; ++c) {
        const float area = __fmul_rn((float)cells[c].ww, (float)cells[c].wh);
        const float mean = __fdiv_rn((float)cost[c], area);
        const float qs = (float)(1.0 / (1.0 + (double)__fdiv_rn(mean, 255.0f)));
        const float qn = __fsub_rn(__fmul_rn(2.0f, qs), 1.0f);
        S.nfc[c] = qn;                            // parked; converted to a budget below
        wsum = __fadd_rn(wsum, qn);
      }
      for (int c = 0; c < nCells; ++c) {
        const float v = ceilf(__fdiv_rn(__fmul_rn((float)nDesired, S.nfc[c]), wsum));
        S.nfc[c] = (1.0f < v) ? v : 1.0f;         // std::max(1.0f, v); NaN -> 1
      }
    }
    int nNoMore = 0, nToDistribute = 0;
    for (int c = 0; c < nCells; ++c) {            // :1082-1097
      const int nKeys = S.nTotal[c];
      const float f = S.nfc[c];
      if ((float)nKeys > f) { S.nRetain[c] = (int)f; S.noMore[c] = 0; }
      else {
        S.nRetain[c] = nKeys;
        nToDistribute = (int)__fadd_rn((float)nToDistribute, __fsub_rn(f, (float)nKeys));
        S.noMore[c] = 1; nNoMore++;
      }
    }
    if (nToDistribute > 0 && nNoMore < nCells) {  // the while loop runs exactly once (:1103-1133)
      for (int c = 0; c < nCells; ++c) {
        if (S.noMore[c]) continue;
        const int nNew = (int)__fadd_rn(S.nfc[c], ceilf(__fdiv_rn((float)nToDistribute, (float)(nCells - nNoMore))));
        if (S.nTotal[c] > nNew) S.nRetain[c] = nNew;
        else { S.nRetain[c] = S.nTotal[c]; nToDistribute += nNew - S.nTotal[c]; S.noMore[c] = 1; nNoMore++; }
      }
    }
    int run = 0;
    for (int c = 0; c < nCells; ++c) { S.prefix[c] = run; run += min(max(S.nRetain[c], 0), S.nTotal[c]); }
    sTotal = run;
  }
  __syncthreads();

  const int total = sTotal;
  SelItem* levelBuf = total <= SEL_LEVEL_CAP ? S.levelBuf
                                             : reinterpret_cast<SelItem*>(fs.workLevel + img * fs.listCapTotal + L.listBase);
  const uint8_t* qual = fs.qual + img * fs.planeBytes + L.planeOff;

  for (int c = warp; c < nCells; c += SEL_WARPS) {
    const int n = S.nTotal[c];
    const int keep = min(max(S.nRetain[c], 0), n);
    if (keep == 0) continue;
    const int stored = S.nStored[c], thr = S.thr[c];
    const uint32_t* list = fs.cellList + img * fs.listCapTotal + cells[c].listOff;
    SelItem* const wbuf = S.cellBuf + (size_t)warp * SEL_CELL_CAP;
    SelItem* buf = n <= SEL_CELL_CAP ? wbuf
                                     : reinterpret_cast<SelItem*>(fs.workCell + img * fs.listCapTotal + cells[c].listOff);
    int run = 0;
    for (int base = 0; base < stored; base += 32) {
      const int i = base + lane;
      uint32_t e = 0;
      bool pass = false;
      if (i < stored) { e = list[i]; pass = unpack_s(e) >= thr; }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (pass) {
        float r = (float)unpack_s(e);
        if (fs.weighted) {
          const float cost = (float)__ldg(qual + (size_t)unpack_y(e) * L.pitch + unpack_x(e));
          r = __fmul_rn(r, introspection_weight(cost));
        }
        buf[run + __popc(m & ((1u << lane) - 1))] = SelItem{__float_as_uint(r), e};
      }
      run += __popc(m);
    }
    __syncwarp();
    if (buf != wbuf) __threadfence_block();
    if (lane == 0 && n > keep) sel_nth_element(buf, keep - 1, n);
    __syncwarp();
    if (buf != wbuf) __threadfence_block();
    const int dst = S.prefix[c];
    for (int i = lane; i < keep; i += 32) levelBuf[dst + i] = buf[i];
    __syncwarp();
  }
  __threadfence_block();
  __syncthreads();

  if (tid == 0) {
    int count = total;
    if (total > L.nDesired) {
      if (L.nDesired == 0) count = 0;
      else { sel_nth_element(levelBuf, L.nDesired - 1, total); count = L.nDesired; }
    }
    sCount = count;
    fs.levelCount[img * MAX_LEVELS + level] = count;
  }
  __threadfence_block();
  __syncthreads();
  const int count = sCount;
  uint2* out = fs.levelKp + img * fs.kpCap + L.kpOff;
  for (int i = tid; i < count; i += blockDim.x) out[i] = make_uint2(levelBuf[i].key, levelBuf[i].val);
}

}  // namespace ivg

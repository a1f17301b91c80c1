// k_frame.cuh — N1: the per-keypoint loops of the stereo Frame constructor that follow extraction, on the
// keypoints still resident on the device:
//   mvKeyQualScore        (introspective_ORB_SLAM/src/Frame.cc:128-143): cost-map sampled at the rounded level-0 position,
//                         qual = 2 * (1 / (1 + cost/256)) - 1  (note /256 here, /255 in the extractor — SURVEY Q6);
//   UndistortKeyPoints    (:696-726): identity for rectified stereo (k1 == 0, the only case the stereo configs use);
//   AssignFeaturesToGrid  (:415-430) + PosInGrid (:670-680): 64 x 48 grid over the image bounds, every cell lists its
//                         keypoint indices in ascending order (the reference push_backs in index order) — emitted as CSR.
// One CTA per frame; a counting sort over 3072 cells in shared memory, per-cell lists put back in index order.
#pragma once
#include "common.cuh"

namespace ivg {

constexpr int GRID_COLS = 64, GRID_ROWS = 48;     // FRAME_GRID_COLS / FRAME_GRID_ROWS, include/Frame.h:43-44

struct FramePostArgs {
  const uint8_t* kp; const int* n;     // [frames][cap] x 28 B, counts
  const uint8_t* cost;                 // level-0 cost-map planes [frames][planeBytes] or null
  size_t planeBytes; int costPitch;
  int cap;
  float minX, minY, invW, invH;        // mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv
  float* qual;                         // [frames][cap]
  int* gridStart;                      // [frames][GRID_COLS*GRID_ROWS + 1], cell = col*GRID_ROWS + row
  int* gridIdx;                        // [frames][cap]
};

__global__ void __launch_bounds__(256) k_frame_post(FramePostArgs A) {
  __shared__ int cnt[GRID_COLS * GRID_ROWS + 1];
  __shared__ int wsum[8];
  const size_t f = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = A.n[f];
  const int NC = GRID_COLS * GRID_ROWS;
  for (int i = tid; i <= NC; i += 256) cnt[i] = 0;
  __syncthreads();
  auto cell_of = [&](int i) {
    const float* k = reinterpret_cast<const float*>(A.kp + (f * A.cap + i) * 28);
    const int px = (int)roundf(__fmul_rn(__fsub_rn(k[0], A.minX), A.invW));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(k[1], A.minY), A.invH));
    return (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? -1 : px * GRID_ROWS + py;
  };
  for (int i = tid; i < N; i += 256) {
    const float* k = reinterpret_cast<const float*>(A.kp + (f * A.cap + i) * 28);
    float q = 1.0f;
    if (A.cost) {
      const int px = (int)roundf(k[0]), py = (int)roundf(k[1]);
      const float cost = (float)A.cost[f * A.planeBytes + (size_t)py * A.costPitch + px];
      const float qs = (float)(1.0 / (1.0 + (double)__fdiv_rn(cost, 256.0f)));
      q = __fsub_rn(__fmul_rn(2.0f, qs), 1.0f);
    }
    A.qual[f * A.cap + i] = q;
    const int c = cell_of(i);
    if (c >= 0) atomicAdd(&cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of the 3072 counts: 12 per thread
  const int per = NC / 256, b0 = tid * per;
  int local = 0;
  for (int i = 0; i < per; ++i) local += cnt[b0 + i];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int base = incl - local;
  for (int w = 0; w < warp; ++w) base += wsum[w];
  int* gs = A.gridStart + f * (NC + 1);
  for (int i = 0; i < per; ++i) { const int c = cnt[b0 + i]; cnt[b0 + i] = base; gs[b0 + i] = base; base += c; }
  if (tid == 255) gs[NC] = base;
  __syncthreads();
  int* gi = A.gridIdx + f * A.cap;
  for (int i = tid; i < N; i += 256) {
    const int c = cell_of(i);
    if (c >= 0) gi[atomicAdd(&cnt[c], 1)] = i;
  }
  __syncthreads();
  // cnt[c] is now the END of cell c; lists are short: put each back in ascending index order
  for (int c = tid; c < NC; c += 256) {
    const int s = gs[c], e = cnt[c];
    for (int a = s + 1; a < e; ++a) {
      const int v = gi[a];
      int b = a;
      while (b > s && gi[b - 1] > v) { gi[b] = gi[b - 1]; --b; }
      gi[b] = v;
    }
  }
}

}  // namespace ivg

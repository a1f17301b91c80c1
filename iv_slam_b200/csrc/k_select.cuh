// k_select.cuh — K3: keypoint selection of ComputeKeyPointsOld, the LIVE selection path of the reference
// (introspective_ORB_SLAM/src/ORBextractor.cc:880-1213; operator() calls it at :1248, the OctTree variant at :1247
// is commented out — SURVEY F1).
//
//   (k_fast_cells in k_fast.cuh produced, per cell: the row-major corner list, the corner counts at iniThFAST /
//    minThFAST for the "<= 3 keypoints => retry with minThFAST" rule (:1045-1052, post-NMS counts) and the cost-map
//    window sum for IV-SLAM's introspection weighting (:976-984).)
//   k_level_select one CTA per (level, frame): warp 0 replays the budget logic (cell weights :942-987, per-cell budgets
//                 :1028-1031/:1085-1096 lane-parallel, the one-shot redistribution loop :1101-1133 — a true recurrence,
//                 SURVEY Q4 — on one lane); each warp then trims cells with the replayed std::nth_element (retainBest +
//                 resize, :1146-1148), and warp 0 (or the whole CTA, one frame at a time) trims the level (:1162-1166).
//                 Output order = reference order.
// Deterministic: no atomics decide any order.
#pragma once
#include "common.cuh"
#include "introselect.h"

namespace ivg {

constexpr int SEL_MAX_CELLS = 1024;      // cells per level the select kernel supports (host check)
#ifndef IVG_SEL_WARPS
#define IVG_SEL_WARPS 4
#endif
constexpr int SEL_WARPS = IVG_SEL_WARPS;            // warps per CTA for large batches (throughput); small batches launch SEL_WARPS_LAT
constexpr int SEL_WARPS_LAT = 32;       // one-frame-at-a-time: only nlevels CTAs exist, so each gets every warp an SM-quarter can hold
// Shared memory is sized per handle (dynamic): FrameSet::selLevelCap level-list entries, selCellCap entries per warp for
// a cell list, selCells per-cell scalars.  Lists longer than the caps are processed in global memory (same code).

// Which of a[first+1], a[mid], a[last-1] libstdc++'s __move_median_to_first swaps to the front (every thread evaluates it
// from three broadcast loads instead of waiting for one thread's compare-and-swap chain).
__device__ __forceinline__ int sel_median3(const SelItem* a, int first, int last, uint32_t& pkey) {
  const int A = first + 1, B = first + (last - first) / 2, C = last - 1;
  const uint32_t ka = a[A].key, kb = a[B].key, kc = a[C].key;      // sel_before(x, y) = x.key > y.key
  int P;
  if (ka > kb) P = kb > kc ? B : (ka > kc ? C : A);
  else P = ka > kc ? A : (kb > kc ? C : B);
  pkey = P == A ? ka : (P == B ? kb : kc);
  return P;
}

// std::nth_element as libstdc++ runs it, executed by a WARP.  Same introselect skeleton as introselect.h (median-of-3 to
// the front, unguarded Hoare partition, narrow, 3-element insertion sort, heap-select when the depth limit trips); only
// the O(n) partition is parallel.  The sequential partition pairs the i-th element from the left that is not before
// the pivot ("L-stopper", key <= pivot) with the i-th element from the right that the pivot is not before ("R-stopper",
// key >= pivot) and swaps them while the left position is below the right one; scans between swaps only visit
// positions no swap has touched, so both stopper sequences can be read off the ORIGINAL array: ranks by ballot/popcount,
// positions scattered by rank into two scratch arrays, m = #{i : F[i] < R[i]} swaps done in parallel, and the cut is
// F[m] if it lies below R[m-1] (the sequential left scan finds it first) else R[m-1] (the scan stops on the element the
// last swap put there).  Produces the identical permutation (tests/test_gpu_parity.py::test_warp_nth_element).
// Partition rounds on a range of at most 32 elements, held one per lane in registers (lane l = position first + l): the
// same round as below, but stopper ranks and the cut come from ballots and popcounts instead of
// scratch arrays in shared memory (one 32-entry scatter remains: the lanes of a swap find each other through it) — a round is a
// chain of ~25 dependent register instructions and one shared-memory round trip instead of seven.  With r = rank of an L-stopper from the left and t = number of R-stoppers above a lane (= rank of an R-stopper
// from the right): F[i] < R[TR-1-i] for the L-stopper of rank r is "t > r" on its own lane, the i-th swap pairs the lanes with
// r == i and t == i, F[m] is the L-stopper with r == m and R[TR-m] the R-stopper with t == m-1.  Runs until the range is down
// to three elements or the depth limit trips, writes the elements back and returns the narrowed range.
__device__ __forceinline__ void warp_nth_rounds_reg(SelItem* a, int nth, int& first, int& last, int& depth, uint16_t* sF, uint16_t* sR, int lane) {
  const int base = first, cnt = last - first, nr = nth - base;
  SelItem it = SelItem{0u, 0u};
  if (lane < cnt) it = a[base + lane];
  uint32_t key = it.key, val = it.val;
  int f = 0, e = cnt;
  const unsigned lt = (1u << lane) - 1u, gt = ~lt & ~(1u << lane);
  while (e - f > 3 && depth > 0) {
    --depth;
    const int A = f + 1, B = f + (e - f) / 2, C = e - 1;
    const uint32_t ka = __shfl_sync(0xffffffffu, key, A), kb = __shfl_sync(0xffffffffu, key, B), kc = __shfl_sync(0xffffffffu, key, C);
    int P;
    if (ka > kb) P = kb > kc ? B : (ka > kc ? C : A);
    else P = ka > kc ? A : (kb > kc ? C : B);
    const uint32_t pkey = P == A ? ka : (P == B ? kb : kc);
    {
      // pivot swap a[first] <-> a[P]
      const uint32_t kf = __shfl_sync(0xffffffffu, key, f), vf = __shfl_sync(0xffffffffu, val, f), vp = __shfl_sync(0xffffffffu, val, P);
      if (lane == f) { key = pkey; val = vp; }
      else if (lane == P) { key = kf; val = vf; }
    }
    const bool in = lane > f && lane < e;
    const bool isL = in && !(key > pkey), isR = in && !(pkey > key);
    const unsigned mL = __ballot_sync(0xffffffffu, isL), mR = __ballot_sync(0xffffffffu, isR);
    const int r = __popc(mL & lt), t = __popc(mR & gt);
    const int m = __popc(__ballot_sync(0xffffffffu, isL && t > r));
    const bool swL = isL && r < m, swR = isR && t < m;            // never both: a position takes part in at most one swap
    // swap partners meet through two 32-entry scratch rows (MATCH.ANY costs ~11 cycles per distinct value: 350 cycles here)
    if (swL) sF[r] = (uint16_t)lane;
    if (swR) sR[t] = (uint16_t)lane;
    const unsigned bF = __ballot_sync(0xffffffffu, isL && r == m), bR = __ballot_sync(0xffffffffu, isR && t == m - 1);
    __syncwarp();
    const int partner = swL ? (int)sR[r] : (swR ? (int)sF[t] : lane);
    __syncwarp();
    key = __shfl_sync(0xffffffffu, key, partner);
    val = __shfl_sync(0xffffffffu, val, partner);
    const int rprev = m > 0 ? __ffs(bR) - 1 : e;
    const int fm = __ffs(bF) - 1;                                   // -1: no L-stopper of rank m (m == TL)
    const int cut = (bF && fm < rprev) ? fm : rprev;
    if (cut <= nr) f = cut; else e = cut;
  }
  if (e - f <= 3) {
    // the closing __insertion_sort of at most three elements, also in registers: an element's place is the number of
    // elements that go before it (larger key, or equal key and earlier position: the sort is stable)
    const uint32_t k0 = __shfl_sync(0xffffffffu, key, f), k1 = __shfl_sync(0xffffffffu, key, min(f + 1, 31)), k2 = __shfl_sync(0xffffffffu, key, min(f + 2, 31));
    const int c = e - f;
    const int r0 = (c > 1 && k1 > k0) + (c > 2 && k2 > k0);
    const int r1 = (k0 >= k1) + (c > 2 && k2 > k1);
    const int o = lane - f;                                    // this lane's place in the range
    const int srcLane = (o >= 0 && o < c) ? f + (r0 == o ? 0 : (c > 1 && r1 == o ? 1 : 2)) : lane;      // the third one takes the place that is left
    key = __shfl_sync(0xffffffffu, key, srcLane);
    val = __shfl_sync(0xffffffffu, val, srcLane);
    f = e;                                                     // nothing left for the caller to sort
  }
  if (lane < cnt) a[base + lane] = SelItem{key, val};
  __syncwarp();
  first = base + f; last = base + e;
}

// The partition rounds of the replay from a given state (first, last, depth) to the end, executed by one warp.
// The pivot swap a[first] <-> a[P] is done by lane 0 while the scan already runs: the scan reads position P as the key that
// is being moved there (the old a[first]) and no other position changes.
__device__ __forceinline__ void warp_nth_rounds(SelItem* a, int nth, int first, int last, int depth, uint16_t* sF, uint16_t* sR, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  while (last - first > 3) {
    if (depth == 0) {
      if (lane == 0) { sel_heap_select(a + first, nth + 1 - first, last - first); sel_swap(a, first, nth); }
      __syncwarp();
      return;
    }
    if (last - first <= 32) { warp_nth_rounds_reg(a, nth, first, last, depth, sF, sR, lane); continue; }
    --depth;
    uint32_t pkey;
    const int P = sel_median3(a, first, last, pkey);
    const SelItem fi = a[first];                               // the element the pivot swap moves to P
    const int lo = first + 1, hi = last;
    int TL = 0, TR = 0;
    for (int base = lo; base < hi; base += 32) {
      const int j = base + lane;
      const bool v = j < hi;
      const uint32_t k = v ? (j == P ? fi.key : a[j].key) : 0u;
      const bool isL = v && !(k > pkey), isR = v && !(pkey > k);
      const unsigned mL = __ballot_sync(0xffffffffu, isL), mR = __ballot_sync(0xffffffffu, isR);
      if (isL) sF[TL + __popc(mL & lt)] = (uint16_t)j;
      if (isR) sR[TR + __popc(mR & lt)] = (uint16_t)j;
      TL += __popc(mL); TR += __popc(mR);
    }
    __syncwarp();                                              // every lane has read what it needs of the old array
    if (lane == 0) { const SelItem pv = a[P]; a[P] = fi; a[first] = pv; }
    __syncwarp();
    int m = 0, x0 = 0, y0 = 0;
    const int lim = min(TL, TR);
    for (int base = 0; base < lim; base += 32) {
      const int i = base + lane;
      const int x = i < lim ? sF[i] : 0, y = i < lim ? sR[TR - 1 - i] : 0;
      if (base == 0) { x0 = x; y0 = y; }
      const unsigned b = __ballot_sync(0xffffffffu, i < lim && x < y);
      m += __popc(b);
      if (b != 0xffffffffu) break;
    }
    if (lane < m) { const SelItem t = a[x0]; a[x0] = a[y0]; a[y0] = t; }
    for (int i = lane + 32; i < m; i += 32) {
      const int x = sF[i], y = sR[TR - 1 - i];
      const SelItem t = a[x]; a[x] = a[y]; a[y] = t;
    }
    const int rprev = m > 0 ? (int)sR[TR - m] : hi;
    const int cut = (m < TL && (int)sF[m] < rprev) ? (int)sF[m] : rprev;
    __syncwarp();
    if (cut <= nth) first = cut; else last = cut;
  }
  if (lane == 0) {
    for (int i = first + 1; i < last; ++i) {
      const SelItem v = a[i];
      if (sel_before(v, a[first])) {
        for (int j = i; j > first; --j) a[j] = a[j - 1];
        a[first] = v;
      } else {
        int j = i;
        while (sel_before(v, a[j - 1])) { a[j] = a[j - 1]; --j; }
        a[j] = v;
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_nth_element(SelItem* a, int nth, int n, uint16_t* sF, uint16_t* sR, int lane) {
  if (n == 0 || nth == n) return;
  warp_nth_rounds(a, nth, 0, n, 2 * (31 - __clz(n)), sF, sR, lane);
}

// The same replay executed by BN_THREADS threads of a CTA (the one-frame-at-a-time configuration: one CTA per level): a
// partition round handles BN_THREADS elements per trip — stopper ranks from warp ballots plus a prefix over the warps'
// counts, m from a counting barrier (the predicate F[i] < R[TR-1-i] is a prefix of trues), swaps in parallel.  Once the
// range is down to BN_MIN elements, warp 0 finishes alone (rounds without barriers).  Identical permutation; exactly the
// first BN_THREADS threads of the CTA call it (n <= 65535, lists in shared memory); they meet on named barrier 1, so a
// barrier costs what 8 warps cost, not what the whole CTA costs.
constexpr int BN_THREADS = 256, BN_MIN = 96;
__device__ __forceinline__ void bn_sync() { asm volatile("bar.sync 1, %0;" ::"n"(BN_THREADS) : "memory"); }
__device__ __forceinline__ int bn_sync_count(bool p) {
  int r;
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %1, 0;\nbar.red.popc.u32 %0, 1, %2, q;\n}" : "=r"(r) : "r"((int)p), "n"(BN_THREADS) : "memory");
  return r;
}

__device__ __forceinline__ void block_nth_element(SelItem* a, int nth, int n, uint16_t* sF, uint16_t* sR, int* wcnt /*[2][2*32]*/) {
  if (n == 0 || nth == n) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = BN_THREADS, nw = nthr >> 5;
  int first = 0, last = n;
  int depth = 2 * (31 - __clz(n));
  const unsigned lt = (1u << lane) - 1u;
  int buf = 0;                                                // wcnt is double buffered: one barrier per trip
  while (last - first > BN_MIN && depth > 0) {
    --depth;
    uint32_t pkey;
    const int P = sel_median3(a, first, last, pkey);
    const uint32_t fkey = a[first].key;
    const int lo = first + 1, hi = last;
    int TL = 0, TR = 0;
    for (int base = lo; base < hi; base += nthr) {
      const int j = base + tid;
      const bool v = j < hi;
      const uint32_t k = v ? (j == P ? fkey : a[j].key) : 0u;
      const bool isL = v && !(k > pkey), isR = v && !(pkey > k);
      const unsigned mL = __ballot_sync(0xffffffffu, isL), mR = __ballot_sync(0xffffffffu, isR);
      int* wc = wcnt + 64 * buf;
      buf ^= 1;
      if (lane == 0) { wc[warp] = __popc(mL); wc[32 + warp] = __popc(mR); }
      bn_sync();
      if (base == lo && tid == 0) sel_swap(a, first, P);      // every thread has read the three candidates, a[first] and a[P]
      // exclusive prefix of this warp's counts over the warps before it, and the totals: one lane per warp + a shuffle scan
      int sl = lane < nw ? wc[lane] : 0, sr = lane < nw ? wc[32 + lane] : 0;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int vl = __shfl_up_sync(0xffffffffu, sl, o), vr = __shfl_up_sync(0xffffffffu, sr, o);
        if (lane >= o) { sl += vl; sr += vr; }
      }
      const int tL = __shfl_sync(0xffffffffu, sl, nw - 1), tR = __shfl_sync(0xffffffffu, sr, nw - 1);
      const int pL = warp ? __shfl_sync(0xffffffffu, sl, warp - 1) : 0, pR = warp ? __shfl_sync(0xffffffffu, sr, warp - 1) : 0;
      if (isL) sF[TL + pL + __popc(mL & lt)] = (uint16_t)j;
      if (isR) sR[TR + pR + __popc(mR & lt)] = (uint16_t)j;
      TL += tL; TR += tR;
    }
    bn_sync();
    int m = 0;
    const int lim = min(TL, TR);
    for (int base = 0; base < lim; base += nthr) {
      const int i = base + tid;
      const bool ok = i < lim && sF[i] < sR[TR - 1 - i];
      const int c = bn_sync_count(ok);
      m += c;
      if (c != nthr) break;                                 // CTA-uniform
    }
    for (int i = tid; i < m; i += nthr) {
      const int x = sF[i], y = sR[TR - 1 - i];
      const SelItem t = a[x]; a[x] = a[y]; a[y] = t;
    }
    const int rprev = m > 0 ? (int)sR[TR - m] : hi;
    const int cut = (m < TL && (int)sF[m] < rprev) ? (int)sF[m] : rprev;
    bn_sync();
    if (cut <= nth) first = cut; else last = cut;
  }
  if (warp == 0) warp_nth_rounds(a, nth, first, last, depth, sF, sR, lane);      // also the depth-limit fallback
  bn_sync();
}

// test hook: one warp runs warp_nth_element on n (key, index) items staged in shared memory
__global__ void k_debug_nth_element(const uint32_t* keys, int n, int nth, uint32_t* order) {
  extern __shared__ __align__(16) unsigned char dsm[];
  SelItem* a = reinterpret_cast<SelItem*>(dsm);
  uint16_t* sF = reinterpret_cast<uint16_t*>(a + n);
  uint16_t* sR = sF + n;
  const int lane = threadIdx.x;
  if (blockDim.x > 32) {          // the CTA-wide replay of the one-frame-at-a-time configuration
    __shared__ int wc[128];
    for (int i = lane; i < n; i += blockDim.x) a[i] = SelItem{keys[i], (uint32_t)i};
    __syncthreads();
    if (lane < BN_THREADS) block_nth_element(a, nth, n, sF, sR, wc);
    __syncthreads();
    for (int i = lane; i < n; i += blockDim.x) order[i] = a[i].val;
    return;
  }
  for (int i = lane; i < n; i += 32) a[i] = SelItem{keys[i], (uint32_t)i};
  __syncwarp();
#ifdef IVG_SEL_CLOCK                      // developer build: cycles of one replay
  const long long t0 = clock64();
#endif
  warp_nth_element(a, nth, n, sF, sR, lane);
#ifdef IVG_SEL_CLOCK
  const long long t1 = clock64();
  if (lane == 0) printf("warp nth_element n %d nth %d: %lld cycles\n", n, nth, t1 - t0);
#endif
  for (int i = lane; i < n; i += 32) order[i] = a[i].val;
}

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
  uint16_t* order;           // cells in the order the warps take them: those whose list has to be trimmed first
  uint16_t* cellScratch;     // [SEL_WARPS][2 * selCellCap]
  uint16_t* levelScratch;    // [2 * selLevelCap]
};

__host__ __device__ inline size_t sel_smem_bytes(int levelCap, int cellCap, int cells, int warps = SEL_WARPS) {
  return (size_t)8 * levelCap + (size_t)8 * warps * cellCap + (size_t)cells * (5 * 4 + 2 + 2) + 16 +
         (size_t)4 * warps * cellCap + (size_t)4 * levelCap + 16;
}

template <int MAX_THREADS>   // two instantiations: 8 warps (register budget of the throughput configuration) and 32 warps (latency)
__global__ void __launch_bounds__(MAX_THREADS) k_level_select(FrameSet fs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sTotal, sCount, sNext;
  const int SEL_LEVEL_CAP = fs.selLevelCap, SEL_CELL_CAP = fs.selCellCap;
  const int nWarps = blockDim.x >> 5;
  SelShared S;
  {
    // offsets, not pointer arithmetic through uintptr_t: the compiler keeps seeing shared memory (LDS/STS, not generic accesses)
    size_t o = 0;
    S.levelBuf = reinterpret_cast<SelItem*>(smem_raw + o); o += (size_t)8 * SEL_LEVEL_CAP;
    S.cellBuf = reinterpret_cast<SelItem*>(smem_raw + o); o += (size_t)8 * nWarps * SEL_CELL_CAP;
    const int nc = fs.selCells;
    S.nTotal = reinterpret_cast<int*>(smem_raw + o); o += 4 * (size_t)nc;
    S.nStored = reinterpret_cast<int*>(smem_raw + o); o += 4 * (size_t)nc;
    S.nRetain = reinterpret_cast<int*>(smem_raw + o); o += 4 * (size_t)nc;
    S.prefix = reinterpret_cast<int*>(smem_raw + o); o += 4 * (size_t)nc;
    S.nfc = reinterpret_cast<float*>(smem_raw + o); o += 4 * (size_t)nc;
    S.thr = smem_raw + o; o += nc;
    S.noMore = smem_raw + o; o += nc;
    S.order = reinterpret_cast<uint16_t*>(smem_raw + o); o += 2 * (size_t)nc;
    o = (o + 15) & ~(size_t)15;                                // smem_raw is 16-byte aligned
    S.cellScratch = reinterpret_cast<uint16_t*>(smem_raw + o); o += (size_t)4 * nWarps * SEL_CELL_CAP;
    S.levelScratch = reinterpret_cast<uint16_t*>(smem_raw + o);
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

  if (fs.weighted) {
    // cell weights (:942-987): the per-cell divisions run in parallel, only the float sum keeps the reference's
    // row-major accumulation order (thread 0), then the budgets are again per cell
    __shared__ float sWsum;
    const uint32_t* cost = fs.cellCost + img * fs.cellCostStride + L.cellBase;
    for (int c = tid; c < nCells; c += blockDim.x) {
      const float area = __fmul_rn((float)cells[c].ww, (float)cells[c].wh);
      const float mean = __fdiv_rn((float)cost[c], area);
      const float qs = (float)(1.0 / (1.0 + (double)__fdiv_rn(mean, 255.0f)));
      S.nfc[c] = __fsub_rn(__fmul_rn(2.0f, qs), 1.0f);      // parked; converted to a budget below
    }
    __syncthreads();
    if (tid == 0) {
      float wsum = 0.0f;
      for (int c = 0; c < nCells; ++c) wsum = __fadd_rn(wsum, S.nfc[c]);
      sWsum = wsum;
    }
    __syncthreads();
    const float wsum = sWsum;
    for (int c = tid; c < nCells; c += blockDim.x) {
      const float v = ceilf(__fdiv_rn(__fmul_rn((float)L.nDesired, S.nfc[c]), wsum));
      S.nfc[c] = (1.0f < v) ? v : 1.0f;           // std::max(1.0f, v); NaN -> 1
    }
    __syncthreads();
  }
  if (warp == 0) {
    // :1082-1097.  Per cell the test and nToRetain are independent; nToDistribute is the reference's running
    // `int += float` (int = (int)((float)int + float)).  When every term is a finite integer-valued float and the sum stays
    // far below 2^24 all those additions are exact, so the running value equals the integer sum: lanes add in parallel.
    // Anything else (NaN / inf budgets from degenerate cost-maps) replays the recurrence on lane 0.
    int nNoMore = 0, isum = 0;
    bool exact = true;
    for (int base = 0; base < nCells; base += 32) {
      const int c = base + lane;
      bool nm = false; int d = 0;
      if (c < nCells) {
        const int nKeys = S.nTotal[c];
        const float f = S.nfc[c];
        if ((float)nKeys > f) { S.nRetain[c] = (int)f; S.noMore[c] = 0; }
        else {
          S.nRetain[c] = nKeys; S.noMore[c] = 1; nm = true;
          const float df = __fsub_rn(f, (float)nKeys);
          if (!(fabsf(df) < 8192.0f) || df != truncf(df)) exact = false; else d = (int)df;
        }
      }
      nNoMore += __popc(__ballot_sync(0xffffffffu, nm));
      isum += d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
    exact = __all_sync(0xffffffffu, exact) && nCells <= SEL_MAX_CELLS;      // every partial sum stays below 1024 * 8192 = 2^23
    int nToDistribute = isum;
    __syncwarp();
    if (!exact) {          // uniform
      nToDistribute = 0;
      if (lane == 0)
        for (int c = 0; c < nCells; ++c)
          if (S.noMore[c]) nToDistribute = (int)__fadd_rn((float)nToDistribute, __fsub_rn(S.nfc[c], (float)S.nTotal[c]));
      nToDistribute = __shfl_sync(0xffffffffu, nToDistribute, 0);
    }
    if (nToDistribute > 0 && nNoMore < nCells) {  // the while loop runs exactly once (:1103-1133)
      // A true recurrence over the cells in order: a cell that cannot absorb its share changes the share of the cells after
      // it.  The running values are warp-uniform; one cell per lane.
      float share = ceilf(__fdiv_rn((float)nToDistribute, (float)(nCells - nNoMore)));
      for (int base = 0; base < nCells; base += 32) {
        const int c = base + lane;
        const bool in = c < nCells;
        bool nm = in ? S.noMore[c] != 0 : true;
        const float f = in ? S.nfc[c] : 0.f;
        const int nt = in ? S.nTotal[c] : 0;
        int nr = in ? S.nRetain[c] : 0;
        // the cells of this chunk that still take part, in order; between two cells that saturate every cell sees the same
        // share, so all of them are evaluated at once and committed up to the first one that saturates
        unsigned pending = __ballot_sync(0xffffffffu, !nm);
        while (pending) {
          const int nNew = (int)__fadd_rn(f, share);
          const unsigned bs = __ballot_sync(0xffffffffu, !(nt > nNew)) & pending;
          const int k = bs ? __ffs(bs) - 1 : 32;
          if (((pending >> lane) & 1u) && lane < k) nr = nNew;
          if (k == 32) break;
          const int dk = __shfl_sync(0xffffffffu, nNew - nt, k);
          if (lane == k) { nr = nt; nm = true; }
          nToDistribute += dk; nNoMore++;
          share = ceilf(__fdiv_rn((float)nToDistribute, (float)(nCells - nNoMore)));
          pending &= ~((2u << k) - 1u);
        }
        if (in) { S.nRetain[c] = nr; S.noMore[c] = nm ? 1 : 0; }
      }
    }
    __syncwarp();
    int run = 0;      // prefix[c] = retained keypoints of the cells before c (warp scan)
    int nHeavy = 0, nLight = 0;
    for (int base = 0; base < nCells; base += 32) {
      const int c = base + lane;
      const int nt = c < nCells ? S.nTotal[c] : 0;
      const int k = c < nCells ? min(max(S.nRetain[c], 0), nt) : 0;
      int incl = k;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (c < nCells) S.prefix[c] = run + incl - k;
      run += __shfl_sync(0xffffffffu, incl, 31);
      // work order of the cells phase: cells that need a trim (the expensive ones) from the front, the others from the back
      const bool heavy = c < nCells && nt > k && k > 0, light = c < nCells && !heavy;
      const unsigned mh = __ballot_sync(0xffffffffu, heavy), ml = __ballot_sync(0xffffffffu, light);
      const unsigned below = (1u << lane) - 1u;
      if (heavy) S.order[nHeavy + __popc(mh & below)] = (uint16_t)c;
      if (light) S.order[nCells - 1 - nLight - __popc(ml & below)] = (uint16_t)c;
      nHeavy += __popc(mh); nLight += __popc(ml);
    }
    if (lane == 0) { sTotal = run; sNext = 0; }
  }
  __syncthreads();

  const int total = sTotal;
  SelItem* levelBuf = total <= SEL_LEVEL_CAP ? S.levelBuf
                                             : reinterpret_cast<SelItem*>(fs.workLevel + img * fs.listCapTotal + L.listBase);
  const uint8_t* qual = fs.qual + img * fs.planeBytes + L.planeOff;

  // the warps take cells as they get free (per-cell cost varies by an order of magnitude); a cell's place in the level list
  // is fixed by prefix[], so the order of processing does not show in the output
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(&sNext, 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= nCells) break;
    const int c = S.order[q];
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
    if (n > keep) {
      // (wbuf, not buf: a pointer the compiler can see is shared memory -> LDS/STS instead of generic loads and stores)
      if (buf == wbuf) warp_nth_element(wbuf, keep - 1, n, S.cellScratch + (size_t)warp * 2 * SEL_CELL_CAP,
                                        S.cellScratch + (size_t)warp * 2 * SEL_CELL_CAP + SEL_CELL_CAP, lane);
      else if (lane == 0) sel_nth_element(buf, keep - 1, n);      // oversized list in global memory: sequential replay
    }
    __syncwarp();
    const int dst = S.prefix[c];
    for (int i = lane; i < keep; i += 32) levelBuf[dst + i] = buf[i];
    __syncwarp();
  }
  __syncthreads();

  if (MAX_THREADS >= 512 && levelBuf == S.levelBuf && total > L.nDesired && L.nDesired > 0) {
    // latency configuration: the level's trim is the longest serial piece of the frame, all 32 warps take part
    __shared__ int sWcnt[128];
    if (tid < BN_THREADS) block_nth_element(S.levelBuf, L.nDesired - 1, total, S.levelScratch, S.levelScratch + SEL_LEVEL_CAP, sWcnt);
    if (tid == 0) { sCount = L.nDesired; fs.levelCount[img * MAX_LEVELS + level] = L.nDesired; }
  } else if (warp == 0) {
    int count = total;
    if (total > L.nDesired) {
      if (L.nDesired == 0) count = 0;
      else {
        if (levelBuf == S.levelBuf) warp_nth_element(S.levelBuf, L.nDesired - 1, total, S.levelScratch, S.levelScratch + SEL_LEVEL_CAP, lane);
        else if (lane == 0) sel_nth_element(levelBuf, L.nDesired - 1, total);
        count = L.nDesired;
      }
    }
    if (lane == 0) { sCount = count; fs.levelCount[img * MAX_LEVELS + level] = count; }
  }
  __syncthreads();
  const int count = sCount;
  uint2* out = fs.levelKp + img * fs.kpCap + L.kpOff;
  for (int i = tid; i < count; i += blockDim.x) out[i] = make_uint2(levelBuf[i].key, levelBuf[i].val);
}

}  // namespace ivg

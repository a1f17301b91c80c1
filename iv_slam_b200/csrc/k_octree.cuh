// k_octree.cuh — optional keypoint-selection mode: ComputeKeyPointsOctTree + DistributeOctTree + ExtractorNode::DivideNode
// (introspective_ORB_SLAM/src/ORBextractor.cc:771-878, :545-769, :487-543).
//
// In the reference this path is compiled but DEAD (operator() calls ComputeKeyPointsOld, :1247-1248); it is offered as
// mode 1 of the handle because the north star names it.  The FAST side is the same kernel as the live path
// (k_fast_cells) run on the OctTree's cell table (30-px cells over the area inset by 16, windows of cell+6, retry with
// minThFAST only when a cell is EMPTY, :813-819).  This kernel replays the quadtree:
//   gather   the corners of all cells, cell row-major and FAST order inside a cell (= vToDistributeKeys, :793-851);
//   tree     one warp walks the reference's node list (std::list semantics: children are pushed to the FRONT, the
//            divided parent is erased) — full-expansion rounds, then the one-by-one phase that splits the largest
//            nodes first — with the 4-way split of a node's keys done cooperatively (ballot ranks, stable);
//   emit     one keypoint per node, the first of maximal response, in list order (:747-766).
// One deliberate difference: the reference sorts (size, ExtractorNode*) pairs (:690), so equal-size nodes are ordered
// by heap address — nondeterministic across runs.  Here the tie is broken by creation order (later
// created = larger), so the result is a pure function of the image (the tests hold the CPU restatement to the same rule).
#pragma once
#include "common.cuh"

namespace ivg {

struct OctNodeDev {
  int16_t ulx, uly, brx, bry;
  uint32_t kb;          // first key slot (in buffer `buf`)
  uint32_t kc;          // number of keys
  int32_t prev, next;   // list links (-1 = none)
  uint32_t seq;         // creation order
  uint8_t buf, noMore, pad0, pad1;
};

constexpr int OCT_SMEM_KEYS = 4096;      // keys kept in shared memory (else the global work area is used)
constexpr int OCT_SMEM_NODES = 1024;     // nodes kept in shared memory

struct OctShared {
  uint32_t keys[OCT_SMEM_KEYS];          // packed (y<<20 | x<<8 | score), level coordinates
  uint32_t idxA[OCT_SMEM_KEYS], idxB[OCT_SMEM_KEYS];
  OctNodeDev nodes[OCT_SMEM_NODES];
  int32_t expand[OCT_SMEM_NODES], prevList[OCT_SMEM_NODES];
};

__global__ void __launch_bounds__(256) k_octree_select(FrameSet fs) {
  extern __shared__ __align__(16) unsigned char osm_raw[];
  OctShared& S = *reinterpret_cast<OctShared*>(osm_raw);
  __shared__ int sN;
  const int level = blockIdx.x;
  const size_t img = blockIdx.y;
  const LevelDev& L = fs.lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nCells = L.nCells;
  const int2* cc = fs.cellCount + img * fs.nCellsTotal + L.cellBase;
  const CellDev* cells = fs.cells + L.cellBase;

  // ---- gather: per-cell counts -> offsets (block scan, offsets parked in the per-cell scratch array) -> ordered copy
  __shared__ int wsum[8];
  uint32_t* cellOff = fs.cellCost + img * fs.cellCostStride + L.cellBase + level;    // nCells + 1 entries
  {
    const int per = (nCells + 255) / 256, b0 = tid * per, b1 = min(b0 + per, nCells);
    int local = 0;
    for (int c = b0; c < b1; ++c) { const int2 v = cc[c]; local += v.y > 0 ? v.y : v.x; }   // FAST(iniTh); if empty FAST(minTh) (:813-819)
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = incl - local;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    for (int c = b0; c < b1; ++c) { const int2 v = cc[c]; cellOff[c] = (uint32_t)base; base += v.y > 0 ? v.y : v.x; }
    if (tid == 255) { cellOff[nCells] = (uint32_t)base; sN = base; }
  }
  __threadfence_block();
  __syncthreads();
  const int nKeys = sN;
  uint32_t* work = reinterpret_cast<uint32_t*>(fs.workLevel + img * fs.listCapTotal + L.listBase);   // 2 words per list slot
  uint32_t* keys = nKeys <= OCT_SMEM_KEYS ? S.keys : work;
  uint32_t* idxA = nKeys <= OCT_SMEM_KEYS ? S.idxA : reinterpret_cast<uint32_t*>(fs.workCell + img * fs.listCapTotal + L.listBase);
  uint32_t* idxB = nKeys <= OCT_SMEM_KEYS ? S.idxB : idxA + L.listCap;
  for (int c = warp; c < nCells; c += 8) {
    const int2 v = cc[c];
    const int thr = v.y > 0 ? fs.iniTh : fs.minTh;
    const int stored = fs.iniTh < fs.minTh ? v.y : v.x;
    const uint32_t* list = fs.cellList + img * fs.listCapTotal + cells[c].listOff;
    int run = (int)cellOff[c];
    for (int base = 0; base < stored; base += 32) {
      const int i = base + lane;
      uint32_t e = 0;
      bool pass = false;
      if (i < stored) { e = list[i]; pass = unpack_s(e) >= thr; }
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (pass) keys[run + __popc(m & ((1u << lane) - 1))] = e;
      run += __popc(m);
    }
  }
  __threadfence_block();
  __syncthreads();
  if (warp != 0) return;

  // ---- tree (warp 0; every lane follows the same control flow, lanes cooperate inside split())
  const int N = L.nDesired;
  const int minB = EDGE - 3;
  const int maxX = L.w - EDGE + 3, maxY = L.h - EDGE + 3;
  const int nodeCap = N + 16 <= OCT_SMEM_NODES ? OCT_SMEM_NODES : 0;
  OctNodeDev* nodes = S.nodes;
  int32_t* expand = S.expand;
  int32_t* prevList = S.prevList;
  if (!nodeCap) {   // oversized budgets: node tables live after the key indices in the global work area
    OctNodeDev* g = reinterpret_cast<OctNodeDev*>(work + L.listCap);     // second half of this level's work area (host checks it fits)
    nodes = g;
    expand = reinterpret_cast<int32_t*>(g + (N + 16));
    prevList = expand + (N + 16);
  }
  const int slots = nodeCap ? OCT_SMEM_NODES : N + 16;
  int count = 0;
  uint2* out = fs.levelKp + img * fs.kpCap + L.kpOff;

  int head = -1, tail = -1, nNodes = 0, freeHead = -1, nextFresh = 0;
  uint32_t seq = 0;
  auto alloc_node = [&]() {
    int id;
    if (freeHead >= 0) {                        // lane 0 reads the link and broadcasts it: it is also the lane that rewrites the node
      id = freeHead;
      int nx = 0;
      if (lane == 0) nx = nodes[id].next;
      freeHead = __shfl_sync(0xffffffffu, nx, 0);
    }
    else id = nextFresh < slots ? nextFresh++ : -1;
    return id;
  };
  auto free_node = [&](int id) { if (lane == 0) nodes[id].next = freeHead; __syncwarp(); freeHead = id; };
  auto unlink = [&](int id) {
    const int p = nodes[id].prev, n = nodes[id].next;
    __syncwarp();
    if (lane == 0) {
      if (p >= 0) nodes[p].next = n;
      if (n >= 0) nodes[n].prev = p;
    }
    if (p < 0) head = n;
    if (n < 0) tail = p;
    --nNodes;
    __syncwarp();
  };
  auto push_front = [&](int id) {
    if (lane == 0) { nodes[id].prev = -1; nodes[id].next = head; if (head >= 0) nodes[head].prev = id; }
    if (head < 0) tail = id;
    head = id;
    ++nNodes;
    __syncwarp();
  };
  auto push_back = [&](int id) {
    if (lane == 0) { nodes[id].next = -1; nodes[id].prev = tail; if (tail >= 0) nodes[tail].next = id; }
    if (tail < 0) head = id;
    tail = id;
    ++nNodes;
    __syncwarp();
  };

  const int spanX = maxX - minB, spanY = maxY - minB;
  const int nIni = (int)roundf(__fdiv_rn((float)spanX, (float)spanY));
  bool ok = nIni >= 1 && nIni + 8 <= slots;
  if (ok) {
    const float hX = __fdiv_rn((float)spanX, (float)nIni);
    // initial nodes: vertical strips; keys binned by (int)(x / hX), order preserved (stable counting sort)
    for (int i = 0; i < nIni; ++i) {
      const int id = alloc_node();
      if (lane == 0) {
        OctNodeDev n{};
        n.ulx = (int16_t)(int)__fmul_rn(hX, (float)i); n.uly = 0;
        n.brx = (int16_t)(int)__fmul_rn(hX, (float)(i + 1)); n.bry = (int16_t)spanY;
        n.kb = 0; n.kc = 0; n.buf = 0; n.noMore = 0; n.seq = seq;
        nodes[id] = n;
      }
      ++seq;
      __syncwarp();
      push_back(id);                                   // ids 0..nIni-1 in order
    }
    // counts per strip
    for (int i = 0; i < nIni; ++i) {
      int cnt = 0;
      for (int base = 0; base < nKeys; base += 32) {
        const int k = base + lane;
        const bool in = k < nKeys && (int)__fdiv_rn((float)(unpack_x(keys[k]) - minB), hX) == i;
        cnt += __popc(__ballot_sync(0xffffffffu, in));
      }
      if (lane == 0) nodes[i].kc = (uint32_t)cnt;
    }
    __syncwarp();
    {
      uint32_t run = 0;
      for (int i = 0; i < nIni; ++i) { const uint32_t c = nodes[i].kc; __syncwarp(); if (lane == 0) nodes[i].kb = run; run += c; }
      __syncwarp();
      for (int i = 0; i < nIni; ++i) {
        uint32_t pos = nodes[i].kb;
        for (int base = 0; base < nKeys; base += 32) {
          const int k = base + lane;
          const bool in = k < nKeys && (int)__fdiv_rn((float)(unpack_x(keys[k]) - minB), hX) == i;
          const unsigned m = __ballot_sync(0xffffffffu, in);
          if (in) idxA[pos + __popc(m & ((1u << lane) - 1))] = (uint32_t)k;
          pos += __popc(m);
        }
      }
      __syncwarp();
    }
    // drop empty strips, mark single-key strips (:581-593)
    for (int id = head; id >= 0;) {
      const int nx = nodes[id].next;
      const uint32_t kc = nodes[id].kc;
      __syncwarp();
      if (kc == 1) { if (lane == 0) nodes[id].noMore = 1; }
      else if (kc == 0) { unlink(id); free_node(id); }
      id = nx;
      __syncwarp();
    }

    // split node `id` into up to four children pushed to the front in the order n1..n4; returns via expand list
    int nExpand = 0;
    auto split = [&](int id, int* nToExpand) {
      const OctNodeDev P = nodes[id];
      __syncwarp();
      const int mx = P.ulx + ((P.brx - P.ulx + 1) >> 1), my = P.uly + ((P.bry - P.uly + 1) >> 1);   // ceil(d/2)
      const uint32_t* src = P.buf ? idxB : idxA;
      uint32_t* dst = P.buf ? idxA : idxB;
      int cnt[4] = {0, 0, 0, 0};
      for (uint32_t base = 0; base < P.kc; base += 32) {
        const uint32_t i = base + lane;
        int c = -1;
        if (i < P.kc) {
          const uint32_t e = keys[src[P.kb + i]];
          const int x = unpack_x(e) - minB, y = unpack_y(e) - minB;
          c = (x < mx ? 0 : 1) + (y < my ? 0 : 2);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) cnt[q] += __popc(__ballot_sync(0xffffffffu, c == q));
      }
      uint32_t off[4];
      off[0] = P.kb; off[1] = off[0] + cnt[0]; off[2] = off[1] + cnt[1]; off[3] = off[2] + cnt[2];
      uint32_t pos[4] = {off[0], off[1], off[2], off[3]};
      for (uint32_t base = 0; base < P.kc; base += 32) {
        const uint32_t i = base + lane;
        int c = -1;
        uint32_t kidx = 0;
        if (i < P.kc) {
          kidx = src[P.kb + i];
          const uint32_t e = keys[kidx];
          const int x = unpack_x(e) - minB, y = unpack_y(e) - minB;
          c = (x < mx ? 0 : 1) + (y < my ? 0 : 2);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned m = __ballot_sync(0xffffffffu, c == q);
          if (c == q) dst[pos[q] + __popc(m & ((1u << lane) - 1))] = kidx;
          pos[q] += __popc(m);
        }
      }
      __syncwarp();
      bool fine = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (cnt[q] == 0) continue;
        const int cid = alloc_node();
        if (cid < 0) { fine = false; break; }
        if (lane == 0) {
          OctNodeDev n{};
          n.ulx = (int16_t)((q & 1) ? mx : P.ulx); n.brx = (int16_t)((q & 1) ? P.brx : mx);
          n.uly = (int16_t)((q & 2) ? my : P.uly); n.bry = (int16_t)((q & 2) ? P.bry : my);
          n.kb = off[q]; n.kc = (uint32_t)cnt[q]; n.buf = P.buf ^ 1; n.noMore = cnt[q] == 1; n.seq = seq;
          nodes[cid] = n;
        }
        ++seq;
        __syncwarp();
        push_front(cid);
        if (cnt[q] > 1) {
          if (nToExpand) ++*nToExpand;
          if (lane == 0) expand[nExpand] = cid;
          ++nExpand;
          __syncwarp();
        }
      }
      return fine;
    };

    bool finish = false;
    while (!finish && ok) {
      const int prevSize = nNodes;
      int nToExpand = 0;
      nExpand = 0;
      for (int id = head; id >= 0 && ok;) {                 // children go to the front: never revisited in this round
        const int nx = nodes[id].next;
        const bool nm = nodes[id].noMore != 0;
        __syncwarp();
        if (!nm) {
          ok = split(id, &nToExpand);
          unlink(id);
          free_node(id);
        }
        id = nx;
      }
      if (!ok) break;
      if (nNodes >= N || nNodes == prevSize) finish = true;
      else if (nNodes + nToExpand * 3 > N) {
        while (!finish && ok) {
          const int prev2 = nNodes;
          const int nPrev = nExpand;
          // sort ascending by (size, creation order): rank sort, keys are unique
          for (int i = lane; i < nPrev; i += 32) {
            const int a = expand[i];
            const uint32_t ka = nodes[a].kc, sa = nodes[a].seq;
            int rank = 0;
            for (int j = 0; j < nPrev; ++j) {
              const int b = expand[j];
              const uint32_t kb2 = nodes[b].kc, sb = nodes[b].seq;
              rank += (kb2 < ka || (kb2 == ka && sb < sa)) ? 1 : 0;
            }
            prevList[rank] = a;
          }
          __syncwarp();
          nExpand = 0;
          for (int j = nPrev - 1; j >= 0 && ok; --j) {
            const int id = prevList[j];
            __syncwarp();
            ok = split(id, nullptr);
            unlink(id);
            free_node(id);
            if (nNodes >= N) break;
          }
          if (nNodes >= N || nNodes == prev2) finish = true;
        }
      }
    }
  }

  // ---- emit: list order, first key of maximal response per node
  if (ok) {
    // positions of the nodes in list order (sequential walk), then lanes pick the best key of their node
    int order = 0;
    for (int id = head; id >= 0; id = nodes[id].next) { if (lane == 0) prevList[order] = id; ++order; }
    __syncwarp();
    count = min(order, L.kpCapLevel);
    for (int i = lane; i < count; i += 32) {
      const OctNodeDev n = nodes[prevList[i]];
      const uint32_t* src = n.buf ? idxB : idxA;
      uint32_t best = keys[src[n.kb]];
      for (uint32_t k = 1; k < n.kc; ++k) {
        const uint32_t e = keys[src[n.kb + k]];
        if (unpack_s(e) > unpack_s(best)) best = e;
      }
      out[i] = make_uint2(__float_as_uint((float)unpack_s(best)), best);
    }
  }
  if (lane == 0) fs.levelCount[img * MAX_LEVELS + level] = ok ? count : 0;
}

}  // namespace ivg

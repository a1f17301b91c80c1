// k_fast.cuh — K2+K3a fused: per-cell FAST-9/16 score, cell-local 3x3 non-max suppression and ordered (row-major)
// compaction of the surviving corners into the cell's list.
//
// Replaces the per-cell cv::FAST(cellImage, kps, th, true) calls of ComputeKeyPointsOld
// (introspective_ORB_SLAM/src/ORBextractor.cc:1045 and :1051) including their emission order, and cv::sum over the
// cost-map window (:976-978).  SURVEY Appendix A.3:
//   * the corner score S (largest threshold at which the pixel is still a 9-arc corner) does not depend on the
//     threshold and "corner at th" <=> S >= th, so ONE score pass serves both iniThFAST and minThFAST; the
//     "<= 3 keypoints => retry with minThFAST" rule (:1047-1052) only needs the two counts, taken here;
//   * the reference runs FAST per cell window, so NMS never sees scores of the neighbouring cell: one CTA = one cell,
//     scores outside the cell's own detect range simply do not exist (zero border);
//   * with a cost-map the detect rows shrink to the stale window height of the last cell row: that is
//     just a different cell table.
//
// Work shape: integer stencil on u8 pixels, no tensor-core shape anywhere.  Phases of one CTA (per band of cell rows):
//   A stage    the cell's pixels (+3 px ring margin) are loaded with aligned 32-bit words and stored in shared memory
//              as one 32-bit word per pixel holding TWO vertically adjacent pixels in 16-bit lanes
//              (pix(x,y) | pix(x,y+1)<<16): every ring sample of a vertical pixel pair is then ONE conflict-free LDS.32
//              that is already in the packed 16x2 layout of the DPX min/max instructions (VIMNMX[3].U16x2);
//   B reject   every pixel pair takes the opposing-pair test on 4 of the 8 ring diameters (any 9-arc contains one end
//              of every diameter): 8 packed differences, 14 packed min/max.  ~4 % of pixels survive; their pair
//              positions are appended to a list (order irrelevant);
//   C score    dense loop over the list: 16 packed differences and the exact score network — min over each 9-arc as a
//              min3 of three 3-minima, max over arcs, both polarities: 88 packed min/max for two pixels;
//   D nms      dense loop over the list: strict 3x3 maximum inside the cell's score map; survivors set a bit;
//   E emit     the bitmap is scanned in row-major order (= cv::FAST's emission order): popcount + scan give every
//              corner its slot in the cell list; the two threshold counts are reduced on the way.
// No atomics decide any order.  Tall cells are processed in bands of rows (2 score rows recomputed per band).
#pragma once
#include "common.cuh"

namespace ivg {

__device__ __forceinline__ unsigned vmin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned vmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }

// T = S + 257 per 16-bit lane; D[k] = 256 + centre - ring[k]
__device__ __forceinline__ unsigned fast_score_pair(const unsigned (&D)[16]) {
  unsigned lo3[16], hi3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo3[k] = vmin3(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
    hi3[k] = vmax3(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
  }
  unsigned mn[16], mx[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    mn[k] = vmin3(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);   // min over the arc k..k+8
    mx[k] = vmax3(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
  }
  unsigned b0 = vmax3(mn[0], mn[1], mn[2]), b1 = vmax3(mn[3], mn[4], mn[5]), b2 = vmax3(mn[6], mn[7], mn[8]);
  unsigned b3 = vmax3(mn[9], mn[10], mn[11]), b4 = vmax3(mn[12], mn[13], mn[14]);
  const unsigned best = vmax3(vmax3(b0, b1, b2), vmax3(b3, b4, mn[15]), 0u);
  unsigned w0 = vmin3(mx[0], mx[1], mx[2]), w1 = vmin3(mx[3], mx[4], mx[5]), w2 = vmin3(mx[6], mx[7], mx[8]);
  unsigned w3 = vmin3(mx[9], mx[10], mx[11]), w4 = vmin3(mx[12], mx[13], mx[14]);
  const unsigned worst = vmin3(vmin3(w0, w1, w2), vmin3(w3, w4, mx[15]), 0xFFFFFFFFu);
  // bright: best-256, dark: 256-worst  =>  S + 257 = max(best, 512 - worst)
  return __vmaxu2(best, 0x02000200u - worst);
}

constexpr int FC_THREADS = 160; // CTA size of k_fast_cells (per-cell fixed costs are paid once per warp: fewer, busier warps)
constexpr int FC_WARPS = FC_THREADS / 32;

__global__ void __launch_bounds__(FC_THREADS) k_fast_cells(FrameSet fs) {
  extern __shared__ __align__(16) unsigned char fsm[];
  __shared__ int wcnt[2][FC_WARPS];
  __shared__ int sred[3][FC_WARPS];

  const CellDev c = fs.cells[blockIdx.x];
  const LevelDev& L = fs.lv[c.level];
  const size_t img = blockIdx.y;
  const uint8_t* pix = fs.pyr + img * fs.planeBytes + L.planeOff;
  uint32_t* list = fs.cellList + img * fs.listCapTotal + c.listOff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cw = c.cw, ch = c.ch;
  const int SP = L.fSP, SS = L.fSS, BH = L.fBH, BW = L.fBW, pitch = L.pitch, segCap = L.fSeg;
  // shared layout: packed pixel pairs | score bytes | survivor bitmap | pair list
  uint32_t* sp = reinterpret_cast<uint32_t*>(fsm);
  const int ssBytes = (SS * (BH + 4) + 15) & ~15, bitBytes = (4 * BW * (BH + 2) + 15) & ~15;
  uint8_t* ss = fsm + (size_t)4 * SP * (BH + 8);
  uint32_t* sbit = reinterpret_cast<uint32_t*>(ss + ssBytes);
  uint16_t* slist = reinterpret_cast<uint16_t*>(ss + ssBytes + bitBytes);
  const int xa = (c.x0 - 3) & ~3;          // global x of staged column 0 (word aligned)
  const int cOff = c.x0 - xa;              // staged column of detect column 0
  const int scoreTh = fs.scoreTh;
  const unsigned kK2 = ((unsigned)scoreTh + 1u) * 0x00010001u;      // th + 1 in both 16-bit lanes
  const int nChunks = (cw + 31) >> 5;

  int running = 0, nIni = 0, nMin = 0, par = 0;

  for (int r0 = 0; r0 < ch; r0 += BH) {
    const int r1 = min(r0 + BH, ch);
    const int s0 = max(r0 - 1, 0), s1 = min(r1 + 1, ch);
    const int nPairs = (s1 - s0 + 1) >> 1;
    const int nWR = 2 * nPairs + 5;

    // ---- A: stage. word row q holds pixel rows (q, q+1) counted from global row gy0
    const int gy0 = c.y0 + s0 - 3;
    {
      // warp = row segment, lane = column group: rows are walked top to bottom with the previous row carried in a
      // register, so every global word is loaded once (plus one seed row per segment); no integer division anywhere
      const int nG4 = SP >> 2;
      const int segRows = (nWR + FC_WARPS - 1) / FC_WARPS;
      const int q0 = warp * segRows, q1 = min(q0 + segRows, nWR);
      if (q0 < q1) {
        for (int g = lane; g < nG4; g += 32) {
          const bool in = xa + 4 * g < pitch;
          const uint8_t* prow = pix + (size_t)(gy0 + q0) * pitch + xa + 4 * g;
          uint32_t a = in ? __ldg(reinterpret_cast<const uint32_t*>(prow)) : 0u;
          uint32_t* dst = sp + q0 * SP + 4 * g;
          for (int q = q0; q < q1; ++q) {
            prow += pitch;
            const uint32_t b = in ? __ldg(reinterpret_cast<const uint32_t*>(prow)) : 0u;
            uint4 o;
            o.x = __byte_perm(a, b, 0x0400) & 0x00FF00FFu;
            o.y = __byte_perm(a, b, 0x0501) & 0x00FF00FFu;
            o.z = __byte_perm(a, b, 0x0602) & 0x00FF00FFu;
            o.w = __byte_perm(a, b, 0x0703) & 0x00FF00FFu;
            *reinterpret_cast<uint4*>(dst) = o;
            dst += SP;
            a = b;
          }
        }
      }
    }
    {
      uint4* z = reinterpret_cast<uint4*>(ss);                // scores + bitmap are contiguous
      for (int i = tid; i < ((ssBytes + bitBytes) >> 4); i += FC_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // ---- B: opposing-pair rejection on diameters 0-8, 2-10, 4-12, 6-14
    uint16_t* wlist = slist + warp * segCap;        // this warp's private segment: no atomics, no ordering needed
    int wn = 0;
    {
      // work item = (row pair p, segment of up to 4 chunks of 32 columns); the inner loop only bumps pointers
      const int nSegs = (nChunks + 3) >> 2;
      const unsigned ltmask = (1u << lane) - 1u;
      int p = 0, sg = warp;                          // item index it = p * nSegs + sg, advanced by FC_WARPS without dividing
      while (sg >= nSegs) { sg -= nSegs; ++p; }
      for (; p < nPairs; sg += FC_WARPS) {
        while (sg >= nSegs) { sg -= nSegs; ++p; }
        if (p >= nPairs) break;
        int x = sg * 128 + lane;
        const uint32_t* ctr = sp + (2 * p + 3) * SP + cOff + x;
        const uint32_t* cp3 = ctr + 3 * SP;  const uint32_t* cm3 = ctr - 3 * SP;     // ring rows +-3, +-2 (row 0 via ctr)
        const uint32_t* cp2 = ctr + 2 * SP;  const uint32_t* cm2 = ctr - 2 * SP;
        const int xend = min(cw, sg * 128 + 128);
        const int ebase = p * 512;
#pragma unroll 1
        for (; x - lane < xend; x += 32, ctr += 32, cp3 += 32, cm3 += 32, cp2 += 32, cm2 += 32) {
          bool pass = false;
          if (x < xend) {
            // directly on the packed pixels (no differences needed for a reject test):
            //   bright possible  <=>  min over diameters of max(ring_k, ring_k+8) >= centre + th + 1
            //   dark possible    <=>  max over diameters of min(ring_k, ring_k+8) <= centre - th - 1
            const unsigned C = ctr[0];
            const unsigned r0 = cp3[0], r8 = cm3[0], r2 = cp2[2], r10 = cm2[-2], r4 = ctr[3], r12 = ctr[-3], r6 = cm2[2], r14 = cp2[-2];
            const unsigned a = __vminu2(vmin3(__vmaxu2(r0, r8), __vmaxu2(r2, r10), __vmaxu2(r4, r12)), __vmaxu2(r6, r14));
            const unsigned b = __vmaxu2(vmax3(__vminu2(r0, r8), __vminu2(r2, r10), __vminu2(r4, r12)), __vminu2(r6, r14));
            // bit 15 of each 16-bit lane: (a >= C + K) and (C >= b + K), K = th + 1; no borrow can cross lanes
            const unsigned X = a + (0x80008000u - kK2) - C;
            const unsigned Y = C + (0x80008000u - kK2) - b;
            pass = ((X | Y) & 0x80008000u) != 0u;
          }
          const unsigned m = __ballot_sync(0xffffffffu, pass);
          if (pass) wlist[wn + __popc(m & ltmask)] = (uint16_t)(ebase + x);   // cw <= 512 enforced by the host
          wn += __popc(m);
        }
      }
    }
    __syncwarp();

    // ---- C: exact score of the surviving pairs
    for (int j = lane; j < wn; j += 32) {
      const int e = wlist[j], p = e >> 9, x = e & 511;
      const uint32_t* ctr = sp + (2 * p + 3) * SP + cOff + x;
      const unsigned C = ctr[0] + 0x01000100u;
      unsigned D[16];
      D[0] = C - ctr[3 * SP];          D[1] = C - ctr[3 * SP + 1];      D[2] = C - ctr[2 * SP + 2];      D[3] = C - ctr[SP + 3];
      D[4] = C - ctr[3];               D[5] = C - ctr[-SP + 3];         D[6] = C - ctr[-2 * SP + 2];     D[7] = C - ctr[-3 * SP + 1];
      D[8] = C - ctr[-3 * SP];         D[9] = C - ctr[-3 * SP - 1];     D[10] = C - ctr[-2 * SP - 2];    D[11] = C - ctr[-SP - 3];
      D[12] = C - ctr[-3];             D[13] = C - ctr[SP - 3];         D[14] = C - ctr[2 * SP - 2];     D[15] = C - ctr[3 * SP - 1];
      const unsigned T = fast_score_pair(D);
      int sA = (int)(T & 0xFFFFu) - 257, sB = (int)(T >> 16) - 257;
      sA = sA >= scoreTh ? sA : 0;
      sB = sB >= scoreTh ? sB : 0;
      uint8_t* o = ss + (2 * p + 1) * SS + 4 + x;
      o[0] = (uint8_t)sA;
      if (s0 + 2 * p + 1 < s1) o[SS] = (uint8_t)sB;
    }
    __syncthreads();

    // ---- D: 3x3 strict maximum inside the cell (zero border = "no score outside the cell")
    for (int j = lane; j < 2 * wn; j += 32) {
      const int e = wlist[j >> 1], p = e >> 9, x = e & 511, half = j & 1;
      const int srow = 2 * p + half;                       // score row relative to s0
      const int row = s0 + srow;                           // detect row of the cell
      if (row < r0 || row >= r1) continue;                 // overlap rows belong to the neighbouring band
      const uint8_t* s = ss + (srow + 1) * SS + 4 + x;
      const int v = s[0];
      if (v == 0) continue;
      const int m = max(max(max((int)s[-1], (int)s[1]), max((int)s[-SS - 1], (int)s[-SS])),
                        max(max((int)s[-SS + 1], (int)s[SS - 1]), max((int)s[SS], (int)s[SS + 1])));
      if (v > m) atomicOr(&sbit[(row - r0) * BW + (x >> 5)], 1u << (x & 31));
    }
    __syncthreads();

    // ---- E: emit in row-major order
    const int nWords = (r1 - r0) * BW;
    for (int base = 0; base < nWords; base += FC_THREADS) {
      const int i = base + tid;
      unsigned bits = i < nWords ? sbit[i] : 0u;
      const int cnt = __popc(bits);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) wcnt[par][warp] = incl;
      __syncthreads();
      int off = running + incl - cnt, tot = 0;
#pragma unroll
      for (int w = 0; w < FC_WARPS; ++w) { const int v = wcnt[par][w]; if (w < warp) off += v; tot += v; }
      running += tot;
      par ^= 1;
      if (bits) {
        const int row = i / BW, xb = (i - row * BW) * 32;
        const uint8_t* srow = ss + (r0 + row - s0 + 1) * SS + 4 + xb;
        while (bits) {
          const int k = __ffs(bits) - 1;
          bits &= bits - 1;
          const int s = srow[k];
          list[off++] = pack_xys(c.x0 + xb + k, c.y0 + r0 + row, s);
          nIni += s >= fs.iniTh;
          nMin += s >= fs.minTh;
        }
      }
    }
    __syncthreads();   // smem is restaged by the next band
  }

  // ---- per-cell counters (+ cost-map window sum for the introspection budgets)
  unsigned csum = 0;
  if (fs.weighted) {
    const uint8_t* q = fs.qual + img * fs.planeBytes + L.planeOff;
    for (int y = warp; y < c.wh; y += FC_WARPS) {
      const uint8_t* qr = q + (size_t)(c.wy + y) * L.pitch + c.wx;
      for (int x = lane; x < c.ww; x += 32) csum += __ldg(qr + x);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    nIni += __shfl_xor_sync(0xffffffffu, nIni, o);
    nMin += __shfl_xor_sync(0xffffffffu, nMin, o);
    csum += __shfl_xor_sync(0xffffffffu, csum, o);
  }
  if (lane == 0) { sred[0][warp] = nIni; sred[1][warp] = nMin; sred[2][warp] = (int)csum; }
  __syncthreads();
  if (tid == 0) {
    int a = 0, b = 0; unsigned s = 0;
#pragma unroll
    for (int w = 0; w < FC_WARPS; ++w) { a += sred[0][w]; b += sred[1][w]; s += (unsigned)sred[2][w]; }
    fs.cellCount[img * fs.nCellsTotal + blockIdx.x] = make_int2(b, a);   // x: corners at minTh, y: corners at iniTh
    fs.cellCost[img * fs.cellCostStride + blockIdx.x] = s;
  }
}

}  // namespace ivg

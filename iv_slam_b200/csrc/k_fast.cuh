// k_fast.cuh — K2+K3a fused: per-cell FAST-9/16 score, cell-local 3x3 non-max suppression and ordered (row-major)
// compaction of the surviving corners into the cell's list.
//
// Replaces the per-cell cv::FAST(cellImage, kps, th, true) calls of ComputeKeyPointsOld
// (introspective_ORB_SLAM/src/ORBextractor.cc:1045 and :1051) including their emission order, and cv::sum over the
// cost-map window (:976-978).  SURVEY Appendix A.3:
//   * the corner score S (largest threshold at which the pixel is still a 9-arc corner) does not depend on the
//     threshold, "corner at th" <=> S >= th, and a corner with S >= th survives the 3x3 NMS of a run at any lower
//     threshold iff it survives at th (the extra neighbours all score below th).  The cell is processed like the
//     reference does (:1045-1052): once at iniThFAST — few pixels pass the reject test at that threshold, so the exact
//     scoring, NMS and emission touch few candidates — and again at minThFAST only when that left <= 3 corners
//     (fs.fastRetry; 0 in OctTree mode, :818-822).  cellCount = (entries in the list, corners at iniTh);
//   * the reference runs FAST per cell window, so NMS never sees scores of the neighbouring cell: one CTA = one cell,
//     scores outside the cell's own detect range simply do not exist (zero border);
//   * with a cost-map the detect rows shrink to the stale window height of the last cell row: that is
//     just a different cell table.
//
// Work shape: integer stencil on u8 pixels, no tensor-core shape anywhere.  Phases of one CTA (per band of cell rows):
//   A stage    the cell's pixels (+3 px ring margin) are loaded with aligned 32-bit words and stored in shared memory
//              as one 32-bit word per HORIZONTAL pixel pair, one pixel per 16-bit lane (pix(2w,y) | pix(2w+1,y)<<16):
//              2 B of shared memory per pixel, and a word is already in the packed 16x2 layout of the DPX min/max
//              instructions (VIMNMX[3].U16x2).  Ring samples at even dx are one conflict-free LDS.32, samples at odd
//              dx are two LDS.32 and one PRMT;
//   B reject   every pixel pair takes the opposing-pair test on the two cardinal ring diameters (any 9-arc contains one
//              end of every diameter): 2 PRMT and 6 packed min/max per pair; a lane takes two adjacent pairs per step so
//              that five 64-bit loads serve both.  ~2 % of the pairs survive at iniTh; each lane appends them to its own
//              list, the lists are merged once per band (order irrelevant);
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

#ifndef IVG_FC_THREADS
#define IVG_FC_THREADS 128
#endif
constexpr int FC_THREADS = IVG_FC_THREADS; // CTA size of k_fast_cells (per-cell fixed costs are paid once per warp: fewer, busier warps)
constexpr int FC_WARPS = FC_THREADS / 32;
#ifndef IVG_FC_THREADS_LAT
#define IVG_FC_THREADS_LAT 256
#endif
constexpr int FC_THREADS_LAT = IVG_FC_THREADS_LAT, FC_WARPS_LAT = FC_THREADS_LAT / 32;
constexpr int FC_SLACK = 272;   // bytes after the staged pixels that B's masked lanes may read (at most 65 words past the last row)

// FCT threads per cell: FC_THREADS in batches (8 CTAs per SM keep every scheduler busy; fewer, busier warps pay the per-cell
// fixed costs less often), FC_THREADS_LAT when one or two frames' cells are all there is (two CTAs per SM: twice the warps on
// each cell halve its critical path).  The per-level band geometry (fBH / fBX / fSeg) depends on the warp count: the host
// passes the matching set in the FrameSet.
template <int FCT>
__global__ void __launch_bounds__(FCT) k_fast_cells(FrameSet fs) {
  constexpr int FCW = FCT / 32;
  extern __shared__ __align__(16) unsigned char fsm[];
  __shared__ int wcnt[2][FCW];
  __shared__ int sred[3][FCW];

  const CellDev c = fs.cells[blockIdx.x];
  const LevelDev& L = fs.lv[c.level];
  const size_t img = blockIdx.y;
  const uint8_t* pix = fs.pyr + img * fs.planeBytes + L.planeOff;
  uint32_t* list = fs.cellList + img * fs.listCapTotal + c.listOff;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cw = c.cw, ch = c.ch;
  const int SP = L.fSP, SS = L.fSS, BH = L.fBH, BW = L.fBW, pitch = L.pitch, segCap = L.fSeg;
  const int sh = L.fShift;                 // list entry = (score row << sh) | pair index
  const int piMask = (1 << sh) - 1;
  // shared layout: packed pixel pairs (+ slack for the masked overreads of B) | score bytes | survivor bitmap | pair list
  uint32_t* sp = reinterpret_cast<uint32_t*>(fsm);
  // a single-band cell needs BH + 6 pixel rows and BH + 2 score rows; bands add one overlap score row per side (L.fBX = 2)
  const int ssBytes = (SS * (BH + 2 + L.fBX) + 15) & ~15, bitBytes = (4 * BW * (BH + 2) + 15) & ~15;
  uint8_t* ss = fsm + (((size_t)4 * SP * (BH + 6 + L.fBX) + 15) & ~(size_t)15) + FC_SLACK;
  uint32_t* sbit = reinterpret_cast<uint32_t*>(ss + ssBytes);
  uint16_t* slist = reinterpret_cast<uint16_t*>(ss + ssBytes + bitBytes);
  const int xa = (c.x0 - 3) & ~3;          // global x of staged column 0 (word aligned)
  const int cOff = c.x0 - xa;              // staged column of detect column 0 (3..6)
  const int lead = cOff & 1;               // 1: pair 0 starts one column left of the detect range
  const int w0 = cOff >> 1;                // staged word of pair 0
  const int nPr = (lead + cw + 1) >> 1;    // pixel pairs per row
  // pass 0 runs at iniTh; a second pass at minTh follows only when the first one left too few corners (and minTh is lower:
  // with iniTh <= minTh the retry is a subset of the first pass and the consumers filter the list by score)
  int scoreTh = fs.iniTh;
  int running = 0, nIniPass = 0, nHigh = 0, par = 0;
  const bool countHigh = fs.iniTh < fs.minTh;     // rare configuration: count the corners at the higher threshold per thread

  for (int pass = 0; pass < 2; ++pass) {
  const unsigned kK2 = ((unsigned)scoreTh + 1u) * 0x00010001u;      // th + 1 in both 16-bit lanes
  const unsigned kB = 0x80008000u - kK2;
  running = 0;

  for (int r0 = 0; r0 < ch; r0 += BH) {
    const int r1 = min(r0 + BH, ch);
    const int s0 = max(r0 - 1, 0), s1 = min(r1 + 1, ch);
    const int nSR = s1 - s0;               // score rows of this band (1 overlap row per side for the NMS)
    const int nR = nSR + 6;

    // ---- A: stage pixel rows gy0 .. gy0+nR-1; warp = row, lane = group of 4 pixels
    const int gy0 = c.y0 + s0 - 3;
    if (pass == 0 || ch > BH) {      // the minTh retry of a single-band cell finds its pixels still staged
      // per-thread source/destination pointers are set up once and bumped per row; two column groups per pass
      const int nG4 = SP >> 1;
      const int rstep = FCW * pitch, dstep = FCW * SP;
      for (int gb = 0; gb < nG4; gb += 64) {
        const int g0 = gb + lane, g1 = g0 + 32;
        const bool st0 = g0 < nG4, st1 = g1 < nG4;
        const bool ld0 = st0 && xa + 4 * g0 < L.w, ld1 = st1 && xa + 4 * g1 < L.w;     // words past the image width (row padding) are never written: stage zeros
        const uint8_t* src = pix + (size_t)(gy0 + warp) * pitch + xa + 4 * g0;
        uint32_t* dst = sp + warp * SP + 2 * g0;
#pragma unroll 2
        for (int q = warp; q < nR; q += FCW, src += rstep, dst += dstep) {
          const uint32_t a0 = ld0 ? __ldg(reinterpret_cast<const uint32_t*>(src)) : 0u;
          const uint32_t a1 = ld1 ? __ldg(reinterpret_cast<const uint32_t*>(src + 128)) : 0u;
          if (st0) *reinterpret_cast<uint2*>(dst) = make_uint2(__byte_perm(a0, 0u, 0x4140), __byte_perm(a0, 0u, 0x4342));
          if (st1) *reinterpret_cast<uint2*>(dst + 64) = make_uint2(__byte_perm(a1, 0u, 0x4140), __byte_perm(a1, 0u, 0x4342));
        }
      }
    }
    {
      uint4* z = reinterpret_cast<uint4*>(ss);                // scores + bitmap are contiguous
      for (int i = tid; i < ((ssBytes + bitBytes) >> 4); i += FCT) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // ---- B: opposing-pair rejection on the two cardinal diameters 0-8 (vertical) and 4-12 (horizontal)
    // Any 9-arc contains one end of every diameter, so a bright (dark) corner needs a bright (dark) end on both.  Two
    // diameters cost 5 shared loads and 12 packed min/max per four pixels and at iniTh they reject all but ~2 % of the
    // pixel pairs; the pairs that pass go straight to the exact score.  A lane takes two adjacent pixel pairs per step
    // (their samples overlap: five 64-bit loads serve both).  Rows are walked from an even word so that the 64-bit loads
    // are aligned; the slots outside [0, nPr) (the word before the first pair, lanes past the end of the row) read staged
    // neighbours / slack and are dropped when the lists are merged.  Every lane appends to its own lane-strided list
    // (entry k of lane l at [32 k + l]): two predicated instructions per pair instead of a ballot-ranked append per step.
    uint16_t* wlist = slist + warp * segCap;        // this warp's private segment: no atomics, no ordering needed
    int wn = 0;
    {
      const unsigned ltmask = (1u << lane) - 1u;
      const int wE = w0 & ~1, pOff = w0 - wE;       // first (even) word walked; slot q holds pair q - pOff
      const int nSteps = (nPr + pOff + 63) >> 6;    // 64 slots per warp step
      // directly on the packed pixels (no differences needed for a reject test):
      //   bright possible  <=>  min over diameters of max(ring_k, ring_k+8) >= centre + th + 1
      //   dark possible    <=>  max over diameters of min(ring_k, ring_k+8) <= centre - th - 1
      auto reject_test = [&](unsigned C, unsigned r0_, unsigned r8, unsigned r4, unsigned r12) {
        const unsigned a = __vminu2(__vmaxu2(r0_, r8), __vmaxu2(r4, r12));
        const unsigned b = __vmaxu2(__vminu2(r0_, r8), __vminu2(r4, r12));
        // bit 15 of each 16-bit lane: (a >= C + K) and (C >= b + K), K = th + 1; no borrow can cross lanes
        const unsigned X = a + kB - C;
        const unsigned Y = C + kB - b;
        return ((X | Y) & 0x80008000u) != 0u;
      };
      uint16_t* lp = wlist + lane;
      const int up3 = 3 * (SP >> 1);                // three rows in 64-bit words (SP is even)
      const uint2* row = reinterpret_cast<const uint2*>(sp + (warp + 3) * SP + wE) + lane;
      unsigned e0 = (unsigned)(warp << sh) + 2u * lane;            // (score row, slot); fits 16 bits: the host bounds the band height
      for (int y = warp; y < nSR; y += FCW, row += FCW * (SP >> 1), e0 += (unsigned)FCW << sh) {
        const uint2* ctr = row;
        unsigned e = e0;
        for (int st = 0; st < nSteps; ++st, ctr += 32, e += 64) {
          const uint2 ca = ctr[-1], cb = ctr[0], cc = ctr[1];        // words W-2 .. W+3 of the centre row
          const uint2 t3 = ctr[up3], b3 = ctr[-up3];
          if (reject_test(cb.x, t3.x, b3.x, __byte_perm(cb.y, cc.x, 0x5432), __byte_perm(ca.x, ca.y, 0x5432))) { *lp = (uint16_t)e; lp += 32; }
          if (reject_test(cb.y, t3.y, b3.y, __byte_perm(cc.x, cc.y, 0x5432), __byte_perm(ca.y, cb.x, 0x5432))) { *lp = (uint16_t)(e + 1); lp += 32; }
        }
      }
      // merge the lane lists into one dense list, in place: row k of the lane-strided layout (entry k of every lane) is
      // read by the whole warp before its survivors are written at [wn, wn + count) <= 32 k, so nothing unread is overwritten
      const int cnt = (int)(lp - (wlist + lane)) >> 5;
      __syncwarp();                                                // the lanes' appends are ordered before the merge's stores
      const int maxc = __reduce_max_sync(0xffffffffu, cnt);
      for (int k = 0; k < maxc; ++k) {
        int v = 0;
        bool has = k < cnt;
        if (has) {
          v = wlist[32 * k + lane];
          const int pi = (v & piMask) - pOff;
          has = pi >= 0 && pi < nPr;
          v -= pOff;                                               // list entry = (score row << sh) | pair index
        }
        const unsigned m = __ballot_sync(0xffffffffu, has);
        __syncwarp();                                              // every lane has read its entry of row k
        if (has) wlist[wn + __popc(m & ltmask)] = (uint16_t)v;
        wn += __popc(m);
        __syncwarp();
      }
    }
    __syncwarp();

    // ---- C: exact score of the surviving pairs
    for (int j = lane; j < wn; j += 32) {
      const int e = wlist[j], y = e >> sh, pi = e & piMask;
      const uint32_t* ctr = sp + (y + 3) * SP + w0 + pi;
      const unsigned C = ctr[0] + 0x01000100u;
      unsigned D[16];
      {
        const uint32_t* r = ctr + 3 * SP;
        const unsigned a = r[-1], b = r[0], d = r[1];
        D[15] = C - __byte_perm(a, b, 0x5432);  D[0] = C - b;  D[1] = C - __byte_perm(b, d, 0x5432);
      }
      {
        const uint32_t* r = ctr - 3 * SP;
        const unsigned a = r[-1], b = r[0], d = r[1];
        D[9] = C - __byte_perm(a, b, 0x5432);   D[8] = C - b;  D[7] = C - __byte_perm(b, d, 0x5432);
      }
      D[14] = C - ctr[2 * SP - 1];   D[2] = C - ctr[2 * SP + 1];
      D[10] = C - ctr[-2 * SP - 1];  D[6] = C - ctr[-2 * SP + 1];
      {
        const uint32_t* r = ctr + SP;
        D[13] = C - __byte_perm(r[-2], r[-1], 0x5432);  D[3] = C - __byte_perm(r[1], r[2], 0x5432);
      }
      D[12] = C - __byte_perm(ctr[-2], ctr[-1], 0x5432);  D[4] = C - __byte_perm(ctr[1], ctr[2], 0x5432);
      {
        const uint32_t* r = ctr - SP;
        D[11] = C - __byte_perm(r[-2], r[-1], 0x5432);  D[5] = C - __byte_perm(r[1], r[2], 0x5432);
      }
      const unsigned T = fast_score_pair(D);
      const int x = 2 * pi - lead;                            // detect column of the low lane
      int sA = (int)(T & 0xFFFFu) - 257, sB = (int)(T >> 16) - 257;
      sA = (sA >= scoreTh && x >= 0) ? sA : 0;
      sB = (sB >= scoreTh && x + 1 < cw) ? sB : 0;
      *reinterpret_cast<uint16_t*>(ss + (y + 1) * SS + 2 + 2 * pi) = (uint16_t)(sA | (sB << 8));
    }
    __syncthreads();

    // ---- D: 3x3 strict maximum inside the cell (zero border = "no score outside the cell")
    for (int j = lane; j < 2 * wn; j += 32) {
      const int e = wlist[j >> 1], y = e >> sh, pi = e & piMask, half = j & 1;
      const int row = s0 + y;                              // detect row of the cell
      if (row < r0 || row >= r1) continue;                 // overlap rows belong to the neighbouring band
      const uint8_t* s = ss + (y + 1) * SS + 2 + 2 * pi + half;
      const int v = s[0];
      if (v == 0) continue;
      const int m = max(max(max((int)s[-1], (int)s[1]), max((int)s[-SS - 1], (int)s[-SS])),
                        max(max((int)s[-SS + 1], (int)s[SS - 1]), max((int)s[SS], (int)s[SS + 1])));
      const int x = 2 * pi - lead + half;
      if (v > m) atomicOr(&sbit[(row - r0) * BW + (x >> 5)], 1u << (x & 31));
    }
    __syncthreads();

    // ---- E: emit in row-major order; thread t owns the bitmap words [t*per, (t+1)*per): one scan per band
    {
      const int nWords = (r1 - r0) * BW;
      const int per = (nWords + FCT - 1) / FCT;
      const int wbeg = tid * per, wend = min(wbeg + per, nWords);
      int cnt = 0;
      for (int i = wbeg; i < wend; ++i) cnt += __popc(sbit[i]);
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
      for (int w = 0; w < FCW; ++w) { const int v = wcnt[par][w]; if (w < warp) off += v; tot += v; }
      running += tot;
      par ^= 1;
      if (cnt) {
        int row = wbeg / BW, xw = wbeg - row * BW;
        for (int i = wbeg; i < wend; ++i) {
          unsigned bits = sbit[i];
          const int xb = xw * 32;
          const uint8_t* srow = ss + (r0 + row - s0 + 1) * SS + 2 + lead + xb;
          while (bits) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            const int s = srow[k];
            list[off++] = pack_xys(c.x0 + xb + k, c.y0 + r0 + row, s);
            if (countHigh) nHigh += s >= fs.minTh;
          }
          if (++xw == BW) { xw = 0; ++row; }
        }
      }
    }
    __syncthreads();   // smem is restaged by the next band
  }
    if (pass == 0) nIniPass = running;
    if (pass == 1 || running > fs.fastRetry || fs.minTh >= fs.iniTh) break;     // CTA-uniform
    scoreTh = fs.minTh;
  }

  // ---- per-cell counters (+ cost-map window sum for the introspection budgets)
  unsigned csum = 0;
  if (fs.weighted) {
    const uint8_t* q = fs.qual + img * fs.planeBytes + L.planeOff;
    // aligned words + DP4A against a byte mask that cuts the window's first and last word to [wx, wx + ww)
    const int xa0 = c.wx & ~3, xe = c.wx + c.ww, nw = (xe - xa0 + 3) >> 2;
    for (int y = warp; y < c.wh; y += FCW) {
      const uint32_t* qr = reinterpret_cast<const uint32_t*>(q + (size_t)(c.wy + y) * L.pitch + xa0);
      for (int k = lane; k < nw; k += 32) {
        const int b0 = xa0 + 4 * k;
        unsigned m = 0x01010101u;
        if (b0 < c.wx) m &= 0x01010101u << (8 * (c.wx - b0));
        if (b0 + 4 > xe) m &= 0x01010101u >> (8 * (b0 + 4 - xe));
        csum = __dp4a(__ldg(qr + k), m, csum);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    if (countHigh) nHigh += __shfl_xor_sync(0xffffffffu, nHigh, o);
    csum += __shfl_xor_sync(0xffffffffu, csum, o);
  }
  if (lane == 0) { sred[0][warp] = nHigh; sred[2][warp] = (int)csum; }
  __syncthreads();
  if (tid == 0) {
    int a = 0; unsigned s = 0;
#pragma unroll
    for (int w = 0; w < FCW; ++w) { a += sred[0][w]; s += (unsigned)sred[2][w]; }
    // x: entries in the list (corners at the lower threshold), y: corners at iniTh
    fs.cellCount[img * fs.nCellsTotal + blockIdx.x] = countHigh ? make_int2(a, nIniPass) : make_int2(running, nIniPass);
    fs.cellCost[img * fs.cellCostStride + blockIdx.x] = s;
  }
}

}  // namespace ivg

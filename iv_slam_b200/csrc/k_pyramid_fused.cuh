// k_pyramid_fused.cuh — the whole resize cascade of one frame in ONE launch (the one-frame-at-a-time configuration).
//
// ComputePyramid / ComputeQualityImagePyramid (introspective_ORB_SLAM/src/ORBextractor.cc:1298-1357) build level l from
// level l-1, so the per-level kernel of k_pyramid.cuh needs nlevels-1 dependent launches: ~6.5 us each on an otherwise
// idle GPU, 45 us of a 120 us frame.  Here a CTA owns the same relative window of EVERY level ("pyramid column"): it
// loads its window of level 0 once, then walks down the cascade in shared memory, level by level, and writes its own
// part of each level to global memory on the way.  Bilinear taps reach one source pixel past a window edge, so the
// window a CTA has to COMPUTE at level l (E_l) is a few pixels larger than the part it OWNS (O_l): E_l = O_l + what
// E_(l+1) reads.  The overlap is recomputed by the neighbouring CTA — bit-identical, because every pixel of every level is
// produced by the same integer arithmetic from the same source pixels whoever computes it.  No inter-CTA dependency, no
// grid barrier, 7 launches become one.
//
// Arithmetic = k_resize_level's (OpenCV's 8-bit fixed-point bilinear path) in its direct 4-tap form:
//   T(sy) = src[sy][sx0]*cx0 + src[sy][sx1]*cx1;   dst = (((cy0*(T(sy0) >> 4)) >> 16) + ((cy1*(T(sy1) >> 4)) >> 16) + 2) >> 2
// The windows (host tables, one entry per level and tile column / tile row) are widened to whole 4-pixel words in x so that
// a thread produces one aligned 32-bit word for shared and global memory; the pixels a window gains that way reuse the
// taps of its edge pixel (nobody reads them), so the widening does not compound down the cascade.
#pragma once
#include "common.cuh"

namespace ivg {

struct PyrSpan { int o0, o1, e0, e1, t0, t1; };   // one tile column (or row) at one level: owned [o0, o1), needed [t0, t1) = own part + what the next
                                                  // level reads, computed [e0, e1) = the needed range widened to whole 4-pixel words (x only)
struct PyrFusedArgs {
  const PyrSpan* spanX; const PyrSpan* spanY;   // [level][tile column], [level][tile row]
  int TX, TY, bufBytes;
  int tapOffX[MAX_LEVELS], tapOffY[MAX_LEVELS]; // entry offsets of level l's tap slices in the shared-memory tap area
};

#ifndef IVG_PF_THREADS
#define IVG_PF_THREADS 512
#endif
constexpr int PF_THREADS = IVG_PF_THREADS;

// i / d and i % d for 0 <= i < 2^20, 0 < d: one float multiply and a fix-up instead of the ~40-instruction integer division
__device__ __forceinline__ void pf_divmod(int i, int d, float inv, int& q, int& r) {
  q = (int)((float)i * inv);
  r = i - q * d;
  if (r < 0) { --q; r += d; }
  else if (r >= d) { ++q; r -= d; }
}

__global__ void __launch_bounds__(PF_THREADS) k_pyramid_fused(FrameSet fs, const PyrFusedArgs A) {
  extern __shared__ __align__(16) unsigned char psm[];
  const int TX = A.TX, TY = A.TY, bufBytes = A.bufBytes;
  const int tx = blockIdx.x % TX, ty = blockIdx.x / TX;
  const int which = (int)blockIdx.y >= fs.nImages ? 1 : 0;           // 0 image, 1 cost-map plane of the same frame
  const int img = (int)blockIdx.y - which * fs.nImages;
  uint8_t* plane = (which ? fs.qual : fs.pyr) + (size_t)img * fs.planeBytes;
  const int tid = threadIdx.x;
  ResizeTap* stap = reinterpret_cast<ResizeTap*>(psm + 2 * (size_t)bufBytes);

  // Everything that costs a global-memory latency happens up front, in two rounds: (1) this CTA's window table of every
  // level, (2) the tap-table slices of every level plus the level-0 window.  Inside the cascade every load is then a
  // shared-memory load.
  __shared__ PyrSpan sSpanX[MAX_LEVELS], sSpanY[MAX_LEVELS];
  __shared__ int4 sSeg[2 * MAX_LEVELS];        // tap slices to stage: (first flat entry, destination offset, source index of entry 0, -)
  __shared__ int2 sSegClamp[2 * MAX_LEVELS];   // first / last source index an entry may use
  __shared__ int sNSeg, sNEnt;
  if (tid < fs.nlevels) sSpanX[tid] = A.spanX[tid * TX + tx];
  else if (tid >= 32 && tid < 32 + fs.nlevels) sSpanY[tid - 32] = A.spanY[(tid - 32) * TY + ty];
  __syncthreads();
  if (tid == 0) {
    int n = 0, e = 0;
    for (int l = 1; l < fs.nlevels; ++l) {
      const LevelDev& D = fs.lv[l];
      sSegClamp[n] = make_int2(D.rtabX + sSpanX[l].t0, D.rtabX + min(sSpanX[l].t1, D.w) - 1);
      sSeg[n++] = make_int4(e, A.tapOffX[l], D.rtabX + sSpanX[l].e0, 0);
      e += sSpanX[l].e1 - sSpanX[l].e0;
      sSegClamp[n] = make_int2(D.rtabY + sSpanY[l].t0, D.rtabY + min(sSpanY[l].t1, D.h) - 1);
      sSeg[n++] = make_int4(e, A.tapOffY[l], D.rtabY + sSpanY[l].e0, 0);
      e += sSpanY[l].e1 - sSpanY[l].e0;
    }
    sNSeg = n; sNEnt = e;
  }
  __syncthreads();
  {
    const int nSeg = sNSeg, nEnt = sNEnt;
    for (int i = tid; i < nEnt; i += PF_THREADS) {
      int s = 0;
      while (s + 1 < nSeg && sSeg[s + 1].x <= i) ++s;
      const int4 g = sSeg[s];
      const int2 c = sSegClamp[s];
      stap[g.y + i - g.x] = fs.rtab[min(max(g.z + i - g.x, c.x), c.y)];
    }
  }

  // level 0 window -> buffer 0
  PyrSpan sx = sSpanX[0], sy = sSpanY[0];
  {
    const LevelDev& S = fs.lv[0];
    const int ew4 = (sx.e1 - sx.e0) >> 2, eh = ew4 > 0 ? sy.e1 - sy.e0 : 0;
    const uint8_t* src = plane + S.planeOff + (size_t)sy.e0 * S.pitch + sx.e0;
    uint32_t* dst = reinterpret_cast<uint32_t*>(psm);
    if (eh > 0) {
      const float inv = 1.0f / (float)ew4;
      int wy, wx, dy, dx;
      pf_divmod(tid, ew4, inv, wy, wx);
      pf_divmod(PF_THREADS, ew4, inv, dy, dx);
      while (wy < eh) {
        dst[wy * ew4 + wx] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)wy * S.pitch) + wx);
        wx += dx; wy += dy;
        if (wx >= ew4) { wx -= ew4; ++wy; }
      }
    }
  }
  __syncthreads();

  for (int l = 1; l < fs.nlevels; ++l) {
    const LevelDev& D = fs.lv[l];
    const PyrSpan dx_ = sSpanX[l], dy_ = sSpanY[l];
    const uint8_t* sbuf = psm + ((l - 1) & 1) * bufBytes;
    uint8_t* dbuf = psm + (l & 1) * bufBytes;
    const int spitch = sx.e1 - sx.e0;                         // source window: columns [sx.e0, sx.e1), rows [sy.e0, sy.e1)
    const int ew4 = (dx_.e1 - dx_.e0) >> 2, eh = dy_.e1 - dy_.e0;
    const ResizeTap* tX = stap + A.tapOffX[l];               // slices: entry i = column dx_.e0 + i / row dy_.e0 + i
    const ResizeTap* tY = stap + A.tapOffY[l];
    uint8_t* gdst = plane + D.planeOff;
    if (ew4 > 0 && eh > 0) {
      const float inv = 1.0f / (float)ew4;
      int wy, wx, sdy, sdx;
      pf_divmod(tid, ew4, inv, wy, wx);
      pf_divmod(PF_THREADS, ew4, inv, sdy, sdx);
      const uint8_t* sb0 = sbuf - sy.e0 * spitch - sx.e0;     // source pixel (x, y) of the previous level at sb0[y * spitch + x]
      // four output pixels (one word) from the previous level's window; two words per trip so that the dependent
      // shared-memory loads of two independent items overlap
      auto item = [&](int iy, int ix) {
        const ResizeTap ty_ = tY[iy];
        const uint8_t* r0 = sb0 + ty_.s0 * spitch;
        const uint8_t* r1 = sb0 + ty_.s1 * spitch;
        const int b0 = ty_.c0, b1 = ty_.c1;
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const ResizeTap t = tX[4 * ix + k];
          const int T0 = r0[t.s0] * t.c0 + r0[t.s1] * t.c1;
          const int T1 = r1[t.s0] * t.c0 + r1[t.s1] * t.c1;
          const int v = (((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2;
          word |= (uint32_t)v << (8 * k);
        }
        return word;
      };
      auto emit = [&](int iy, int ix, uint32_t word) {
        const int y = dy_.e0 + iy, x4 = dx_.e0 + 4 * ix;
        reinterpret_cast<uint32_t*>(dbuf)[iy * ew4 + ix] = word;
        if (y >= dy_.o0 && y < dy_.o1 && x4 >= dx_.o0 && x4 < dx_.o1 && x4 < D.w)
          *reinterpret_cast<uint32_t*>(gdst + (size_t)y * D.pitch + x4) = word;
      };
      while (wy < eh) {
        int wy2 = wy + sdy, wx2 = wx + sdx;
        if (wx2 >= ew4) { wx2 -= ew4; ++wy2; }
        const bool two = wy2 < eh;
        const uint32_t wa = item(wy, wx);
        const uint32_t wb = two ? item(wy2, wx2) : 0u;
        emit(wy, wx, wa);
        if (two) emit(wy2, wx2, wb);
        wx = wx2 + sdx; wy = wy2 + sdy;
        if (wx >= ew4) { wx -= ew4; ++wy; }
      }
    }
    sx = dx_; sy = dy_;
    __syncthreads();
  }
}

}  // namespace ivg

// k_pyramid_fused.cuh — the whole resize cascade of one frame in ONE launch (the one-frame-at-a-time configuration).
//
// ComputePyramid / ComputeQualityImagePyramid (introspective_ORB_SLAM/src/ORBextractor.cc:1298-1357) build level l from
// level l-1, so the per-level kernel of k_pyramid.cuh needs nlevels-1 dependent launches: ~6.5 us each on an otherwise
// idle GPU, 45 us per frame where this kernel takes 14.  Here a CTA owns the same relative window of EVERY level ("pyramid column"): it
// loads its window of level 0 once, then walks down the cascade in shared memory, level by level, and writes its own
// part of each level to global memory on the way.  Bilinear taps reach one source pixel past a window edge, so the
// window a CTA has to COMPUTE at level l (E_l) is a few pixels larger than the part it OWNS (O_l): E_l = O_l + what
// E_(l+1) reads.  The overlap is recomputed by the neighbouring CTA — bit-identical, because every pixel of every level is
// produced by the same integer arithmetic from the same source pixels whoever computes it.  No inter-CTA dependency, no
// grid barrier, 7 launches become one.
//
// Arithmetic = k_resize_level's (OpenCV's 8-bit fixed-point bilinear path), per level the same two
// separable passes over shared memory: T(sy) = src[sy][sx0]*cx0 + src[sy][sx1]*cx1 for every source row the window's
// rows read (kept as T >> 4), then dst = (((cy0*(T(sy0) >> 4)) >> 16) + ((cy1*(T(sy1) >> 4)) >> 16) + 2) >> 2.
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
  int tOff;                                     // byte offset of the horizontal-pass buffer (T >> 4, two 16-bit values per word)
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

  // Everything that costs a global-memory latency happens up front and at once: warp w stages tap slice w (x or y taps of
  // one level, read through this CTA's window entry of that level), the first threads copy the window table of every level
  // to shared memory, and all threads load the level-0 window.  Inside the cascade every load is a shared-memory load.
  __shared__ PyrSpan sSpanX[MAX_LEVELS], sSpanY[MAX_LEVELS];
  {
    const int warp = tid >> 5, lane = tid & 31, nSeg = 2 * (fs.nlevels - 1);
    for (int sgi = warp; sgi < nSeg; sgi += PF_THREADS / 32) {
      const int l = 1 + (sgi >> 1);
      const bool isY = sgi & 1;
      const PyrSpan sp = isY ? A.spanY[l * TY + ty] : A.spanX[l * TX + tx];
      const LevelDev& D = fs.lv[l];
      const int rt = isY ? D.rtabY : D.rtabX, dim = isY ? D.h : D.w;
      const int lo = rt + sp.t0, hi = rt + min(sp.t1, dim) - 1, src0 = rt + sp.e0, len = sp.e1 - sp.e0;      // entries outside the needed range reuse its edge
      ResizeTap* dst = stap + (isY ? A.tapOffY[l] : A.tapOffX[l]);
      for (int i = lane; i < len; i += 32) dst[i] = fs.rtab[min(max(src0 + i, lo), hi)];
    }
  }
  if (tid < fs.nlevels) sSpanX[tid] = A.spanX[tid * TX + tx];
  else if (tid >= 32 && tid < 32 + fs.nlevels) sSpanY[tid - 32] = A.spanY[(tid - 32) * TY + ty];

  // level 0 window -> buffer 0 (four independent loads in flight per thread)
  PyrSpan sx = A.spanX[tx], sy = A.spanY[ty];
  {
    const LevelDev& S = fs.lv[0];
    const int ew4 = (sx.e1 - sx.e0) >> 2, eh = ew4 > 0 ? sy.e1 - sy.e0 : 0;
    const uint8_t* src = plane + S.planeOff + (size_t)sy.e0 * S.pitch + sx.e0;
    uint32_t* dst = reinterpret_cast<uint32_t*>(psm);
    if (eh > 0) {
      const float inv = __fdividef(1.0f, (float)ew4);
      int wy, wx, dy, dx;
      pf_divmod(tid, ew4, inv, wy, wx);
      pf_divmod(PF_THREADS, ew4, inv, dy, dx);
      while (wy < eh) {
        int py[4], px[4];
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          py[k] = wy; px[k] = wx;
          wx += dx; wy += dy;
          if (wx >= ew4) { wx -= ew4; ++wy; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (py[k] < eh) v[k] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)py[k] * S.pitch) + px[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (py[k] < eh) dst[py[k] * ew4 + px[k]] = v[k];
      }
    }
  }
  __syncthreads();

  uint32_t* sT = reinterpret_cast<uint32_t*>(psm + A.tOff);
  for (int l = 1; l < fs.nlevels; ++l) {
    const LevelDev& D = fs.lv[l];
    const PyrSpan dx_ = sSpanX[l], dy_ = sSpanY[l];
    const uint8_t* sbuf = psm + ((l - 1) & 1) * bufBytes;
    uint8_t* dbuf = psm + (l & 1) * bufBytes;
    const int spitch = sx.e1 - sx.e0;                         // source window: columns [sx.e0, sx.e1), rows [sy.e0, sy.e1)
    const int ew4 = (dx_.e1 - dx_.e0) >> 2, eh = dy_.e1 - dy_.e0;
    const int ew2 = 2 * ew4;                                  // column pairs = words of a T row
    const ResizeTap* tX = stap + A.tapOffX[l];               // slices: entry i = column dx_.e0 + i / row dy_.e0 + i
    const ResizeTap* tY = stap + A.tapOffY[l];
    uint8_t* gdst = plane + D.planeOff;
    const bool any = ew4 > 0 && eh > 0;
    const int ry0 = any ? tY[0].s0 : 0, nsr = any ? tY[eh - 1].s1 - ry0 + 1 : 0;      // source rows the window's rows read
    if (any) {
      // Horizontal pass over those source rows (k_resize_level's: a thread owns two adjacent columns, one funnel shift and
      // two IDP.2A per row; T >> 4 kept in 16 bits): each source row is filtered once, not once per output row.
      const float inv = __fdividef(1.0f, (float)ew2);
      int ph, pd;
      pf_divmod(tid, ew2, inv, ph, pd);
      const int RP = PF_THREADS / ew2;                        // row phases (windows are at most ~130 columns wide)
      if (ph < RP) {
        const ResizeTap tA = tX[2 * pd], tB = tX[2 * pd + 1];
        const unsigned cA = (unsigned)(uint16_t)tA.c0 | ((unsigned)(uint16_t)tA.c1 << 16);
        const unsigned cB = (unsigned)(uint16_t)tB.c0 | ((unsigned)(uint16_t)tB.c1 << 16);
        const int a0 = tA.s0 - sx.e0, dB = (int)tB.s0 - (int)tA.s0;
        const int shA = 8 * (a0 & 3), shB = 8 * min(max(dB, 0), 2);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(sbuf + (ry0 - sy.e0 + ph) * spitch) + (a0 >> 2);
        uint32_t* o = sT + ph * ew2 + pd;
        const int pstep = RP * (spitch >> 2), ostep = RP * ew2;
        if (dB <= 2) {
          for (int r = ph; r < nsr; r += RP, p += pstep, o += ostep) {
            const unsigned w = __funnelshift_r(p[0], p[1], shA);
            const unsigned TA = __dp2a_lo(cA, w, 0u), TB = __dp2a_lo(cB, w >> shB, 0u);
            *o = (TA >> 4) | ((TB >> 4) << 16);
          }
        } else {      // scale factors above 2: the second column has its own window
          const int b0 = tB.s0 - sx.e0, wB = (b0 >> 2) - (a0 >> 2), shB2 = 8 * (b0 & 3);
          for (int r = ph; r < nsr; r += RP, p += pstep, o += ostep) {
            const unsigned TA = __dp2a_lo(cA, __funnelshift_r(p[0], p[1], shA), 0u);
            const unsigned TB = __dp2a_lo(cB, __funnelshift_r(p[wB], p[wB + 1], shB2), 0u);
            *o = (TA >> 4) | ((TB >> 4) << 16);
          }
        }
      }
    }
    // Vertical pass: four output pixels (one word) per item into the next window and, for the owned part, global memory
    int wy = 0, wx = 0, sdy = 0, sdx = 0;
    if (any) {
      const float inv = __fdividef(1.0f, (float)ew4);
      pf_divmod(tid, ew4, inv, wy, wx);
      pf_divmod(PF_THREADS, ew4, inv, sdy, sdx);
    }
    __syncthreads();
    if (any) {
      while (wy < eh) {
        const ResizeTap t = tY[wy];
        const uint2 P = *reinterpret_cast<const uint2*>(sT + (t.s0 - ry0) * ew2 + 2 * wx);
        const uint2 Q = *reinterpret_cast<const uint2*>(sT + (t.s1 - ry0) * ew2 + 2 * wx);
        const unsigned b0 = (unsigned)(int)t.c0, b1 = (unsigned)(int)t.c1;
        const unsigned v0 = (__umulhi(b0, P.x << 16) + __umulhi(b1, Q.x << 16) + 2u) >> 2;
        const unsigned v1 = (__umulhi(b0, P.x & 0xFFFF0000u) + __umulhi(b1, Q.x & 0xFFFF0000u) + 2u) >> 2;
        const unsigned v2 = (__umulhi(b0, P.y << 16) + __umulhi(b1, Q.y << 16) + 2u) >> 2;
        const unsigned v3 = (__umulhi(b0, P.y & 0xFFFF0000u) + __umulhi(b1, Q.y & 0xFFFF0000u) + 2u) >> 2;
        const uint32_t word = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
        const int y = dy_.e0 + wy, x4 = dx_.e0 + 4 * wx;
        reinterpret_cast<uint32_t*>(dbuf)[wy * ew4 + wx] = word;
        if (y >= dy_.o0 && y < dy_.o1 && x4 >= dx_.o0 && x4 < dx_.o1 && x4 < D.w)
          *reinterpret_cast<uint32_t*>(gdst + (size_t)y * D.pitch + x4) = word;
        wx += sdx; wy += sdy;
        if (wx >= ew4) { wx -= ew4; ++wy; }
      }
    }
    sx = dx_; sy = dy_;
    // the next level's horizontal pass reads dbuf (written above) and overwrites sT (read above)
    __syncthreads();
    }
}

}  // namespace ivg

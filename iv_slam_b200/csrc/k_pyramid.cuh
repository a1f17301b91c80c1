// k_pyramid.cuh — K1: one pyramid level from the previous one.
// Replaces cv::resize(INTER_LINEAR) inside ORBextractor::ComputePyramid / ComputeQualityImagePyramid
// (introspective_ORB_SLAM/src/ORBextractor.cc:1298-1357, resize calls at :1311 and :1341).
// Arithmetic = OpenCV's 8-bit fixed-point bilinear path: taps and Q11 coefficients are
// precomputed on the host exactly as OpenCV derives them; the kernel does the integer part, separably:
//   stage       the source rectangle of a 128x64 output tile lands in shared memory through ONE TMA box load
//               (cp.async.bulk.tensor.3d over x, y, frame of the source level, signalled on an mbarrier; every source
//               byte is read from L2/HBM once per tile).  Scale factors whose source box would exceed the 256-element
//               TMA box limit use a plain 32-bit load loop instead;
//   horizontal  T[sy][d] = src[sy][sx0]*cx0 + src[sy][sx1]*cx1 for every staged source row, kept as (T >> 4) in 16 bits
//               (the only form the vertical pass uses) — each source row is filtered once, not once per output row; two
//               adjacent columns per thread from one 4-byte window, one IDP.2A each;
//   vertical    dst = (((cy0*T0) >> 16) + ((cy1*T1) >> 16) + 2) >> 2 with (c*T) >> 16 as the high word of c * (T << 16)
//               (IMAD.HI), 4 pixels per thread, one aligned 32-bit store.
// HBM-bound stencil (read level l-1, write level l); the cascade is strictly sequential across levels.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace ivg {

#ifndef IVG_RZ_H
#define IVG_RZ_H 64
#endif
constexpr int RZ_W = 128, RZ_H = IVG_RZ_H;
constexpr int RZ_RPS = RZ_H / 8;           // output rows per 32-thread row segment

// Level-0 ingest: frames arrive from the host as ONE contiguous copy (row-pitched DMA of 1241-byte rows runs at a third
// of the PCIe rate); this kernel lays them out in the row-pitched level-0 plane.  One aligned 32-bit store per thread,
// source bytes come from two aligned words and a funnel shift (rows of an unpitched frame are not word aligned).
__global__ void __launch_bounds__(256) k_ingest(const uint8_t* __restrict__ stage, uint8_t* __restrict__ plane, size_t planeBytes,
                                                int w, int h, int pitch) {
  const int x16 = (blockIdx.x * 256 + threadIdx.x) * 16;     // 16 destination bytes per thread (one 128-bit store)
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x16 >= w) return;
  const size_t a = f * (size_t)w * h + (size_t)y * w + x16;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(stage) + (a >> 2);
  const int sh = 8 * (int)(a & 3);
  const size_t last = (f * (size_t)w * h + (size_t)h * w - 1) >> 2;     // word holding this frame's last byte: nothing past it is read
  uint32_t v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = ((a >> 2) + i <= last && (i < 4 || sh)) ? __ldg(s + i) : 0u;
  uint4 o;
  o.x = __funnelshift_r(v[0], v[1], sh); o.y = __funnelshift_r(v[1], v[2], sh);
  o.z = __funnelshift_r(v[2], v[3], sh); o.w = __funnelshift_r(v[3], v[4], sh);
  *reinterpret_cast<uint4*>(plane + f * planeBytes + (size_t)y * pitch + x16) = o;      // bytes past w land in the row padding
}

// N3 hand-off, float form: the introspection CNN's output as it leaves the network (float in [0, 1], contiguous H x W per
// frame) becomes the u8 cost-map plane on the device — what Examples/Stereo/stereo_kitti.cc:513-514 does with
// `(cost_img * 255.0).to(torch::kByte)`: a float multiply, then torch's float -> uint8 cast (through int64: truncation toward
// zero, modulo 256).  Four pixels per thread, one aligned 32-bit store.
__global__ void __launch_bounds__(256) k_cost_from_f32(const float* __restrict__ src, size_t frameFloats, size_t strideFloats,
                                                       uint8_t* __restrict__ plane, size_t planeBytes, int w, int h, int pitch) {
  const int x4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const size_t f = blockIdx.z;
  if (x4 >= w) return;
  const float* s = src + f * frameFloats + (size_t)y * strideFloats + x4;
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (x4 + k < w) o |= (uint32_t)(uint8_t)(long long)__fmul_rn(__ldg(s + k), 255.0f) << (8 * k);
  *reinterpret_cast<uint32_t*>(plane + f * planeBytes + (size_t)y * pitch + x4) = o;       // bytes past w land in the row padding
}

// blockIdx.z < nImages: image planes; blockIdx.z >= nImages (weighted batches only): the cost-map planes of the same
// level — both pyramids of a level go out in one launch.
__global__ void __launch_bounds__(256) k_resize_level(FrameSet fs, int level, const __grid_constant__ TmaMaps maps,
                                                      const __grid_constant__ TmaMaps mapsQ, int useTma) {
  extern __shared__ __align__(128) unsigned char rsm[];
  __shared__ __align__(8) uint64_t bar;
  const LevelDev& D = fs.lv[level];
  const LevelDev& S = fs.lv[level - 1];
  const int x0 = blockIdx.x * RZ_W, y0 = blockIdx.y * RZ_H;
  const int which = (int)blockIdx.z >= fs.nImages ? 1 : 0;           // 0 image, 1 cost-map
  const int img = (int)blockIdx.z - which * fs.nImages;
  uint8_t* plane = (which ? fs.qual : fs.pyr) + (size_t)img * fs.planeBytes;
  const uint8_t* src = plane + S.planeOff;
  uint8_t* dst = plane + D.planeOff;
  const ResizeTap* tx = fs.rtab + D.rtabX;
  const ResizeTap* ty = fs.rtab + D.rtabY;
  const int tid = threadIdx.x;
  const int x1 = min(x0 + RZ_W, D.w) - 1, y1 = min(y0 + RZ_H, D.h) - 1;
  const int sxa = tx[x0].s0 & ~15, sxe = tx[x1].s1;          // staged source columns [sxa, sxe] (16-byte aligned start for TMA)
  const int sya = ty[y0].s0, sye = ty[y1].s1;                // staged source rows [sya, sye]
  const int nW = (sxe - sxa) / 4 + 1, nR = sye - sya + 1;
  const int SPB = D.rzPitch;                                 // staged bytes per source row (host-computed bound, multiple of 4)
  uint8_t* spx = rsm;
  uint16_t* sT = reinterpret_cast<uint16_t*>(rsm + (((size_t)SPB * D.rzRows + 15) & ~(size_t)15));

  // the tile's vertical taps go to shared memory while the source box is in flight: (row offsets into sT, weights)
  __shared__ int4 sTapY[RZ_H];
  if (useTma) {
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar, (uint32_t)(SPB * D.rzRows));
      tma_load_3d(spx, which ? &mapsQ.m[level] : &maps.m[level], &bar, sxa, sya, img);
    }
    if (tid < RZ_H && y0 + tid <= y1) { const ResizeTap t = ty[y0 + tid]; sTapY[tid] = make_int4((t.s0 - sya) * RZ_W, (t.s1 - sya) * RZ_W, t.c0, t.c1); }
    mbar_wait(&bar, 0);
  } else {
    if (tid < RZ_H && y0 + tid <= y1) { const ResizeTap t = ty[y0 + tid]; sTapY[tid] = make_int4((t.s0 - sya) * RZ_W, (t.s1 - sya) * RZ_W, t.c0, t.c1); }
    for (int i = tid; i < nR * nW; i += 256) {
      const int r = i / nW, g = i - r * nW;
      const int gx = sxa + 4 * g;
      reinterpret_cast<uint32_t*>(spx + r * SPB)[g] = gx < S.pitch ? __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(sya + r) * S.pitch + gx)) : 0u;
    }
    __syncthreads();
  }
  {
    // Horizontal pass: a thread owns TWO adjacent output columns (64 column pairs x 4 row phases).  With a scale above 1 their
    // four source pixels lie inside the four bytes that start at the first column's left tap, so a row costs two aligned word
    // loads, one funnel shift to that window, one more shift to the second column's taps and two IDP.2A (16-bit coefficients
    // x 8-bit pixels); the two 16-bit results (T >> 4) are stored as one word.
    const int pd = tid & 63, ph = tid >> 6;
    const int xA = min(x0 + 2 * pd, D.w - 1), xB = min(x0 + 2 * pd + 1, D.w - 1);        // columns past the width only feed row padding
    const ResizeTap tA = tx[xA], tB = tx[xB];
    const unsigned cA = (unsigned)(uint16_t)tA.c0 | ((unsigned)(uint16_t)tA.c1 << 16);
    const unsigned cB = (unsigned)(uint16_t)tB.c0 | ((unsigned)(uint16_t)tB.c1 << 16);
    const int a0 = tA.s0 - sxa;                                   // byte offset of the window in a staged row
    const int dB = (int)tB.s0 - (int)tA.s0;                        // 1 or 2 for scales up to 2 (0: both columns clamped at the right edge)
    const int shA = 8 * (a0 & 3), shB = 8 * min(max(dB, 0), 2);    // the right tap of a clamped column has weight 0
    if (x0 + 2 * pd <= x1) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(spx + ph * SPB) + (a0 >> 2);
      uint32_t* o = reinterpret_cast<uint32_t*>(sT) + ph * (RZ_W / 2) + pd;
      if (dB <= 2) {
        for (int r = ph; r < nR; r += 4, p += SPB, o += 2 * RZ_W) {      // SPB words = 4 rows of SPB bytes
          const unsigned w = __funnelshift_r(p[0], p[1], shA);
          const unsigned TA = __dp2a_lo(cA, w, 0u), TB = __dp2a_lo(cB, w >> shB, 0u);
          *o = (TA >> 4) | ((TB >> 4) << 16);
        }
      } else {      // scale factors above 2: the second column has its own window
        const int b0 = tB.s0 - sxa, wB = (b0 >> 2) - (a0 >> 2), shB2 = 8 * (b0 & 3);
        for (int r = ph; r < nR; r += 4, p += SPB, o += 2 * RZ_W) {
          const unsigned TA = __dp2a_lo(cA, __funnelshift_r(p[0], p[1], shA), 0u);
          const unsigned TB = __dp2a_lo(cB, __funnelshift_r(p[wB], p[wB + 1], shB2), 0u);
          *o = (TA >> 4) | ((TB >> 4) << 16);
        }
      }
    }
  }
  __syncthreads();
  {
    const int g = tid & 31, seg = tid >> 5;                  // 32 groups of 4 columns x 8 row segments of RZ_RPS rows
    const int gx = x0 + 4 * g;
    if (gx < D.pitch && gx <= x1) {
#pragma unroll
      for (int rr = 0; rr < RZ_RPS; ++rr) {
        const int y = y0 + seg * RZ_RPS + rr;
        if (y > y1) break;
        const int4 t = sTapY[seg * RZ_RPS + rr];
        const uint2 A = *reinterpret_cast<const uint2*>(sT + t.x + 4 * g);
        const uint2 B = *reinterpret_cast<const uint2*>(sT + t.y + 4 * g);
        const unsigned b0 = (unsigned)t.z, b1 = (unsigned)t.w;
        // (b * T) >> 16 as the high word of b * (T << 16): one multiply per tap, no shift
        const unsigned v0 = (__umulhi(b0, A.x << 16) + __umulhi(b1, B.x << 16) + 2u) >> 2;
        const unsigned v1 = (__umulhi(b0, A.x & 0xFFFF0000u) + __umulhi(b1, B.x & 0xFFFF0000u) + 2u) >> 2;
        const unsigned v2 = (__umulhi(b0, A.y << 16) + __umulhi(b1, B.y << 16) + 2u) >> 2;
        const unsigned v3 = (__umulhi(b0, A.y & 0xFFFF0000u) + __umulhi(b1, B.y & 0xFFFF0000u) + 2u) >> 2;
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * D.pitch + gx) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
      }
    }
  }
}

}  // namespace ivg

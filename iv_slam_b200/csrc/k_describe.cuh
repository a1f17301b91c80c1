// k_describe.cuh — K4+K6 fused: intensity-centroid orientation and 256-bit rotated-BRIEF descriptor.
//
// Replaces computeOrientation/IC_Angle (introspective_ORB_SLAM/src/ORBextractor.cc:478-485, :78-105) and
// computeDescriptors/computeOrbDescriptor (:1215-1222, :108-148), plus the final bookkeeping of operator()
// (:1263-1295): level-major concatenation, pt *= scale for levels > 0, size/octave/class_id fields.
//
// One CTA = DK_SLOTS keypoint slots (64 for batches, 16 when a launch would otherwise leave most SMs idle: one frame at a
// time), one warp per keypoint (DK_SLOTS/8 keypoints per warp), three phases:
//   moments  IC_Angle reads the UNBLURRED level: lanes = 3 rows x 9 aligned words, 11 coalesced load instructions per
//            keypoint, reduced with DP4A against the in-circle byte masks of the rows (the umax table); integer moments
//            are exact and order-free;
//   angle    ONE lane per keypoint: cv::fastAtan2's float polynomial without FMA and cos/sin of the angle
//            evaluated in double and rounded to float — the CTA's keypoints share one pass through the double-precision code
//            instead of every warp repeating it (the reference calls glibc cosf/sinf; measured disagreement of the two
//            is ~1 descriptor bit in 5e7, SURVEY Q9 — the only source of non-identical descriptor bits);
//   brief    the descriptor reads the BLURRED level: each lane owns one descriptor byte = 8 pattern pairs = 16 rotated
//            samples; rotation uses float mul/add without contraction and cvRound = round-half-even (A.5).
// The 37x37 footprint of a keypoint in the blurred level is staged per warp in shared memory by ONE TMA box load
// (64 x 37 bytes, cp.async.bulk.tensor.3d, 16-byte aligned start, signalled on a per-warp mbarrier; two buffers per
// warp so the next keypoint's patch is in flight while the current one is sampled): a direct gather from global
// touched up to 32 sectors per load instruction and made this kernel L1-bound (ncu: l1tex 87 %).  The pattern table
// is staged as float2, transposed so that lane-strided reads are conflict-free.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace ivg {

__device__ float2 g_patternT[512];       // [k*32 + byte] = pattern point (byte*16 + k) as floats, filled by the host

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float sc = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
  const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
  const float ax = fabsf(x), ay = fabsf(y);
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, 2.2204460492503131e-16f));
    const float c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, 2.2204460492503131e-16f));
    const float c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

#ifndef IVG_DK_SLOTS_BATCH
#define IVG_DK_SLOTS_BATCH 64
#endif
constexpr int DK_SLOTS_BATCH = IVG_DK_SLOTS_BATCH;       // keypoint slots per CTA, throughput configuration
#ifndef IVG_DK_SLOTS_LAT
#define IVG_DK_SLOTS_LAT 16
#endif
constexpr int DK_SLOTS_LAT = IVG_DK_SLOTS_LAT;         // latency configuration (small batches): 4x the CTAs, 2 keypoints per warp
constexpr int DK_PR = 18;                // |rotated pattern offset| <= 18 (pattern radius 13*sqrt(2) rounds to 18)
constexpr int DK_BOXW = 64, DK_BOXH = 2 * DK_PR + 1;   // TMA box: 64 x 37 bytes
constexpr int DK_BOXN = 48;                             // narrow box for keypoints whose 37 columns start within 11 bytes of a 16-byte boundary

template <int DK_SLOTS>
__global__ void __launch_bounds__(256) k_orient_describe(FrameSet fs, const __grid_constant__ TmaMaps maps, const __grid_constant__ TmaMaps mapsN) {
  __shared__ __align__(128) uint8_t spatch[8][2][DK_BOXW * DK_BOXH + 64];     // +64 keeps every buffer 128-byte aligned
  __shared__ __align__(8) uint64_t bars[8][2];
  __shared__ float2 spat[512];
  __shared__ uint32_t sOnes[33 * 8];        // in-circle byte masks of the 31 patch rows (+2 empty rows), 8 words of 4 columns each
  __shared__ int sCx[DK_SLOTS], sCy[DK_SLOTS], sLevel[DK_SLOTS], sOut[DK_SLOTS];
  __shared__ float sM01[DK_SLOTS], sM10[DK_SLOTS], sA[DK_SLOTS], sB[DK_SLOTS], sResp[DK_SLOTS];
  __shared__ float sRec[DK_SLOTS][7];       // the keypoint's cv::KeyPoint record, assembled by the lane that computes its angle

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t img = blockIdx.y;
  const int* lc = fs.levelCount + img * MAX_LEVELS;
  for (int i = tid; i < 512; i += 256) spat[i] = g_patternT[i];
  for (int i = tid; i < 33 * 8; i += 256) {
    // umax[|v|] of a radius-15 circular patch (src/ORBextractor.cc:458-475) = {15,15,15,15,14,14,14,13,13,12,11,10,9,8,6,3}, packed
    // in nibbles; the host checks its own table against it
    const int v = i / 8 - HALF_PATCH, k = i & 7;
    const int d = abs(v) <= HALF_PATCH ? (int)((0x3689ABCDDEEEFFFFull >> (4 * abs(v))) & 15ull) : -1;
    uint32_t o = 0;
    for (int bb = 0; bb < 4; ++bb)
      if (abs(4 * k + bb - HALF_PATCH) <= d) o |= 1u << (8 * bb);
    sOnes[i] = o;
  }
  if (tid < 16) mbar_init(&bars[tid >> 1][tid & 1], 1);

  if (tid < DK_SLOTS) {
    const int slot = blockIdx.x * DK_SLOTS + tid;
    int level = -1, cx = 0, cy = 0, outIdx = 0;
    float resp = 0.f;
    if (slot < fs.kpCap) {
      int l = 0;
      for (int k = 1; k < fs.nlevels; ++k)
        if (slot >= fs.lv[k].kpOff) l = k;
      const int i = slot - fs.lv[l].kpOff;
      if (i < lc[l]) {
        int base = 0;
        for (int k = 0; k < l; ++k) base += lc[k];
        const uint2 rec = fs.levelKp[img * fs.kpCap + slot];
        level = l; cx = unpack_x(rec.y); cy = unpack_y(rec.y); outIdx = base + i; resp = __uint_as_float(rec.x);
      }
    }
    sLevel[tid] = level; sCx[tid] = cx; sCy[tid] = cy; sOut[tid] = outIdx; sResp[tid] = resp;
    if (blockIdx.x == 0 && tid == 0) {
      int n = 0;
      for (int l = 0; l < fs.nlevels; ++l) n += lc[l];
      fs.outN[img] = n;
    }
  }
  __syncthreads();

  // patch loads for this warp's first two keypoints go out now and land while the moments are computed
  auto issue_patch = [&](int q) {
    const int j = warp * (DK_SLOTS / 8) + q;
    const int level = sLevel[j];
    if (level < 0 || lane != 0) return;
    // three quarters of the keypoints fit a 48-byte-wide box: its 48-byte row pitch spreads the rotated samples over all 32
    // banks (a 64-byte pitch only ever uses two bank offsets per column group) and moves a quarter fewer bytes
    const int xa = (sCx[j] - DK_PR) & ~15;
    const bool narrow = sCx[j] - DK_PR - xa + 2 * DK_PR + 1 <= DK_BOXN;
    mbar_expect_tx(&bars[warp][q & 1], (narrow ? DK_BOXN : DK_BOXW) * DK_BOXH);
    tma_load_3d(spatch[warp][q & 1], narrow ? &mapsN.m[level] : &maps.m[level], &bars[warp][q & 1], xa, sCy[j] - DK_PR, (int)img);
  };
  issue_patch(0);
  if (DK_SLOTS / 8 > 1) issue_patch(1);

  // ---- moments: lane = (row r of 3, word k of 9): every load instruction fetches three 36-byte row segments (the rows'
  // 31 patch bytes as aligned words), 11 instructions cover the 31 rows.  A word is funnel-shifted with its right
  // neighbour (one shuffle) to a fixed alignment (byte 0 of word 0 = column u = -15) and reduced with DP4A against the
  // row's in-circle byte mask (sOnes: the umax table as masks):
  //   s = sum of the in-circle pixels of the word,  t = sum of (byte index) * pixel
  //   m10 = sum over words of t + (4k - 15) * s,    m01 = sum over words of v * s.     Integers: exact, order free.
  // Keypoints are >= 19 px from the border, so rows y-15..y+17 and the 9 words stay inside the level's pitched plane.
  const int mr = lane / 9, mk = lane - 9 * mr;
  const bool mact = lane < 27;
  uint32_t mOnes[11];                       // this lane's in-circle masks of its 11 (row, word) positions: constant for all keypoints
#pragma unroll
  for (int it = 0; it < 11; ++it) mOnes[it] = (mact && mk < 8) ? sOnes[(3 * it + mr) * 8 + mk] : 0u;
#pragma unroll 1
  for (int q = 0; q < DK_SLOTS / 8; ++q) {
    const int j = warp * (DK_SLOTS / 8) + q;
    const int level = sLevel[j];
    if (level < 0) continue;
    const LevelDev& L = fs.lv[level];
    const int xs = sCx[j] - HALF_PATCH;                     // first patch column
    const int pitch = L.pitch;
    const uint8_t* p = fs.pyr + img * fs.planeBytes + L.planeOff + (size_t)(sCy[j] - HALF_PATCH + (mact ? mr : 0)) * pitch + (xs & ~3) + 4 * (mact ? mk : 0);
    const int sh = 8 * (xs & 3);
    uint32_t w[11];
#pragma unroll
    for (int it = 0; it < 11; ++it) w[it] = __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)(3 * it) * pitch));
    int S = 0, T = 0, m01 = 0;
#pragma unroll
    for (int it = 0; it < 11; ++it) {
      const uint32_t wn = __shfl_down_sync(0xffffffffu, w[it], 1);
      const uint32_t x = __funnelshift_r(w[it], wn, sh);
      const uint32_t o = mOnes[it];
      const int s = (int)__dp4a(x, o, 0u);
      T = (int)__dp4a(x, (o * 0xFFu) & 0x03020100u, (unsigned)T);
      S += s;
      m01 += (3 * it + mr - HALF_PATCH) * s;
    }
    int m10 = T + (4 * mk - HALF_PATCH) * S;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    if (lane == 0) { sM01[j] = (float)m01; sM10[j] = (float)m10; }
  }
  __syncthreads();

  // ---- angle, cos, sin: one lane per keypoint
  if (tid < DK_SLOTS && sLevel[tid] >= 0) {
    const float angle = fast_atan2_deg(sM01[tid], sM10[tid]);
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    double sn, cs;
    sincos((double)rad, &sn, &cs);
    sA[tid] = (float)cs; sB[tid] = (float)sn;
    // the 28-byte output record (src/ORBextractor.cc:1286-1291: pt *= scale for levels > 0, size, angle, response, octave, class_id)
    const int level = sLevel[tid];
    const LevelDev& L = fs.lv[level];
    sRec[tid][0] = level ? __fmul_rn((float)sCx[tid], L.scale) : (float)sCx[tid];
    sRec[tid][1] = level ? __fmul_rn((float)sCy[tid], L.scale) : (float)sCy[tid];
    sRec[tid][2] = L.sizeField;
    sRec[tid][3] = angle;
    sRec[tid][4] = sResp[tid];
    sRec[tid][5] = __int_as_float(level);
    sRec[tid][6] = __int_as_float(-1);
  }
  __syncthreads();

  // ---- rotated BRIEF + output record
  // A lane's 16 pattern points are the same for every keypoint: they stay in registers for the whole phase, packed as two bf16
  // per word (the coordinates are integers of magnitude <= 13: exact in bf16, and bf16 -> f32 is a shift / a mask), instead of
  // 16 shared-memory loads of 8 bytes per keypoint (a third of the kernel's shared-memory wavefronts)
  uint32_t pp[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float2 p = spat[k * 32 + lane];
    pp[k] = (__float_as_uint(p.x) >> 16) | (__float_as_uint(p.y) & 0xFFFF0000u);
  }
  int useCount[2] = {0, 0};
#pragma unroll
  for (int q = 0; q < DK_SLOTS / 8; ++q) {
    const int j = warp * (DK_SLOTS / 8) + q;
    const int level = sLevel[j];
    if (level < 0) {                                      // empty slot: its buffer is free for slot q+2 right away
      if (q + 2 < DK_SLOTS / 8) issue_patch(q + 2);
      continue;
    }
    const int cx = sCx[j], cy = sCy[j];
    mbar_wait(&bars[warp][q & 1], (useCount[q & 1]++) & 1);   // parity = loads already consumed from this buffer
    const int cOff = cx - DK_PR - ((cx - DK_PR) & ~15);
    const int pw = cOff + 2 * DK_PR + 1 <= DK_BOXN ? DK_BOXN : DK_BOXW;       // row pitch of this keypoint's patch
    const uint8_t* bctr = spatch[warp][q & 1] + DK_PR * pw + cOff + DK_PR;
    const float a = sA[j], b = sB[j];
    unsigned val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 p0 = make_float2(__uint_as_float(pp[2 * k] << 16), __uint_as_float(pp[2 * k] & 0xFFFF0000u));
      const float2 p1 = make_float2(__uint_as_float(pp[2 * k + 1] << 16), __uint_as_float(pp[2 * k + 1] & 0xFFFF0000u));
      const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a)));
      const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b)));
      const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p1.y, a)));
      const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(p1.x, a), __fmul_rn(p1.y, b)));
      const int t0 = bctr[iy0 * pw + ix0], t1 = bctr[iy1 * pw + ix1];
      val |= (t0 < t1 ? 1u : 0u) << k;
    }
    __syncwarp();                                         // every lane is done with this buffer
    if (q + 2 < DK_SLOTS / 8) issue_patch(q + 2);
    const int outIdx = sOut[j];
    fs.outDesc[(img * fs.kpCap + outIdx) * 32 + lane] = (uint8_t)val;
    if (lane < 7) reinterpret_cast<float*>(fs.outKp + (img * fs.kpCap + outIdx) * 28)[lane] = sRec[j][lane];
  }
}

}  // namespace ivg

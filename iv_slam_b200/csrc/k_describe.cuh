// k_describe.cuh — K4+K6 fused: intensity-centroid orientation and 256-bit rotated-BRIEF descriptor,
// one warp per keypoint.
//
// Replaces computeOrientation/IC_Angle (introspective_ORB_SLAM/src/ORBextractor.cc:478-485, :78-105) and
// computeDescriptors/computeOrbDescriptor (:1215-1222, :108-148), plus the final bookkeeping of operator()
// (:1263-1295): level-major concatenation, pt *= scale for levels > 0, size/octave/class_id fields.
//   * IC_Angle reads the UNBLURRED level: 31 lanes each own one column u of the circular patch (rows bounded by umax),
//     integer moments are exact and order-free, the angle is cv::fastAtan2's float polynomial without FMA;
//   * the descriptor reads the BLURRED level: each lane owns one descriptor byte = 8 pattern pairs = 16 rotated
//     samples; rotation uses float mul/add without contraction and cvRound = round-half-even (A.5); cos/sin are
//     evaluated in double and rounded to float (the reference calls glibc cosf/sinf; measured disagreement of the
//     two is ~1 descriptor bit in 5e7, SURVEY Q9 — this is the only source of non-identical descriptor bits).
// The 37x37 footprint of one keypoint is read through L1 (read-only path); pattern table is staged in shared
// memory transposed so that lane-strided reads are conflict-free.
#pragma once
#include "common.cuh"

namespace ivg {

__constant__ int8_t c_pattern[1024];     // 256 x (x0,y0,x1,y1)
__constant__ int c_umax[16];

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

struct KpRecord { float x, y, size, angle, response; int octave, class_id; };

__global__ void __launch_bounds__(256) k_orient_describe(FrameSet fs) {
  __shared__ int16_t spat[512];          // [k][lane]: sample k (0..15) of descriptor byte `lane`, x | y<<8
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 512; i += 256) {
    const int byte = i >> 4, k = i & 15;   // pattern point index i = byte*16 + k
    spat[k * 32 + byte] = (int16_t)((uint8_t)c_pattern[2 * i] | ((int)c_pattern[2 * i + 1] << 8));
  }
  __syncthreads();

  const size_t img = blockIdx.y;
  const int slot = blockIdx.x * 8 + warp;
  if (slot >= fs.kpCap) return;
  const int* lc = fs.levelCount + img * MAX_LEVELS;
  int level = 0, outBase = 0;
#pragma unroll 1
  for (int l = 1; l < fs.nlevels; ++l)
    if (slot >= fs.lv[l].kpOff) level = l;
  for (int l = 0; l < level; ++l) outBase += lc[l];
  const LevelDev& L = fs.lv[level];
  const int i = slot - L.kpOff;
  if (slot == 0 && lane == 0) {
    int n = 0;
    for (int l = 0; l < fs.nlevels; ++l) n += lc[l];
    fs.outN[img] = n;
  }
  if (i >= lc[level]) return;
  const uint2 rec = fs.levelKp[img * fs.kpCap + slot];
  const int cx = unpack_x(rec.y), cy = unpack_y(rec.y);
  const size_t frameOff = img * fs.planeBytes + L.planeOff;

  // ---- IC_Angle on the unblurred level
  const uint8_t* ctr = fs.pyr + frameOff + (size_t)cy * L.pitch + cx;
  int m10 = 0, m01 = 0;
  {
    const int u = lane - HALF_PATCH;
    if (lane <= 2 * HALF_PATCH) {
      const int au = abs(u);
#pragma unroll 1
      for (int v = -HALF_PATCH; v <= HALF_PATCH; ++v) {
        if (au <= c_umax[abs(v)]) {
          const int val = __ldg(ctr + v * L.pitch + u);
          m10 += u * val;
          m01 += v * val;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- rotated BRIEF on the blurred level
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  const float rad = __fmul_rn(angle, factorPI);
  const float a = (float)cos((double)rad), b = (float)sin((double)rad);
  const uint8_t* bctr = fs.blur + frameOff + (size_t)cy * L.pitch + cx;
  unsigned val = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int p0 = spat[(2 * k) * 32 + lane], p1 = spat[(2 * k + 1) * 32 + lane];
    const float x0 = (float)(int8_t)(p0 & 0xFF), y0 = (float)(p0 >> 8);
    const float x1 = (float)(int8_t)(p1 & 0xFF), y1 = (float)(p1 >> 8);
    const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
    const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
    const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
    const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
    const int t0 = __ldg(bctr + iy0 * L.pitch + ix0), t1 = __ldg(bctr + iy1 * L.pitch + ix1);
    val |= (t0 < t1 ? 1u : 0u) << k;
  }
  const int outIdx = outBase + i;
  fs.outDesc[(img * fs.kpCap + outIdx) * 32 + lane] = (uint8_t)val;
  if (lane == 0) {
    KpRecord r;
    r.x = level ? __fmul_rn((float)cx, L.scale) : (float)cx;
    r.y = level ? __fmul_rn((float)cy, L.scale) : (float)cy;
    r.size = L.sizeField;
    r.angle = angle;
    r.response = __uint_as_float(rec.x);
    r.octave = level;
    r.class_id = -1;
    float* o = reinterpret_cast<float*>(fs.outKp + (img * fs.kpCap + outIdx) * 28);
    o[0] = r.x; o[1] = r.y; o[2] = r.size; o[3] = r.angle; o[4] = r.response;
    reinterpret_cast<int*>(o)[5] = r.octave; reinterpret_cast<int*>(o)[6] = r.class_id;
  }
}

}  // namespace ivg

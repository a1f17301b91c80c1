// common.cuh — device-visible layout descriptors shared by all kernels of the stereo front-end.
//
// HBM layout (one "frame set" per extractor handle, B = reserved batch):
//   plane[which][b]   which in {image pyramid, blurred pyramid, cost-map pyramid};
//                     every plane holds all levels of one frame back to back, each level row-pitched
//                     (pitch = width rounded up to 64 B, so rows are 16 B aligned for vector access).
//   cellList[b]       u32 per possible FAST corner of every cell, row-major inside the cell (y<<20 | x<<8 | score)
//   cellCount[b]      int2 per cell: entries in the cell's list, corners at iniTh;  cellCost[b]: u32 cost-map sum of the window
//   levelKp[b]        uint2 per kept keypoint per level (response bits, packed y/x/score) + levelCount[b][level]
//   outKp/outDesc[b]  final cv::KeyPoint-layout records and 32-byte descriptors, reference order; outN[b]
//   uRight/depth/sad  stereo results per left keypoint
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ivg {

constexpr int MAX_LEVELS = 12;
constexpr int EDGE = 19;          // EDGE_THRESHOLD, src/ORBextractor.cc:75
constexpr int HALF_PATCH = 15;    // HALF_PATCH_SIZE, :74
constexpr int PATCH = 31;         // PATCH_SIZE, :73


struct LevelDev {
  int w, h, pitch;
  unsigned planeOff;      // byte offset of this level inside a frame plane
  int maxBX, maxBY;       // w-19, h-19
  int nDesired;           // mnFeaturesPerLevel[level]
  int nfeaturesCell;      // ceil(nDesired / nCells)
  int cols, rows, cellW, cellH, nCells;
  int cellBase;           // first cell index of this level in the cell tables
  int kpOff;              // offset of this level inside levelKp (prefix sum of kpCapLevel)
  int kpCapLevel;         // keypoints this level can emit: nDesired (live path) or nDesired + 3 (OctTree mode)
  int fSP, fSS, fBH, fBW, fSeg, fShift; // k_fast_cells: staged words per row, score bytes per row, rows per band, bitmap words per row, list entries per warp
  int fBX;                // 2 when the level's cells are processed in several bands (one overlap score row per side), else 0
  int btBase, btX, btY;   // blur tile numbering
  int rzPitch, rzRows;    // k_resize_level: staged source bytes per row / rows of one output tile (this level as destination)
  int rtabX, rtabY;       // offsets into the resize tap tables (level >= 1)
  unsigned listBase;      // first cellList slot of this level
  unsigned listCap;       // cellList slots of this level
  float scale, invScale;
  float sizeField;        // (int)(31*scale)
};

struct CellDev {          // one FAST cell of ComputeKeyPointsOld (src/ORBextractor.cc:989-1023)
  int level;
  int x0, y0, cw, ch;     // detect range (window minus its 3-px FAST margin), level coordinates
  int wx, wy, ww, wh;     // cost-map averaging window of the budget pass (:976-978) = the cell's nominal FAST window
  unsigned listOff;       // first cellList slot
  unsigned listCap;
};

struct ResizeTap { uint16_t s0, s1; int16_t c0, c1; };   // two source indices and Q11 coefficients

struct FrameSet {
  int nlevels, nImages, weighted;
  int iniTh, minTh, scoreTh;
  int fastRetry;           // k_fast_cells redoes a cell at minTh when the iniTh pass leaves <= this many corners (3: live path :1047, 0: OctTree :818)
  int nCellsTotal, kpCap;
  int cellCostStride;      // entries per frame in cellCost (nCellsTotal + slack used by the OctTree gather)
  int btTotal;
  int selLevelCap, selCellCap, selCells;   // k_level_select shared-memory capacities
  unsigned listCapTotal;
  size_t planeBytes;
  uint8_t* pyr; uint8_t* blur; uint8_t* qual;                    // [nImages][planeBytes]
  const CellDev* cells;                                           // active variant
  const ResizeTap* rtab;
  const uint32_t* blurTiles;                                      // [btTotal] level<<28 | tileX<<14 | tileY
  uint32_t* cellList;      // [nImages][listCapTotal]
  int2* cellCount;         // [nImages][nCellsTotal]
  uint32_t* cellCost;      // [nImages][nCellsTotal]
  uint2* workCell;         // [nImages][listCapTotal]  global fallback for per-cell selection
  uint2* workLevel;        // [nImages][listCapTotal]  global fallback for per-level selection
  uint2* levelKp;          // [nImages][kpCap]
  int* levelCount;         // [nImages][MAX_LEVELS]
  uint8_t* outKp;          // [nImages][kpCap] x 28 B (ivg_keypoint)
  uint8_t* outDesc;        // [nImages][kpCap] x 32 B
  int* outN;               // [nImages]
  LevelDev lv[MAX_LEVELS];
};

__device__ __forceinline__ uint32_t pack_xys(int x, int y, int s) { return ((uint32_t)y << 20) | ((uint32_t)x << 8) | (uint32_t)s; }
__device__ __forceinline__ int unpack_x(uint32_t p) { return (p >> 8) & 0xFFF; }
__device__ __forceinline__ int unpack_y(uint32_t p) { return p >> 20; }
__device__ __forceinline__ int unpack_s(uint32_t p) { return p & 0xFF; }

}  // namespace ivg

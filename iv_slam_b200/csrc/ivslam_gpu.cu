// ivslam_gpu.cu — host side of the C ABI declared in include/ivslam_gpu.h.
//
// Mirrors ORB_SLAM2::ORBextractor (introspective_ORB_SLAM/include/ORBextractor.h:54-128, src/ORBextractor.cc:411-476
// constructor tables, :1224-1296 operator()) and the orchestration of the stereo Frame constructor
// (src/Frame.cc:115-125 two extractions, :193 ComputeStereoMatches).  All pixel work is done by the sm_100a kernels
// in k_*.cuh; this file builds the per-shape tables (level sizes, cell grids, x/y detect flags, bilinear taps — the
// float/double mix of the reference is kept so the integer tables come out identical), owns the device workspace
// and sequences the launches on the handle's stream.  No CPU fallback exists: without a usable GPU every entry
// point returns an error.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ivslam_gpu.h"
#include "common.cuh"
#include "tma.cuh"
#include "k_blur.cuh"
#include "k_describe.cuh"
#include "k_fast.cuh"
#include "k_frame.cuh"
#include "k_pyramid.cuh"
#include "k_pyramid_fused.cuh"
#include "k_select.cuh"
#include "k_octree.cuh"
#include "k_prologue.cuh"
#include "k_project.cuh"
#include "k_bow.cuh"
#include "k_stereo.cuh"

using namespace ivg;

namespace {

#ifndef IVG_FAST_SMEM_KB
#define IVG_FAST_SMEM_KB 31
#endif
constexpr int EAGER_INDEX_MAX_BATCH = 8;     // batches up to this size are treated as the one-frame-at-a-time (latency) use
constexpr size_t FAST_SMEM_BUDGET = IVG_FAST_SMEM_KB * 1024;   // k_fast_cells stages at most this much per CTA (taller cells are banded)

thread_local std::string g_cuda_err;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_cuda_err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
      return IVG_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

const int8_t kPatternHost[1024] = {
#include "../../include/ivslam_brief_pattern.inc"
};

inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n && p) return IVG_OK;
    release();
    if (count == 0) count = 1;
    CK(cudaMalloc(&p, count * sizeof(T)));
    n = count;
    return IVG_OK;
  }
  void release() { if (p && !view) cudaFree(p); p = nullptr; n = 0; view = false; }
  bool view = false;                       // p points into another DevBuf's allocation
  void view_of(void* q, size_t count) { release(); p = static_cast<T*>(q); n = count; view = true; }
};

}  // namespace

// One-frame-at-a-time stereo (the drop-in use: two extractor threads, then ComputeStereoMatches): once two handles have been
// matched, they are linked; from then on the handle whose run is enqueued SECOND enqueues the matcher right behind it with
// the remembered calibration, and its results land in pinned staging while the host is still joining its threads.
// ivg_stereo_match then only has to wait and copy (if the generations and the calibration still agree — otherwise the
// speculative result is dropped and the call runs as usual).
struct StereoLink {
  std::mutex mu;
  ivg_extractor* left = nullptr; ivg_extractor* right = nullptr;
  float mbf = 0.f, maxD = 0.f;
  unsigned long long genL = 0, genR = 0;          // generation of each side's last enqueued run
  unsigned long long specGenL = 0, specGenR = 0;  // generations the last (speculative or explicit) matcher launch consumed
  bool specValid = false; int specN = 0;
};

struct ivg_extractor {
  int device = 0;
  std::shared_ptr<StereoLink> link;     // set by the first single-frame ivg_stereo_match of this handle with a partner
  unsigned long long runGen = 0;        // bumped by every upload and run: identifies what the device buffers currently hold
  void* specHost = nullptr; size_t specHostBytes = 0;   // pinned staging of the matcher's uRight / depth (left handle)
  int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
  double scaleFactor = 1.2;
  bool enableIntrospection = false;
  int kpMode = 0;                       // 0: ComputeKeyPointsOld (live in the reference), 1: ComputeKeyPointsOctTree (dead there)
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> nPerLevel, umax;
  int kpCap = 0;

  // Streams: `stream` runs every kernel (it may be shared with other handles, ivg_share_stream: kernels of different
  // handles then never overlap — co-running two of these kernels costs ~20 % at large batches); copyIn / copyOut carry
  // the H2D / D2H DMA so copies overlap the kernels of other chunks.  Events order the three.
  cudaStream_t stream = nullptr, copyIn = nullptr, copyOut = nullptr;
  bool ownsStream = true;
  cudaEvent_t evDone = nullptr, evT0 = nullptr, evT1 = nullptr;
  cudaEvent_t evH2D = nullptr;          // copyIn: staged frames have landed
  cudaEvent_t evIngest = nullptr;       // stream: staging buffer consumed (next upload may overwrite it)
  cudaEvent_t evKernels = nullptr;      // stream: results of the last run are complete
  cudaEvent_t evD2H = nullptr;          // copyOut: results of the last run have been read (next run may overwrite them)
  cudaEvent_t evStereo = nullptr;       // stream (left handle): matcher finished reading both handles
  cudaEvent_t evD2Hs = nullptr;         // copyOut (left handle): uRight/depth have been read
  bool eagerIndex = false;              // this handle is the right eye of single-frame stereo calls: build the matcher's row index as the tail of its own run
  bool indexValid = false;              // sortedR / rowStart of THIS handle describe its current keypoints
  cudaStream_t aux = nullptr;           // side stream for the blur of small batches (forked from / joined into `stream`)
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  cudaEvent_t evConsumed = nullptr;     // recorded (on the matcher's stream) when another handle's kernels have read our buffers
  cudaEvent_t waitFor = nullptr;        // set to evConsumed (our own event) when it must complete before we overwrite our buffers
  long long launches = 0;

  // shape-dependent state
  int W = 0, H = 0, maxBatch = 0;
  bool shapeReady = false;
  FrameSet fs{};                        // template (pointers filled, nImages/weighted set per run)
  int curBatch = 0;
  bool curWeighted = false;
  bool haveCost = false;                // a cost-map was supplied with the current batch (used by ivg_frame_postprocess even without introspection)
  bool haveResults = false, havePyramid = false;
  std::vector<CellDev> cellsPlain, cellsWeighted;
  DevBuf<uint8_t> pyr, blur, qual, outKp, outDesc, stageImg, stageCost;
  DevBuf<uint8_t> outAll; size_t outDescOff = 0, outNOff = 0;                // owner of outKp | outDesc | outN (views)
  DevBuf<uint8_t> projIn; DevBuf<uint2> projCand; DevBuf<int> projInt;   // N2 scratch
  void* outHost = nullptr; size_t outHostBytes = 0;                         // pinned staging for the results of the synchronous calls
  void* projHost = nullptr; size_t projHostBytes = 0;                       // N2 pinned staging (inputs, then match[] + nmatches)
  bool haveGrid = false, haveStereo = false;
  DevBuf<float> mapX, mapY;                   // N4: rectification maps of ivg_set_rectify_maps
  int mapW = 0, mapH = 0;
  size_t fastSmem = 0, fastSmemLat = 0, resizeSmem = 0, selSmem = 0, selSmemLat = 0;
  int forceCfg = 0;                     // test hook (ivg_debug_force_config): 0 auto, 1 throughput kernels, 2 one-frame kernels
  int fastLat[MAX_LEVELS][3] = {};      // fBH, fBX, fSeg of every level for the FC_THREADS_LAT configuration of k_fast_cells
  TmaMaps blurMaps{};                   // per level: 160 x 38 x 1 boxes over the image-pyramid planes (k_gauss7)
  TmaMaps resizeMaps{}, resizeMapsQ{};  // per destination level l >= 1: source boxes over level l-1 of the image / cost-map planes
  bool resizeTma[MAX_LEVELS] = {false};
  TmaMaps descMapsN{};                  // the same with 48 x 37 boxes
  DevBuf<PyrSpan> pyrSpanX, pyrSpanY;   // k_pyramid_fused window tables (level-major)
  int pyrTX = 0, pyrTY = 0; size_t pyrFusedBuf = 0, pyrFusedTOff = 0, pyrFusedSmem = 0; int pyrTapOffX[MAX_LEVELS] = {0}, pyrTapOffY[MAX_LEVELS] = {0};   // 0 tiles: the fused cascade is not available for this shape
  TmaMaps descMaps{};                   // per level: 64 x 37 x 1 boxes over the blurred planes (k_orient_describe)
  DevBuf<CellDev> dCellsPlain, dCellsWeighted;
  DevBuf<ResizeTap> rtab;
  DevBuf<uint32_t> cellList, cellCost, blurTiles;
  DevBuf<int2> cellCount;
  DevBuf<uint2> workCell, workLevel, levelKp;
  DevBuf<int> levelCount, outN, sad, nExt, rowStart;
  DevBuf<uint4> sortedR;
  DevBuf<float> uRight, depth, kpQual;
  DevBuf<float> udAll;                  // owner of uRight | depth (views)
  DevBuf<int> gridStart, gridIdx;
  // stereo on caller-supplied keypoints
  DevBuf<uint8_t> extKpL, extDescL, extKpR, extDescR;
  DevBuf<float> extU, extD; DevBuf<int> extS;        // results of ivg_stereo_match_keypoints (kept across calls)
  bool graphMode = false;
  bool graphEager = false;
  cudaGraphExec_t graphExec = nullptr;   // captured kernel sequence of one run (re-captured when batch / mode / buffers change)
  int graphBatch = 0, graphLaunches = 0;
  bool graphWeighted = false;
  // per-kernel CUDA-event profile (bench.py roofline): events bracket every launch while enabled
  bool profile = false;
  std::vector<cudaEvent_t> profEv;     // pairs
  std::vector<int> profKid;
  size_t profUsed = 0;
  double profMs[IVG_NUM_KERNELS] = {0};
  long long profCnt[IVG_NUM_KERNELS] = {0};
};

namespace {

struct ProfScope {   // brackets one kernel launch with two events when profiling is on
  ivg_extractor* h; size_t slot; bool on;
  ProfScope(ivg_extractor* h_, int kid) : h(h_), slot(0), on(h_->profile) {
    h->launches++;
    if (!on) return;
    if (h->profUsed + 2 > h->profEv.size()) {
      for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->profEv.push_back(e); }
    }
    slot = h->profUsed; h->profUsed += 2;
    h->profKid.resize(h->profEv.size() / 2);
    h->profKid[slot / 2] = kid;
    cudaEventRecord(h->profEv[slot], h->stream);
  }
  ~ProfScope() { if (on) cudaEventRecord(h->profEv[slot + 1], h->stream); }
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// u8 tensor (x, y, frame) over one pyramid level of a plane: TMA boxes of boxW x boxH x 1, zero fill outside the image
int make_level_map(CUtensorMap* out, uint8_t* base, int w, int h, int pitch, size_t planeBytes, int frames, int boxW, int boxH) {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;      // two extractor threads (the reference's left/right threads) may build shapes concurrently
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (PFN_encodeTiled)p;
  });
  if (!fn) { g_cuda_err = "cuTensorMapEncodeTiled entry point not available"; return IVG_ERR_CUDA; }
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)planeBytes};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_cuda_err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return IVG_ERR_CUDA; }
  return IVG_OK;
}

void drop_graph(ivg_extractor* h) {
  if (h->graphExec) { cudaGraphExecDestroy(h->graphExec); h->graphExec = nullptr; }
}

int build_tables(ivg_extractor* h) {
  // src/ORBextractor.cc:417-432 (float tables; the scaleFactor member is double, include/ORBextractor.h:108)
  const int nl = h->nlevels;
  h->scale.resize(nl); h->sigma2.resize(nl); h->invScale.resize(nl); h->invSigma2.resize(nl);
  h->scale[0] = 1.f; h->sigma2[0] = 1.f;
  for (int i = 1; i < nl; ++i) { h->scale[i] = (float)(h->scale[i - 1] * h->scaleFactor); h->sigma2[i] = h->scale[i] * h->scale[i]; }
  for (int i = 0; i < nl; ++i) { h->invScale[i] = 1.f / h->scale[i]; h->invSigma2[i] = 1.f / h->sigma2[i]; }
  // :437-452
  h->nPerLevel.resize(nl);
  float factor = (float)(1.0f / h->scaleFactor);
  float nDesired = h->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) { h->nPerLevel[l] = cv_round(nDesired); sum += h->nPerLevel[l]; nDesired *= factor; }
  h->nPerLevel[nl - 1] = std::max(h->nfeatures - sum, 0);
  h->kpCap = 0;
  for (int l = 0; l < nl; ++l) h->kpCap += h->nPerLevel[l];
  // :458-475
  h->umax.assign(HALF_PATCH + 1, 0);
  int vmax = (int)std::floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1);
  int vmin = (int)std::ceil(HALF_PATCH * std::sqrt(2.f) / 2);
  const double hp2 = HALF_PATCH * HALF_PATCH;
  for (int v = 0; v <= vmax; ++v) h->umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (int v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
    while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
    h->umax[v] = v0; ++v0;
  }
  return IVG_OK;
}

void make_taps(int S, int D, std::vector<ResizeTap>& out) {
  const double inv_scale = (double)D / S;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < D; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= S - 1) { f = 0.f; s = S - 1; }
    ResizeTap t;
    t.s0 = (uint16_t)s; t.s1 = (uint16_t)std::min(s + 1, S - 1);
    t.c0 = (int16_t)cv_round((1.f - f) * 2048.f); t.c1 = (int16_t)cv_round(f * 2048.f);
    out.push_back(t);
  }
}

// Per-shape geometry: level sizes (src/ORBextractor.cc:1302-1303), cell grids (:884-907), detect flags, tap tables.
int build_shape(ivg_extractor* h, int W, int H, int batch) {
  const int nl = h->nlevels;
  if (W > 4095 || H > 4095) return IVG_ERR_INVALID;     // 12-bit packed coordinates
  FrameSet fs{};
  fs.nlevels = nl; fs.iniTh = std::min(std::max(h->iniTh, 0), 255); fs.minTh = std::min(std::max(h->minTh, 0), 255);
  fs.scoreTh = std::min(fs.iniTh, fs.minTh);
  std::vector<ResizeTap> taps;
  h->cellsPlain.clear(); h->cellsWeighted.clear();
  size_t planeOff = 0;
  unsigned listOff = 0;
  int kpOff = 0, btBase = 0;
  size_t fastSmem = 0, fastSmemLat = 0, resizeSmem = 0;
  const float imageRatio = (float)W / H;
  for (int l = 0; l < nl; ++l) {
    LevelDev& L = fs.lv[l];
    const float sc = h->invScale[l];
    L.w = cv_round((float)W * sc); L.h = cv_round((float)H * sc);
    L.pitch = (int)align_up(L.w, 64);
    L.planeOff = (unsigned)planeOff;
    planeOff += align_up((size_t)L.pitch * L.h, 256);
    L.scale = h->scale[l]; L.invScale = h->invScale[l];
    L.sizeField = (float)(int)(PATCH * h->scale[l]);
    L.nDesired = h->nPerLevel[l];
    L.kpCapLevel = L.nDesired + (h->kpMode == 1 ? 3 : 0);
    L.kpOff = kpOff; kpOff += L.kpCapLevel;
    // grid
    L.maxBX = L.w - EDGE; L.maxBY = L.h - EDGE;
    const int Wd = L.maxBX - EDGE, Hd = L.maxBY - EDGE;
    if (h->kpMode == 1) {
      // ComputeKeyPointsOctTree (:771-797): 30-px cells over the area inset by 16; the detect ranges (window minus its
      // 3-px FAST margin) tile [19, dim-19) exactly like the live path, trailing cells may be empty
      const float width = (float)(L.w - 2 * (EDGE - 3)), height = (float)(L.h - 2 * (EDGE - 3));
      L.cols = (int)(width / 30.f); L.rows = (int)(height / 30.f);
      if (L.cols < 1 || L.rows < 1 || Wd < 1 || Hd < 1) return IVG_ERR_GEOMETRY;
      L.cellW = (int)std::ceil(width / L.cols); L.cellH = (int)std::ceil(height / L.rows);
      L.nCells = L.rows * L.cols;
    } else {
      L.cols = (int)std::sqrt((float)L.nDesired / (5 * imageRatio));
      L.rows = (int)(imageRatio * L.cols);
      if (L.cols < 1 || L.rows < 1 || Wd < 1 || Hd < 1) return IVG_ERR_GEOMETRY;
      L.cellW = (int)std::ceil((float)Wd / L.cols);
      L.cellH = (int)std::ceil((float)Hd / L.rows);
      L.nCells = L.rows * L.cols;
      if ((L.cols - 1) * L.cellW > Wd || (L.rows - 1) * L.cellH > Hd) return IVG_ERR_GEOMETRY;
      if (L.nCells > SEL_MAX_CELLS) return IVG_ERR_CAPACITY;
    }
    L.nfeaturesCell = (int)std::ceil((float)L.nDesired / L.nCells);
    L.cellBase = (int)h->cellsPlain.size();
    L.btX = (L.w + BL_W - 1) / BL_W; L.btY = (L.h + BL_H - 1) / BL_H;
    L.btBase = btBase; btBase += L.btX * L.btY;
    // k_fast_cells shared-memory geometry: one 32-bit word per horizontal pixel pair (2 B per pixel) + 1 B per score
    L.fSP = (int)align_up(L.cellW + 9, 4) / 2;
    L.fSS = (int)align_up(L.cellW + 6, 4);
    L.fBW = (L.cellW + 31) / 32;
    {
      const int slots = (((L.cellW + 2) / 2 + 1 + 63) / 64) * 64;                         // pair slots a warp walks per row (64 per step)
      auto bytes = [&](int bh, int* seg, int warps) {
        const int bx = bh < L.cellH ? 2 : 0;                                              // banded cells carry one overlap score row per side
        const size_t ssBytes = align_up((size_t)L.fSS * (bh + 2 + bx), 16), bitBytes = align_up((size_t)4 * L.fBW * (bh + 2), 16);
        *seg = ((bh + bx + warps - 1) / warps) * slots;                                   // per-warp pair list: its rows, every slot
        return align_up((size_t)4 * L.fSP * (bh + 6 + bx), 16) + FC_SLACK + ssBytes + bitBytes + (size_t)2 * warps * *seg;
      };
      L.fShift = 1;
      while ((1 << L.fShift) < slots) ++L.fShift;                                         // list entry = (score row << fShift) | pair slot, 16 bits
      if ((65536 >> L.fShift) < 8) return IVG_ERR_CAPACITY;                               // cells wider than ~16k px
      // band geometry for both CTA sizes (k_fast.cuh): the batch set lives in the FrameSet, the one-frame set is patched in at launch
      for (int lat = 0; lat < 2; ++lat) {
        const int warps = lat ? FC_WARPS_LAT : FC_WARPS;
        int seg = 0;
        int bh = std::min(L.cellH, (65536 >> L.fShift) - 2);
        while (bh > 4 && bytes(bh, &seg, warps) > FAST_SMEM_BUDGET) --bh;                 // taller cells are processed in bands
        const size_t need = bytes(bh, &seg, warps);
        if (lat) { h->fastLat[l][0] = bh; h->fastLat[l][1] = bh < L.cellH ? 2 : 0; h->fastLat[l][2] = seg; fastSmemLat = std::max(fastSmemLat, need); }
        else { L.fBH = bh; L.fBX = bh < L.cellH ? 2 : 0; L.fSeg = seg; fastSmem = std::max(fastSmem, need); }
      }
    }
    const int hYlast = Hd - (L.rows - 1) * L.cellH + 6;       // window height of the last row (:951, :995)
    const int chW = std::max(hYlast - 6, 0);                    // weighted: every row searches this many rows
    // cells, row-major
    L.listBase = listOff;
    for (int i = 0; i < L.rows; ++i)
      for (int j = 0; j < L.cols; ++j) {
        CellDev c{};
        c.level = l;
        c.x0 = EDGE + j * L.cellW; c.cw = (j < L.cols - 1 ? c.x0 + L.cellW : L.maxBX) - c.x0;
        c.y0 = EDGE + i * L.cellH; c.ch = (i < L.rows - 1 ? c.y0 + L.cellH : L.maxBY) - c.y0;
        if (h->kpMode == 1) {   // every cell is clipped at the search area; cells past it are empty (their windows are skipped, :801,:810)
          c.cw = std::max(0, std::min(c.x0 + L.cellW, L.maxBX) - c.x0);
          c.ch = std::max(0, std::min(c.y0 + L.cellH, L.maxBY) - c.y0);
          if (c.cw == 0 || c.ch == 0) { c.cw = 0; c.ch = 0; c.x0 = EDGE; c.y0 = EDGE; }
        }
        c.wx = c.x0 - 3; c.ww = j < L.cols - 1 ? L.cellW + 6 : L.maxBX + 3 - c.wx;
        c.wy = c.y0 - 3; c.wh = i < L.rows - 1 ? L.cellH + 6 : L.maxBY + 3 - c.wy;
        c.listOff = listOff;
        c.listCap = (unsigned)(((c.cw + 1) / 2) * ((c.ch + 1) / 2));
        if (h->kpMode == 1 && c.cw == 0) { c.ww = 1; c.wh = 1; c.wx = EDGE; c.wy = EDGE; }
        listOff += c.listCap;
        h->cellsPlain.push_back(c);
        CellDev cw = c;
        cw.ch = chW;
        h->cellsWeighted.push_back(cw);
      }
    L.listCap = listOff - L.listBase;
    // taps
    if (l > 0) {
      L.rtabX = (int)taps.size(); make_taps(fs.lv[l - 1].w, L.w, taps);
      L.rtabY = (int)taps.size(); make_taps(fs.lv[l - 1].h, L.h, taps);
      int maxW = 1, maxR = 1;
      for (int x0 = 0; x0 < L.w; x0 += RZ_W) {
        const int x1 = std::min(x0 + RZ_W, L.w) - 1;
        const int sxa = taps[L.rtabX + x0].s0 & ~15, sxe = taps[L.rtabX + x1].s1;
        maxW = std::max(maxW, (sxe - sxa) / 4 + 1);
      }
      for (int y0 = 0; y0 < L.h; y0 += RZ_H) {
        const int y1 = std::min(y0 + RZ_H, L.h) - 1;
        maxR = std::max(maxR, taps[L.rtabY + y1].s1 - taps[L.rtabY + y0].s0 + 1);
      }
      L.rzPitch = (int)align_up((size_t)maxW * 4, 16); L.rzRows = maxR;        // = TMA box (bytes x rows)
      resizeSmem = std::max(resizeSmem, align_up((size_t)L.rzPitch * L.rzRows, 128) + (size_t)L.rzRows * RZ_W * 2);
    }
  }
  std::vector<PyrSpan> spanX, spanY;
  {
    // k_pyramid_fused: windows of every level for every tile column / row (see k_pyramid_fused.cuh)
    // level-0 tile of a CTA: 80 x 64 measured best for one KITTI frame (96 CTAs; 100 x 80, 64 x 56, 48 x 48 are 2-3 us slower:
    // fewer CTAs leave SMs idle, smaller tiles recompute more halo and pay the per-level fixed cost on more CTAs)
    const int TX = std::max(1, (W + 79) / 80), TY = std::max(1, (H + 63) / 64);
    spanX.assign((size_t)nl * TX, PyrSpan{0, 0, 0, 0, 0, 0}); spanY.assign((size_t)nl * TY, PyrSpan{0, 0, 0, 0, 0, 0});
    auto build = [&](int T, bool isX, std::vector<PyrSpan>& out) {
      for (int t = 0; t < T; ++t) {
        for (int l = 1; l < nl; ++l) {
          const int dim = isX ? fs.lv[l].w : fs.lv[l].h;
          auto split = [&](int k) { const int v = (int)((long long)k * dim / T); return k == T ? (isX ? (int)align_up(dim, 4) : dim) : (isX ? (v & ~3) : v); };
          out[(size_t)l * T + t].o0 = split(t); out[(size_t)l * T + t].o1 = split(t + 1);
        }
        for (int l = nl - 1; l >= 0; --l) {
          PyrSpan& s = out[(size_t)l * T + t];
          int e0 = s.o0, e1 = s.o1;
          const int dim = isX ? fs.lv[l].w : fs.lv[l].h;
          e1 = std::min(e1, dim);                         // the owned x-range of the last column ends on a word boundary past the width
          if (l + 1 < nl) {
            const PyrSpan& up = out[(size_t)(l + 1) * T + t];
            if (up.t1 > up.t0) {
              const ResizeTap* tp = taps.data() + (isX ? fs.lv[l + 1].rtabX : fs.lv[l + 1].rtabY);
              const int a = tp[up.t0].s0, b = tp[up.t1 - 1].s1 + 1;
              if (e1 > e0) { e0 = std::min(e0, a); e1 = std::max(e1, b); } else { e0 = a; e1 = b; }
            }
          }
          s.t0 = e0; s.t1 = e1;                           // needed range; computed range = whole words around it
          if (isX) { e0 &= ~3; e1 = std::min((int)align_up(e1, 4), fs.lv[l].pitch); }
          s.e0 = e0; s.e1 = e1;
        }
      }
    };
    build(TX, true, spanX);
    build(TY, false, spanY);
    size_t buf = 16, tapEntries = 0;
    for (int l = 0; l < nl; ++l) {
      int mw = 0, mh = 0;
      for (int t = 0; t < TX; ++t) mw = std::max(mw, spanX[(size_t)l * TX + t].e1 - spanX[(size_t)l * TX + t].e0);
      for (int t = 0; t < TY; ++t) mh = std::max(mh, spanY[(size_t)l * TY + t].e1 - spanY[(size_t)l * TY + t].e0);
      buf = std::max(buf, align_up((size_t)mw * mh, 16));
      h->pyrTapOffX[l] = (int)tapEntries; tapEntries += (size_t)mw;
      h->pyrTapOffY[l] = (int)tapEntries; tapEntries += (size_t)mh;
    }
    size_t tWords = 0;                                    // horizontal-pass buffer: (column pairs) x (source rows a window's rows read)
    for (int l = 1; l < nl; ++l) {
      int mw = 0, mr = 0;
      for (int t = 0; t < TX; ++t) mw = std::max(mw, spanX[(size_t)l * TX + t].e1 - spanX[(size_t)l * TX + t].e0);
      for (int t = 0; t < TY; ++t) {
        const PyrSpan& s = spanY[(size_t)l * TY + t];
        if (s.e1 > s.e0) mr = std::max(mr, taps[fs.lv[l].rtabY + std::min(s.e1, fs.lv[l].h) - 1].s1 - taps[fs.lv[l].rtabY + s.e0].s0 + 1);
      }
      tWords = std::max(tWords, (size_t)(mw / 2) * mr);
    }
    h->pyrTX = TX; h->pyrTY = TY; h->pyrFusedBuf = buf;
    h->pyrFusedTOff = align_up(2 * buf + tapEntries * sizeof(ResizeTap), 16);
    h->pyrFusedSmem = h->pyrFusedTOff + tWords * 4;
    if (h->pyrFusedSmem > 160 * 1024 || nl < 2 || tWords == 0) { h->pyrTX = 0; h->pyrTY = 0; }
  }
  fs.planeBytes = align_up(planeOff, (size_t)fs.lv[0].pitch * 4);   // multiple of the level-0 pitch: batched 3-D copies
  while (fs.planeBytes % fs.lv[0].pitch) fs.planeBytes += 256;
  fs.listCapTotal = listOff;
  if (h->kpMode == 1)
    for (int l = 0; l < nl; ++l) {   // k_octree_select: oversized budgets keep their node tables in the level's work area
      const LevelDev& L = fs.lv[l];
      if (L.nDesired + 16 > OCT_SMEM_NODES && (size_t)L.listCap * 4 < (size_t)(L.nDesired + 16) * (sizeof(OctNodeDev) + 8)) return IVG_ERR_CAPACITY;
    }
  fs.nCellsTotal = (int)h->cellsPlain.size();
  fs.kpCap = kpOff;
  fs.btTotal = btBase;
  {
    // k_level_select shared memory: level list up to 2x the largest per-level budget (the reference itself reserves
    // nDesired*2, :1136), a modest per-warp cell list, per-cell scalars; longer lists fall back to global memory
    int maxDesired = 1, maxCells = 1;
    for (int l = 0; l < nl; ++l) { maxDesired = std::max(maxDesired, fs.lv[l].nDesired); maxCells = std::max(maxCells, fs.lv[l].nCells); }
    fs.selLevelCap = std::min((int)align_up(2 * maxDesired + 64, 64), 8192);
    fs.selCellCap = 128;
    fs.selCells = (int)align_up(maxCells, 4);
    h->selSmem = sel_smem_bytes(fs.selLevelCap, fs.selCellCap, fs.selCells);
    h->selSmemLat = sel_smem_bytes(fs.selLevelCap, fs.selCellCap, fs.selCells, SEL_WARPS_LAT);
  }
  h->fastSmem = fastSmem; h->fastSmemLat = fastSmemLat; h->resizeSmem = resizeSmem;
  if (fastSmem > 200 * 1024 || fastSmemLat > 200 * 1024 || resizeSmem > 200 * 1024) return IVG_ERR_CAPACITY;
  if (fs.kpCap > 65535) return IVG_ERR_CAPACITY;           // stereo packs the right index in 16 bits

  const size_t B = (size_t)batch;
  int rc;
  if ((rc = h->pyr.alloc(B * fs.planeBytes))) return rc;
  if ((rc = h->blur.alloc(B * fs.planeBytes))) return rc;
  if (h->enableIntrospection && (rc = h->qual.alloc(B * fs.planeBytes))) return rc;
  if ((rc = h->dCellsPlain.alloc(h->cellsPlain.size()))) return rc;
  if ((rc = h->dCellsWeighted.alloc(h->cellsWeighted.size()))) return rc;
  if ((rc = h->rtab.alloc(taps.size()))) return rc;
  if ((rc = h->pyrSpanX.alloc(spanX.size())) || (rc = h->pyrSpanY.alloc(spanY.size()))) return rc;
  std::vector<uint32_t> btab;
  for (int l = 0; l < nl; ++l)
    for (int ty = 0; ty < fs.lv[l].btY; ++ty)
      for (int tx = 0; tx < fs.lv[l].btX; ++tx) btab.push_back(((uint32_t)l << 28) | ((uint32_t)tx << 14) | (uint32_t)ty);
  if ((rc = h->blurTiles.alloc(btab.size()))) return rc;
  if ((rc = h->cellList.alloc(B * fs.listCapTotal))) return rc;
  fs.cellCostStride = fs.nCellsTotal + MAX_LEVELS + 1;
  if ((rc = h->cellCost.alloc(B * fs.cellCostStride))) return rc;
  if ((rc = h->cellCount.alloc(B * fs.nCellsTotal))) return rc;
  if ((rc = h->workCell.alloc(B * fs.listCapTotal))) return rc;
  if ((rc = h->workLevel.alloc(B * fs.listCapTotal))) return rc;
  if ((rc = h->levelKp.alloc(B * fs.kpCap))) return rc;
  if ((rc = h->levelCount.alloc(B * MAX_LEVELS))) return rc;
  {
    // keypoint records | descriptors | counts in ONE allocation: a full batch goes to the host in one copy
    const size_t descOff = align_up(B * fs.kpCap * 28, 256), nOff = descOff + align_up(B * fs.kpCap * 32, 256);
    h->outKp.release(); h->outDesc.release(); h->outN.release();
    if ((rc = h->outAll.alloc(nOff + B * sizeof(int)))) return rc;
    h->outKp.view_of(h->outAll.p, B * fs.kpCap * 28);
    h->outDesc.view_of(h->outAll.p + descOff, B * fs.kpCap * 32);
    h->outN.view_of(h->outAll.p + nOff, B);
    h->outDescOff = descOff; h->outNOff = nOff;
  }
  h->uRight.release(); h->depth.release();                     // uRight | depth in one allocation, for the same reason
  if ((rc = h->udAll.alloc(2 * B * fs.kpCap))) return rc;
  h->uRight.view_of(h->udAll.p, B * fs.kpCap);
  h->depth.view_of(h->udAll.p + B * fs.kpCap, B * fs.kpCap);
  if ((rc = h->sad.alloc(B * fs.kpCap))) return rc;
  CK(cudaMemcpyAsync(h->dCellsPlain.p, h->cellsPlain.data(), h->cellsPlain.size() * sizeof(CellDev), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->dCellsWeighted.p, h->cellsWeighted.data(), h->cellsWeighted.size() * sizeof(CellDev), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->blurTiles.p, btab.data(), btab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  if (!taps.empty()) CK(cudaMemcpyAsync(h->rtab.p, taps.data(), taps.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->pyrSpanX.p, spanX.data(), spanX.size() * sizeof(PyrSpan), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->pyrSpanY.p, spanY.data(), spanY.size() * sizeof(PyrSpan), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->outAll.p, 0, h->outAll.n, h->stream));      // counts start at 0; record slots a frame does not fill and the alignment gaps are copied out with the block
  CK(cudaMemsetAsync(h->levelCount.p, 0, B * MAX_LEVELS * sizeof(int), h->stream));
  CK(cudaStreamSynchronize(h->stream));   // host vectors go out of scope

  fs.pyr = h->pyr.p; fs.blur = h->blur.p; fs.qual = h->qual.p;
  fs.rtab = h->rtab.p; fs.blurTiles = h->blurTiles.p;
  fs.cellList = h->cellList.p; fs.cellCount = h->cellCount.p; fs.cellCost = h->cellCost.p;
  fs.workCell = h->workCell.p; fs.workLevel = h->workLevel.p; fs.levelKp = h->levelKp.p;
  fs.levelCount = h->levelCount.p; fs.outKp = h->outKp.p; fs.outDesc = h->outDesc.p; fs.outN = h->outN.p;
  for (int l = 1; l < nl; ++l) {
    const LevelDev& D = fs.lv[l];
    const LevelDev& S = fs.lv[l - 1];
    h->resizeTma[l] = D.rzPitch <= 256 && D.rzRows <= 256;
    if (!h->resizeTma[l]) continue;
    if ((rc = make_level_map(&h->resizeMaps.m[l], h->pyr.p + S.planeOff, S.w, S.h, S.pitch, fs.planeBytes, batch, D.rzPitch, D.rzRows))) return rc;
    if (h->enableIntrospection &&
        (rc = make_level_map(&h->resizeMapsQ.m[l], h->qual.p + S.planeOff, S.w, S.h, S.pitch, fs.planeBytes, batch, D.rzPitch, D.rzRows)))
      return rc;
  }
  for (int l = 0; l < nl; ++l)
    if ((rc = make_level_map(&h->descMaps.m[l], h->blur.p + fs.lv[l].planeOff, fs.lv[l].w, fs.lv[l].h, fs.lv[l].pitch, fs.planeBytes,
                             batch, DK_BOXW, DK_BOXH)) ||
        (rc = make_level_map(&h->descMapsN.m[l], h->blur.p + fs.lv[l].planeOff, fs.lv[l].w, fs.lv[l].h, fs.lv[l].pitch, fs.planeBytes,
                             batch, DK_BOXN, DK_BOXH)))
      return rc;
  for (int l = 0; l < nl; ++l)
    if ((rc = make_level_map(&h->blurMaps.m[l], h->pyr.p + fs.lv[l].planeOff, fs.lv[l].w, fs.lv[l].h, fs.lv[l].pitch, fs.planeBytes,
                             batch, BL_BOXW, BL_PH)))
      return rc;
  h->fs = fs;
  h->kpCap = fs.kpCap;
  h->W = W; h->H = H; h->maxBatch = batch; h->shapeReady = true;
  h->haveResults = false; h->havePyramid = false; h->curBatch = 0;
  return IVG_OK;
}

int ensure_shape(ivg_extractor* h, int W, int H, int batch) {
  if (W <= 0 || H <= 0 || batch <= 0) return IVG_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  if (h->shapeReady && h->W == W && h->H == H && batch <= h->maxBatch) return IVG_OK;
  if (!h->shapeReady && h->W == W && h->H == H) batch = std::max(batch, h->maxBatch);   // mode change: keep the reserved batch
  CK(cudaStreamSynchronize(h->stream));
  const int b = (h->shapeReady && h->W == W && h->H == H) ? std::max(batch, h->maxBatch) : batch;
  h->shapeReady = false;
  drop_graph(h);
  return build_shape(h, W, H, b);
}

int honour_wait(ivg_extractor* h) {
  if (h->waitFor) { CK(cudaStreamWaitEvent(h->stream, h->waitFor, 0)); h->waitFor = nullptr; }
  return IVG_OK;
}

// one-frame (latency) configuration of a kernel?  `small` = the launch would leave most of the GPU idle in the throughput configuration
static inline bool one_frame_cfg(const ivg_extractor* h, bool small) { return h->forceCfg == 2 || (h->forceCfg == 0 && small); }

FrameSet active_fs(const ivg_extractor* h) {
  FrameSet fs = h->fs;
  fs.nImages = h->curBatch;
  fs.weighted = h->curWeighted ? 1 : 0;
  fs.cells = (h->curWeighted && h->kpMode == 0) ? h->dCellsWeighted.p : h->dCellsPlain.p;
  fs.fastRetry = h->kpMode == 1 ? 0 : 3;
  return fs;
}

int launch_pyramid(ivg_extractor* h, const FrameSet& fs) {
  const int planes = fs.nImages * (fs.weighted ? 2 : 1);
  static const bool noFused = getenv("IVSLAM_NO_FUSED_PYRAMID") != nullptr;      // developer A/B switch
  if (h->pyrTX > 0 && !noFused && one_frame_cfg(h, (long long)planes * h->pyrTX * h->pyrTY <= 2 * 148) && (long long)planes * h->pyrTX * h->pyrTY <= 65535) {
    // one frame at a time: the whole cascade in one launch (k_pyramid_fused.cuh) instead of nlevels-1 dependent ones
    ProfScope ps(h, IVG_K_RESIZE);
    PyrFusedArgs A{};
    A.spanX = h->pyrSpanX.p; A.spanY = h->pyrSpanY.p; A.TX = h->pyrTX; A.TY = h->pyrTY; A.bufBytes = (int)h->pyrFusedBuf; A.tOff = (int)h->pyrFusedTOff;
    for (int l = 0; l < MAX_LEVELS; ++l) { A.tapOffX[l] = h->pyrTapOffX[l]; A.tapOffY[l] = h->pyrTapOffY[l]; }
    k_pyramid_fused<<<dim3(h->pyrTX * h->pyrTY, planes), PF_THREADS, h->pyrFusedSmem, h->stream>>>(fs, A);
    CK(cudaGetLastError());
    return IVG_OK;
  }
  for (int l = 1; l < fs.nlevels; ++l) {
    dim3 grid((fs.lv[l].w + RZ_W - 1) / RZ_W, (fs.lv[l].h + RZ_H - 1) / RZ_H, fs.nImages * (fs.weighted ? 2 : 1));   // image (+ cost-map) planes
    { ProfScope ps(h, IVG_K_RESIZE); k_resize_level<<<grid, 256, h->resizeSmem, h->stream>>>(fs, l, h->resizeMaps, h->resizeMapsQ, h->resizeTma[l] ? 1 : 0); }
  }
  CK(cudaGetLastError());
  return IVG_OK;
}

int launch_extract_kernels(ivg_extractor* h, const FrameSet& fs) {
  int rc = launch_pyramid(h, fs);
  if (rc) return rc;
  // one frame at a time: the blur only needs the pyramid, so it runs on a side stream next to FAST + selection and joins
  // before the descriptors (a fork/join inside the captured graph in graph mode).  Big batches fill the GPU with either
  // kernel alone and co-running them was measured slower, so they stay on one stream.
  const bool fork = h->aux && !h->profile && fs.btTotal * fs.nImages <= 4 * 148;
  if (fork) {
    CK(cudaEventRecord(h->evFork, h->stream));
    CK(cudaStreamWaitEvent(h->aux, h->evFork, 0));
    h->launches++;
    k_gauss7<<<dim3(fs.btTotal, fs.nImages), 256, 0, h->aux>>>(fs, h->blurMaps);
    CK(cudaEventRecord(h->evJoin, h->aux));
  }
  {
    ProfScope ps(h, IVG_K_FAST);
    if (one_frame_cfg(h, (long long)fs.nCellsTotal * fs.nImages <= 4 * 148)) {      // one or two KITTI frames: more warps per cell (measured: 4 frames and up prefer 128)
      FrameSet fl = fs;
      for (int l = 0; l < fs.nlevels; ++l) { fl.lv[l].fBH = h->fastLat[l][0]; fl.lv[l].fBX = h->fastLat[l][1]; fl.lv[l].fSeg = h->fastLat[l][2]; }
      k_fast_cells<FC_THREADS_LAT><<<dim3(fs.nCellsTotal, fs.nImages), FC_THREADS_LAT, h->fastSmemLat, h->stream>>>(fl);
    } else {
      k_fast_cells<FC_THREADS><<<dim3(fs.nCellsTotal, fs.nImages), FC_THREADS, h->fastSmem, h->stream>>>(fs);
    }
  }
  if (!fork) { ProfScope ps(h, IVG_K_BLUR); k_gauss7<<<dim3(fs.btTotal, fs.nImages), 256, 0, h->stream>>>(fs, h->blurMaps); }
  if (h->kpMode == 1) { ProfScope ps(h, IVG_K_SELECT); k_octree_select<<<dim3(fs.nlevels, fs.nImages), 256, sizeof(OctShared), h->stream>>>(fs); }
  else {
    // few CTAs (one per level and frame): give each every warp it can use; big batches fill the GPU with 8-warp CTAs
    const bool lat = one_frame_cfg(h, fs.nlevels * fs.nImages <= 2 * 148) && h->selSmemLat <= 200 * 1024;
    ProfScope ps(h, IVG_K_SELECT);
    if (lat) k_level_select<SEL_WARPS_LAT * 32><<<dim3(fs.nlevels, fs.nImages), SEL_WARPS_LAT * 32, h->selSmemLat, h->stream>>>(fs);
    else k_level_select<SEL_WARPS * 32><<<dim3(fs.nlevels, fs.nImages), SEL_WARPS * 32, h->selSmem, h->stream>>>(fs);
  }
  if (fork) CK(cudaStreamWaitEvent(h->stream, h->evJoin, 0));
  {
    ProfScope ps(h, IVG_K_DESCRIBE);
    // one frame at a time: 16 keypoint slots per CTA, so that ~2000 keypoints spread over 125 CTAs instead of 32
    if (one_frame_cfg(h, (fs.kpCap + DK_SLOTS_BATCH - 1) / DK_SLOTS_BATCH * fs.nImages < 2 * 148))
      k_orient_describe<DK_SLOTS_LAT><<<dim3((fs.kpCap + DK_SLOTS_LAT - 1) / DK_SLOTS_LAT, fs.nImages), 256, 0, h->stream>>>(fs, h->descMaps, h->descMapsN);
    else
      k_orient_describe<DK_SLOTS_BATCH><<<dim3((fs.kpCap + DK_SLOTS_BATCH - 1) / DK_SLOTS_BATCH, fs.nImages), 256, 0, h->stream>>>(fs, h->descMaps, h->descMapsN);
  }
  if (h->eagerIndex && fs.nImages <= EAGER_INDEX_MAX_BATCH) {
    // Right eye of one-frame-at-a-time stereo: the matcher's (octave, row) index only depends on this handle's keypoints,
    // so it is built here, off the critical path of ivg_stereo_match (it overlaps this eye's download and the host's join).
    StereoArgs A{};
    A.kpR = h->outKp.p; A.nR = h->outN.p; A.cap = fs.kpCap; A.nRows = fs.lv[0].h; A.nLevels = fs.nlevels;
    A.sorted = h->sortedR.p; A.rowStart = h->rowStart.p;
    ProfScope ps(h, IVG_K_STEREO);
    k_stereo_index<<<fs.nImages, 1024, ((size_t)A.nRows * A.nLevels + 1) * sizeof(int), h->stream>>>(A);      // eager: a handful of frames
  }
  CK(cudaGetLastError());
  return IVG_OK;
}

int launch_extract(ivg_extractor* h) {
  const FrameSet fs = active_fs(h);
  if (h->eagerIndex && fs.nImages <= EAGER_INDEX_MAX_BATCH) {      // buffers first: no allocation inside a stream capture
    const size_t nBins = (size_t)fs.lv[0].h * fs.nlevels;
    int rc;
    if ((nBins + 1) * sizeof(int) > 200 * 1024) h->eagerIndex = false;
    else if ((rc = h->sortedR.alloc((size_t)fs.nImages * fs.kpCap)) || (rc = h->rowStart.alloc((size_t)fs.nImages * (nBins + 1)))) return rc;
  }
  if (h->graphMode && !h->profile) {
    // one graph launch instead of ~11 kernel launches: matters for the one-frame-at-a-time (drop-in) use
    if (!h->graphExec || h->graphBatch != fs.nImages || h->graphWeighted != (fs.weighted != 0) || h->graphEager != h->eagerIndex) {
      drop_graph(h);
      cudaGraph_t g = nullptr;
      const long long before = h->launches;
      CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      int rc = launch_extract_kernels(h, fs);
      cudaError_t e = cudaStreamEndCapture(h->stream, &g);
      if (rc) { if (g) cudaGraphDestroy(g); return rc; }
      if (e != cudaSuccess) { g_cuda_err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return IVG_ERR_CUDA; }
      e = cudaGraphInstantiate(&h->graphExec, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) { g_cuda_err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); h->graphExec = nullptr; return IVG_ERR_CUDA; }
      h->graphLaunches = (int)(h->launches - before);
      h->launches = before;
      h->graphBatch = fs.nImages; h->graphWeighted = fs.weighted != 0; h->graphEager = h->eagerIndex;
    }
    CK(cudaGraphLaunch(h->graphExec, h->stream));
    h->launches += h->graphLaunches;
  } else {
    int rc = launch_extract_kernels(h, fs);
    if (rc) return rc;
  }
  h->haveResults = true; h->havePyramid = true; h->haveGrid = false; h->haveStereo = false;
  h->indexValid = h->eagerIndex && fs.nImages <= EAGER_INDEX_MAX_BATCH;
  return IVG_OK;
}

int init_device_constants(int device) {
  CK(cudaSetDevice(device));
  {
    std::vector<float2> pt(512);
    for (int i = 0; i < 512; ++i) {   // pattern point i = byte*16 + k  ->  slot k*32 + byte
      const int byte = i >> 4, k = i & 15;
      pt[k * 32 + byte] = make_float2((float)kPatternHost[2 * i], (float)kPatternHost[2 * i + 1]);
    }
    CK(cudaMemcpyToSymbol(g_patternT, pt.data(), sizeof(float2) * 512));
  }
  CK(cudaFuncSetAttribute(k_level_select<SEL_WARPS * 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_level_select<SEL_WARPS_LAT * 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_fast_cells<FC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_fast_cells<FC_THREADS_LAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_octree_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OctShared)));
  CK(cudaFuncSetAttribute(k_resize_level, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_pyramid_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_stereo_index, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return IVG_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* ivg_strerror(int s) {
  switch (s) {
    case IVG_OK: return "ok";
    case IVG_ERR_INVALID: return "invalid argument";
    case IVG_ERR_GEOMETRY: return "image too small for the reference cell grid";
    case IVG_ERR_CAPACITY: return "capacity exceeded";
    case IVG_ERR_CUDA: return "CUDA error";
    case IVG_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case IVG_ERR_STATE: return "invalid call order / state";
  }
  return "unknown";
}

const char* ivg_last_cuda_error(void) { return g_cuda_err.c_str(); }

int ivg_device_info(int device, char* name, int name_cap, int* sm, int* sm_count) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { g_cuda_err = "no CUDA device"; return IVG_ERR_NO_DEVICE; }
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  if (name && name_cap > 0) { std::strncpy(name, p.name, name_cap - 1); name[name_cap - 1] = 0; }
  if (sm) *sm = p.major * 10 + p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  return p.major == 10 ? IVG_OK : IVG_ERR_NO_DEVICE;   // the library carries sm_100a code only
}

int ivg_extractor_create(ivg_extractor** out, int device, int nfeatures, float scaleFactor, int nlevels,
                         int iniThFAST, int minThFAST, int enableIntrospection) {
  if (!out) return IVG_ERR_INVALID;
  *out = nullptr;
  if (nfeatures < 1 || nlevels < 1 || nlevels > MAX_LEVELS || !(scaleFactor > 1.0f)) return IVG_ERR_INVALID;
  if (device < 0 || device >= 64) return IVG_ERR_INVALID;     // per-device constants are tracked in a 64-entry table
  int rc = ivg_device_info(device, nullptr, 0, nullptr, nullptr);
  if (rc) return rc;
  CK(cudaSetDevice(device));
  ivg_extractor* h = new ivg_extractor();
  h->device = device; h->nfeatures = nfeatures; h->scaleFactor = scaleFactor; h->nlevels = nlevels;
  h->iniTh = iniThFAST; h->minTh = minThFAST; h->enableIntrospection = enableIntrospection != 0;
  build_tables(h);
  // one-time per-device constants (pattern) — per device in a multi-GPU process
  static std::mutex mu;
  static bool done[64] = {false};
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!done[device]) {
      rc = init_device_constants(device);
      if (rc) { delete h; return rc; }
      done[device] = true;
    }
  }
  {
    static const int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};   // compiled into k_orient_describe
    for (int i = 0; i < 16; ++i)
      if (h->umax[i] != kUmax[i]) { g_cuda_err = "umax table mismatch"; delete h; return IVG_ERR_INVALID; }
  }
  bool ok = cudaStreamCreateWithFlags(&h->copyIn, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->copyOut, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking) == cudaSuccess;
  for (cudaEvent_t* e : {&h->evH2D, &h->evIngest, &h->evKernels, &h->evD2H, &h->evStereo, &h->evD2Hs, &h->evConsumed, &h->evFork, &h->evJoin})
    ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
  if (!ok || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->evDone, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&h->evT0) != cudaSuccess || cudaEventCreate(&h->evT1) != cudaSuccess) {
    g_cuda_err = "stream/event creation failed"; delete h; return IVG_ERR_CUDA;
  }
  *out = h;
  return IVG_OK;
}

void ivg_extractor_destroy(ivg_extractor* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->link) {      // the partner keeps the (now dead) link; it is replaced at its next explicit match
    std::lock_guard<std::mutex> lk(h->link->mu);
    if (h->link->left == h) h->link->left = nullptr;
    if (h->link->right == h) h->link->right = nullptr;
    h->link->specValid = false;
  }
  h->link.reset();
  if (h->specHost) { cudaFreeHost(h->specHost); h->specHost = nullptr; }
  // a shared kernel stream belongs to another handle (which may already be gone): the cudaFree calls below synchronise
  // the device anyway, so only streams this handle owns are touched
  if (h->stream && h->ownsStream) cudaStreamSynchronize(h->stream);
  if (h->copyIn) cudaStreamSynchronize(h->copyIn);
  if (h->copyOut) cudaStreamSynchronize(h->copyOut);
  h->pyr.release(); h->blur.release(); h->qual.release(); h->outKp.release(); h->outDesc.release(); h->outN.release(); h->outAll.release(); h->stageImg.release(); h->stageCost.release(); h->mapX.release(); h->mapY.release(); h->projIn.release(); h->projCand.release(); h->projInt.release(); if (h->projHost) { cudaFreeHost(h->projHost); h->projHost = nullptr; } if (h->outHost) { cudaFreeHost(h->outHost); h->outHost = nullptr; }
  h->dCellsPlain.release(); h->dCellsWeighted.release(); h->rtab.release(); h->pyrSpanX.release(); h->pyrSpanY.release(); h->cellList.release(); h->cellCost.release(); h->blurTiles.release();
  h->cellCount.release(); h->workCell.release(); h->workLevel.release(); h->levelKp.release(); h->levelCount.release();
  h->kpQual.release(); h->gridStart.release(); h->gridIdx.release();
  h->outN.release(); h->sad.release(); h->nExt.release(); h->uRight.release(); h->depth.release(); h->udAll.release();
  h->extKpL.release(); h->extDescL.release(); h->extKpR.release(); h->extDescR.release(); h->extU.release(); h->extD.release(); h->extS.release(); h->sortedR.release(); h->rowStart.release();
  for (cudaEvent_t e : h->profEv) cudaEventDestroy(e);
  drop_graph(h);
  if (h->evDone) cudaEventDestroy(h->evDone);
  if (h->evT0) cudaEventDestroy(h->evT0);
  if (h->evT1) cudaEventDestroy(h->evT1);
  for (cudaEvent_t e : {h->evH2D, h->evIngest, h->evKernels, h->evD2H, h->evStereo, h->evD2Hs, h->evConsumed, h->evFork, h->evJoin}) if (e) cudaEventDestroy(e);
  if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
  if (h->copyIn) cudaStreamDestroy(h->copyIn);
  if (h->copyOut) cudaStreamDestroy(h->copyOut);
  if (h->stream && h->ownsStream) cudaStreamDestroy(h->stream);
  delete h;
}

int ivg_extractor_reserve(ivg_extractor* h, int width, int height, int max_batch) {
  if (!h) return IVG_ERR_INVALID;
  return ensure_shape(h, width, height, max_batch);
}

int ivg_get_levels(const ivg_extractor* h) { return h ? h->nlevels : 0; }
float ivg_get_scale_factor(const ivg_extractor* h) { return h ? (float)h->scaleFactor : 0.f; }
int ivg_get_scale_table(const ivg_extractor* h, int which, float* out) {
  if (!h || !out || which < 0 || which > 3) return IVG_ERR_INVALID;
  const std::vector<float>& t = which == 0 ? h->scale : which == 1 ? h->invScale : which == 2 ? h->sigma2 : h->invSigma2;
  for (int i = 0; i < h->nlevels; ++i) out[i] = t[i];
  return IVG_OK;
}
int ivg_get_features_per_level(const ivg_extractor* h, int* out) {
  if (!h || !out) return IVG_ERR_INVALID;
  for (int i = 0; i < h->nlevels; ++i) out[i] = h->nPerLevel[i];
  return IVG_OK;
}
int ivg_max_keypoints(const ivg_extractor* h) { return h ? h->kpCap : 0; }

int ivg_extractor_set_mode(ivg_extractor* h, int mode) {
  if (!h || (mode != 0 && mode != 1)) return IVG_ERR_INVALID;
  if (mode == h->kpMode) return IVG_OK;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->kpMode = mode;
  h->kpCap = 0;
  for (int l = 0; l < h->nlevels; ++l) h->kpCap += h->nPerLevel[l] + (mode == 1 ? 3 : 0);
  h->shapeReady = false;      // tables and capacities are rebuilt on the next call
  h->haveResults = false;
  drop_graph(h);
  return IVG_OK;
}

int ivg_set_batch(ivg_extractor* h, int n, int width, int height, int with_cost) {
  if (!h) return IVG_ERR_INVALID;
  int rc = ensure_shape(h, width, height, n);
  if (rc) return rc;
  h->curBatch = n;
  h->runGen++;
  h->curWeighted = with_cost && h->enableIntrospection;   // src/ORBextractor.cc:1231
  h->haveCost = with_cost != 0;
  if (with_cost) {
    // cost-map planes are also needed by ivg_frame_postprocess on handles without introspection (mvKeyQualScore is
    // computed from the map regardless of the flag, Frame.cc:128-139).  build_shape only sizes `qual` for introspection
    // handles, so size it here for the current shape and batch every time (a no-op once it is large enough).
    if ((rc = h->qual.alloc((size_t)h->maxBatch * h->fs.planeBytes))) return rc;
    h->fs.qual = h->qual.p;
  }
  return IVG_OK;
}

int ivg_device_input(ivg_extractor* h, int index, int which, void** dev_ptr, size_t* pitch) {
  if (!h || !h->shapeReady || index < 0 || index >= h->maxBatch || !dev_ptr) return IVG_ERR_INVALID;
  uint8_t* base = which == 0 ? h->pyr.p : which == 2 ? h->qual.p : nullptr;
  if (!base) return IVG_ERR_INVALID;
  *dev_ptr = base + (size_t)index * h->fs.planeBytes;
  if (pitch) *pitch = h->fs.lv[0].pitch;
  return IVG_OK;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
static int copy_frames_in(ivg_extractor* h, uint8_t* plane, DevBuf<uint8_t>& stage, int n, const uint8_t* src, size_t stride, size_t frame_bytes,
                          bool src_on_device = false) {
  const FrameSet& fs = h->fs;
  const size_t dpitch = fs.lv[0].pitch;
  if (src_on_device) {
    // frames already in device memory (e.g. the introspection CNN's cost-map): no staging copy at all when they are
    // contiguous and word aligned — the ingest kernel reads them in place; otherwise a device-to-device pitched copy
    if (stride == (size_t)h->W && frame_bytes == (size_t)h->W * h->H && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
      dim3 grid(((h->W + 15) / 16 + 255) / 256, h->H, n);
      k_ingest<<<grid, 256, 0, h->stream>>>(src, plane, fs.planeBytes, h->W, h->H, (int)dpitch);
      h->launches++;
      CK(cudaGetLastError());
    } else {
      for (int f = 0; f < n; ++f)
        CK(cudaMemcpy2DAsync(plane + (size_t)f * fs.planeBytes, dpitch, src + (size_t)f * frame_bytes, stride, h->W, h->H,
                             cudaMemcpyDeviceToDevice, h->stream));
    }
    return IVG_OK;
  }
  if (stride == (size_t)h->W && frame_bytes == (size_t)h->W * h->H) {
    // contiguous frames: one linear DMA + on-device re-pitch
    const size_t bytes = (size_t)n * frame_bytes;
    int rc = stage.alloc(bytes + 16);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(h->copyIn, h->evIngest, 0));          // the previous ingest has consumed the staging buffer
    CK(cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, h->copyIn));
    CK(cudaEventRecord(h->evH2D, h->copyIn));
    CK(cudaStreamWaitEvent(h->stream, h->evH2D, 0));
    dim3 grid(((h->W + 15) / 16 + 255) / 256, h->H, n);
    k_ingest<<<grid, 256, 0, h->stream>>>(stage.p, plane, fs.planeBytes, h->W, h->H, (int)dpitch);
    h->launches++;
    CK(cudaGetLastError());
    return IVG_OK;
  }
  if (n > 1 && stride > 0 && frame_bytes % stride == 0 && fs.planeBytes % dpitch == 0) {
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(src), stride, h->W, frame_bytes / stride);
    p.dstPtr = make_cudaPitchedPtr(plane, dpitch, h->W, fs.planeBytes / dpitch);
    p.extent = make_cudaExtent(h->W, h->H, n);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3DAsync(&p, h->stream));
  } else {
    for (int f = 0; f < n; ++f)
      CK(cudaMemcpy2DAsync(plane + (size_t)f * fs.planeBytes, dpitch, src + (size_t)f * frame_bytes, stride, h->W, h->H,
                           cudaMemcpyHostToDevice, h->stream));
  }
  return IVG_OK;
}

int ivg_upload_batch(ivg_extractor* h, int n, const uint8_t* images, int width, int height, size_t stride,
                     size_t frame_bytes, const uint8_t* costs, size_t cost_stride, size_t cost_frame_bytes) {
  if (!h || !images || n < 1 || stride < (size_t)width) return IVG_ERR_INVALID;
  int rc = ivg_set_batch(h, n, width, height, costs != nullptr);
  if (rc) return rc;
  if ((rc = honour_wait(h))) return rc;
  if ((rc = copy_frames_in(h, h->pyr.p, h->stageImg, n, images, stride, frame_bytes))) return rc;
  if (h->haveCost) {
    if (cost_stride < (size_t)width) return IVG_ERR_INVALID;
    if ((rc = copy_frames_in(h, h->qual.p, h->stageCost, n, costs, cost_stride, cost_frame_bytes))) return rc;
  }
  CK(cudaEventRecord(h->evIngest, h->stream));
  return IVG_OK;
}

int ivg_upload_batch_device(ivg_extractor* h, int n, const uint8_t* d_images, int width, int height, size_t stride,
                            size_t frame_bytes, const uint8_t* d_costs, size_t cost_stride, size_t cost_frame_bytes) {
  if (!h || !d_images || n < 1 || stride < (size_t)width) return IVG_ERR_INVALID;
  int rc = ivg_set_batch(h, n, width, height, d_costs != nullptr);
  if (rc) return rc;
  if ((rc = honour_wait(h))) return rc;
  if ((rc = copy_frames_in(h, h->pyr.p, h->stageImg, n, d_images, stride, frame_bytes, true))) return rc;
  if (h->haveCost) {
    if (cost_stride < (size_t)width) return IVG_ERR_INVALID;
    if ((rc = copy_frames_in(h, h->qual.p, h->stageCost, n, d_costs, cost_stride, cost_frame_bytes, true))) return rc;
  }
  return IVG_OK;
}

int ivg_upload_batch_device_cost_f32(ivg_extractor* h, int n, const uint8_t* d_images, int width, int height, size_t stride,
                                     size_t frame_bytes, const float* d_costs, size_t cost_stride_floats, size_t cost_frame_floats) {
  if (!h || !d_images || !d_costs || n < 1 || stride < (size_t)width || cost_stride_floats < (size_t)width) return IVG_ERR_INVALID;
  int rc = ivg_set_batch(h, n, width, height, 1);
  if (rc) return rc;
  if ((rc = honour_wait(h))) return rc;
  if ((rc = copy_frames_in(h, h->pyr.p, h->stageImg, n, d_images, stride, frame_bytes, true))) return rc;
  dim3 grid(((width + 3) / 4 + 255) / 256, height, n);
  k_cost_from_f32<<<grid, 256, 0, h->stream>>>(d_costs, cost_frame_floats, cost_stride_floats, h->qual.p, h->fs.planeBytes, width, height, h->fs.lv[0].pitch);
  h->launches++;
  CK(cudaGetLastError());
  return IVG_OK;
}

// ---------------------------------------------------------------------------------------- N2: SearchByProjection
namespace {
struct ProjUpload {   // packs the caller's host arrays into one pinned staging buffer -> one H2D copy
  ivg_extractor* h; size_t off = 0; std::vector<std::pair<size_t, std::pair<const void*, size_t>>> items;
  size_t add(const void* src, size_t bytes) { const size_t o = off; items.push_back({o, {src, bytes}}); off = (off + bytes + 15) & ~(size_t)15; return o; }
  int commit() {
    int rc = h->projIn.alloc(off + 16);
    if (rc) return rc;
    const size_t need = std::max(off + 16, ((size_t)h->fs.kpCap + 1) * sizeof(int));   // also holds the results (match[] + nmatches)
    if (h->projHostBytes < need) {
      CK(cudaStreamSynchronize(h->stream));
      if (h->projHost) cudaFreeHost(h->projHost);
      h->projHost = nullptr; h->projHostBytes = 0;
      CK(cudaMallocHost(&h->projHost, 2 * need));
      h->projHostBytes = 2 * need;
    }
    for (auto& it : items)
      if (it.second.first && it.second.second) std::memcpy((uint8_t*)h->projHost + it.first, it.second.first, it.second.second);
    if (off) CK(cudaMemcpyAsync(h->projIn.p, h->projHost, off, cudaMemcpyHostToDevice, h->stream));
    return IVG_OK;
  }
};
}  // namespace

static int proj_common(ivg_extractor* h, int index, ProjArgs& A, int n, float minX, float maxX, float minY, float maxY, int* match, int cap, int* nmatches) {
  if (!h->haveResults || !h->haveGrid) return IVG_ERR_STATE;              // needs ivg_frame_postprocess_batch (the 64x48 grid)
  if (index < 0 || index >= h->curBatch || n < 0 || !match || !nmatches) return IVG_ERR_INVALID;
  if (!(maxX > minX) || !(maxY > minY)) return IVG_ERR_INVALID;
  if (cap < h->fs.kpCap) return IVG_ERR_CAPACITY;
  const int K = h->fs.kpCap;
  int rc;
  if ((rc = h->projCand.alloc((size_t)std::max(n, 1) * K)) || (rc = h->projInt.alloc((size_t)std::max(n, 1) * 12 + K + 64))) return rc;
  A.kp = h->outKp.p; A.desc = h->outDesc.p; A.uRight = h->haveStereo ? h->uRight.p : nullptr; A.nPtr = h->outN.p; A.index = index; A.cap = K;
  A.gridStart = h->gridStart.p; A.gridIdx = h->gridIdx.p;
  A.minX = minX; A.minY = minY; A.maxX = maxX; A.maxY = maxY;
  A.invW = (float)GRID_COLS / (maxX - minX); A.invH = (float)GRID_ROWS / (maxY - minY);
  for (int l = 0; l < h->nlevels; ++l) A.scale[l] = h->scale[l];
  A.nLevels = h->nlevels;
  A.n = n;
  A.cand = h->projCand.p; A.candStride = K;
  int* ip = h->projInt.p;                     // layout: candCount[n] | tent[8n] (16-B aligned) | accIdx[n] | accBin[n bytes] | match[K] | nmatches
  A.candCount = ip;
  const size_t tentOff = ((size_t)n + 3) & ~(size_t)3;
  A.tent = reinterpret_cast<uint4*>(ip + tentOff);
  A.accIdx = ip + tentOff + 8 * (size_t)n;
  A.accBin = reinterpret_cast<int8_t*>(ip + tentOff + 9 * (size_t)n);
  A.match = ip + tentOff + 9 * (size_t)n + ((size_t)n + 3) / 4;
  A.nmatches = A.match + K;
  if (n > 0) { ProfScope ps(h, IVG_K_PROJ_CAND); k_proj_candidates<<<(n + 7) / 8, 256, 0, h->stream>>>(A); }
  { ProfScope ps(h, IVG_K_PROJ_RESOLVE); k_proj_resolve<<<1, 32, 0, h->stream>>>(A); }
  CK(cudaGetLastError());
  const size_t outBytes = ((size_t)K + 1) * sizeof(int);
  if (h->projHostBytes < outBytes) return IVG_ERR_STATE;     // sized by ProjUpload::commit
  CK(cudaMemcpyAsync(h->projHost, A.match, outBytes, cudaMemcpyDeviceToHost, h->stream));   // match[K] and nmatches are contiguous
  CK(cudaStreamSynchronize(h->stream));
  std::memcpy(match, h->projHost, (size_t)K * sizeof(int));
  *nmatches = ((const int*)h->projHost)[K];
  for (int i = K; i < cap; ++i) match[i] = -1;
  return IVG_OK;
}

int ivg_search_by_projection_last(ivg_extractor* cur, int index, int n, const float* world_pos, const uint8_t* desc, const int* octave,
                                  const float* angle, const uint8_t* flags, const float* Rcw, const float* tcw, float fx, float fy, float cx,
                                  float cy, float mbf, float minX, float maxX, float minY, float maxY, int mode, float th,
                                  int check_orientation, int* match, int cap, int* nmatches) {
  if (!cur || n < 0 || (n > 0 && (!world_pos || !desc || !octave || !angle || !flags)) || !Rcw || !tcw || mode < 0 || mode > 2) return IVG_ERR_INVALID;
  CK(cudaSetDevice(cur->device));
  ProjUpload up{cur};
  const size_t oW = up.add(world_pos, (size_t)n * 12), oD = up.add(desc, (size_t)n * 32), oO = up.add(octave, (size_t)n * 4),
               oA = up.add(angle, (size_t)n * 4), oF = up.add(flags, (size_t)n);
  int rc = up.commit();
  if (rc) return rc;
  ProjArgs A{};
  const uint8_t* b = cur->projIn.p;
  A.world = reinterpret_cast<const float*>(b + oW); A.pdesc = b + oD; A.octave = reinterpret_cast<const int*>(b + oO);
  A.angle = reinterpret_cast<const float*>(b + oA); A.flags = b + oF;
  for (int i = 0; i < 9; ++i) A.Rcw[i] = Rcw[i];
  for (int i = 0; i < 3; ++i) A.tcw[i] = tcw[i];
  A.fx = fx; A.fy = fy; A.cx = cx; A.cy = cy; A.mbf = mbf; A.mode = mode; A.th = th; A.nnratio = 0.f; A.checkOri = check_orientation != 0;
  return proj_common(cur, index, A, n, minX, maxX, minY, maxY, match, cap, nmatches);
}

int ivg_search_by_projection_map(ivg_extractor* cur, int index, int n, const float* proj, const float* view_cos, const int* level,
                                 const uint8_t* desc, const uint8_t* flags, const uint8_t* cur_blocked, float minX, float maxX, float minY,
                                 float maxY, float th, float nnratio, int* match, int cap, int* nmatches) {
  if (!cur || n < 0 || (n > 0 && (!proj || !view_cos || !level || !desc || !flags))) return IVG_ERR_INVALID;
  CK(cudaSetDevice(cur->device));
  ProjUpload up{cur};
  const size_t oP = up.add(proj, (size_t)n * 12), oV = up.add(view_cos, (size_t)n * 4), oL = up.add(level, (size_t)n * 4),
               oD = up.add(desc, (size_t)n * 32), oF = up.add(flags, (size_t)n), oB = up.add(cur_blocked, cur_blocked ? (size_t)cur->fs.kpCap : 0);
  int rc = up.commit();
  if (rc) return rc;
  ProjArgs A{};
  const uint8_t* b = cur->projIn.p;
  A.proj = reinterpret_cast<const float*>(b + oP); A.viewCos = reinterpret_cast<const float*>(b + oV); A.octave = reinterpret_cast<const int*>(b + oL);
  A.pdesc = b + oD; A.flags = b + oF; A.curBlocked = cur_blocked ? b + oB : nullptr;
  A.mode = 3; A.th = th; A.nnratio = nnratio; A.checkOri = 0;
  return proj_common(cur, index, A, n, minX, maxX, minY, maxY, match, cap, nmatches);
}

int ivg_search_by_bow(ivg_extractor* cur, int index, int n, const uint8_t* desc, const float* angle, const uint8_t* flags, const int* node_slot,
                      int n_nodes, const int* node_start, const int* node_idx, float nnratio, int check_orientation, int* match, int cap,
                      int* nmatches) {
  if (!cur || n < 0 || n_nodes < 0 || !match || !nmatches || (n > 0 && (!desc || !angle || !flags || !node_slot)) ||
      (n_nodes > 0 && (!node_start || !node_idx)))
    return IVG_ERR_INVALID;
  if (!cur->haveResults) return IVG_ERR_STATE;
  if (index < 0 || index >= cur->curBatch) return IVG_ERR_INVALID;
  if (cap < cur->fs.kpCap) return IVG_ERR_CAPACITY;
  CK(cudaSetDevice(cur->device));
  const int K = cur->fs.kpCap;
  const int total = n_nodes > 0 ? node_start[n_nodes] : 0;
  if (total < 0 || (n_nodes > 0 && node_start[0] != 0)) return IVG_ERR_INVALID;
  for (int s = 0; s < n_nodes; ++s) if (node_start[s + 1] < node_start[s]) return IVG_ERR_INVALID;   // CSR must be non-decreasing
  for (int i = 0; i < total; ++i) if (node_idx[i] < 0 || node_idx[i] >= K) return IVG_ERR_INVALID;    // the kernels index keypoints with these
  // points grouped by node slot, caller's order kept inside a slot (nodes never interact: a keypoint of F is in one node)
  std::vector<int> ptStart((size_t)n_nodes + 1, 0), ptIdx((size_t)std::max(n, 1));
  for (int i = 0; i < n; ++i) if (node_slot[i] >= 0 && node_slot[i] < n_nodes) ++ptStart[(size_t)node_slot[i] + 1];
  for (int s = 0; s < n_nodes; ++s) ptStart[(size_t)s + 1] += ptStart[s];
  {
    std::vector<int> cur_(ptStart.begin(), ptStart.end() - 1);
    for (int i = 0; i < n; ++i) if (node_slot[i] >= 0 && node_slot[i] < n_nodes) ptIdx[(size_t)cur_[node_slot[i]]++] = i;
  }
  ProjUpload up{cur};
  const size_t oD = up.add(desc, (size_t)n * 32), oA = up.add(angle, (size_t)n * 4), oF = up.add(flags, (size_t)n),
               oNS = up.add(node_start, (size_t)(n_nodes + 1) * 4), oNI = up.add(node_idx, (size_t)total * 4),
               oPS = up.add(ptStart.data(), ptStart.size() * 4), oPI = up.add(ptIdx.data(), (size_t)n * 4);
  int rc = up.commit();
  if (rc) return rc;
  if ((rc = cur->projInt.alloc((size_t)K + 64 + 2 * (size_t)std::max(n, 1)))) return rc;
  BowArgs A{};
  const uint8_t* b = cur->projIn.p;
  A.kp = cur->outKp.p; A.desc = cur->outDesc.p; A.nPtr = cur->outN.p; A.index = index; A.cap = K;
  A.n = n; A.pdesc = b + oD; A.angle = reinterpret_cast<const float*>(b + oA); A.flags = b + oF;
  A.nNodes = n_nodes; A.nodeStart = reinterpret_cast<const int*>(b + oNS); A.nodeIdx = reinterpret_cast<const int*>(b + oNI);
  A.ptStart = reinterpret_cast<const int*>(b + oPS); A.ptIdx = reinterpret_cast<const int*>(b + oPI);
  A.nnratio = nnratio; A.checkOri = check_orientation != 0;
  int* ip = cur->projInt.p;                   // layout: match[K] | nmatches | hist[30] (+pad to 64) | accIdx[n] | accBin[n bytes]
  A.match = ip; A.nmatches = ip + K; A.hist = ip + K + 1; A.accIdx = ip + K + 64; A.accBin = reinterpret_cast<int8_t*>(ip + K + 64 + n);
  CK(cudaMemsetAsync(ip, 0xFF, (size_t)K * 4, cur->stream));                        // match = -1
  CK(cudaMemsetAsync(ip + K, 0, 64 * 4, cur->stream));                             // nmatches, hist = 0
  if (n > 0) CK(cudaMemsetAsync(A.accBin, 0xFF, (size_t)n, cur->stream));            // accBin = -1
  if (n_nodes > 0 && n > 0) { ProfScope ps(cur, IVG_K_PROJ_CAND); k_bow_match<<<(n_nodes + 7) / 8, 256, 0, cur->stream>>>(A); }
  { ProfScope ps(cur, IVG_K_PROJ_RESOLVE); k_bow_finish<<<1, 256, 0, cur->stream>>>(A); }
  CK(cudaGetLastError());
  const size_t outBytes = ((size_t)K + 1) * sizeof(int);
  if (cur->projHostBytes < outBytes) return IVG_ERR_STATE;     // sized by ProjUpload::commit
  CK(cudaMemcpyAsync(cur->projHost, A.match, outBytes, cudaMemcpyDeviceToHost, cur->stream));   // match[K] and nmatches are contiguous
  CK(cudaStreamSynchronize(cur->stream));
  std::memcpy(match, cur->projHost, (size_t)K * sizeof(int));
  *nmatches = ((const int*)cur->projHost)[K];
  for (int i = K; i < cap; ++i) match[i] = -1;
  return IVG_OK;
}

// ---------------------------------------------------------------------------------------- N4: input prologue
int ivg_set_rectify_maps(ivg_extractor* h, const float* mapx, const float* mapy, int width, int height, size_t stride_floats) {
  if (!h) return IVG_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (!mapx || !mapy) { h->mapW = h->mapH = 0; return IVG_OK; }
  if (width < 1 || height < 1 || stride_floats < (size_t)width) return IVG_ERR_INVALID;
  int rc;
  const size_t n = (size_t)width * height;
  if ((rc = h->mapX.alloc(n)) || (rc = h->mapY.alloc(n))) return rc;
  CK(cudaMemcpy2D(h->mapX.p, (size_t)width * 4, mapx, stride_floats * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice));
  CK(cudaMemcpy2D(h->mapY.p, (size_t)width * 4, mapy, stride_floats * 4, (size_t)width * 4, height, cudaMemcpyHostToDevice));
  h->mapW = width; h->mapH = height;
  return IVG_OK;
}

static int prologue_in(ivg_extractor* h, uint8_t* plane, DevBuf<uint8_t>& stage, int n, const uint8_t* src, int sw, int sh, size_t stride,
                       size_t frame_bytes, int cn, int rgb) {
  const size_t rowB = (size_t)sw * cn, fb = rowB * sh;
  int rc = stage.alloc((size_t)n * fb + 16);
  if (rc) return rc;
  CK(cudaStreamWaitEvent(h->copyIn, h->evIngest, 0));            // the previous prologue has consumed the staging buffer
  if (stride == rowB && frame_bytes == fb) {
    CK(cudaMemcpyAsync(stage.p, src, (size_t)n * fb, cudaMemcpyHostToDevice, h->copyIn));
  } else {
    for (int f = 0; f < n; ++f)
      CK(cudaMemcpy2DAsync(stage.p + (size_t)f * fb, rowB, src + (size_t)f * frame_bytes, stride, rowB, sh, cudaMemcpyHostToDevice, h->copyIn));
  }
  CK(cudaEventRecord(h->evH2D, h->copyIn));
  CK(cudaStreamWaitEvent(h->stream, h->evH2D, 0));
  PrologueArgs P{};
  P.src = stage.p; P.srcFrameBytes = fb; P.sw = sw; P.sh = sh; P.cn = cn; P.rgb = rgb;
  P.mapx = h->mapW ? h->mapX.p : nullptr; P.mapy = h->mapW ? h->mapY.p : nullptr;
  P.plane = plane; P.planeBytes = h->fs.planeBytes; P.W = h->W; P.H = h->H; P.pitch = h->fs.lv[0].pitch;
  dim3 grid(((h->W + 3) / 4 + 255) / 256, h->H, n);
  { ProfScope ps(h, IVG_K_PROLOGUE); k_prologue<<<grid, 256, 0, h->stream>>>(P); }
  CK(cudaGetLastError());
  return IVG_OK;
}

int ivg_upload_batch_raw(ivg_extractor* h, int n, const uint8_t* frames, int src_width, int src_height, size_t stride, size_t frame_bytes,
                         int channels, int rgb_order, const uint8_t* costs, size_t cost_stride, size_t cost_frame_bytes) {
  if (!h || !frames || n < 1 || src_width < 1 || src_height < 1) return IVG_ERR_INVALID;
  if (channels != 1 && channels != 3 && channels != 4) return IVG_ERR_INVALID;
  if (stride < (size_t)src_width * channels) return IVG_ERR_INVALID;
  const int W = h->mapW ? h->mapW : src_width, H = h->mapW ? h->mapH : src_height;
  int rc = ivg_set_batch(h, n, W, H, costs != nullptr);
  if (rc) return rc;
  if ((rc = honour_wait(h))) return rc;
  if ((rc = prologue_in(h, h->pyr.p, h->stageImg, n, frames, src_width, src_height, stride, frame_bytes, channels, rgb_order != 0))) return rc;
  if (h->haveCost) {   // the cost-map goes through the left camera's maps too (stereo_kitti.cc:519-521)
    if (cost_stride < (size_t)src_width) return IVG_ERR_INVALID;
    if ((rc = prologue_in(h, h->qual.p, h->stageCost, n, costs, src_width, src_height, cost_stride, cost_frame_bytes, 1, 0))) return rc;
  }
  CK(cudaEventRecord(h->evIngest, h->stream));
  return IVG_OK;
}

static void maybe_speculate_stereo(ivg_extractor* h);

int ivg_run_batch(ivg_extractor* h) {
  if (!h || !h->shapeReady || h->curBatch < 1) return IVG_ERR_STATE;
  CK(cudaSetDevice(h->device));
  int rc = honour_wait(h);
  if (rc) return rc;
  CK(cudaStreamWaitEvent(h->stream, h->evD2H, 0));               // results of the previous run have been read
  CK(cudaStreamWaitEvent(h->stream, h->evD2Hs, 0));
  if ((rc = launch_extract(h))) return rc;
  CK(cudaEventRecord(h->evKernels, h->stream));
  h->runGen++;
  maybe_speculate_stereo(h);
  return IVG_OK;
}

static int ensure_out_host(ivg_extractor* h, size_t bytes) {
  if (h->outHostBytes >= bytes) return IVG_OK;
  if (h->outHost) { cudaFreeHost(h->outHost); h->outHost = nullptr; h->outHostBytes = 0; }
  CK(cudaHostAlloc(&h->outHost, bytes, cudaHostAllocPortable));
  h->outHostBytes = bytes;
  return IVG_OK;
}

int ivg_download_batch(ivg_extractor* h, ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
  if (!h || !h->haveResults) return IVG_ERR_STATE;
  if (cap < h->fs.kpCap) return IVG_ERR_CAPACITY;
  const int n = h->curBatch;
  const size_t k = h->fs.kpCap;
  CK(cudaStreamWaitEvent(h->copyOut, h->evKernels, 0));
  if (keypoints) CK(cudaMemcpy2DAsync(keypoints, (size_t)cap * 28, h->outKp.p, k * 28, k * 28, n, cudaMemcpyDeviceToHost, h->copyOut));
  if (descriptors) CK(cudaMemcpy2DAsync(descriptors, (size_t)cap * 32, h->outDesc.p, k * 32, k * 32, n, cudaMemcpyDeviceToHost, h->copyOut));
  if (n_out) CK(cudaMemcpyAsync(n_out, h->outN.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->copyOut));
  CK(cudaEventRecord(h->evD2H, h->copyOut));
  return IVG_OK;
}

int ivg_sync(ivg_extractor* h) {
  if (!h) return IVG_ERR_INVALID;
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->copyIn));
  CK(cudaStreamSynchronize(h->copyOut));
  return IVG_OK;
}

int ivg_share_stream(ivg_extractor* h, ivg_extractor* owner) {
  if (!h || !owner || h == owner || h->device != owner->device) return IVG_ERR_INVALID;
  CK(cudaStreamSynchronize(h->stream));
  drop_graph(h);
  if (h->ownsStream) cudaStreamDestroy(h->stream);
  h->stream = owner->stream;
  h->ownsStream = false;
  return IVG_OK;
}

// Results of the current run into the handle's pinned staging: one device-to-host copy when the run is the batch the buffers
// were sized for (records | descriptors | counts are one block), three otherwise.  Leaves evD2H recorded on copyOut.
static int staged_download_enqueue(ivg_extractor* h) {
  const int n = h->curBatch;
  const size_t k = h->fs.kpCap;
  const bool whole = n == h->maxBatch;
  const size_t descOff = whole ? h->outDescOff : (size_t)n * k * 28, nOff = whole ? h->outNOff : (size_t)n * k * 60;
  int rc = ensure_out_host(h, nOff + (size_t)n * sizeof(int));
  if (rc) return rc;
  uint8_t* stg = (uint8_t*)h->outHost;
  CK(cudaStreamWaitEvent(h->copyOut, h->evKernels, 0));
  if (whole) {
    CK(cudaMemcpyAsync(stg, h->outAll.p, nOff + (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->copyOut));
  } else {
    CK(cudaMemcpyAsync(stg, h->outKp.p, (size_t)n * k * 28, cudaMemcpyDeviceToHost, h->copyOut));
    CK(cudaMemcpyAsync(stg + descOff, h->outDesc.p, (size_t)n * k * 32, cudaMemcpyDeviceToHost, h->copyOut));
    CK(cudaMemcpyAsync(stg + nOff, h->outN.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->copyOut));
  }
  CK(cudaEventRecord(h->evD2H, h->copyOut));
  return IVG_OK;
}
// Waits for that copy and hands the caller the records each frame really produced.
static int staged_download_collect(ivg_extractor* h, ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
  const int n = h->curBatch;
  const size_t k = h->fs.kpCap;
  const bool whole = n == h->maxBatch;
  const size_t descOff = whole ? h->outDescOff : (size_t)n * k * 28, nOff = whole ? h->outNOff : (size_t)n * k * 60;
  const uint8_t* stg = (const uint8_t*)h->outHost;
  const int* cnt = reinterpret_cast<const int*>(stg + nOff);
  CK(cudaEventSynchronize(h->evD2H));      // results are on the host (=> the kernels and the upload before them are done); a
                                           // speculative matcher that the other eye's thread may have queued behind is not waited for
  for (int f = 0; f < n; ++f) {
    const int m = std::min(std::max(cnt[f], 0), (int)k);
    n_out[f] = cnt[f];
    std::memcpy(reinterpret_cast<uint8_t*>(keypoints) + (size_t)f * cap * 28, stg + (size_t)f * k * 28, (size_t)m * 28);
    std::memcpy(descriptors + (size_t)f * cap * 32, stg + descOff + (size_t)f * k * 32, (size_t)m * 32);
  }
  return IVG_OK;
}

int ivg_extract_batch(ivg_extractor* h, int n, const uint8_t* images, int width, int height, size_t stride,
                      size_t frame_bytes, const uint8_t* costs, size_t cost_stride, size_t cost_frame_bytes,
                      ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
  if (!h) return IVG_ERR_INVALID;
  if (cap < h->kpCap) return IVG_ERR_CAPACITY;
  int rc = ivg_upload_batch(h, n, images, width, height, stride, frame_bytes, costs, cost_stride, cost_frame_bytes);
  if (rc) return rc;
  if ((rc = ivg_run_batch(h))) return rc;
  // large pinned result buffers: straight DMA into them.  Small results (one frame at a time) and pageable buffers go through
  // the handle's pinned staging: one device-to-host copy for records, descriptors and counts, then a memcpy of what was produced
  // (a device-to-host copy straight into pageable memory goes through the driver's bounce buffers and blocks).
  const bool smallWhole = n == h->maxBatch && (size_t)n * h->fs.kpCap * 60 <= ((size_t)1 << 20);
  if (!keypoints || !descriptors || !n_out || (!smallWhole && is_pinned(keypoints))) {
    if ((rc = ivg_download_batch(h, keypoints, descriptors, cap, n_out))) return rc;
    return ivg_sync(h);
  }
  if ((rc = staged_download_enqueue(h))) return rc;
  return staged_download_collect(h, keypoints, descriptors, cap, n_out);
}

// Both eyes of one stereo frame and the matcher from ONE host thread: everything is queued back to back on the two handles'
// streams and the host waits once per result.  See include/ivslam_gpu.h.
int ivg_extract_stereo(ivg_extractor* left, ivg_extractor* right, const uint8_t* image_left, const uint8_t* image_right,
                       int width, int height, size_t stride, const uint8_t* cost_left, size_t cost_stride,
                       ivg_keypoint* kp_left, uint8_t* desc_left, int* n_left, ivg_keypoint* kp_right, uint8_t* desc_right, int* n_right,
                       float mbf, float maxD, float* uRight, float* depth, int cap) {
  if (!left || !right || left == right || !image_left || !image_right || width <= 0 || height <= 0) return IVG_ERR_INVALID;
  if (!kp_left || !desc_left || !n_left || !kp_right || !desc_right || !n_right || !uRight || !depth) return IVG_ERR_INVALID;
  if (cap < left->kpCap || cap < right->kpCap) return IVG_ERR_CAPACITY;
  const size_t fb = (size_t)height * stride;
  int rc;
  if ((rc = ivg_upload_batch(left, 1, image_left, width, height, stride, fb, cost_left, cost_stride, (size_t)height * cost_stride))) return rc;
  if ((rc = ivg_run_batch(left))) return rc;
  if ((rc = staged_download_enqueue(left))) return rc;   // before the matcher's results queue up on the same copy stream
  // Frame::Frame hands the same cost-map to both extraction threads (src/Frame.cc:116-117); it weights an eye only if that extractor
  // was built with introspection — the reference builds the right one without (src/Tracking.cc:182-183), so it normally stays home
  const uint8_t* cost_right = right->enableIntrospection ? cost_left : nullptr;
  if ((rc = ivg_upload_batch(right, 1, image_right, width, height, stride, fb, cost_right, cost_stride, (size_t)height * cost_stride))) return rc;
  if ((rc = ivg_run_batch(right))) return rc;          // a linked pair queues the matcher here (maybe_speculate_stereo)
  if ((rc = staged_download_enqueue(right))) return rc;
  if ((rc = ivg_stereo_match_batch(left, right, mbf, maxD, uRight, depth, cap, 1))) return rc;   // collects the queued matcher, or runs it and links the pair
  if ((rc = staged_download_collect(left, kp_left, desc_left, cap, n_left))) return rc;
  return staged_download_collect(right, kp_right, desc_right, cap, n_right);
}

int ivg_extract(ivg_extractor* h, const uint8_t* image, int width, int height, size_t stride,
                const uint8_t* cost, size_t cost_stride, ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
  if (!h || !n_out) return IVG_ERR_INVALID;
  *n_out = 0;
  if (!image || width <= 0 || height <= 0) return IVG_OK;     // empty image: silent return (src/ORBextractor.cc:1227-1228)
  return ivg_extract_batch(h, 1, image, width, height, stride, (size_t)height * stride, cost, cost_stride,
                           (size_t)height * cost_stride, keypoints, descriptors, cap, n_out);
}

int ivg_compute_pyramid(ivg_extractor* h, const uint8_t* image, int width, int height, size_t stride) {
  if (!h || !image) return IVG_ERR_INVALID;
  int rc = ivg_upload_batch(h, 1, image, width, height, stride, (size_t)height * stride, nullptr, 0, 0);
  if (rc) return rc;
  const FrameSet fs = active_fs(h);
  if ((rc = launch_pyramid(h, fs))) return rc;
  h->havePyramid = true;
  return ivg_sync(h);
}

int ivg_level_size(const ivg_extractor* h, int level, int* width, int* height) {
  if (!h || !h->shapeReady || level < 0 || level >= h->nlevels) return IVG_ERR_INVALID;
  if (width) *width = h->fs.lv[level].w;
  if (height) *height = h->fs.lv[level].h;
  return IVG_OK;
}

int ivg_get_pyramid_level(ivg_extractor* h, int index, int level, int which, uint8_t* dst, size_t dst_stride) {
  if (!h || !h->havePyramid || level < 0 || level >= h->nlevels || index < 0 || index >= h->curBatch || !dst) return IVG_ERR_STATE;
  const LevelDev& L = h->fs.lv[level];
  const uint8_t* base = which == 0 ? h->pyr.p : which == 1 ? h->blur.p : which == 2 ? h->qual.p : nullptr;
  if (!base || dst_stride < (size_t)L.w) return IVG_ERR_INVALID;
  if (which == 2 && !h->curWeighted) return IVG_ERR_STATE;
  CK(cudaMemcpy2DAsync(dst, dst_stride, base + (size_t)index * h->fs.planeBytes + L.planeOff, L.pitch, L.w, L.h,
                       cudaMemcpyDeviceToHost, h->stream));
  return ivg_sync(h);
}

int ivg_get_level_keypoints(ivg_extractor* h, int index, int level, float* x, float* y, float* response, int cap, int* n_out) {
  if (!h || !h->haveResults || level < 0 || level >= h->nlevels || index < 0 || index >= h->curBatch || !n_out) return IVG_ERR_STATE;
  const LevelDev& L = h->fs.lv[level];
  int cnt = 0;
  CK(cudaMemcpyAsync(&cnt, h->levelCount.p + (size_t)index * MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *n_out = cnt;
  if (cnt > cap) return IVG_ERR_CAPACITY;
  std::vector<uint2> tmp(cnt);
  if (cnt) {
    CK(cudaMemcpyAsync(tmp.data(), h->levelKp.p + (size_t)index * h->fs.kpCap + L.kpOff, cnt * sizeof(uint2), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  for (int i = 0; i < cnt; ++i) {
    float r; std::memcpy(&r, &tmp[i].x, 4);
    if (x) x[i] = (float)((tmp[i].y >> 8) & 0xFFF);
    if (y) y[i] = (float)(tmp[i].y >> 20);
    if (response) response[i] = r;
  }
  return IVG_OK;
}

// ---------------------------------------------------------------------------------------- stereo
static int stereo_launch(ivg_extractor* left, ivg_extractor* right, StereoArgs& A, int nPairs) {
  const FrameSet fs = active_fs(left);
  int rc;
  A.nLevels = left->nlevels;
  const size_t nBins = (size_t)A.nRows * A.nLevels;
  if ((nBins + 1) * sizeof(int) > 200 * 1024) return IVG_ERR_CAPACITY;
  if (right->indexValid && A.kpR == right->outKp.p && nPairs == right->curBatch && A.cap == right->fs.kpCap) {
    A.sorted = right->sortedR.p; A.rowStart = right->rowStart.p;      // built by the right handle at the end of its own run
  } else {
    if ((rc = left->sortedR.alloc((size_t)nPairs * A.cap)) || (rc = left->rowStart.alloc((size_t)nPairs * (nBins + 1)))) return rc;
    A.sorted = left->sortedR.p; A.rowStart = left->rowStart.p;
    ProfScope ps(left, IVG_K_STEREO);
    k_stereo_index<<<nPairs, one_frame_cfg(left, nPairs <= EAGER_INDEX_MAX_BATCH) ? 1024 : 256, (nBins + 1) * sizeof(int), left->stream>>>(A);
  }
  {
    ProfScope ps(left, IVG_K_STEREO);
    if (one_frame_cfg(left, (long long)nPairs * ((A.cap + SM_WARPS * SM_KP - 1) / (SM_WARPS * SM_KP)) <= 148))      // less than one CTA per SM: spread the keypoints wider
      k_stereo_match<SM_KP_LAT><<<dim3((A.cap + SM_WARPS * SM_KP_LAT - 1) / (SM_WARPS * SM_KP_LAT), nPairs), 32 * SM_WARPS, 0, left->stream>>>(fs, A);
    else
      k_stereo_match<SM_KP><<<dim3((A.cap + SM_WARPS * SM_KP - 1) / (SM_WARPS * SM_KP), nPairs), 32 * SM_WARPS, 0, left->stream>>>(fs, A);
  }
  { ProfScope ps(left, IVG_K_MEDIAN); k_stereo_median<<<nPairs, one_frame_cfg(left, nPairs <= EAGER_INDEX_MAX_BATCH) ? 1024 : 256, 0, left->stream>>>(A); }
  CK(cudaGetLastError());
  return IVG_OK;
}

// matcher kernels of the current results of (left, right) on left's stream, ordered after right's work; leaves copyOut
// of the left handle waiting for them
static int stereo_enqueue(ivg_extractor* left, ivg_extractor* right, float mbf, float maxD, int n, float* hostU = nullptr, float* hostD = nullptr) {
  CK(cudaEventRecord(right->evDone, right->stream));
  CK(cudaStreamWaitEvent(left->stream, right->evDone, 0));
  CK(cudaStreamWaitEvent(left->stream, left->evD2Hs, 0));        // previous uRight/depth have been read
  StereoArgs A{};
  A.kpL = left->outKp.p; A.descL = left->outDesc.p; A.nL = left->outN.p;
  A.kpR = right->outKp.p; A.descR = right->outDesc.p; A.nR = right->outN.p;
  A.pyrL = left->pyr.p; A.pyrR = right->pyr.p; A.planeBytes = left->fs.planeBytes;
  A.cap = left->fs.kpCap; A.nRows = left->fs.lv[0].h; A.mbf = mbf; A.maxD = maxD;
  A.uRight = left->uRight.p; A.depth = left->depth.p; A.sad = left->sad.p; A.bestDist = nullptr;
  A.hostU = hostU; A.hostD = hostD;
  int rc = stereo_launch(left, right, A, n);
  if (rc) return rc;
  left->haveStereo = true;
  if (n <= EAGER_INDEX_MAX_BATCH) right->eagerIndex = true;      // from its next run on, the right handle builds the index itself
  // the right handle must not overwrite its pyramids/keypoints before the matcher has read them
  CK(cudaEventRecord(left->evStereo, left->stream));
  CK(cudaEventRecord(right->evConsumed, left->stream));          // the right handle's own event: survives the left handle
  right->waitFor = right->evConsumed;
  CK(cudaStreamWaitEvent(left->copyOut, left->evStereo, 0));
  return IVG_OK;
}

// pinned staging of uRight | depth on the left handle (speculative and synchronous matcher results)
static int ensure_spec_host(ivg_extractor* L, size_t bytes) {
  if (L->specHostBytes >= bytes) return IVG_OK;
  CK(cudaStreamSynchronize(L->copyOut));                         // nothing may still be landing in the old buffer
  if (L->specHost) { cudaFreeHost(L->specHost); L->specHost = nullptr; L->specHostBytes = 0; }
  CK(cudaHostAlloc(&L->specHost, bytes, cudaHostAllocPortable));
  L->specHostBytes = bytes;
  return IVG_OK;
}

static bool stereo_compatible(const ivg_extractor* left, const ivg_extractor* right) {
  return left->haveResults && right->haveResults && left->device == right->device && left->W == right->W && left->H == right->H &&
         left->nlevels == right->nlevels && left->curBatch == right->curBatch && left->fs.planeBytes == right->fs.planeBytes &&
         left->fs.kpCap == right->fs.kpCap;
}

// called at the end of every run: the second of two linked handles to get here launches the matcher speculatively
static void maybe_speculate_stereo(ivg_extractor* h) {
  std::shared_ptr<StereoLink> link = h->link;
  if (!link || h->profile || h->curBatch > EAGER_INDEX_MAX_BATCH) return;
  std::lock_guard<std::mutex> lk(link->mu);
  ivg_extractor *L = link->left, *R = link->right;
  if (!L || !R || (h != L && h != R)) return;
  (h == L ? link->genL : link->genR) = h->runGen;
  if (!(link->genL > link->specGenL && link->genR > link->specGenR)) return;      // one side has no new run yet
  if (link->genL != L->runGen || link->genR != R->runGen || L->profile || R->profile || !stereo_compatible(L, R)) return;
  const int n = L->curBatch;
  const size_t k = L->fs.kpCap, bytes = 2 * (size_t)n * k * 4;
  if (ensure_spec_host(L, bytes) != IVG_OK) { cudaGetLastError(); return; }
  link->specValid = false;
  // the matcher's last kernel writes uRight | depth straight into the pinned staging (zero copy): the result is on the host when the
  // kernel is done, without a device-to-host copy on a second stream behind it
  float* stg = (float*)L->specHost;
  if (stereo_enqueue(L, R, link->mbf, link->maxD, n, stg, stg + (size_t)n * k) != IVG_OK) return;
  if (cudaEventRecord(L->evD2Hs, L->stream) != cudaSuccess) { cudaGetLastError(); return; }
  link->specGenL = link->genL; link->specGenR = link->genR; link->specN = n; link->specValid = true;
}

int ivg_stereo_match_batch(ivg_extractor* left, ivg_extractor* right, float mbf, float maxD,
                           float* uRight, float* depth, int cap, int sync) {
  if (!left || !right || !left->haveResults || !right->haveResults) return IVG_ERR_STATE;
  if (!stereo_compatible(left, right)) return IVG_ERR_STATE;
  if (cap < left->fs.kpCap) return IVG_ERR_CAPACITY;
  CK(cudaSetDevice(left->device));
  const int n = left->curBatch;
  const size_t k = left->fs.kpCap;
  int rc;
  const bool small = n <= EAGER_INDEX_MAX_BATCH && sync && uRight && depth;
  if (small && left->link && left->link == right->link) {
    // the matcher may already be running (or done): launched by whichever extractor thread finished second
    StereoLink& K = *left->link;
    std::lock_guard<std::mutex> lk(K.mu);
    if (K.specValid && K.left == left && K.right == right && K.specGenL == left->runGen && K.specGenR == right->runGen &&
        K.mbf == mbf && K.maxD == maxD && K.specN == n) {
      CK(cudaEventSynchronize(left->evD2Hs));
      const float* stg = (const float*)left->specHost;
      for (int f = 0; f < n; ++f) {
        std::memcpy(uRight + (size_t)f * cap, stg + (size_t)f * k, k * 4);
        std::memcpy(depth + (size_t)f * cap, stg + (size_t)n * k + (size_t)f * k, k * 4);
      }
      return IVG_OK;
    }
    K.specValid = false;
  }
  if ((rc = stereo_enqueue(left, right, mbf, maxD, n))) return rc;
  if (small) {
    // remember the pairing and the calibration: the next frame's matcher is launched by the extractor threads themselves
    std::shared_ptr<StereoLink> link = left->link;
    if (!link || link != right->link || link->left != left || link->right != right) {
      link = std::make_shared<StereoLink>();
      link->left = left; link->right = right;
      left->link = link; right->link = link;
    }
    std::lock_guard<std::mutex> lk(link->mu);
    link->mbf = mbf; link->maxD = maxD; link->specValid = false;
    link->genL = link->specGenL = left->runGen; link->genR = link->specGenR = right->runGen;
  }
  if (sync && uRight && depth && !is_pinned(uRight)) {           // pageable mvuRight / mvDepth: through the pinned staging (see ivg_extract_batch)
    // (its own staging, not outHost: the keypoint records of ivg_extract_stereo may still be on their way into that one)
    if ((rc = ensure_spec_host(left, 2 * (size_t)n * k * 4))) return rc;
    float* stg = (float*)left->specHost;
    if (n == left->maxBatch) {
      CK(cudaMemcpyAsync(stg, left->udAll.p, 2 * (size_t)n * k * 4, cudaMemcpyDeviceToHost, left->copyOut));
    } else {
      CK(cudaMemcpyAsync(stg, left->uRight.p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, left->copyOut));
      CK(cudaMemcpyAsync(stg + (size_t)n * k, left->depth.p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, left->copyOut));
    }
    CK(cudaEventRecord(left->evD2Hs, left->copyOut));
    if ((rc = ivg_sync(left))) return rc;
    for (int f = 0; f < n; ++f) {
      std::memcpy(uRight + (size_t)f * cap, stg + (size_t)f * k, k * 4);
      std::memcpy(depth + (size_t)f * cap, stg + (size_t)n * k + (size_t)f * k, k * 4);
    }
    return IVG_OK;
  }
  if (uRight) CK(cudaMemcpy2DAsync(uRight, (size_t)cap * 4, left->uRight.p, k * 4, k * 4, n, cudaMemcpyDeviceToHost, left->copyOut));
  if (depth) CK(cudaMemcpy2DAsync(depth, (size_t)cap * 4, left->depth.p, k * 4, k * 4, n, cudaMemcpyDeviceToHost, left->copyOut));
  CK(cudaEventRecord(left->evD2Hs, left->copyOut));
  if (sync) return ivg_sync(left);
  return IVG_OK;
}

int ivg_stereo_match(ivg_extractor* left, ivg_extractor* right, float mbf, float maxD, float* uRight, float* depth, int cap) {
  if (!left || !right) return IVG_ERR_INVALID;
  if (left->curBatch != 1 || right->curBatch != 1) return IVG_ERR_STATE;
  return ivg_stereo_match_batch(left, right, mbf, maxD, uRight, depth, cap, 1);
}

int ivg_stereo_match_keypoints(ivg_extractor* left, ivg_extractor* right, const ivg_keypoint* kL, int nL, const uint8_t* dL,
                               const ivg_keypoint* kR, int nR, const uint8_t* dR, float mbf, float maxD, float* uRight, float* depth) {
  if (!left || !right || !left->havePyramid || !right->havePyramid) return IVG_ERR_STATE;
  if (left->device != right->device || left->W != right->W || left->H != right->H || left->nlevels != right->nlevels ||
      left->fs.planeBytes != right->fs.planeBytes)
    return IVG_ERR_STATE;
  if (nL < 0 || nR < 0 || nR > 65535 || (nL && (!kL || !dL)) || (nR && (!kR || !dR))) return IVG_ERR_INVALID;
  if (nL == 0) return IVG_OK;
  CK(cudaSetDevice(left->device));
  int rc;
  const int cap = std::max(nL, std::max(nR, 1));
  if ((rc = left->extKpL.alloc((size_t)cap * 28)) || (rc = left->extDescL.alloc((size_t)cap * 32)) ||
      (rc = left->extKpR.alloc((size_t)cap * 28)) || (rc = left->extDescR.alloc((size_t)cap * 32)) || (rc = left->nExt.alloc(2)))
    return rc;
  DevBuf<float>&u = left->extU, &d = left->extD; DevBuf<int>& s = left->extS;
  if ((rc = u.alloc(cap)) || (rc = d.alloc(cap)) || (rc = s.alloc(cap))) return rc;
  CK(cudaEventRecord(right->evDone, right->stream));
  CK(cudaStreamWaitEvent(left->stream, right->evDone, 0));
  cudaStream_t st = left->stream;
  const int cnt[2] = {nL, nR};
  CK(cudaMemcpyAsync(left->nExt.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(left->extKpL.p, kL, (size_t)nL * 28, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(left->extDescL.p, dL, (size_t)nL * 32, cudaMemcpyHostToDevice, st));
  if (nR) {
    CK(cudaMemcpyAsync(left->extKpR.p, kR, (size_t)nR * 28, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(left->extDescR.p, dR, (size_t)nR * 32, cudaMemcpyHostToDevice, st));
  }
  StereoArgs A{};
  A.kpL = left->extKpL.p; A.descL = left->extDescL.p; A.nL = left->nExt.p;
  A.kpR = left->extKpR.p; A.descR = left->extDescR.p; A.nR = left->nExt.p + 1;
  A.pyrL = left->pyr.p; A.pyrR = right->pyr.p; A.planeBytes = left->fs.planeBytes;
  A.cap = cap; A.nRows = left->fs.lv[0].h; A.mbf = mbf; A.maxD = maxD;
  A.uRight = u.p; A.depth = d.p; A.sad = s.p; A.bestDist = nullptr;
  rc = stereo_launch(left, right, A, 1);
  if (!rc && uRight) { cudaError_t e = cudaMemcpyAsync(uRight, u.p, (size_t)nL * 4, cudaMemcpyDeviceToHost, st); if (e != cudaSuccess) rc = IVG_ERR_CUDA; }
  if (!rc && depth) { cudaError_t e = cudaMemcpyAsync(depth, d.p, (size_t)nL * 4, cudaMemcpyDeviceToHost, st); if (e != cudaSuccess) rc = IVG_ERR_CUDA; }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { g_cuda_err = cudaGetErrorString(e); return IVG_ERR_CUDA; }
  return rc;
}

// ---------------------------------------------------------------------------------------- N1: rest of the Frame ctor
int ivg_frame_postprocess_batch(ivg_extractor* h, float minX, float maxX, float minY, float maxY,
                                float* keyQualScore, int* gridStart, int* gridIndices, int cap, int sync) {
  if (!h || !h->haveResults) return IVG_ERR_STATE;
  if (!(maxX > minX) || !(maxY > minY)) return IVG_ERR_INVALID;
  if (cap < h->fs.kpCap) return IVG_ERR_CAPACITY;
  CK(cudaSetDevice(h->device));
  const int n = h->curBatch;
  const int NC = GRID_COLS * GRID_ROWS;
  int rc;
  if ((rc = h->kpQual.alloc((size_t)h->maxBatch * h->fs.kpCap)) || (rc = h->gridIdx.alloc((size_t)h->maxBatch * h->fs.kpCap)) ||
      (rc = h->gridStart.alloc((size_t)h->maxBatch * (NC + 1))))
    return rc;
  FramePostArgs A{};
  A.kp = h->outKp.p; A.n = h->outN.p;
  A.cost = h->haveCost ? h->qual.p : nullptr;             // level 0 of the cost-map plane = the map itself
  A.planeBytes = h->fs.planeBytes; A.costPitch = h->fs.lv[0].pitch; A.cap = h->fs.kpCap;
  A.minX = minX; A.minY = minY;
  A.invW = (float)GRID_COLS / (maxX - minX);               // mfGridElementWidthInv  (src/Frame.cc:213)
  A.invH = (float)GRID_ROWS / (maxY - minY);               // mfGridElementHeightInv (src/Frame.cc:216)
  A.qual = h->kpQual.p; A.gridStart = h->gridStart.p; A.gridIdx = h->gridIdx.p;
  CK(cudaStreamWaitEvent(h->stream, h->evD2H, 0));
  k_frame_post<<<n, 256, 0, h->stream>>>(A); h->launches++;
  h->haveGrid = true;
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->evKernels, h->stream));
  CK(cudaStreamWaitEvent(h->copyOut, h->evKernels, 0));
  const size_t k = h->fs.kpCap;
  if (keyQualScore) CK(cudaMemcpy2DAsync(keyQualScore, (size_t)cap * 4, h->kpQual.p, k * 4, k * 4, n, cudaMemcpyDeviceToHost, h->copyOut));
  if (gridIndices) CK(cudaMemcpy2DAsync(gridIndices, (size_t)cap * 4, h->gridIdx.p, k * 4, k * 4, n, cudaMemcpyDeviceToHost, h->copyOut));
  if (gridStart) CK(cudaMemcpyAsync(gridStart, h->gridStart.p, (size_t)n * (NC + 1) * 4, cudaMemcpyDeviceToHost, h->copyOut));
  CK(cudaEventRecord(h->evD2H, h->copyOut));
  if (sync) return ivg_sync(h);
  return IVG_OK;
}

// ---------------------------------------------------------------------------------------- measurement helpers
int ivg_timer_start(ivg_extractor* h) { if (!h) return IVG_ERR_INVALID; CK(cudaEventRecord(h->evT0, h->stream)); return IVG_OK; }
int ivg_timer_stop(ivg_extractor* h) { if (!h) return IVG_ERR_INVALID; CK(cudaEventRecord(h->evT1, h->stream)); return IVG_OK; }
int ivg_timer_elapsed_ms(ivg_extractor* h, float* ms) {
  if (!h || !ms) return IVG_ERR_INVALID;
  CK(cudaEventSynchronize(h->evT1));
  CK(cudaEventElapsedTime(ms, h->evT0, h->evT1));
  return IVG_OK;
}
long long ivg_launch_count(const ivg_extractor* h) { return h ? h->launches : 0; }
int ivg_host_alloc(void** ptr, size_t bytes) { if (!ptr) return IVG_ERR_INVALID; CK(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable)); return IVG_OK; }
int ivg_host_free(void* ptr) { CK(cudaFreeHost(ptr)); return IVG_OK; }
int ivg_flush_l2(ivg_extractor* h, size_t bytes) {
  if (!h) return IVG_ERR_INVALID;
  static thread_local DevBuf<uint8_t> scratch;
  int rc = scratch.alloc(bytes);
  if (rc) return rc;
  CK(cudaMemsetAsync(scratch.p, 0x5a, bytes, h->stream));
  return IVG_OK;
}
int ivg_debug_force_config(ivg_extractor* h, int mode) {
  if (!h || mode < 0 || mode > 2) return IVG_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graph(h);                       // a captured graph holds the kernels of the old configuration
  h->forceCfg = mode;
  return IVG_OK;
}
static int debug_nth_element(int device, const uint32_t* keys, int n, int nth, uint32_t* order, int threads);
int ivg_debug_nth_element(int device, const uint32_t* keys, int n, int nth, uint32_t* order) { return debug_nth_element(device, keys, n, nth, order, 32); }
int ivg_debug_nth_element_block(int device, const uint32_t* keys, int n, int nth, uint32_t* order) { return debug_nth_element(device, keys, n, nth, order, 1024); }
static int debug_nth_element(int device, const uint32_t* keys, int n, int nth, uint32_t* order, int threads) {
  if (!keys || !order || n < 1 || n > 4000 || nth < 0 || nth >= n) return IVG_ERR_INVALID;
  int rc = ivg_device_info(device, nullptr, 0, nullptr, nullptr);
  if (rc) return rc;
  CK(cudaSetDevice(device));
  DevBuf<uint32_t> dk, dord;
  if ((rc = dk.alloc(n)) || (rc = dord.alloc(n))) return rc;
  CK(cudaMemcpy(dk.p, keys, (size_t)n * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)n * 8 + (size_t)n * 4 + 16;
  CK(cudaFuncSetAttribute(k_debug_nth_element, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  k_debug_nth_element<<<1, threads, smem>>>(dk.p, n, nth, dord.p);
  CK(cudaGetLastError());
  CK(cudaMemcpy(order, dord.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  dk.release(); dord.release();
  return IVG_OK;
}

int ivg_profile_enable(ivg_extractor* h, int enable) {
  if (!h) return IVG_ERR_INVALID;
  CK(cudaStreamSynchronize(h->stream));
  h->profile = enable != 0;
  h->profUsed = 0;
  for (int k = 0; k < IVG_NUM_KERNELS; ++k) { h->profMs[k] = 0; h->profCnt[k] = 0; }
  return IVG_OK;
}
int ivg_profile_read(ivg_extractor* h, double* ms, long long* launches) {
  if (!h) return IVG_ERR_INVALID;
  CK(cudaStreamSynchronize(h->stream));
  for (size_t s = 0; s + 1 < h->profUsed; s += 2) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, h->profEv[s], h->profEv[s + 1]));
    const int k = h->profKid[s / 2];
    h->profMs[k] += t; h->profCnt[k]++;
  }
  h->profUsed = 0;
  for (int k = 0; k < IVG_NUM_KERNELS; ++k) { if (ms) ms[k] = h->profMs[k]; if (launches) launches[k] = h->profCnt[k]; }
  return IVG_OK;
}
int ivg_set_graph_mode(ivg_extractor* h, int enable) {
  if (!h) return IVG_ERR_INVALID;
  h->graphMode = enable != 0;
  if (!h->graphMode) drop_graph(h);
  return IVG_OK;
}

}  // extern "C"

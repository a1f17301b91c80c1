/* ivslam_gpu.h — C ABI of the B200-native (sm_100a) IV-SLAM stereo front-end.
 *
 * This is the drop-in boundary for the reference's data-parallel hot path.  The reference
 * (ut-amrl/IV_SLAM) has no FFI layer; the seam is two C++ signatures inside libORB_SLAM2.so:
 *
 *   ORB_SLAM2::ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels,
 *                                         int iniThFAST, int minThFAST, bool enableIntrospection)
 *                                              introspective_ORB_SLAM/include/ORBextractor.h:57-58
 *   void ORBextractor::operator()(cv::InputArray image, cv::InputArray mask,
 *                                 std::vector<cv::KeyPoint>&, cv::OutputArray descriptors)
 *                                              include/ORBextractor.h:65-67, src/ORBextractor.cc:1224-1296
 *   void Frame::ComputeStereoMatches()         include/Frame.h:163,  src/Frame.cc:758-932
 *
 * Every entry point below names the reference interface it replaces.  shim/ holds the C++
 * classes with the reference's exact signatures built on this ABI.
 *
 * Conventions: plain pointers and sizes only; int status return (0 = IVG_OK, negative = error,
 * text from ivg_strerror); no exceptions cross the ABI; caller owns all host buffers.
 * One handle owns one CUDA stream and its device workspace.  Different handles may be driven
 * concurrently from different host threads (the reference runs the left and right extractor
 * on two std::threads, src/Frame.cc:115-125); a single handle is not re-entrant.
 * There is no CPU fallback: every call fails with IVG_ERR_CUDA / IVG_ERR_NO_DEVICE when no
 * sm_100-class GPU is usable.
 */
#ifndef IVSLAM_GPU_H_
#define IVSLAM_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVG_OK 0
#define IVG_ERR_INVALID (-1)    /* bad argument / handle state                                  */
#define IVG_ERR_GEOMETRY (-2)   /* image too small for the reference's cell grid (it would divide by zero) */
#define IVG_ERR_CAPACITY (-3)   /* caller buffer or reserved batch too small                       */
#define IVG_ERR_CUDA (-4)       /* CUDA runtime error (see ivg_last_cuda_error)                    */
#define IVG_ERR_NO_DEVICE (-5)  /* no CUDA device / not an sm_100 part                             */
#define IVG_ERR_STATE (-6)      /* call order: nothing extracted yet, batch sizes differ, ...      */

typedef struct ivg_extractor ivg_extractor;

/* Same memory layout as cv::KeyPoint (28 bytes) so a shim can memcpy into std::vector<cv::KeyPoint>. */
typedef struct ivg_keypoint {
  float x, y;        /* pt, level-0 pixels (level coordinates * scale factor of the octave)        */
  float size;        /* (int)(31 * scale[octave])                     src/ORBextractor.cc:1138    */
  float angle;       /* degrees in [0,360), cv::fastAtan2 of the intensity centroid  :78-105       */
  float response;    /* FAST score (x introspection weight when a cost-map is given) :1058-1080    */
  int32_t octave;
  int32_t class_id;  /* always -1                                                                  */
} ivg_keypoint;

const char* ivg_strerror(int status);
const char* ivg_last_cuda_error(void);
/* Library/device probe: returns IVG_OK and fills name/sm (e.g. 100) when `device` is usable. */
int ivg_device_info(int device, char* name, int name_cap, int* sm, int* sm_count);

/* ---- ORBextractor::ORBextractor (include/ORBextractor.h:57-58, src/ORBextractor.cc:411-476) ---- */
int ivg_extractor_create(ivg_extractor** out, int device, int nfeatures, float scaleFactor, int nlevels,
                         int iniThFAST, int minThFAST, int enableIntrospection);
void ivg_extractor_destroy(ivg_extractor* h);
/* Keypoint-selection path.  mode 0 (default) = ComputeKeyPointsOld, the path the reference actually runs
 * (src/ORBextractor.cc:1248).  mode 1 = ComputeKeyPointsOctTree + DistributeOctTree (src/ORBextractor.cc:771-878,
 * :545-769), which the reference compiles but never calls (:1247 is commented out).  In mode 1 the reference's one
 * nondeterministic tie-break (sorting nodes by heap address, :690) is replaced by creation order; a level may emit up
 * to 3 keypoints more than its budget, as in the reference.  Changes ivg_max_keypoints(); call before extracting. */
int ivg_extractor_set_mode(ivg_extractor* h, int mode);

/* Pre-allocates the device workspace for `max_batch` images of width x height (otherwise done lazily on the
 * first call and whenever the shape or batch grows). */
int ivg_extractor_reserve(ivg_extractor* h, int width, int height, int max_batch);

/* GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:69-91).  which: 0 scale, 1 inverse scale, 2 sigma^2,
 * 3 inverse sigma^2.  out must hold nlevels floats. */
int ivg_get_levels(const ivg_extractor* h);
float ivg_get_scale_factor(const ivg_extractor* h);
int ivg_get_scale_table(const ivg_extractor* h, int which, float* out);
int ivg_get_features_per_level(const ivg_extractor* h, int* out);
int ivg_max_keypoints(const ivg_extractor* h);   /* capacity one image can produce = sum of features per level */

/* ---- ORBextractor::operator() (src/ORBextractor.cc:1224-1296), one image ----
 * image: 8-bit gray, `stride` bytes per row.  cost: the IV-SLAM cost-map ("mask" argument), same size, or NULL.
 * The cost-map is used only when the handle was created with enableIntrospection (src/ORBextractor.cc:1231).
 * Writes *n_out keypoints (reference order: level-major, cell row-major, nth_element permutation) and
 * n_out x 32 descriptor bytes.  An empty image (NULL / zero size) returns IVG_OK with *n_out = 0 (:1227-1228).
 * Synchronous: results are in the host buffers on return. */
int ivg_extract(ivg_extractor* h, const uint8_t* image, int width, int height, size_t stride,
                const uint8_t* cost, size_t cost_stride,
                ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out);

/* ---- both eyes of one stereo frame and the matcher in ONE call from ONE thread ----
 * What Frame::Frame does with two std::threads (src/Frame.cc:115-125: ExtractORB on the left and right image) followed by
 * ComputeStereoMatches (:127).  Uploads, kernels and downloads of the two handles are queued back to back on their streams, the
 * matcher right behind them, and the host waits once per result: no thread creation, no two threads contending for the driver.
 * Outputs as ivg_extract (per eye) and ivg_stereo_match (uRight / depth: cap floats, -1 where there is no match).
 * cost_left: the cost-map both ExtractORBWeighted threads receive (src/Frame.cc:116-117) or NULL; it weights an eye only if that
 * handle was created with enableIntrospection — the reference creates the right one without (src/Tracking.cc:182-183). */
int ivg_extract_stereo(ivg_extractor* left, ivg_extractor* right, const uint8_t* image_left, const uint8_t* image_right,
                       int width, int height, size_t stride, const uint8_t* cost_left, size_t cost_stride,
                       ivg_keypoint* kp_left, uint8_t* desc_left, int* n_left,
                       ivg_keypoint* kp_right, uint8_t* desc_right, int* n_right,
                       float mbf, float maxD, float* uRight, float* depth, int cap);

/* ---- the same for a batch of n equally-shaped images (frame-parallel data path) ----
 * images: n frames, frame f starts at images + f*frame_bytes, rows `stride` bytes apart.  costs likewise or NULL.
 * keypoints: n*cap records, descriptors: n*cap*32 bytes, n_out: n ints.  Frame f writes at f*cap.
 * The three phases are exposed separately so a caller can keep inputs resident or overlap copies:
 *   ivg_upload_batch  (async H2D on the handle's stream; host memory should be pinned for real overlap)
 *   ivg_run_batch     (async, kernels only, inputs = whatever was uploaded / written through ivg_device_input)
 *   ivg_download_batch(async D2H of keypoints/descriptors/counts into the given buffers)
 *   ivg_sync          (blocks until the stream is idle; reports deferred CUDA errors)
 * ivg_extract_batch = upload + run + download + sync. */
int ivg_extract_batch(ivg_extractor* h, int n, const uint8_t* images, int width, int height, size_t stride,
                      size_t frame_bytes, const uint8_t* costs, size_t cost_stride, size_t cost_frame_bytes,
                      ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out);
int ivg_upload_batch(ivg_extractor* h, int n, const uint8_t* images, int width, int height, size_t stride,
                     size_t frame_bytes, const uint8_t* costs, size_t cost_stride, size_t cost_frame_bytes);
/* Same as ivg_upload_batch for frames that already live in DEVICE memory on the handle's device (the
 * introspection CNN's cost-map is produced on the GPU; this avoids the GPU->CPU->GPU round trip of the reference,
 * Examples/Stereo/stereo_kitti.cc:494-521).  Contiguous, 4-byte aligned frames are read in place by the ingest kernel.
 * The caller guarantees the producer has finished writing (or was enqueued on a stream this handle is ordered after). */
int ivg_upload_batch_device(ivg_extractor* h, int n, const uint8_t* d_images, int width, int height, size_t stride,
                            size_t frame_bytes, const uint8_t* d_costs, size_t cost_stride, size_t cost_frame_bytes);
/* The same with the cost-maps as the introspection CNN emits them: FLOAT frames in device memory (n frames of height rows,
 * cost_stride_floats apart).  The conversion the example driver asks libtorch for before it copies the map to the host,
 * `(cost_img * 255.0).to(torch::kByte)` (Examples/Stereo/stereo_kitti.cc:513-514: float multiply, truncation toward zero,
 * modulo 256), happens on the way into the cost-map plane. */
int ivg_upload_batch_device_cost_f32(ivg_extractor* h, int n, const uint8_t* d_images, int width, int height, size_t stride,
                                     size_t frame_bytes, const float* d_costs, size_t cost_stride_floats, size_t cost_frame_floats);
/* SURVEY §8(f) N4 — the input prologue, fused into the upload.
 * ivg_set_rectify_maps: the CV_32FC1 maps of cv::initUndistortRectifyMap (Examples/Stereo/stereo_kitti.cc:284-343,
 *   stereo_euroc.cc:247-254), uploaded once; width x height is the rectified (output) size.  NULL maps clear them.
 * ivg_upload_batch_raw: frames as they come from the camera / image file — 1, 3 or 4 interleaved 8-bit channels,
 *   rgb_order != 0 when R comes first (Tracking::mbRGB).  On the device each frame goes through
 *   cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) when maps are set (stereo_kitti.cc:463-464) and then through
 *   cvtColor(.., CV_{BGR,RGB,BGRA,RGBA}2GRAY) (src/Tracking.cc:278-294) into pyramid level 0; the optional cost-maps
 *   (1 channel, source size) are remapped with the same maps (stereo_kitti.cc:519-521).  Results are bit-identical to
 *   OpenCV 4.13's fixed-point remap / cvtColor.  Then ivg_run_batch / ivg_download_batch as usual. */
int ivg_set_rectify_maps(ivg_extractor* h, const float* mapx, const float* mapy, int width, int height, size_t stride_floats);
int ivg_upload_batch_raw(ivg_extractor* h, int n, const uint8_t* frames, int src_width, int src_height, size_t stride, size_t frame_bytes,
                         int channels, int rgb_order, const uint8_t* costs, size_t cost_stride, size_t cost_frame_bytes);
int ivg_run_batch(ivg_extractor* h);
int ivg_download_batch(ivg_extractor* h, ivg_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out);
int ivg_sync(ivg_extractor* h);
/* Makes `h` launch its kernels on `owner`'s stream (same device; `owner` must outlive `h`).  Handles that share a
 * stream never run kernels concurrently — measured ~20 % faster at large batches than letting the left and right
 * extraction overlap — while their H2D/D2H copies, which use per-handle copy streams, still overlap the kernels. */
int ivg_share_stream(ivg_extractor* h, ivg_extractor* owner);

/* Device-side level-0 input plane of frame `index` (pitch in *pitch): lets a producer already on the GPU (e.g. the
 * introspection CNN's cost-map) write inputs without a host round trip.  which: 0 image, 2 cost-map.
 * Call ivg_set_batch first to declare how many frames / whether cost-maps are present. */
int ivg_set_batch(ivg_extractor* h, int n, int width, int height, int with_cost);
int ivg_device_input(ivg_extractor* h, int index, int which, void** dev_ptr, size_t* pitch);

/* ---- mvImagePyramid / mvQualityImagePyramid (include/ORBextractor.h:91-92; read by src/Frame.cc:765,855,867,872) ----
 * Copies one level of frame `index` of the last batch to host.  which: 0 image pyramid, 1 blurred level
 * (the GaussianBlur working copy, src/ORBextractor.cc:1276-1277), 2 quality (cost-map) pyramid. */
int ivg_level_size(const ivg_extractor* h, int level, int* width, int* height);
int ivg_get_pyramid_level(ivg_extractor* h, int index, int level, int which, uint8_t* dst, size_t dst_stride);
/* Per-level keypoints in level coordinates before scaling (x, y, response), for stage-wise parity tests. */
int ivg_get_level_keypoints(ivg_extractor* h, int index, int level, float* x, float* y, float* response, int cap, int* n_out);

/* ---- Frame::ComputeStereoMatches (src/Frame.cc:758-932) ----
 * Uses what the two handles hold on the device after their last extract/run: both image pyramids, keypoints and
 * descriptors (frame f of `left` is matched against frame f of `right`).  mbf = Camera.bf; maxD = the disparity
 * limit the reference derives as mbf/mb (it reads mb before assigning it; nominal value = fx).
 * uRight/depth: n*cap floats each, -1 where there is no match (mvuRight / mvDepth).  No surviving match = no-op
 * (the reference indexes an empty vector there).  Runs on left's stream after right's work completes.
 * ivg_stereo_match = frame 0 only, synchronous.  The batch form is async (ivg_sync(left) to wait) unless sync != 0. */
int ivg_stereo_match(ivg_extractor* left, ivg_extractor* right, float mbf, float maxD,
                     float* uRight, float* depth, int cap);
int ivg_stereo_match_batch(ivg_extractor* left, ivg_extractor* right, float mbf, float maxD,
                           float* uRight, float* depth, int cap, int sync);
/* Same matcher on caller-supplied keypoints/descriptors (what Frame holds in mvKeys/mvKeysRight/mDescriptors*),
 * against the pyramids of frame 0 resident in the two handles.  Synchronous. */
int ivg_stereo_match_keypoints(ivg_extractor* left, ivg_extractor* right,
                               const ivg_keypoint* kL, int nL, const uint8_t* dL,
                               const ivg_keypoint* kR, int nR, const uint8_t* dR,
                               float mbf, float maxD, float* uRight, float* depth);
/* Pyramid only (no detection) for frame 0: stages mvImagePyramid for ivg_stereo_match_keypoints. Synchronous. */
int ivg_compute_pyramid(ivg_extractor* h, const uint8_t* image, int width, int height, size_t stride);

/* ---- N1 (next row): the keypoint loops of the stereo Frame constructor after extraction ----
 * mvKeyQualScore (src/Frame.cc:128-143), UndistortKeyPoints for rectified input (identity, :696-700), AssignFeaturesToGrid
 * + PosInGrid (:415-430, :670-680) for every frame of the last batch, on the keypoints still on the device.
 * (minX, maxX, minY, maxY) = mnMinX.. of ComputeImageBounds (:728-756; 0, cols, 0, rows without distortion).
 * keyQualScore: frames*cap floats (1.0 when the batch had no cost-map; the cost-map is used whenever one was uploaded,
 * with or without introspection, like the reference).  Grid as CSR per frame: gridStart has 64*48+1 entries, cell =
 * col*48 + row (mGrid[col][row]); gridIndices lists keypoint indices, ascending inside a cell.  Distorted input
 * (k1 != 0) is not supported.  Async unless sync != 0. */
int ivg_frame_postprocess_batch(ivg_extractor* h, float minX, float maxX, float minY, float maxY,
                                float* keyQualScore, int* gridStart, int* gridIndices, int cap, int sync);

/* ---- N2 (next row): the per-frame Hamming-search consumers in Track() ----
 * Both work on frame `index` of the handle's last batch (keypoints, descriptors on the device; mvuRight if
 * ivg_stereo_match_batch ran, else treated as -1) and need the 64x48 grid of ivg_frame_postprocess_batch (call it first,
 * with the same image bounds).  MapPoint/Frame objects are flattened by the caller; flags[i] bit0 = the point takes part,
 * bit1 = pMP->Observations() > 0 (it then blocks the keypoint it takes for all later points, like the sequential loops
 * of the reference).  match[i2] = index of the point assigned to current keypoint i2 by this call, or -1; *nmatches is
 * the function's return value.  cur_blocked[i2] != 0 marks keypoints whose mvpMapPoints entry has observations on
 * entry (may be NULL).  Synchronous.
 *
 * ivg_search_by_projection_last = ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
 *   (src/ORBmatcher.cc:1372-1519).  flags bit0 = LastFrame.mvpMapPoints[i] && !mvbOutlier[i]; world_pos = GetWorldPos();
 *   desc = GetDescriptor(); octave/angle = LastFrame.mvKeys[i].octave / mvKeysUn[i].angle; Rcw (row-major 3x3), tcw from
 *   CurrentFrame.mTcw; mode 0 = neither forward nor backward (or bMono), 1 = bForward, 2 = bBackward (:1394-1395);
 *   check_orientation = ORBmatcher::mbCheckOrientation.
 * ivg_search_by_projection_map = ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th) (:45-133).
 *   flags bit0 = mbTrackInView && !isBad(); proj = (mTrackProjX, mTrackProjY, mTrackProjXR); view_cos = mTrackViewCos;
 *   level = mnTrackScaleLevel; nnratio = ORBmatcher::mfNNratio. */
int ivg_search_by_projection_last(ivg_extractor* cur, int index, int n, const float* world_pos, const uint8_t* desc, const int* octave,
                                  const float* angle, const uint8_t* flags, const float* Rcw, const float* tcw, float fx, float fy, float cx,
                                  float cy, float mbf, float minX, float maxX, float minY, float maxY, int mode, float th,
                                  int check_orientation, int* match, int cap, int* nmatches);
int ivg_search_by_projection_map(ivg_extractor* cur, int index, int n, const float* proj, const float* view_cos, const int* level,
                                 const uint8_t* desc, const uint8_t* flags, const uint8_t* cur_blocked, float minX, float maxX, float minY,
                                 float maxY, float th, float nnratio, int* match, int cap, int* nmatches);

/* ivg_search_by_bow = ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 *   (src/ORBmatcher.cc:165-294) on frame `index` (F).  The caller walks the two DBoW2 feature vectors as the reference does
 *   (:185-282) and hands over, in that traversal order, the key-frame keypoints of the shared vocabulary nodes: desc =
 *   pKF->mDescriptors.row(realIdxKF), angle = pKF->mvKeysUn[realIdxKF].angle, flags bit0 = pMP && !pMP->isBad(),
 *   node_slot[i] = which of F's node lists the point is matched against; F's lists as CSR (node_start[n_nodes + 1],
 *   node_idx = the vIndicesF entries).  Any keypoint assigned is skipped by all later points (:219-220); acceptance is
 *   bestDist1 <= TH_LOW (50) && bestDist1 < nnratio * bestDist2; then the rotation-histogram vote.  match / nmatches as
 *   above (match[i2] = index of the point, i.e. vpMapPointMatches[i2] = vpMapPointsKF[realIdxKF of that point]).
 *   Does not need the grid. */
int ivg_search_by_bow(ivg_extractor* cur, int index, int n, const uint8_t* desc, const float* angle, const uint8_t* flags, const int* node_slot,
                      int n_nodes, const int* node_start, const int* node_idx, float nnratio, int check_orientation, int* match, int cap,
                      int* nmatches);

/* ---- measurement helpers (bench.py) ---- */
/* CUDA-event timer on the handle's stream: start records an event, stop records another, elapsed waits for it. */
int ivg_timer_start(ivg_extractor* h);
int ivg_timer_stop(ivg_extractor* h);
int ivg_timer_elapsed_ms(ivg_extractor* h, float* ms);
/* Kernel launches issued by this handle since creation (our own kernels only; copies are not counted). */
long long ivg_launch_count(const ivg_extractor* h);
/* Pinned host memory for real async copies. */
int ivg_host_alloc(void** ptr, size_t bytes);
int ivg_host_free(void* ptr);
/* Writes `bytes` of device memory on the handle's stream (L2 flush between timed iterations). */
int ivg_flush_l2(ivg_extractor* h, size_t bytes);
/* Per-kernel device time: while enabled every kernel launch of this handle is bracketed by two CUDA events on the
 * handle's stream.  ivg_profile_read waits for the stream and returns accumulated milliseconds and launch counts per
 * kernel id since the last enable (arrays of IVG_NUM_KERNELS). */
#define IVG_K_RESIZE 0
#define IVG_K_FAST 1
#define IVG_K_BLUR 2
#define IVG_K_SELECT 3
#define IVG_K_DESCRIBE 4
#define IVG_K_STEREO 5
#define IVG_K_MEDIAN 6
#define IVG_K_PROLOGUE 7   /* N4: k_prologue (remap + cvtColor ingest) */
#define IVG_K_PROJ_CAND 8      /* N2: k_proj_candidates */
#define IVG_K_PROJ_RESOLVE 9   /* N2: k_proj_resolve */
#define IVG_NUM_KERNELS 10
int ivg_profile_enable(ivg_extractor* h, int enable);
int ivg_profile_read(ivg_extractor* h, double* ms, long long* launches);
/* Test hook: runs the warp-parallel replay of libstdc++'s std::nth_element(first, first+nth, last, key-greater) used by the
 * selection kernel on n keys (n <= 4000) and returns the resulting permutation (original indices in their new order). */
int ivg_debug_nth_element(int device, const uint32_t* keys, int n, int nth, uint32_t* order);
/* the same replay by a 1024-thread CTA (what k_level_select uses for the level trim of a single frame) */
int ivg_debug_nth_element_block(int device, const uint32_t* keys, int n, int nth, uint32_t* order);
/* Test hook: every kernel of the path exists in a throughput configuration (batches) and a one-frame configuration, chosen
 * by grid size at launch.  mode 0 = automatic (default), 1 = always the throughput kernels, 2 = the one-frame kernels wherever
 * their preconditions hold.  Results are identical in every mode; the tests run odd geometries through both. */
int ivg_debug_force_config(ivg_extractor* h, int mode);
/* When enabled (default off) run_batch wraps the kernel sequence of a batch in a CUDA graph that is re-used while
 * shape/batch stay the same. */
int ivg_set_graph_mode(ivg_extractor* h, int enable);

#ifdef __cplusplus
}
#endif
#endif /* IVSLAM_GPU_H_ */

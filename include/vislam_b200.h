/*
 * vislam_b200.h — C ABI of the B200-native frame-tracking path (libvislam_b200.so).
 *
 * This is the drop-in boundary for the hot path of MecatronicaUSB/vi-slam: the reference has no FFI,
 * its "operator API" is the C++ classes Matcher / Camera / VISystem.  The class mirrors
 * in vi-slam_b200/host/ call ONLY the functions declared here; each entry cites the reference code it
 * replaces (file:line under the reference tree).
 *
 * Conventions
 *   - every function returns an int status: 0 = VSB_OK, negative = error (vsb_error_string()).
 *   - pointers are DEVICE pointers unless the parameter name starts with h_ (host, ideally pinned).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on
 *     that stream unless documented otherwise; no hidden host synchronisation.
 *   - batched calls process `count` independent frame pairs laid out with uniform strides; per-problem
 *     feature counts may be given as device int32 arrays (NULL = every problem uses the max count).
 *   - pose layout: {qx, qy, qz, qw, tx, ty, tz} float32 (Sophus::SE3f storage order).
 *   - there is no CPU fallback: without a CUDA device every compute entry returns VSB_ERR_CUDA.
 */
#ifndef VISLAM_B200_H_
#define VISLAM_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VSB_OK 0
#define VSB_ERR_INVALID (-1)     /* bad argument */
#define VSB_ERR_CUDA (-2)        /* CUDA runtime error (see vsb_last_cuda_error) */
#define VSB_ERR_UNSUPPORTED (-3) /* mode not implemented */
#define VSB_ERR_CAPACITY (-4)    /* a size exceeds a compiled-in or context limit */

#define VSB_MAX_LEVELS 5         /* Frame keeps 5 pyramid levels, Camera.hpp:46 */
#define VSB_MAX_GN_FEATURES 200  /* Camera.cpp:382 */
#define VSB_MAX_TRACE 64

typedef struct vsb_ctx vsb_ctx_t;

/* ---- context ------------------------------------------------------------------------------------ */
int vsb_version(void);
const char* vsb_error_string(int status);
/* A context owns scratch memory that the stand-alone compute entries (vsb_knn2_*, vsb_match_filter, vsb_gn_solve, vsb_orb_*)
 * use internally: calls on ONE context must not overlap in time on different streams (use one context per stream, or
 * the tracker, whose host entry keeps per-stream workspaces for its two streams). */
int vsb_ctx_create(int device, vsb_ctx_t** ctx);
int vsb_ctx_destroy(vsb_ctx_t* ctx);
/* Tuning knobs (defaults in parentheses; the environment variables VSB_KNN_IMPL / VSB_GN_THREADS set the same at
 * context creation): "knn_impl" = 0 POPC kernel on the INT pipe, 1 / 2 tcgen05 int8 kernels (2: packed 16x2 epilogue),
 * 3 / 4 / 5 4-bit (mxf4) kernels (5: persistent, bulk-copied tiles), (6) chosen by the set size: 5 from 768 descriptors up, else 2; "gn_threads" = threads per frame pair of the GN solver, 64 / 128 / 256 / 512 / 1024 ((0) = chosen from the
 * batch size); "gn_impl" = (1) the tracker solves reference-mode problems with gn_track.cu, 0 = always gn_solve.cu;
 * "gn_stage_bytes" = shared-memory budget for the staged current-image level of gn_track.cu ((8192); 0 = none);
 * "gn_cluster" = (1) a batch of fewer frame pairs than 0.7 x the SMs gives each pair a thread-block cluster of 2 / 4 / 8
 *   blocks that share the sweep and exchange partial sums through distributed shared memory; 0 = never, 2 / 4 / 8 = always
 *   that size.  "gn_cluster_threads" = (0: by batch size) | 256 | 512 threads per block of that kernel.
 * "gn_tail" = (1) the pairs of the last partial wave of a large batch get more threads each, 0 = one launch.
 * "gn_dedup" = (1) candidate points of the small levels are merged per distinct pixel with a multiplicity, 0 = one record per
 *   point; "gn_variant" = (0) | 1 | 2 | 3 earlier forms of the tracker's solver kept for comparison;
 *   "knn_l2_impl" = (1) float kNN as tensor-core distance GEMM + exact re-check, 0 = exact FP64 kernel; "pyr_impl" = (1)
 *   register-blocked pyramid kernels, 0 = shared-memory tile kernel.
 * ORB front end: "orb_scratch_mb" = budget of one detector workspace in MB ((32768): 2000 752x480 frames with one block of
 *   scratch per pyramid level); "orb_lp" = (1) the pyramid levels of a batch run on separate streams when one block of scratch
 *   per level fits the workspace, 0 = one stream; "orb_impl" / "fast_impl" = (0) bit masks that switch the PREVIOUS form of a
 *   kernel back on for comparison (orb_impl 1 per-pixel resize, 2 per-warp sin / cos, 4 word-wise blur loader, 8 register
 *   resize; fast_impl 1 per-word compaction and tile loader).
 * Results are identical for every setting. */
int vsb_ctx_option(vsb_ctx_t* ctx, const char* name, int value);
const char* vsb_last_cuda_error(vsb_ctx_t* ctx);
int vsb_sm_count(vsb_ctx_t* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long vsb_launch_count(vsb_ctx_t* ctx);

/* Optional per-kernel timing: CUDA events recorded on the launching stream around every kernel launch.
 * kernel ids are 0 .. vsb_kernel_count()-1 (vsb_kernel_name).  vsb_profile_read synchronises on the events
 * recorded so far and returns the accumulated device time and launch count of one kernel. */
int vsb_kernel_count(void);
const char* vsb_kernel_name(int kernel_id);
int vsb_profile_enable(vsb_ctx_t* ctx, int on);
int vsb_profile_reset(vsb_ctx_t* ctx);
int vsb_profile_read(vsb_ctx_t* ctx, int kernel_id, double* total_ms, long long* launches);
/* INT-pipe ceiling of this GPU, measured: 32-bit POPC (independent chains, all SMs) per second. */
int vsb_popc_peak(vsb_ctx_t* ctx, double* popc_per_s, void* stream);

/* Device / pinned-host memory, copies and streams for hosts that do not link the CUDA runtime themselves
 * (the C++ class mirrors in vi-slam_b200/host/ are plain g++ code; they replace the cv::Mat / cuda::GpuMat
 * upload()/download() calls of MatcherGPU.cpp:48-50 and CameraGPU.cpp).  Copies are asynchronous on `stream`. */
int vsb_malloc(vsb_ctx_t* ctx, size_t bytes, void** dptr);
int vsb_free(vsb_ctx_t* ctx, void* dptr);
int vsb_host_alloc(vsb_ctx_t* ctx, size_t bytes, void** hptr);
int vsb_host_free(vsb_ctx_t* ctx, void* hptr);
int vsb_upload(vsb_ctx_t* ctx, void* dst, const void* h_src, size_t bytes, void* stream);
int vsb_upload_2d(vsb_ctx_t* ctx, void* dst, size_t dst_pitch, const void* h_src, size_t src_pitch,
                  size_t width_bytes, size_t rows, void* stream);
int vsb_download(vsb_ctx_t* ctx, void* h_dst, const void* src, size_t bytes, void* stream);
int vsb_copy(vsb_ctx_t* ctx, void* dst, const void* src, size_t bytes, void* stream);
int vsb_memset(vsb_ctx_t* ctx, void* dst, int value, size_t bytes, void* stream);
int vsb_stream_create(vsb_ctx_t* ctx, void** stream);
int vsb_stream_destroy(vsb_ctx_t* ctx, void* stream);
int vsb_stream_sync(vsb_ctx_t* ctx, void* stream);

/* ---- Matcher ------------------------------------------------------------------------------------ */
/* Replaces Matcher::computeMatches (src/Matcher.cpp:83-94) / MatcherGPU::computeGPUMatches
 * (src/MatcherGPU.cpp:44-66) with BFMatcher(NORM_HAMMING): both knnMatch(...,2) calls from ONE
 * pass over the distance matrix.
 *   d1: [count][n1_max][32] u8 (previous key-frame descriptors), d2: [count][n2_max][32] u8.
 *   n1/n2: optional device int32[count] with the true row counts (<= max).
 *   idx12/dist12: [count][n1_max][2] (neighbours of d1 rows in d2); idx21/dist21: [count][n2_max][2].
 * Order: (distance asc, train index asc) — bit-exact with cv::BFMatcher incl. ties; slots without a
 * neighbour (fewer than 2 train rows) get idx = -1, dist = 0.  dist holds the integer popcount as float. */
int vsb_knn2_hamming(vsb_ctx_t* ctx, const uint8_t* d1, int n1_max, const int32_t* n1,
                     const uint8_t* d2, int n2_max, const int32_t* n2, int count,
                     int32_t* idx12, float* dist12, int32_t* idx21, float* dist21, void* stream);

/* Same for float descriptors, BFMatcher(NORM_L2) (Matcher.cpp:55): dim floats per row,
 * distance = sqrtf(sum (a-b)^2). */
int vsb_knn2_l2(vsb_ctx_t* ctx, const float* d1, int n1_max, const int32_t* n1,
                const float* d2, int n2_max, const int32_t* n2, int dim, int count,
                int32_t* idx12, float* dist12, int32_t* idx21, float* dist21, void* stream);

/* Replaces Matcher::computeBestMatches + getGoodMatches (Matcher.cpp:353-367, 295-303):
 * nnFilter (ratio), symmetry test, sort by y, sqrt(n_cells) x sqrt(n_cells) grid best.
 *   kp1_xy: [count][n1_max][2] f32 key-point coordinates of the previous key frame.
 *   sym_mode: 0 de-facto reference behaviour (2->1 ratio test has no effect, App. B-1), 1 intended.
 *   good_q/good_t/good_d: [count][good_cap]; n_good, n_sym: [count].  Order = reference order
 *   (band-major, column-minor).  good_cap must be >= floor(sqrt(n_cells))^2. */
int vsb_match_filter(vsb_ctx_t* ctx, const int32_t* idx12, const float* dist12, int n1_max, const int32_t* n1,
                     const int32_t* idx21, const float* dist21, int n2_max, const int32_t* n2,
                     const float* kp1_xy, int count, int w, int h, int n_cells, float ratio, int sym_mode,
                     int32_t* good_q, int32_t* good_t, float* good_d, int good_cap,
                     int32_t* n_good, int32_t* n_sym, void* stream);

/* The same chain one public Matcher method at a time (the reference exposes them separately over public
 * vectors, Matcher.hpp:34-47); the class mirror calls these, the tracker uses the fused vsb_match_filter.
 *
 * Matcher::nnFilter (Matcher.cpp:148-169): keep[c][i] = 0 where the reference clears the row
 * (fewer than 2 neighbours, or d0 > ratio * d1 evaluated in double with ratio = (double)0.8f). */
int vsb_nn_filter(vsb_ctx_t* ctx, const int32_t* idx, const float* dist, int n_max, const int32_t* n, int count,
                  double ratio, uint8_t* keep, void* stream);
/* Matcher::computeSymMatches after its two nnFilter calls (Matcher.cpp:113-143).  keep12 / keep21 are the
 * nnFilter masks; sym_mode 0 ignores keep21 (de-facto: cleared rows are still read, App. B-1), 1 requires it.
 * sym_q / sym_t / sym_d: [count][n1_max] compacted in ascending query order = Matcher::matches; n_sym: [count]. */
int vsb_sym_matches(vsb_ctx_t* ctx, const int32_t* idx12, const float* dist12, const uint8_t* keep12, int n1_max,
                    const int32_t* n1, const int32_t* idx21, const uint8_t* keep21, int n2_max, const int32_t* n2,
                    int count, int sym_mode, int32_t* sym_q, int32_t* sym_t, float* sym_d, int32_t* n_sym,
                    void* stream);
/* Matcher::sortMatches (Matcher.cpp:329-352): order[c][r] = position of the r-th smallest key, ties keep the
 * input order (cv::sortIdx ascending; stable by decision).  keys/order: [count][cap]. */
int vsb_sort_keys(vsb_ctx_t* ctx, const float* keys, int cap, const int32_t* n, int count, int32_t* order,
                  void* stream);
/* Matcher::bestMatchesFilter (Matcher.cpp:171-244) over a y-sorted match list (Matcher::sortedMatches):
 * list_q/list_t/list_d [count][list_cap], n_list [count]; good_* [count][good_cap] in reference order;
 * good_pos (optional) = position of each winner in the input list. */
int vsb_grid_best(vsb_ctx_t* ctx, const int32_t* list_q, const int32_t* list_t, const float* list_d, int list_cap,
                  const int32_t* n_list, const float* kp1_xy, int n1_max, int count, int w, int h, int n_cells,
                  int32_t* good_q, int32_t* good_t, float* good_d, int32_t* good_pos, int good_cap,
                  int32_t* n_good, void* stream);

/* ---- Camera ------------------------------------------------------------------------------------- */
typedef struct {
    int levels;                       /* <= VSB_MAX_LEVELS */
    int w[VSB_MAX_LEVELS], h[VSB_MAX_LEVELS];
    int64_t offset[VSB_MAX_LEVELS];   /* pixel offset of level l inside one frame's packed pyramid (level 0 = 0) */
    int64_t frame_stride;             /* pixels per frame in the packed pyramid (256-aligned) */
} vsb_pyr_layout_t;

/* Packed-pyramid layout for a w x h frame: level sizes follow cv::resize(0.5) (cvRound), Camera.cpp:68-70. */
int vsb_pyr_layout(int w, int h, int levels, vsb_pyr_layout_t* out);

/* Replaces Camera::Update (Camera.cpp:63-72): builds levels 0..levels-1 of `count` frames.
 *   img: [count] frames, row pitch `pitch` bytes, frame stride `img_stride` bytes.
 *   pyr: [count][layout.frame_stride] u8 (level 0 is copied, as the reference does). */
int vsb_pyramid_build(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int count,
                      const vsb_pyr_layout_t* layout, uint8_t* pyr, void* stream);

/* Replaces Camera::computeGradient (Camera.cpp:167-184): Scharr x/y, scale 3, CV_16S, reflect-101, every level.
 *   gx, gy: [count][layout.frame_stride] int16;  gmag (optional, may be NULL): |gx|/2+|gy|/2 u8 image. */
int vsb_gradient_build(vsb_ctx_t* ctx, const uint8_t* pyr, int count, const vsb_pyr_layout_t* layout,
                       int16_t* gx, int16_t* gy, uint8_t* gmag, void* stream);

/* Replaces Camera::ObtainPatchesPointsPreviousFrame (Camera.cpp:358-409).
 *   good_xy: [count][good_cap][2] f32 level-0 key points (prev.nextGoodMatches), n_good: [count].
 *   lw/lh: Camera::w_size/h_size tables (w>>l, Camera.cpp:44-47).
 *   cand: [count][levels][cand_cap][4] f32 rows (x, y, 1, 1) = Frame::candidatePoints; n_cand: [count][levels].
 *   cand_cap >= 121 * min(max n_good, 200). */
int vsb_candidates_build(vsb_ctx_t* ctx, const float* good_xy, int good_cap, const int32_t* n_good, int count,
                         int levels, const int* lw, const int* lh,
                         float* cand, int cand_cap, int32_t* n_cand, void* stream);

/* Gathers key points of the good matches (Matcher::getGoodMatches, Matcher.cpp:295-303):
 * out_xy[c][m] = kp_xy[c][good_idx[c][m]]. */
int vsb_gather_keypoints(vsb_ctx_t* ctx, const float* kp_xy, int n_max, const int32_t* good_idx, int good_cap,
                         const int32_t* n_good, int count, float* out_xy, void* stream);

/* ---- VISystem ----------------------------------------------------------------------------------- */
typedef struct {
    float fx, fy, cx, cy, invfx, invfy;
    int w, h;   /* InitializePyramid's size table (w>>l); image rows/cols come from the pyramid layout */
} vsb_intr_t;

/* Replaces VISystem::InitializePyramid (VISystem.cpp:1451-1493) — host-side, no device work. */
int vsb_init_pyramid(int w, int h, float fx, float fy, float cx, float cy, vsb_intr_t out[VSB_MAX_LEVELS]);

typedef struct {
    int first_lvl;       /* 3      VISystem.cpp:1119 */
    int last_lvl;        /* 0      VISystem.cpp:1120 */
    int max_iterations;  /* 10     VISystem.cpp:1117 */
    float epsilon;       /* 0.001  VISystem.cpp:1115 */
    float z_factor;      /* 0.002  VISystem.cpp:1121 */
    int weight_mode;     /* 0 identity (reference, :1343)   2 Huber (north-star extension) */
    int sample_mode;     /* 0 nearest round() (reference, :1321)   1 bilinear (north-star extension) */
    float huber_k;
    int grad_mode;       /* 0 read gx/gy images   1 Scharr evaluated on the fly from the previous image */
    int accum_mode;      /* must be 0: exact FP32 products accumulated in FP64, as cv::gemm does (round 1's mode 1, FP32
                            per-thread partials, missed the 1e-5 tolerance and was removed; VSB_ERR_UNSUPPORTED) */
} vsb_gn_opts_t;
void vsb_gn_default_opts(vsb_gn_opts_t* o);

typedef struct {
    int lvl, iter, n_valid, updated;
    float error;
    float pose[7];
    float delta[6];
} vsb_gn_trace_t;

/* Replaces VISystem::EstimatePoseFeatures (VISystem.cpp:1113-1448) incl. WarpFunctionSE3 (:1495-1558),
 * IdentityWeights (:1561-1565) and the Sophus SE3f update (se3.hpp:723-742, 285-321), for `count`
 * independent frame pairs, one thread block per pair, all levels and iterations inside the kernel.
 *   prev_pyr/cur_pyr: packed pyramids of the previous / current frame of pair c at
 *                     base + c * pair_stride_pixels (so consecutive frames of a sequence can be
 *                     addressed with prev = pyr, cur = pyr + frame_stride, pair stride = frame_stride).
 *   prev_gx/prev_gy : packed gradients of the previous frame (may be NULL when grad_mode == 1).
 *   cand/n_cand     : as produced by vsb_candidates_build.
 *   pose_in/pose_out: [count][7].   trace (optional): [count][VSB_MAX_TRACE], n_trace: [count]. */
int vsb_gn_solve(vsb_ctx_t* ctx, const uint8_t* prev_pyr, const uint8_t* cur_pyr,
                 const int16_t* prev_gx, const int16_t* prev_gy, int64_t pair_stride_pixels,
                 const vsb_pyr_layout_t* layout, const float* cand, int cand_cap, const int32_t* n_cand,
                 const vsb_intr_t K[VSB_MAX_LEVELS], const float* pose_in, const vsb_gn_opts_t* opts, int count,
                 float* pose_out, vsb_gn_trace_t* trace, int32_t* n_trace, void* stream);

/* VISystem::WarpFunctionSE3 (VISystem.cpp:1495-1558) on its own: pts/out are n x 4 f32 rows (x, y, z, 1). */
int vsb_warp_se3(vsb_ctx_t* ctx, const float* pts, int n, const float pose[7], const vsb_intr_t* K, float* out,
                 void* stream);
/* The pose update of a GN iteration, batched on the device: out[i] = pose[i] * exp(delta[i]) (VISystem.cpp:1421;
 * se3.hpp:723-742, 285-321).  pose / out: n x 7 f32 {qx,qy,qz,qw,tx,ty,tz}, delta: n x 6 f32 (upsilon, omega); device. */
int vsb_se3_update_batch(vsb_ctx_t* ctx, const float* pose, const float* delta, int n, float* out, void* stream);
/* Sophus SE3f::exp (se3.hpp:723-742) and SE3f::matrix (se3.hpp:253-268, row-major 4x4).  Host-side. */
int vsb_se3_exp(const float delta[6], float pose[7]);
/* Sophus SE3f(rotation matrix, translation) (so3.hpp:422-427 -> Eigen Quaternion(Matrix3)); r row-major. Host-side. */
int vsb_se3_from_rt(const float r[9], const float t[3], float pose[7]);
int vsb_se3_matrix(const float pose[7], float m[16]);

/* Initial pose exactly as VISystem.cpp:1135-1168 forms it (host-side helper):
 * pose0 = SE3(RPY2rotationMatrix(-rotationMatrix2RPY(imu2cam^T * R_imu_res * imu2cam)), -t_res). */
int vsb_initial_pose(const float imu2cam[9], const float r_imu_res[9], const float t_res[3], float pose[7]);
/* Sophus SE3f group product as VISystem::Track uses it (VISystem.cpp:1599; se3.hpp:285-321). Host-side. */
int vsb_se3_mul(const float a[7], const float b[7], float out[7]);

/* ---- whole tracking step ------------------------------------------------------------------------ */
typedef struct {
    int w, h, n_feat_max, desc_bytes;   /* desc_bytes = 32 (ORB).  float descriptors: desc_bytes = 4*dim */
    int norm;                           /* 1 Hamming, 0 L2 */
    int n_cells;                        /* calibration `num_cells` (49 in calibrationEUROC.xml:54) */
    float ratio;                        /* 0.8f, Matcher.cpp:103 */
    int sym_mode;
    float fx, fy, cx, cy;
    vsb_gn_opts_t gn;
    int max_pairs;                      /* capacity of one tracker batch */
} vsb_tracker_cfg_t;
typedef struct vsb_tracker vsb_tracker_t;

/* A tracker owns the device scratch for `max_pairs` frame pairs (pyramids, kNN results, matches,
 * candidates) so the whole loop of VISystemGPU::AddFrameGPU (VISystemGPU.cpp:144-169) —
 * Update -> computeGPUGoodMatches -> computeGradient -> ObtainPatchesPointsPreviousFrame ->
 * EstimatePoseFeatures — runs without returning to the host between stages. */
int vsb_tracker_create(vsb_ctx_t* ctx, const vsb_tracker_cfg_t* cfg, vsb_tracker_t** out);
int vsb_tracker_destroy(vsb_tracker_t* t);

/* ---- feature detection, first stage ---------------------------------------------------------
 * FAST-9/16 corners with non-maximum suppression, the key-point detector inside cv::ORB (src/Camera.cpp:124-129,
 * ORB::create) and cv::cuda::ORB (src/CameraGPU.cpp:99-104): cv::FAST(img, kps, threshold, nonmax, TYPE_9_16) for `count`
 * frames.  img: device, frames of h rows x pitch bytes, img_stride bytes apart.  Outputs (device): kp_xy
 * [count][cap][2] int32 (x, y) and kp_score [count][cap] int32 (cv::KeyPoint::response; 0 without suppression), in
 * cv::FAST's row-major order; n_kp [count] = number of corners FOUND (only the first cap are stored). */
int vsb_fast_detect(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count,
                    int threshold, int nonmax, int cap, int32_t* kp_xy, int32_t* kp_score, int32_t* n_kp, void* stream);

/* ---- feature detection + description: cv::ORB, one pyramid level --------------------------------
 * What cv::ORB::create(nfeatures, 1.2f, 1, 31, 0, 2, ORB::HARRIS_SCORE, 31, fast_threshold)->detectAndCompute(img, noArray(),
 * keypoints, descriptors) computes — the detector of Camera::detectAndComputeFeatures (src/Camera.cpp:79-86, setDetector
 * :124-129) and CameraGPU (src/CameraGPU.cpp:99-104) restricted to nlevels = 1 — for `count` frames: FAST-9/16 with
 * suppression, border filter (31), retainBest(2 n) on the FAST score, Harris response (7x7, k 0.04), retainBest(n),
 * intensity-centroid angle, 7x7 sigma-2 blur, steered rBRIEF.  img as in vsb_fast_detect.  Outputs (device), per frame up to
 * cap entries in row-major (y, x) order: kp_xy [count][cap][2] int32, kp_resp [count][cap] (Harris response =
 * cv::KeyPoint::response), kp_angle [count][cap] degrees (cv::KeyPoint::angle), desc [count][cap][32] (may be NULL: detect
 * only); n_kp [count] = key points selected (ties at the two thresholds are kept as OpenCV keeps them, so it can exceed
 * nfeatures; only the first cap are stored).  Bit-identical to cv2 4.13 through the oracle (oracle/orb.c).
 * w, h must exceed 62 (the border filter leaves nothing otherwise: VSB_ERR_INVALID). */
int vsb_orb_detect_compute(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count,
                           int nfeatures, int fast_threshold, int cap, int32_t* kp_xy, float* kp_resp, float* kp_angle,
                           uint8_t* desc, int32_t* n_kp, void* stream);

/* The same with the detector's scale pyramid — cv::ORB::create(nfeatures, scale_factor, nlevels, 31, 0, 2, HARRIS_SCORE, 31,
 * fast_threshold), i.e. exactly what the reference's ORB::create(n) runs with scale_factor 1.2f, nlevels 8: level l is the
 * INTER_LINEAR_EXACT resize of level l-1 to cvRound(size / scale_l), scale_l = (float)pow((double)scale_factor, l); every
 * level runs the one-level pipeline with cv::ORB's per-level feature budget.  Outputs per frame, level by level (row-major
 * inside a level): kp_xy [count][cap][2] FLOAT = level coordinates * scale_l (cv::KeyPoint::pt), kp_octave [count][cap] (may
 * be NULL), kp_resp, kp_angle, desc (may be NULL); cv::KeyPoint::size is 31 * scale_l.  Bit-identical to cv2 4.13 through the
 * oracle. */
int vsb_orb_detect_compute_pyr(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count,
                               int nfeatures, float scale_factor, int nlevels, int fast_threshold, int cap, float* kp_xy,
                               int32_t* kp_octave, float* kp_resp, float* kp_angle, uint8_t* desc, int32_t* n_kp, void* stream);

/* Work counters of the solver since the last call (then reset): out[0] = frame pairs solved,
 * out[1] = GN iterations (error evaluations), out[2] = sum over iterations of the candidate points visited
 * (SURVEY.md §8d's  sum_l K_l * P_l), out[3] = accepted pose updates.  Synchronises the device. */
int vsb_tracker_stats(vsb_tracker_t* t, long long out[4]);
/* Per-iteration trace of the solver inside the tracker (parity tests: the north-star tolerance is stated per iteration).
 * trace: device [cfg.max_pairs][VSB_MAX_TRACE], n_trace: device [cfg.max_pairs]; both NULL switches it off.  The
 * device entries (vsb_track_pairs / vsb_track_sequence / _orb) called afterwards fill them for their pairs. */
int vsb_tracker_set_trace(vsb_tracker_t* t, vsb_gn_trace_t* trace, int32_t* n_trace);
/* Bytes the last vsb_track_sequence_host call copied host->device and device->host, and its chunk count
 * (out[0..2]); the frame shared by two consecutive chunks is uploaded with both. */
int vsb_tracker_host_traffic(vsb_tracker_t* t, long long out[3]);

/* Tracks `n_frames - 1` consecutive pairs of a sequence resident on the DEVICE.
 *   frames [n_frames][h][w] u8, desc [n_frames][n_feat_max][desc_bytes], kp_xy [n_frames][n_feat_max][2] f32,
 *   n_feat (optional) int32[n_frames], pose_prior [n_frames-1][7].
 *   Outputs: pose [n_frames-1][7] (relative pose prev->cur of every pair), n_good (optional) [n_frames-1].
 * n_frames - 1 <= cfg.max_pairs. */
int vsb_track_sequence(vsb_tracker_t* t, const uint8_t* frames, const uint8_t* desc, const float* kp_xy,
                       const int32_t* n_feat, const float* pose_prior, int n_frames,
                       float* pose, int32_t* n_good, void* stream);

/* The same from images alone: every frame first goes through cv::ORB::create(nfeatures) on the device
 * (vsb_orb_detect_compute_pyr: 8 levels, factor 1.2) — Camera::detectAndComputeFeatures (src/Camera.cpp:84-93) — and its key
 * points and descriptors feed the matcher without leaving the device.  The tracker must be configured for ORB descriptors
 * (norm 1, desc_bytes 32); a frame with more key points than cfg.n_feat_max keeps the first n_feat_max.  n_feat_out
 * (optional, device) [n_frames] = key points used per frame. */
int vsb_track_sequence_orb(vsb_tracker_t* t, const uint8_t* frames, const float* pose_prior, int n_frames, int nfeatures,
                           float* pose, int32_t* n_good, int32_t* n_feat_out, void* stream);

/* Same with HOST buffers (the end-to-end entry the class mirrors and bench.py's e2e leg use): frames,
 * descriptors, key points and priors are copied host->device in chunks of cfg.max_pairs pairs on two
 * streams so the copy of chunk i+1 overlaps the kernels of chunk i; poses are copied back.
 * Synchronous: returns when h_pose is complete. */
int vsb_track_sequence_host(vsb_tracker_t* t, const uint8_t* h_frames, const uint8_t* h_desc,
                            const float* h_kp_xy, const int32_t* h_n_feat, const float* h_pose_prior,
                            int n_frames, float* h_pose, int32_t* h_n_good);

/* The loop from images alone with HOST buffers: frames and priors are copied host->device in chunks of cfg.max_pairs
 * pairs on a copy stream while ORB (cv::ORB::create(nfeatures), 8 levels, factor 1.2), the matcher and the solver run on
 * the previous chunk — what the reference's CameraGPU path does per frame (upload, cv::cuda::ORB, match;
 * src/CameraGPU.cpp:71-117, src/VISystemGPU.cpp:144-169), for a whole sequence.  h_n_good [n_frames - 1] and h_n_feat
 * [n_frames] (key points used per frame) are optional.  Synchronous: returns when h_pose is complete. */
int vsb_track_sequence_orb_host(vsb_tracker_t* t, const uint8_t* h_frames, const float* h_pose_prior, int n_frames,
                                int nfeatures, float* h_pose, int32_t* h_n_good, int32_t* h_n_feat);

/* Tracks `count` independent frame pairs resident on the DEVICE (BASELINE config 5):
 *   prev/cur [count][h][w] u8, d1/d2 [count][n_feat_max][desc_bytes], kp1 [count][n_feat_max][2]. */
int vsb_track_pairs(vsb_tracker_t* t, const uint8_t* prev, const uint8_t* cur, const uint8_t* d1,
                    const uint8_t* d2, const float* kp1_xy, const int32_t* n1, const int32_t* n2,
                    const float* pose_prior, int count, float* pose, int32_t* n_good, void* stream);

/* The same with HOST buffers (end-to-end form of BASELINE config 5): chunks of cfg.max_pairs pairs alternate between two
 * streams, so the upload of chunk i + 1 overlaps the kernels of chunk i; poses (and n_good, optional) are copied back.
 * h_n1 / h_n2: both NULL (every set holds n_feat_max descriptors) or both given.  Synchronous. */
int vsb_track_pairs_host(vsb_tracker_t* t, const uint8_t* h_prev, const uint8_t* h_cur, const uint8_t* h_d1,
                         const uint8_t* h_d2, const float* h_kp1_xy, const int32_t* h_n1, const int32_t* h_n2,
                         const float* h_pose_prior, int count, float* h_pose, int32_t* h_n_good);

#ifdef __cplusplus
}
#endif
#endif

/*
 * vso.h — CPU ORACLE for the vi-slam frame-tracking hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is a plain-C restatement of the reference's algorithm for the path
 *   Matcher (kNN k=2 + ratio + symmetry + sort + grid)  ->  Camera (pyramid, Scharr, candidates)
 *   ->  VISystem::EstimatePoseFeatures (Gauss-Newton photometric SE3 solve).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (libvislam_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY STATUS: **pinned by the reference's own sources run here** (oracle/_ref, recipe: oracle/Makefile target `ref`).
 * MecatronicaUSB/vi-slam ships no tests, golden vectors or fixtures, and its build needs OpenCV 3.2 + contrib, ROS
 * Kinetic, Eigen and ceres, none of which exist here.  Its translation units for this path are therefore
 * compiled UNMODIFIED, in place, against stand-in headers written for the purpose:
 *   - src/Matcher.cpp against the type shim oracle/cvshim -> nnFilter / computeSymMatches / sortMatches /
 *     bestMatchesFilter / getGoodMatches (tests/golden/matcher_ref.npz);
 *   - src/VISystem.cpp + Camera.cpp (+ CameraModel, Matcher, Plus, Imu) against the functional OpenCV stand-in
 *     oracle/refshim -> Camera::Update / computeGradient / computeGoodMatches / ObtainPatchesPointsPreviousFrame,
 *     VISystem::InitializePyramid / WarpFunctionSE3 / EstimatePoseFeatures / TukeyFunctionWeights
 *     (tests/golden/visystem_ref.npz, tests/test_ref_visystem.py: bit for bit, also live on full-size pairs);
 *   - the third-party primitives the stand-ins restate (BFMatcher::knnMatch, resize(0.5), Scharr scale 3, addWeighted,
 *     invert, solve) are checked against Python cv2 4.13 (tests/test_oracle_cv2.py, the _cv2.npz fixtures in tests/golden).
 * NOT pinned by the reference: the Sophus SE3 exp / compose arithmetic (vendored Sophus needs Eigen; restated here from
 * se3.hpp / so3.hpp and checked against scipy) and OpenCV's float convertTo / MatExpr folding rules (restated from
 * OpenCV 3.2's matop.cpp / convert.cpp; cv2's Python API does not expose them).
 *
 * All citations are file:line under /root/reference.
 */
#ifndef VSO_H_
#define VSO_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VSO_MAX_LEVELS 5

/* ---- Matcher (src/Matcher.cpp) -------------------------------------------------------------- */

/* BFMatcher(NORM_HAMMING)::knnMatch(q, t, out, 2)  (call sites Matcher.cpp:86,88).
 * idx/dist are nq x 2.  Order: (distance asc, train index asc).  When nt < 2 the missing slots get
 * idx = -1, dist = 0 (cv returns shorter lists).  nbytes = descriptor length in bytes (32 for ORB). */
void vso_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                      int32_t* idx, float* dist);

/* BFMatcher(NORM_L2)::knnMatch k=2 for float descriptors (Matcher.cpp:55).  distance =
 * sqrtf((float) sum_d (double)(q-t)^2), differences taken in float as cv::normL2Sqr does. */
void vso_knn2_l2(const float* q, int nq, const float* t, int nt, int dim, int32_t* idx, float* dist);

/* Matcher::nnFilter (Matcher.cpp:148-169).  keep[i] = 0 when the row is cleared. ratio is the
 * double 0.800000011920929 when called as the reference does (0.8f widened, Matcher.cpp:103). */
void vso_nn_filter(const int32_t* idx, const float* dist, int n, double ratio, uint8_t* keep);

/* Matcher::computeSymMatches (Matcher.cpp:96-144).
 * mode 0 = de-facto (stale aux2 rows still readable, App. B-1), 1 = intended (aux2 row must survive).
 * Outputs (capacity n1): queryIdx, trainIdx, distance; returns the number of matches. */
int vso_sym_matches(const int32_t* idx1, const float* dist1, int n1,
                    const int32_t* idx2, const float* dist2, int n2,
                    double ratio, int mode, int32_t* mq, int32_t* mt, float* md);

/* Matcher::sortMatches (Matcher.cpp:329-352): stable ascending by kp1[query].y.  order[] = permutation. */
void vso_sort_matches(const int32_t* mq, int n, const float* kp1_xy, int32_t* order);

/* Matcher::bestMatchesFilter (Matcher.cpp:171-244) on the sorted matches.  Returns count (<= r*r). */
int vso_grid_filter(const int32_t* mq, const int32_t* mt, const float* md, const int32_t* order, int n,
                    const float* kp1_xy, int w, int h, int n_cells,
                    int32_t* gq, int32_t* gt, float* gd);

/* Camera::computeGoodMatches (Camera.cpp:146-157) = computeMatches + computeBestMatches + getGoodMatches.
 * norm: 0 = L2 float descriptors (dim floats), 1 = Hamming (dim bytes).  Returns the number of good
 * matches; gq/gt/gd need capacity n1.  n_sym (optional) receives nSymMatches. */
int vso_match_pipeline(const void* d1, int n1, const void* d2, int n2, int dim, int norm,
                       const float* kp1_xy, int w, int h, int n_cells, double ratio, int mode,
                       int32_t* gq, int32_t* gt, float* gd, int* n_sym);

/* ---- Camera (src/Camera.cpp) ---------------------------------------------------------------- */

/* cv::resize(src, dst, Size(), 0.5, 0.5) (Camera.cpp:69): INTER_LINEAR; dst size = cvRound(dim/2).
 * Exact halving takes OpenCV's area fast path (a+b+c+d+2)>>2; otherwise 11-bit fixed-point bilinear. */
void vso_pyr_size(int w, int h, int* dw, int* dh);
void vso_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst);

/* cv::Scharr(src, dst, CV_16S, dx, dy, scale=3, 0, BORDER_DEFAULT) (Camera.cpp:171-172). */
void vso_scharr3(const uint8_t* src, int w, int h, int16_t* gx, int16_t* gy);
/* the (unused by GN) |gx|/2+|gy|/2 image (Camera.cpp:174-180) */
void vso_grad_mag(const int16_t* gx, const int16_t* gy, int n, uint8_t* g);

/* Camera::ObtainPatchesPointsPreviousFrame (Camera.cpp:358-409) for ONE level.  good_xy = prev
 * nextGoodMatches keypoints (level-0 pixels), nf = their count (capped at 200 inside).
 * lw/lh = Camera::w_size/h_size[lvl] (= w>>lvl, Camera.cpp:44-47).  out = rows of (x,y,1,1) f32,
 * capacity 121*min(nf,200).  Returns the number of rows. */
int vso_candidates(const float* good_xy, int nf, int lvl, int lw, int lh, float* out);

/* ---- VISystem (src/VISystem.cpp) ------------------------------------------------------------ */

typedef struct {
    float fx, fy, cx, cy, invfx, invfy;
    int w, h;
} vso_intr_t;

/* VISystem::InitializePyramid (VISystem.cpp:1451-1493). */
void vso_init_pyramid(int w, int h, float fx, float fy, float cx, float cy, vso_intr_t out[VSO_MAX_LEVELS]);

/* pose layout everywhere: {qx, qy, qz, qw, tx, ty, tz} (Sophus::SE3f storage order). */
void vso_se3_exp(const float delta[6], float pose[7]);          /* se3.hpp:723-742, so3.hpp:534-568 */
void vso_se3_mul(const float a[7], const float b[7], float out[7]); /* se3.hpp:285-321, so3.hpp:338-353 */
void vso_se3_matrix(const float pose[7], float m[16]);          /* se3.hpp:253-268 (row-major 4x4) */
void vso_rpy_to_rot(const double rpy[3], float r[9]);           /* Plus.cpp:182-220 */
void vso_rot_to_rpy(const float r[9], double rpy[3]);           /* Plus.cpp:56-83 */
void vso_rot_to_quat(const float r[9], float q[4]);             /* Eigen Quaternion(Matrix3) */
int vso_se3_from_rt(const float r[9], const float t[3], float pose[7]); /* SE3(R, t), se3.hpp:438-440; returns 0 */
/* VISystem.cpp:1135-1168: pose0 = SE3(RPY2rot(-rot2RPY(imu2cam^T R_imu imu2cam)), -t_res) */
void vso_initial_pose(const float imu2cam[9], const float r_imu_res[9], const float t_res[3], float pose[7]);

/* VISystem::WarpFunctionSE3 (VISystem.cpp:1495-1558).  pts/out: n x 4 f32. */
void vso_warp(const float* pts, int n, const float pose[7], const vso_intr_t* K, float* out);

/* 6x6 float inverse as cv::Mat::inv() (DECOMP_LU) does it: returns 0 and zeros when singular. */
int vso_inv6(const float a[36], float out[36]);
void vso_tukey_weights(const float* r, int n, float* w);         /* VISystem.cpp:1797-1870 */
int vso_solve6(const float a[36], const float rhs[6], float out[6]); /* cv::solve(DECOMP_LU): what A.inv() * b evaluates to */

typedef struct {
    int first_lvl;      /* 3   VISystem.cpp:1119 */
    int last_lvl;       /* 0   VISystem.cpp:1120 */
    int max_iterations; /* 10  VISystem.cpp:1117 */
    float epsilon;      /* 0.001f VISystem.cpp:1115 */
    float z_factor;     /* 0.002f VISystem.cpp:1121 */
    int weight_mode;    /* 0 identity (reference, :1343), 1 Tukey (:1797-1826), 2 Huber (extension) */
    int sample_mode;    /* 0 nearest round() (reference, :1321), 1 bilinear (extension) */
    float huber_k;      /* Huber threshold in intensity levels (extension) */
} vso_gn_opts_t;

typedef struct {
    int lvl, iter, n_valid, updated;  /* updated = 1 when a pose update followed this evaluation */
    float error;                      /* (1/n) sum w r^2, VISystem.cpp:1347-1350 */
    float pose[7];                    /* pose AFTER this iteration (unchanged when updated == 0) */
    float delta[6];
} vso_gn_trace_t;

typedef struct {
    const uint8_t* prev_img[VSO_MAX_LEVELS];
    const uint8_t* cur_img[VSO_MAX_LEVELS];
    const int16_t* prev_gx[VSO_MAX_LEVELS];
    const int16_t* prev_gy[VSO_MAX_LEVELS];
    const float* cand[VSO_MAX_LEVELS]; /* n x 4 */
    int n_cand[VSO_MAX_LEVELS];
    int img_w[VSO_MAX_LEVELS], img_h[VSO_MAX_LEVELS]; /* actual Mat sizes (rows/cols tests, :1299) */
} vso_gn_frames_t;

/* VISystem::EstimatePoseFeatures (VISystem.cpp:1113-1448) without the drawing / imshow / waitKey /
 * debug warps (:1225-1266).  trace capacity = trace_cap entries; returns entries written. */
int vso_gn_solve(const vso_gn_frames_t* f, const vso_intr_t K[VSO_MAX_LEVELS], const float pose_in[7],
                 const vso_gn_opts_t* opts, float pose_out[7], vso_gn_trace_t* trace, int trace_cap);

/* Whole frame-pair tracking step as VISystemGPU::AddFrameGPU strings it together
 * (VISystemGPU.cpp:144-169): match -> candidates -> GN.  Pyramids/gradients are inputs. */
typedef struct {
    int n_good, n_sym;
    int n_cand[VSO_MAX_LEVELS];
    float pose[7];
} vso_track_result_t;

#ifdef __cplusplus
}
#endif
/* FAST-9/16 corners (cv::FAST as used inside cv::ORB / cv::cuda::ORB, Camera.cpp:124-129, CameraGPU.cpp:99-104):
 * up to cap corners (x, y) + score in row-major order; returns the total found.  oracle/fast.c */
int vso_fast9(const uint8_t* img, int w, int h, int pitch, int threshold, int nonmax, int32_t* out_xy, int32_t* out_score,
              int cap);

/* cv::ORB, one pyramid level (oracle/orb.c; pinned against cv2 4.13): the stages and the whole detectAndCompute */
void vso_orb_harris(const uint8_t* img, int w, int h, int pitch, const int32_t* xy, int n, float* resp);
void vso_orb_ic_angle(const uint8_t* img, int w, int h, int pitch, const int32_t* xy, int n, float* angle_deg);
void vso_orb_gauss_kernel(float k[7]);
void vso_orb_blur(const uint8_t* img, int w, int h, int pitch, uint8_t* out);
void vso_orb_describe(const uint8_t* blurred, int w, int h, int pitch, const int32_t* xy, const float* angle_deg, int n,
                      uint8_t* desc);
int vso_orb_detect_compute(const uint8_t* img, int w, int h, int pitch, int nfeatures, int fast_threshold, int32_t* out_xy,
                           float* out_resp, float* out_angle, uint8_t* out_desc, int cap);

/* cv::resize(INTER_LINEAR_EXACT) for 8-bit images and the multi-level cv::ORB built on it (oracle/orb.c) */
void vso_resize_linear_exact(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh);
float vso_orb_level_scale(float scale_factor, int level);
void vso_orb_level_budget(int nfeatures, float scale_factor, int nlevels, int* per_level);
int vso_orb_detect_compute_pyr(const uint8_t* img, int w, int h, int pitch, int nfeatures, float scale_factor, int nlevels,
                               int fast_threshold, float* out_xy, int32_t* out_octave, float* out_resp, float* out_angle,
                               uint8_t* out_desc, int cap);

#endif

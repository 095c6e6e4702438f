// oracle/refshim/ref_visystem_capi.cpp — C entry around the reference's own Camera / VISystem classes (TEST INFRASTRUCTURE).
// Built by `make -C oracle ref` together with /root/reference/src/{VISystem,Camera,CameraModel,Matcher,Plus,Imu}.cpp
// (unmodified, compiled where they lie) into oracle/_ref/libref_visystem.so.  It drives one frame pair through the
// reference's own front end and solver the way VISystemGPU::AddFrameGPU / VISystem::AddFrame sequence them
// (VISystemGPU.cpp:144-169, VISystem.cpp:286-330): Camera::Update -> computeGradient -> saveFrame (previous frame),
// Camera::Update -> computeGradient (current frame), good matches (given, or Camera::computeGoodMatches from given key
// points + descriptors), Camera::ObtainPatchesPointsPreviousFrame, VISystem::EstimatePoseFeatures.
// The per-iteration error the reference prints (VISystem.cpp:1351-1352) is parsed back from std::cout.
#include "VISystem.hpp"
#include <sstream>

namespace {
struct CoutCapture {
    std::streambuf* old;
    std::ostringstream sink;
    CoutCapture() : old(std::cout.rdbuf()) { std::cout.rdbuf(sink.rdbuf()); }
    ~CoutCapture() { std::cout.rdbuf(old); }
};
cv::Mat gray(const uint8_t* p, int w, int h) {
    cv::Mat m(h, w, CV_8UC1);
    memcpy(m.data, p, (size_t)w * h);
    return m;
}
cv::Mat intrinsics(const float K4[4]) {
    cv::Mat K = cv::Mat::eye(3, 3, CV_32FC1);
    K.at<float>(0, 0) = K4[0]; K.at<float>(1, 1) = K4[1]; K.at<float>(0, 2) = K4[2]; K.at<float>(1, 2) = K4[3];
    return K;
}
}  // namespace

extern "C" {

// Outputs (any may be NULL): pyr_prev / pyr_cur = the five levels back to back (tight rows); gx_prev / gy_prev likewise in
// int16; lvl_w / lvl_h [5]; cand = rows (x, y, z, 1) of levels 0..4 back to back, n_cand[5]; pose_out = qx qy qz qw tx ty tz;
// trace = (lvl, iteration, error) triples as printed, n_trace.  The kp/desc inputs are used when n_good < 0.
// Returns the number of out-of-bounds Mat::at accesses the solver made (undefined behaviour upstream, served as
// zeros by the shim; 0 for a well-defined run), or -1 with the message in err (the shim throws where OpenCV would).
int ref_track_pair(const uint8_t* img_prev, const uint8_t* img_cur, int w, int h, const float K4[4], int n_cells,
                   const float* good_prev_xy, const float* good_cur_xy, int n_good,
                   const float* kp_prev_xy, const uint8_t* desc_prev, int n_prev,
                   const float* kp_cur_xy, const uint8_t* desc_cur, int n_cur,
                   const float imu2cam[9], const float r_imu_res[9], const float t_res[3],
                   uint8_t* pyr_prev, uint8_t* pyr_cur, int16_t* gx_prev, int16_t* gy_prev, int* lvl_w, int* lvl_h,
                   float* cand, int cand_cap_rows, int* n_cand,
                   float* good_out_prev_xy, float* good_out_cur_xy, int* n_good_out,
                   float* pose_out, float* trace, int trace_cap, int* n_trace, char* err, int err_cap) {
    try {
        CoutCapture cap;
        vi::VISystem sys;
        sys.InitializeCamera(USE_ORB, USE_BRUTE_FORCE_HAMMING, w, h, n_cells, 5);
        sys.InitializePyramid(w, h, intrinsics(K4));

        sys.camera.Update(gray(img_prev, w, h));
        sys.camera.computeGradient();
        if (n_good < 0) {
            sys.camera.currentFrame->keypoints.resize(n_prev);
            for (int i = 0; i < n_prev; i++) sys.camera.currentFrame->keypoints[i].pt = cv::Point2f(kp_prev_xy[2 * i], kp_prev_xy[2 * i + 1]);
            cv::Mat d(n_prev, 32, CV_8UC1);
            memcpy(d.data, desc_prev, (size_t)n_prev * 32);
            sys.camera.currentFrame->descriptors = d;
        }
        sys.camera.saveFrame();
        Frame* prev = sys.camera.frameList.back();

        sys.camera.Update(gray(img_cur, w, h));
        sys.camera.computeGradient();
        Frame* cur = sys.camera.currentFrame;
        if (n_good < 0) {
            cur->keypoints.resize(n_cur);
            for (int i = 0; i < n_cur; i++) cur->keypoints[i].pt = cv::Point2f(kp_cur_xy[2 * i], kp_cur_xy[2 * i + 1]);
            cv::Mat d(n_cur, 32, CV_8UC1);
            memcpy(d.data, desc_cur, (size_t)n_cur * 32);
            cur->descriptors = d;
            sys.camera.computeGoodMatches();
        } else {
            prev->nextGoodMatches.resize(n_good);
            cur->prevGoodMatches.resize(n_good);
            for (int i = 0; i < n_good; i++) {
                prev->nextGoodMatches[i].pt = cv::Point2f(good_prev_xy[2 * i], good_prev_xy[2 * i + 1]);
                cur->prevGoodMatches[i].pt = cv::Point2f(good_cur_xy[2 * i], good_cur_xy[2 * i + 1]);
            }
        }
        if (n_good_out) {
            *n_good_out = (int)prev->nextGoodMatches.size();
            for (size_t i = 0; i < prev->nextGoodMatches.size(); i++) {
                if (good_out_prev_xy) { good_out_prev_xy[2 * i] = prev->nextGoodMatches[i].pt.x; good_out_prev_xy[2 * i + 1] = prev->nextGoodMatches[i].pt.y; }
                if (good_out_cur_xy) { good_out_cur_xy[2 * i] = cur->prevGoodMatches[i].pt.x; good_out_cur_xy[2 * i + 1] = cur->prevGoodMatches[i].pt.y; }
            }
        }
        sys.camera.ObtainPatchesPointsPreviousFrame();
        // EstimatePoseFeatures warps candidateDebugPoints too (VISystem.cpp:1225); with none, Mat::col(0) of an empty matrix throws
        sys.camera.ObtainDebugPointsPreviousFrame();

        size_t off = 0;
        for (int l = 0; l < 5; l++) {
            const cv::Mat& a = prev->grayImage[l];
            const cv::Mat& b = cur->grayImage[l];
            if (lvl_w) lvl_w[l] = a.cols;
            if (lvl_h) lvl_h[l] = a.rows;
            for (int r = 0; r < a.rows; r++) {
                if (pyr_prev) memcpy(pyr_prev + off + (size_t)r * a.cols, a.ptr(r), a.cols);
                if (pyr_cur) memcpy(pyr_cur + off + (size_t)r * a.cols, b.ptr(r), a.cols);
                if (gx_prev) memcpy(gx_prev + off + (size_t)r * a.cols, prev->gradientX[l].ptr(r), 2 * (size_t)a.cols);
                if (gy_prev) memcpy(gy_prev + off + (size_t)r * a.cols, prev->gradientY[l].ptr(r), 2 * (size_t)a.cols);
            }
            off += (size_t)a.rows * a.cols;
        }
        int crow = 0;
        for (int l = 0; l < 5; l++) {
            const cv::Mat& c = prev->candidatePoints[l];
            if (n_cand) n_cand[l] = c.rows;
            for (int r = 0; r < c.rows; r++, crow++)
                if (cand && crow < cand_cap_rows) memcpy(cand + 4 * (size_t)crow, c.ptr(r), 16);
        }

        // initial pose inputs (VISystem.cpp:1135-1165)
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                sys.imu2camRotation(i, j) = imu2cam[3 * i + j];
                sys.imuCore.residual_rotationMatrix(i, j) = r_imu_res[3 * i + j];
                sys.imuCore.final_rotationMatrix(i, j) = i == j ? 1.f : 0.f;
            }
        sys.TranslationResidual = cv::Mat::zeros(3, 1, CV_32FC1);
        for (int i = 0; i < 3; i++) sys.TranslationResidual.at<float>(i, 0) = t_res[i];

        cap.sink.str("");
        cv::Mat::oob_reads() = 0;
        sys.EstimatePoseFeatures(prev, cur);
        const long long oob = cv::Mat::oob_reads();
        for (int i = 0; i < 7; i++) pose_out[i] = prev->rigid_transformation_.data()[i];   // qx qy qz qw tx ty tz (se3.hpp data())

        // "lvl = <l>Error it <k> =<e>    Last Error it  =<le>"
        int nt = 0;
        std::istringstream in(cap.sink.str());
        std::string line;
        while (std::getline(in, line)) {
            int l, k;
            float e;
            if (sscanf(line.c_str(), "lvl = %dError it %d =%g", &l, &k, &e) == 3) {
                if (trace && nt < trace_cap) { trace[3 * nt] = (float)l; trace[3 * nt + 1] = (float)k; trace[3 * nt + 2] = e; }
                nt++;
            }
        }
        if (n_trace) *n_trace = nt;
        return (int)std::min<long long>(oob, 1 << 30);
    } catch (const std::exception& e) {
        if (err && err_cap > 0) { strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return -1;
    }
}

// VISystem::WarpFunctionSE3 on its own (VISystem.cpp:1495-1558): pts rows (x, y, z, 1), pose qx qy qz qw tx ty tz
int ref_warp(const float* pts, int n, const float pose[7], int w, int h, const float K4[4], int lvl, float* out) {
    try {
        CoutCapture cap;
        vi::VISystem sys;
        sys.InitializePyramid(w, h, intrinsics(K4));
        cv::Mat p(n, 4, CV_32FC1);
        memcpy(p.data, pts, (size_t)n * 16);
        vi::SE3 T;
        for (int i = 0; i < 7; i++) T.data()[i] = pose[i];
        cv::Mat r = sys.WarpFunctionSE3(p, T, lvl);
        for (int i = 0; i < n; i++) memcpy(out + 4 * (size_t)i, r.ptr(i), 16);
        return 0;
    } catch (const std::exception&) { return -1; }
}

// VISystem::TukeyFunctionWeights (VISystem.cpp:1797-1826) on a residual column
int ref_tukey(const float* r, int n, float* wout) {
    try {
        CoutCapture cap;
        vi::VISystem sys;
        cv::Mat R(n, 1, CV_32FC1);
        memcpy(R.data, r, (size_t)n * 4);
        cv::Mat W = sys.TukeyFunctionWeights(R);
        for (int i = 0; i < n; i++) wout[i] = W.at<float>(i, 0);
        return 0;
    } catch (const std::exception&) { return -1; }
}
}

// Camera.hpp — drop-in mirror of the reference's Frame / Camera / CameraGPU (include/Camera.hpp:32-173,
// include/CameraGPU.hpp:18-40).  A Frame keeps its pyramid, Scharr gradients and candidate points on the device
// (packed-pyramid layout of include/vislam_b200.h); the reference's public cv::Mat vectors are filled from the
// device when Camera::mirror_host is set (default on: a caller that reads frame->grayImage[l] keeps working).
//
// Feature detection / description (Camera::detectFeatures, detectAndComputeFeatures): with the ORB
// detector (USE_ORB, what the reference's CPU and GPU front ends select for binary descriptors) the frame's level-0 image,
// already on the device, goes through vsb_orb_detect_compute_pyr — cv::ORB::create(orb_nfeatures) exactly (8 levels, factor
// 1.2; 200 features in Camera as src/Camera.cpp:127, 1000 in CameraGPU as src/CameraGPU.cpp:99), key points level by level in
// row-major order.  The other detectors (KAZE/AKAZE/SIFT/SURF) are not part of this library: their features enter through
// Camera::setFeatures or a Camera::featureProvider callback (either also overrides ORB), and the detect* methods report what
// was provided.
#ifndef VISLAM_CAMERA_HPP_
#define VISLAM_CAMERA_HPP_

#include <functional>
#include <memory>
#include <vector>

#include "vislam/Matcher.hpp"
#include "vislam/compat.hpp"
#include "vislam/device.hpp"

// include/Camera.hpp:20-27
enum detectorType { USE_KAZE, USE_AKAZE, USE_ORB, USE_SIFT, USE_SURF };

class Frame {
public:
    Frame();
    ~Frame();

    std::vector<cv::Mat> grayImage = std::vector<cv::Mat>(5);
    std::vector<cv::Mat> gradientX = std::vector<cv::Mat>(5);
    std::vector<cv::Mat> gradientY = std::vector<cv::Mat>(5);
    std::vector<cv::Mat> gradient = std::vector<cv::Mat>(5);

    std::vector<cv::KeyPoint> keypoints;
    std::vector<cv::KeyPoint> prevGoodMatches;
    std::vector<cv::KeyPoint> nextGoodMatches;
    cv::Mat descriptors;
    std::vector<cv::Mat> prevPatches;
    std::vector<cv::Mat> nextPatches;
    std::vector<cv::Mat> candidatePoints = std::vector<cv::Mat>(5);
    std::vector<cv::KeyPoint> debugKeypoints;
    std::vector<cv::Mat> candidateDebugPoints = std::vector<cv::Mat>(5);

    int idFrame;
    double imageTime;
    vi::SE3 rigid_transformation_;

    bool obtainedGradients;
    bool obtainedGoodMatches;
    bool isKeyFrame;

    // ---- device state (additions) ----
    vsb_pyr_layout_t layout;          // packed pyramid layout of this frame (5 levels)
    vi::DevBuf d_pyr, d_gx, d_gy, d_gmag, d_cand, d_ncand;
    int cand_cap;                     // rows per level in d_cand
    int n_cand[VSB_MAX_LEVELS];       // candidate rows per level (host copy)
    bool pyr_on_device, grad_on_device, cand_on_device;
};

class Camera {
public:
    Camera();
    Camera(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path);
    virtual ~Camera() {}
    void initializate(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path);
    void Update(cv::Mat _grayImage);

    void setDetector(int _detector);
    int detectFeatures();
    int detectAndComputeFeatures();

    void setMatcher(int _matcher);
    void computeDescriptors();
    void computeGoodMatches();

    void computeGradient();

    bool addKeyframe();

    void saveFrame();
    void ObtainDebugPointsPreviousFrame();
    void ObtainPatchesPointsPreviousFrame();

    void printStatistics();

    std::vector<Frame*> frameList;
    std::vector<cv::Mat> Residuals;
    std::vector<cv::DMatch> goodMatches;
    Frame* currentFrame;

    int w_residual;
    int h_residual;
    int detectorType;
    int matcherType;

    Matcher matcher;
    int nPointsDetect;
    int nBestMatches;
    int n_cells;

    std::vector<int> w_size = std::vector<int>(5);
    std::vector<int> h_size = std::vector<int>(5);
    int w_patch, h_patch;

    double elapsed_detect;
    double elapsed_descriptors;
    double elapsed_computeGoodMatches;
    double elapsed_computeGradient;
    double elapsed_computePatches;

    double elapsed_detect_mean;
    double elapsed_descriptors_mean;
    double elapsed_computeGoodMatches_mean;
    double elapsed_computeGradient_mean;
    double elapsed_computePatches_mean;
    double nPointsDetect_mean;
    double nBestMatches_mean;
    int num_images;

    double elapsed_detect_sum;
    double elapsed_descriptors_sum;
    double elapsed_computeGoodMatches_sum;
    double elapsed_computeGradient_sum;
    double elapsed_computePatches_sum;
    double nPointsDetect_sum;
    double nBestMatches_sum;

    // ---- additions (not in the reference) ----
    // key points + descriptors of the current frame, replacing detector->detectAndCompute (Camera.cpp:84-93)
    void setFeatures(const std::vector<cv::KeyPoint>& keypoints, const cv::Mat& descriptors);
    std::function<void(const cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors)> featureProvider;
    bool mirror_host;   // fill the public cv::Mat members of Frame from the device (default true)
    int orb_nfeatures;  // cv::ORB::create(n): 200 in Camera (src/Camera.cpp:127), 1000 in CameraGPU (src/CameraGPU.cpp:99)
    int detectOrbOnDevice(bool describe);   // fills currentFrame->keypoints (and descriptors); returns the key-point count
    bool verbose;

protected:
    void match_with(Matcher& m, bool gpu_entry);
    void stats_accumulate();
    std::shared_ptr<struct CameraOrbBuffers> orb_bufs;   // device outputs of detectOrbOnDevice, kept between frames (six allocations per frame otherwise)
};

// include/CameraGPU.hpp:18-40
class CameraGPU : public Camera {
public:
    CameraGPU();
    CameraGPU(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_patch);
    void initializateCameraGPU(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_patch);
    void setGPUDetector(int _detector);
    void detectGPUFeatures();
    void setGPUMatcher(int _matcher);
    int detectAndComputeGPUFeatures();
    void computeGPUGoodMatches();
    bool addGPUKeyframe();

    MatcherGPU matcherGPU;
    bool useGPU;
};

#endif

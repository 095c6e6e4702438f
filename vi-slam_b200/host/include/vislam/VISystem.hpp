// VISystem.hpp — drop-in mirror of vi::VISystem / vi::VISystemGPU for the frame-tracking path
// (include/VISystem.hpp:40-148, include/VISystemGPU.hpp:17-30): InitializePyramid, EstimatePoseFeatures,
// WarpFunctionSE3, IdentityWeights, Track, AddFrame / AddFrameGPU and the public pose state.
//
// VISystemGPU::InitializeSystemGPU (VISystemGPU.cpp:39-129) reads the reference's calibration XML (vi::CameraModel, without
// undistortion) and sets up the pyramid, the initial pose state, imuCore and the camera; AddFrame / AddFrameGPU run
// imuCore.setImuData + estimate() on the IMU samples they are given and feed residual_rotationMatrix into the GN prior
// (VISystemGPU.cpp:148-149, VISystem.cpp:1135).  A caller that has no samples (empty vectors) sets RotationResidualImu
// itself.  Out of scope: undistortion, RANSAC / triangulation experiments (F2FRansac, Triangulate, ...),
// drawing.  TranslationResidual comes from setGtRes, as upstream.
#ifndef VISLAM_VISYSTEM_HPP_
#define VISLAM_VISYSTEM_HPP_

#include <string>
#include <vector>

#include "vislam/Camera.hpp"
#include "vislam/CameraModel.hpp"
#include "vislam/Imu.hpp"
#include "vislam/Plus.hpp"
#include "vislam/compat.hpp"
#include "vislam/device.hpp"

#define PYRAMID_LEVELS 5


namespace vi {

class VISystem {
public:
    VISystem();
    VISystem(int argc, char* argv[]);
    virtual ~VISystem();

    void InitializeCamera(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path);
    void InitializePyramid(int _width, int _height, cv::Mat _K);
    void EstimatePoseFeatures(Frame* _previous_frame, Frame* _current_frame);
    bool AddFrame(cv::Mat _currentImage, std::vector<cv::Point3d> _imuAngularVelocity,
                  std::vector<cv::Point3d> _imuAcceleration);
    bool AddFrame(cv::Mat _currentImage, std::vector<cv::Point3d> _imuAngularVelocity,
                  std::vector<cv::Point3d> _imuAcceleration, cv::Point3d _gtPosition);
    void FreeLastFrame();
    void Track();
    cv::Mat IdentityWeights(int _num_residuals);
    cv::Mat TukeyFunctionWeights(cv::Mat _input);
    cv::Mat WarpFunctionSE3(cv::Mat _points2warp, SE3 _rigid_transformation, int _lvl);
    void setGtRes(cv::Mat TranslationResGT, cv::Mat RotationGT);
    void Calibration(std::string _calibration_path);        // VISystem.cpp:208-221

    bool initialized, distortion_valid, depth_available;
    int num_keyframes;
    int num_max_keyframes;
    int min_features;
    int start_index;

    int h, w, h_input, w_input;
    float fx, fy, cx, cy;

    cv::Point3d positionImu, velocityImu, accImu;
    Quaterniond qOrientationImu;
    cv::Point3d RPYOrientationImu;

    cv::Point3d positionCam, velocityCam, accCam;
    Quaterniond qOrientationCam;
    cv::Point3d RPYOrientationCam;

    cv::Matx33f imu2camRotation;
    cv::Point3d imu2camTranslation;
    cv::Mat imu2camTransformation, world2imuTransformation;
    cv::Matx33f world2imuRotation;
    CameraModel* camera_model;
    Imu imuCore;

    SE3 final_poseCam;
    SE3 final_poseImu;
    SE3 current_poseCam;
    SE3 current_poseImu;
    cv::Point3d prev_gtPosition, current_gtPosition, current_gtTraslation;

    Camera camera;

    cv::Mat currentImage, prevImage;
    cv::Mat K;

    std::vector<int> w_ = std::vector<int>(PYRAMID_LEVELS);
    std::vector<int> h_ = std::vector<int>(PYRAMID_LEVELS);
    std::vector<float> fx_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> fy_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> cx_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> cy_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> invfx_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> invfy_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> invcx_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<float> invcy_ = std::vector<float>(PYRAMID_LEVELS);
    std::vector<cv::Mat> K_ = std::vector<cv::Mat>(PYRAMID_LEVELS);

    cv::Mat TranslationResidual;
    cv::Matx33f RotationResidual;
    cv::Matx33f RotationResCam;
    cv::Matx33f init_rotationMatrix, final_rotationMatrix;
    cv::Point3f translationResEst;
    int nPointsLastKeyframe;
    int nPointsCurrentImage;
    bool lastImageWasKeyframe;
    bool currentImageIsKeyframe;

    // ---- additions (not in the reference) ----
    cv::Matx33f RotationResidualImu;    // stands in for imuCore.residual_rotationMatrix (VISystem.cpp:1135)
    vsb_gn_opts_t gn_options;           // the literals of VISystem.cpp:1115-1121 by default
    bool track_from_estimate;           // Track() composes the GN estimate instead of RotationResCam / translationResEst
                                        // (the reference ignores the GN result there); default false
    std::vector<vsb_gn_trace_t> last_trace;   // per-iteration record of the last EstimatePoseFeatures (when keep_trace)
    bool keep_trace;
    bool verbose;
    bool imu_ready;                     // imuCore has its initial orientation (first batch of samples seen)

protected:
    void update_imu_prior(std::vector<cv::Point3d>& w, std::vector<cv::Point3d>& a);
    void fill_intrinsics(vsb_intr_t out[VSB_MAX_LEVELS]) const;
    virtual Camera& active_camera() { return camera; }   // the camera whose frameList Track() reads
    DevBuf d_pose_, d_trace_, d_ntrace_, d_cand_, d_ncand_, d_pts_;
};

class VISystemGPU : public VISystem {
public:
    VISystemGPU();
    VISystemGPU(int argc, char* argv[]);
    ~VISystemGPU();
    void InitializeSystemGPU(std::string _calPath, cv::Point3d _iniPosition, cv::Point3d _iniVelocity, cv::Point3d _iniRPY,
                             cv::Mat image);
    void InitializeCameraGPU(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path);
    void AddFrameGPU(cv::Mat _currentImage, std::vector<cv::Point3d> _imuAngularVelocity,
                     std::vector<cv::Point3d> _imuAcceleration);
    void FreeLastFrameGPU();

    CameraGPU cameraGPU;

protected:
    Camera& active_camera() override { return cameraGPU; }
};

}  // namespace vi

#endif

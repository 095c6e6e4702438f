// CameraModel.hpp — the calibration / configuration file of the reference (include/CameraModel.hpp:20-146,
// src/CameraModel.cpp:16-101): an OpenCV FileStorage XML with the keys listed in SURVEY.md §5, read here by a small parser
// (OpenCV is not a dependency of this library).  What the frame-tracking path needs of it: image size, the pinhole
// intrinsics, imu2cam0Transformation, the IMU rate and the system knobs (num_cells, detector, matcher, ...).
// Undistortion (getOptimalNewCameraMatrix / initUndistortRectifyMap / remap, CameraModel.cpp:85-103) is OpenCV calib3d work
// upstream of the path and is NOT done here: a file with non-zero `rectification` coefficients is refused with an
// exception — feed undistorted frames with their pinhole calibration.
#ifndef VISLAM_CAMERAMODEL_HPP_
#define VISLAM_CAMERAMODEL_HPP_
#include <string>

#include "compat.hpp"

namespace vi {

class CameraModel {
public:
    CameraModel();
    ~CameraModel();
    void GetCameraModel(std::string _calibrationPath);      // throws std::runtime_error where the reference exit()s
    const cv::Mat& GetK() const { return output_intrinsic_camera_; }
    const cv::Mat& GetOriginalK() const { return original_intrinsic_camera_; }
    int GetOutputWidth() const { return out_width_; }
    int GetOutputHeight() const { return out_height_; }
    int GetInputWidth() const { return in_width_; }
    int GetInputHeight() const { return in_height_; }
    bool IsValid() const { return valid_; }                 // "rectification on": always false here (see above)

    float camera_frecuency;
    float imu_frecuency;
    int min_features;
    int num_max_keyframes;
    int start_index;
    int use_gt;
    int use_ros;
    int detector, matcher;
    int num_cells, length_patch;
    cv::Mat imu2cam0Transformation;

private:
    cv::Mat output_intrinsic_camera_, original_intrinsic_camera_;
    float input_calibration_[4];
    float dist_coeffs_[4];
    int out_width_, out_height_;
    int in_width_, in_height_;
    bool valid_;
};

}  // namespace vi
#endif

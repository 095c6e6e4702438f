// CameraModel.cpp — reads the reference's calibration XML (src/CameraModel.cpp:16-101; format: cv::FileStorage, e.g.
// calibration/calibrationEUROC.xml) without OpenCV.
#include "vislam/CameraModel.hpp"

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace vi {
namespace {

// text between <key ...> and </key> with XML comments removed; empty when the key is absent
std::string element(const std::string& xml, const std::string& key) {
    size_t a = 0;
    while ((a = xml.find("<" + key, a)) != std::string::npos) {
        const char next = xml[a + key.size() + 1];
        if (next == '>' || next == ' ' || next == '\t' || next == '\n') break;
        a += key.size();
    }
    if (a == std::string::npos) return "";
    const size_t open_end = xml.find('>', a);
    const size_t b = xml.find("</" + key + ">", open_end);
    if (open_end == std::string::npos || b == std::string::npos) return "";
    return xml.substr(open_end + 1, b - open_end - 1);
}

std::string strip_comments(const std::string& s) {
    std::string out;
    size_t i = 0;
    while (i < s.size()) {
        const size_t c = s.find("<!--", i);
        if (c == std::string::npos) { out += s.substr(i); break; }
        out += s.substr(i, c - i);
        const size_t e = s.find("-->", c);
        if (e == std::string::npos) break;
        i = e + 3;
    }
    return out;
}

bool scalar(const std::string& xml, const std::string& key, double& v) {
    const std::string e = element(xml, key);
    if (e.empty()) return false;
    char* end = nullptr;
    v = std::strtod(e.c_str(), &end);
    return end != e.c_str();
}

// <key type_id="opencv-matrix"><rows>r</rows><cols>c</cols><dt>f</dt><data>...</data></key> -> CV_32F matrix
cv::Mat matrix(const std::string& xml, const std::string& key) {
    const std::string e = element(xml, key);
    double r = 0, c = 0;
    if (e.empty() || !scalar(e, "rows", r) || !scalar(e, "cols", c)) return cv::Mat();
    std::istringstream in(element(e, "data"));
    cv::Mat m = cv::Mat::zeros((int)r, (int)c, CV_32FC1);
    for (int i = 0; i < (int)r; i++)
        for (int j = 0; j < (int)c; j++) {
            double v = 0;
            if (!(in >> v)) throw std::runtime_error("calibration file: matrix '" + key + "' has too few values");
            m.at<float>(i, j) = (float)v;
        }
    return m;
}

}  // namespace

CameraModel::CameraModel()
    : camera_frecuency(0), imu_frecuency(0), min_features(0), num_max_keyframes(0), start_index(0), use_gt(0), use_ros(0),
      detector(0), matcher(0), num_cells(0), length_patch(0), out_width_(0), out_height_(0), in_width_(0), in_height_(0),
      valid_(false) {
    for (int i = 0; i < 4; i++) input_calibration_[i] = dist_coeffs_[i] = 0.f;
}
CameraModel::~CameraModel() {}

void CameraModel::GetCameraModel(std::string _calibration_path) {
    std::ifstream f(_calibration_path.c_str());
    if (!f) throw std::runtime_error("calibration file not found: " + _calibration_path + " (cannot operate without calibration)");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string xml = strip_comments(ss.str());
    double v = 0;
    auto geti = [&](const char* key, int& dst) { if (scalar(xml, key, v)) dst = (int)v; };
    auto getf = [&](const char* key, float& dst) { if (scalar(xml, key, v)) dst = (float)v; };
    geti("in_width", in_width_); geti("in_height", in_height_);                       // CameraModel.cpp:25-42
    geti("out_width", out_width_); geti("out_height", out_height_);
    const cv::Mat calibration_values = matrix(xml, "calibration_values");
    const cv::Mat distortion_values = matrix(xml, "rectification");
    imu2cam0Transformation = matrix(xml, "imu2cam0Transformation");
    getf("camera_frecuency", camera_frecuency); getf("imu_frecuency", imu_frecuency);
    geti("min_features", min_features); geti("num_max_keyframes", num_max_keyframes); geti("start_index", start_index);
    geti("use_gt", use_gt); geti("use_ros", use_ros); geti("num_cells", num_cells); geti("length_patch", length_patch);
    geti("detector", detector); geti("matcher", matcher);
    if (calibration_values.rows != 1 || calibration_values.cols != 4) throw std::runtime_error("calibration file: calibration_values must be 1 x 4");
    if (imu2cam0Transformation.rows != 4 || imu2cam0Transformation.cols != 4) throw std::runtime_error("calibration file: imu2cam0Transformation must be 4 x 4");
    for (int i = 0; i < 4; i++) {
        input_calibration_[i] = calibration_values.at<float>(0, i);
        dist_coeffs_[i] = (distortion_values.rows == 1 && distortion_values.cols == 4) ? distortion_values.at<float>(0, i) : 0.f;
    }
    if (input_calibration_[2] < 1 && input_calibration_[3] < 1) {                      // normalised intrinsics, :61-69
        input_calibration_[0] *= in_width_; input_calibration_[1] *= in_height_;
        input_calibration_[2] *= in_width_; input_calibration_[3] *= in_height_;
    }
    original_intrinsic_camera_ = cv::Mat::zeros(3, 3, CV_32FC1);                      // :72-76
    original_intrinsic_camera_.at<float>(0, 0) = input_calibration_[0];
    original_intrinsic_camera_.at<float>(1, 1) = input_calibration_[1];
    original_intrinsic_camera_.at<float>(0, 2) = input_calibration_[2];
    original_intrinsic_camera_.at<float>(1, 2) = input_calibration_[3];
    original_intrinsic_camera_.at<float>(2, 2) = 1.f;
    if (dist_coeffs_[0] != 0.f)                                                        // :79-103 would rectify
        throw std::runtime_error("calibration file carries distortion coefficients: undistortion (OpenCV calib3d, "
                                 "CameraModel.cpp:85-103) is outside this library — feed undistorted frames and zero `rectification`");
    valid_ = false;
    output_intrinsic_camera_ = original_intrinsic_camera_;
}

}  // namespace vi

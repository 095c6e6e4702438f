// VISystem.cpp — class mirror of vi::VISystem / vi::VISystemGPU for the tracking path (src/VISystem.cpp,
// src/VISystemGPU.cpp).
//   InitializePyramid      VISystem.cpp:1451-1493   -> vsb_init_pyramid (host arithmetic of the C ABI)
//   EstimatePoseFeatures   VISystem.cpp:1113-1448   -> vsb_initial_pose + vsb_gn_solve (one kernel: all levels / iterations)
//   WarpFunctionSE3        VISystem.cpp:1495-1558   -> vsb_warp_se3
//   Track                  VISystem.cpp:1567-1635   -> vsb_se3_from_rt + vsb_se3_mul
//   AddFrame / AddFrameGPU VISystem.cpp:290-299 / VISystemGPU.cpp:144-169: Update -> keyframe -> estimate -> Track
// No CPU fallback: vi::DeviceError without a CUDA device.
#include "vislam/VISystem.hpp"

#include <cmath>
#include <iostream>

using cv::Mat;
using cv::Matx33f;
using cv::Point3d;
using std::vector;

namespace vi {

VISystem::VISystem()
    : initialized(false), distortion_valid(false), depth_available(false), num_keyframes(0), num_max_keyframes(10),
      min_features(0), start_index(0), h(0), w(0), h_input(0), w_input(0), fx(0), fy(0), cx(0), cy(0),
      imu2camRotation(Matx33f::eye()), world2imuRotation(Matx33f::eye()), camera_model(nullptr), RotationResidual(Matx33f::eye()), RotationResCam(Matx33f::eye()),
      init_rotationMatrix(Matx33f::eye()), final_rotationMatrix(Matx33f::eye()), nPointsLastKeyframe(0),
      nPointsCurrentImage(0), lastImageWasKeyframe(false), currentImageIsKeyframe(false),
      RotationResidualImu(Matx33f::eye()),
      track_from_estimate(false), keep_trace(false), verbose(false), imu_ready(false) {
    vsb_gn_default_opts(&gn_options);
    TranslationResidual = Mat::zeros(3, 1, CV_32F);
}

VISystem::VISystem(int, char*[]) : VISystem() {}   // VISystem.cpp:36-46: the reference only starts a ROS node here

VISystem::~VISystem() { delete camera_model; }

void VISystem::Calibration(std::string _calibration_path) {   // VISystem.cpp:208-221 (throws where the reference exit()s)
    delete camera_model;
    camera_model = new CameraModel();
    camera_model->GetCameraModel(_calibration_path);
    w = camera_model->GetOutputWidth();
    h = camera_model->GetOutputHeight();
    if (w % 2 != 0 || h % 2 != 0) throw std::invalid_argument("Output image dimensions must be multiples of 32");   // :216-220
}

// imuCore.setImuData + estimate() (VISystem.cpp:284-285, VISystemGPU.cpp:142-143): the residual rotation of the frame interval
// becomes the rotation prior of the Gauss-Newton solve (VISystem.cpp:1135).  The first batch of samples also gives the
// filter its initial orientation — Imu::initializate(yaw, velocity, w, a), which InitializeSystem calls upstream
// (VISystem.cpp:143); VISystemGPU.cpp:119 names a one-argument overload that does not exist.  Empty vectors leave
// RotationResidualImu as the caller set it.
void VISystem::update_imu_prior(vector<Point3d>& w_meas, vector<Point3d>& a_meas) {
    if (w_meas.empty() || a_meas.empty()) return;
    if (imuCore.timeStep <= 0) imuCore.createPublisher(1.0 / 200.0);        // calibration default (imu_frecuency 200)
    if (!imu_ready) {
        imuCore.initializate(RPYOrientationImu.z, velocityImu, w_meas, a_meas);
        imu_ready = true;
    }
    imuCore.setImuData(w_meas, a_meas);
    imuCore.estimate();
    RotationResidualImu = imuCore.residual_rotationMatrix;
}

void VISystem::InitializeCamera(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path) {
    camera.initializate(_detector, _matcher, _w_size, _h_size, _num_cells, _length_path);   // VISystem.cpp:1637-1640
}

void VISystem::InitializePyramid(int _width, int _height, Mat _K) {   // VISystem.cpp:1451-1493
    if (_K.empty() || _K.type() != CV_32F || _K.rows < 3 || _K.cols < 3)
        throw std::invalid_argument("VISystem::InitializePyramid: K must be a 3x3 CV_32FC1 matrix");
    vsb_intr_t I[VSB_MAX_LEVELS];
    const int rc = vsb_init_pyramid(_width, _height, _K.at<float>(0, 0), _K.at<float>(1, 1), _K.at<float>(0, 2),
                                    _K.at<float>(1, 2), I);
    if (rc != VSB_OK) throw std::invalid_argument("VISystem::InitializePyramid: bad image size");
    w = _width; h = _height;
    fx = I[0].fx; fy = I[0].fy; cx = I[0].cx; cy = I[0].cy;
    K = _K;
    for (int lvl = 0; lvl < PYRAMID_LEVELS; lvl++) {
        w_[lvl] = I[lvl].w; h_[lvl] = I[lvl].h;
        fx_[lvl] = I[lvl].fx; fy_[lvl] = I[lvl].fy; cx_[lvl] = I[lvl].cx; cy_[lvl] = I[lvl].cy;
        invfx_[lvl] = I[lvl].invfx; invfy_[lvl] = I[lvl].invfy;
        invcx_[lvl] = 1 / cx_[lvl]; invcy_[lvl] = 1 / cy_[lvl];
        if (lvl == 0) { K_[0] = _K; continue; }
        K_[lvl] = Mat::zeros(3, 3, CV_32F);
        K_[lvl].at<float>(0, 0) = fx_[lvl]; K_[lvl].at<float>(1, 1) = fy_[lvl]; K_[lvl].at<float>(2, 2) = 1.f;
        K_[lvl].at<float>(0, 2) = cx_[lvl]; K_[lvl].at<float>(1, 2) = cy_[lvl];
    }
}

void VISystem::fill_intrinsics(vsb_intr_t out[VSB_MAX_LEVELS]) const {
    for (int l = 0; l < VSB_MAX_LEVELS; l++) {
        out[l].fx = fx_[l]; out[l].fy = fy_[l]; out[l].cx = cx_[l]; out[l].cy = cy_[l];
        out[l].invfx = invfx_[l]; out[l].invfy = invfy_[l];
        out[l].w = w_[l]; out[l].h = h_[l];
    }
}

void VISystem::setGtRes(Mat TranslationResGT, Mat RotationResGT) {   // VISystem.cpp:415-419
    TranslationResidual = TranslationResGT;
    // RPY2rotationMatrix(rotationMatrix2RPY(R)): re-orthonormalises through the Euler angles (Plus.cpp:56-83,182-220)
    if (!RotationResGT.empty() && RotationResGT.type() == CV_32F && RotationResGT.rows >= 3 && RotationResGT.cols >= 3) {
        const double r11 = RotationResGT.at<float>(0, 0), r21 = RotationResGT.at<float>(1, 0), r31 = RotationResGT.at<float>(2, 0);
        const double r32 = RotationResGT.at<float>(2, 1), r33 = RotationResGT.at<float>(2, 2);
        const double yaw = std::atan2(r21, r11), pitch = std::atan2(-r31, std::sqrt(r32 * r32 + r33 * r33)), roll = std::atan2(r32, r33);
        const double c1 = std::cos(roll), s1 = std::sin(roll), c2 = std::cos(pitch), s2 = std::sin(pitch), c3 = std::cos(yaw), s3 = std::sin(yaw);
        RotationResidual = Matx33f((float)(c3 * c2), (float)(c3 * s2 * s1 - s3 * c1), (float)(c3 * s2 * c1 + s3 * s1),
                                   (float)(s3 * c2), (float)(s3 * s2 * s1 + c3 * c1), (float)(s3 * s2 * c1 - c3 * s1),
                                   (float)(-s2), (float)(c2 * s1), (float)(c2 * c1));
    }
}

void VISystem::EstimatePoseFeatures(Frame* prev, Frame* cur) {   // VISystem.cpp:1113-1448
    if (!prev || !cur) throw std::invalid_argument("VISystem::EstimatePoseFeatures: null frame");
    if (!prev->pyr_on_device || !cur->pyr_on_device)
        throw std::logic_error("VISystem::EstimatePoseFeatures: frames must come from Camera::Update");
    if (TranslationResidual.empty() || TranslationResidual.type() != CV_32F || TranslationResidual.total() < 3)
        throw std::invalid_argument("VISystem::EstimatePoseFeatures: TranslationResidual must hold 3 floats (setGtRes)");
    Device& dev = Device::get();
    void* st = dev.stream();
    // initial pose: SE3(RPY2rotationMatrix(-rotationMatrix2RPY(imu2cam^T R_imu_res imu2cam)), (-sx,-sy,-sz)), :1135-1162
    const float t_res[3] = {TranslationResidual.ptr<float>()[0],
                            TranslationResidual.rows >= 3 ? TranslationResidual.at<float>(1, 0) : TranslationResidual.ptr<float>()[1],
                            TranslationResidual.rows >= 3 ? TranslationResidual.at<float>(2, 0) : TranslationResidual.ptr<float>()[2]};
    float pose0[7];
    dev.check(vsb_initial_pose(imu2camRotation.val, RotationResidualImu.val, t_res, pose0), "vsb_initial_pose");

    // candidate points of the previous frame: on the device when Camera::ObtainPatchesPointsPreviousFrame made
    // them, otherwise taken from the public Mats (a caller may have filled candidatePoints by hand)
    const float* d_cand = nullptr;
    const int32_t* d_ncand = nullptr;
    int cand_cap = 0;
    if (prev->cand_on_device) {
        d_cand = prev->d_cand.as<float>();
        d_ncand = prev->d_ncand.as<int32_t>();
        cand_cap = prev->cand_cap;
    } else {
        int32_t nc[VSB_MAX_LEVELS];
        for (int l = 0; l < VSB_MAX_LEVELS; l++) {
            const Mat& m = prev->candidatePoints[l];
            if (!m.empty() && (m.type() != CV_32F || m.cols != 4))
                throw std::invalid_argument("VISystem::EstimatePoseFeatures: candidatePoints must be N x 4 CV_32FC1");
            nc[l] = m.empty() ? 0 : m.rows;
            if (nc[l] > cand_cap) cand_cap = nc[l];
        }
        if (cand_cap == 0) cand_cap = 1;
        float* dc = static_cast<float*>(d_cand_.reserve(sizeof(float) * 4 * (size_t)cand_cap * VSB_MAX_LEVELS));
        int32_t* dn = static_cast<int32_t*>(d_ncand_.reserve(sizeof(nc)));
        for (int l = 0; l < VSB_MAX_LEVELS; l++)
            if (nc[l] > 0)
                dev.check(vsb_upload_2d(dev.ctx(), dc + (size_t)4 * cand_cap * l, 16, prev->candidatePoints[l].data,
                                        prev->candidatePoints[l].step, 16, (size_t)nc[l], st), "candidate upload");
        dev.check(vsb_upload(dev.ctx(), dn, nc, sizeof(nc), st), "upload");
        d_cand = dc;
        d_ncand = dn;
    }
    vsb_gn_opts_t o = gn_options;
    if (!prev->grad_on_device) o.grad_mode = 1;   // no gradient images: evaluate the same Scharr at the candidate points
    vsb_intr_t I[VSB_MAX_LEVELS];
    fill_intrinsics(I);
    float* d_pose = static_cast<float*>(d_pose_.reserve(sizeof(float) * 14));
    vsb_gn_trace_t* d_trace = keep_trace ? static_cast<vsb_gn_trace_t*>(d_trace_.reserve(sizeof(vsb_gn_trace_t) * VSB_MAX_TRACE)) : nullptr;
    int32_t* d_ntrace = keep_trace ? static_cast<int32_t*>(d_ntrace_.reserve(sizeof(int32_t))) : nullptr;
    dev.check(vsb_upload(dev.ctx(), d_pose, pose0, sizeof(pose0), st), "upload");
    dev.check(vsb_gn_solve(dev.ctx(), prev->d_pyr.as<uint8_t>(), cur->d_pyr.as<uint8_t>(),
                           prev->grad_on_device ? prev->d_gx.as<int16_t>() : nullptr,
                           prev->grad_on_device ? prev->d_gy.as<int16_t>() : nullptr, 0, &prev->layout, d_cand, cand_cap,
                           d_ncand, I, d_pose, &o, 1, d_pose + 7, d_trace, d_ntrace, st), "vsb_gn_solve");
    float pose[7];
    dev.check(vsb_download(dev.ctx(), pose, d_pose + 7, sizeof(pose), st), "download");
    last_trace.clear();
    if (keep_trace) {
        int32_t nt = 0;
        dev.check(vsb_download(dev.ctx(), &nt, d_ntrace, sizeof(nt), st), "download");
        dev.sync();
        last_trace.resize((size_t)nt);
        if (nt > 0) dev.check(vsb_download(dev.ctx(), last_trace.data(), d_trace, sizeof(vsb_gn_trace_t) * nt, st), "download");
    }
    dev.sync();
    prev->rigid_transformation_ = SE3(pose);   // :1445
}

Mat VISystem::IdentityWeights(int _num_residuals) { return Mat::ones(_num_residuals, 1, CV_32F); }   // VISystem.cpp:1561-1565

Mat VISystem::TukeyFunctionWeights(Mat) {   // VISystem.cpp:1797-1826: its only call site is commented out upstream (:1344)
    throw std::logic_error("VISystem::TukeyFunctionWeights: dead code in the reference (VISystem.cpp:1344); not provided");
}

Mat VISystem::WarpFunctionSE3(Mat _points2warp, SE3 _rigid_transformation, int _lvl) {   // VISystem.cpp:1495-1558
    if (_lvl < 0 || _lvl >= PYRAMID_LEVELS) throw std::invalid_argument("VISystem::WarpFunctionSE3: bad level");
    if (_points2warp.empty()) return Mat();
    if (_points2warp.type() != CV_32F || _points2warp.cols != 4)
        throw std::invalid_argument("VISystem::WarpFunctionSE3: points must be N x 4 CV_32FC1");
    Device& dev = Device::get();
    void* st = dev.stream();
    const int n = _points2warp.rows;
    float* d = static_cast<float*>(d_pts_.reserve(sizeof(float) * 8 * (size_t)n));
    dev.check(vsb_upload_2d(dev.ctx(), d, 16, _points2warp.data, _points2warp.step, 16, (size_t)n, st), "upload");
    vsb_intr_t I[VSB_MAX_LEVELS];
    fill_intrinsics(I);
    dev.check(vsb_warp_se3(dev.ctx(), d, n, _rigid_transformation.data(), &I[_lvl], d + (size_t)4 * n, st), "vsb_warp_se3");
    Mat out(n, 4, CV_32F);
    dev.check(vsb_download(dev.ctx(), out.data, d + (size_t)4 * n, sizeof(float) * 4 * n, st), "download");
    dev.sync();
    return out;
}

void VISystem::Track() {   // VISystem.cpp:1567-1635
    const std::vector<Frame*>& frames = active_camera().frameList;
    if (track_from_estimate && frames.size() >= 2) {
        current_poseCam = frames[frames.size() - 2]->rigid_transformation_;
    } else {
        const float t[3] = {translationResEst.x, translationResEst.y, translationResEst.z};
        float p[7];
        vsb_se3_from_rt(RotationResCam.val, t, p);
        current_poseCam = SE3(p);
    }
    final_poseCam = final_poseCam * current_poseCam;
    const SE3::Vec3 t = final_poseCam.translation();
    positionCam.x = t(0); positionCam.y = t(1); positionCam.z = t(2);
    const SE3::Quat q = final_poseCam.unit_quaternion();
    qOrientationCam.x = q.x(); qOrientationCam.y = q.y(); qOrientationCam.z = q.z(); qOrientationCam.w = q.w();
    RPYOrientationCam = toRPY(qOrientationCam);
    const float zero[3] = {0.f, 0.f, 0.f};
    float pi[7];
    vsb_se3_from_rt(RotationResidualImu.val, zero, pi);
    current_poseImu = SE3(pi);
    final_poseImu = final_poseImu * current_poseImu;
    const SE3::Vec3 t2 = final_poseImu.translation();
    positionImu.x = -t2(0); positionImu.y = -t2(2); positionImu.z = -t2(1);
    const SE3::Quat q2 = final_poseImu.unit_quaternion();
    qOrientationImu.x = q2.x(); qOrientationImu.y = q2.y(); qOrientationImu.z = q2.z(); qOrientationImu.w = q2.w();
    RPYOrientationImu = toRPY(qOrientationImu);
}

void VISystem::FreeLastFrame() {   // VISystem.cpp:407-412 (the reference runs the destructor without delete)
    if (camera.frameList.empty()) return;
    delete camera.frameList[0];
    camera.frameList.erase(camera.frameList.begin());
}

// The built CPU executable's AddFrame (VISystem.cpp:290-405) is an interactive experiment (imshow / waitKey(-1),
// keyframes picked from the keyboard).  The tracking loop the path is quoted on is VISystemGPU::AddFrameGPU
// (VISystemGPU.cpp:144-169); AddFrame runs that same sequence on `camera`.
bool VISystem::AddFrame(Mat _currentImage, vector<Point3d> _imuAngularVelocity, vector<Point3d> _imuAcceleration) {
    update_imu_prior(_imuAngularVelocity, _imuAcceleration);
    prevImage = currentImage;
    currentImage = _currentImage.clone();
    camera.Update(_currentImage);
    nPointsCurrentImage = camera.detectAndComputeFeatures();
    bool key = false;
    if (nPointsCurrentImage > 1 && camera.frameList.size() != 0) {
        camera.num_images++;
        camera.computeGoodMatches();
        camera.computeGradient();
        camera.ObtainPatchesPointsPreviousFrame();
        camera.saveFrame();
        camera.nBestMatches = (int)camera.matcher.goodMatches.size();
        key = true;
    } else if (nPointsCurrentImage > 1) {
        camera.computeGradient();
        camera.saveFrame();
        key = true;
    }
    num_keyframes = (int)camera.frameList.size();
    if (key && camera.frameList.size() > 1) {
        if (num_keyframes > num_max_keyframes) FreeLastFrame();
        EstimatePoseFeatures(camera.frameList[camera.frameList.size() - 2], camera.frameList[camera.frameList.size() - 1]);
        Track();
    }
    lastImageWasKeyframe = currentImageIsKeyframe = key;
    return key;
}

bool VISystem::AddFrame(Mat _currentImage, vector<Point3d> _imuAngularVelocity, vector<Point3d> _imuAcceleration,
                        Point3d _gtPosition) {   // VISystem.cpp:290-299
    prev_gtPosition = current_gtPosition;
    current_gtPosition = _gtPosition;
    current_gtTraslation = current_gtPosition - prev_gtPosition;
    return AddFrame(_currentImage, _imuAngularVelocity, _imuAcceleration);
}

// ---- VISystemGPU (src/VISystemGPU.cpp) --------------------------------------------------------------------------
VISystemGPU::VISystemGPU() : VISystem() {}
VISystemGPU::VISystemGPU(int argc, char* argv[]) : VISystem(argc, argv) {}
VISystemGPU::~VISystemGPU() {}

// VISystemGPU.cpp:39-129.  Differences, all forced by what is outside the library: no undistortion maps (CameraModel refuses
// distorted calibrations), imuCore receives its initial orientation with the first batch of samples (update_imu_prior).
void VISystemGPU::InitializeSystemGPU(std::string _calPath, Point3d _iniPosition, Point3d _iniVelocity, Point3d _iniRPY, Mat image) {
    currentImage = image;
    Calibration(_calPath);
    imu2camTransformation = camera_model->imu2cam0Transformation;
    const Mat r = transformationMatrix2rotationMatrix(imu2camTransformation);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) imu2camRotation(i, j) = r.at<float>(i, j);
    imu2camTranslation = transformationMatrix2position(imu2camTransformation);
    K = camera_model->GetK();
    w_input = camera_model->GetInputWidth();
    h_input = camera_model->GetInputHeight();
    distortion_valid = camera_model->IsValid();          // always false: frames are taken as undistorted
    w = w_input;
    h = h_input;
    InitializePyramid(w, h, K);
    initialized = true;
    positionImu = _iniPosition;                          // :97-104
    velocityImu = _iniVelocity;
    RPYOrientationImu = _iniRPY;
    qOrientationImu = toQuaternion(_iniRPY.x, _iniRPY.y, _iniRPY.z);
    world2imuTransformation = RPYAndPosition2transformationMatrix(RPYOrientationImu, positionImu);
    const Mat wr = transformationMatrix2rotationMatrix(world2imuTransformation);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) world2imuRotation(i, j) = wr.at<float>(i, j);
    final_poseImu = SE3((float)qOrientationImu.w, (float)qOrientationImu.x, (float)qOrientationImu.y, (float)qOrientationImu.z,
                        0.f, 0.f, 0.f);
    // camera pose: T_imu2cam [p 1]^T, R_imu2cam v, R_imu2cam R_world2imu (:108-114)
    {
        const Mat& T = imu2camTransformation;
        const float p[4] = {(float)positionImu.x, (float)positionImu.y, (float)positionImu.z, 1.f};
        double o[3];
        for (int i = 0; i < 3; i++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += (double)T.at<float>(i, k) * p[k];       // cv::gemm: float in, double accumulate
            o[i] = (float)s;
        }
        positionCam = Point3d(o[0], o[1], o[2]);
        const cv::Point3f v = imu2camRotation * cv::Point3f((float)velocityImu.x, (float)velocityImu.y, (float)velocityImu.z);
        velocityCam = Point3d(v.x, v.y, v.z);
    }
    RPYOrientationCam = rotationMatrix2RPY(imu2camRotation * world2imuRotation);
    qOrientationCam = toQuaternion(RPYOrientationCam.x, RPYOrientationCam.y, RPYOrientationCam.z);
    final_poseCam = SE3((float)qOrientationCam.w, (float)qOrientationCam.x, (float)qOrientationCam.y, (float)qOrientationCam.z,
                        (float)-positionCam.x, (float)-positionCam.z, (float)-positionCam.y);
    imuCore.createPublisher(1.0 / (camera_model->imu_frecuency));    // :117-119
    imuCore.setImuInitialVelocity(_iniVelocity);
    imu_ready = false;
    num_max_keyframes = camera_model->min_features;                   // :122 (sic: the reference assigns min_features)
    min_features = camera_model->min_features;
    start_index = camera_model->start_index;
    InitializeCameraGPU(camera_model->detector, camera_model->matcher, w, h, camera_model->num_cells, camera_model->length_patch);
}

void VISystemGPU::InitializeCameraGPU(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path) {
    cameraGPU.initializateCameraGPU(_detector, _matcher, _w_size, _h_size, _num_cells, _length_path);   // VISystemGPU.cpp:136-139
}

void VISystemGPU::FreeLastFrameGPU() {   // VISystemGPU.cpp:171-175
    if (cameraGPU.frameList.empty()) return;
    delete cameraGPU.frameList[0];
    cameraGPU.frameList.erase(cameraGPU.frameList.begin());
}

void VISystemGPU::AddFrameGPU(Mat _currentImage, vector<Point3d> _imuAngularVelocity, vector<Point3d> _imuAcceleration) {
    update_imu_prior(_imuAngularVelocity, _imuAcceleration);   // VISystemGPU.cpp:142-143; the rest is :144-169
    prevImage = currentImage;
    currentImage = _currentImage.clone();
    cameraGPU.Update(_currentImage);
    cameraGPU.addGPUKeyframe();
    num_keyframes = (int)cameraGPU.frameList.size();
    if (cameraGPU.frameList.size() > 1) {
        if (num_keyframes > num_max_keyframes) FreeLastFrameGPU();
        Frame* prev = cameraGPU.frameList[cameraGPU.frameList.size() - 2];
        Frame* cur = cameraGPU.frameList[cameraGPU.frameList.size() - 1];
        EstimatePoseFeatures(prev, cur);
        Track();
    }
}

}  // namespace vi

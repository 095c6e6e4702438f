"""CPU-side checks of the C++ class mirrors (vi-slam_b200/host): the library builds with plain g++ against the
C ABI only, exports the reference's class entry points, and fails loudly — no CPU fallback — when no CUDA
device is visible."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "vi-slam_b200", "host")
PKG = os.path.join(ROOT, "vi-slam_b200", "vislam_b200")


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "vi-slam_b200")])
    subprocess.check_call(["make", "-s", "-C", HOST])


def test_host_library_builds_and_exports_reference_methods():
    _build()
    so = os.path.join(PKG, "libvislam_host.so")
    assert os.path.exists(so)
    syms = subprocess.run(["nm", "-DC", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    for name in ["Matcher::computeMatches()", "Matcher::computeSymMatches()", "Matcher::nnFilter(", "Matcher::sortMatches()",
                 "Matcher::bestMatchesFilter(int)", "Matcher::computeBestMatches(int)", "Matcher::getGoodMatches(",
                 "MatcherGPU::computeGPUMatches()", "Camera::Update(", "Camera::computeGoodMatches()",
                 "Camera::computeGradient()", "Camera::ObtainPatchesPointsPreviousFrame()", "Camera::addKeyframe()",
                 "CameraGPU::computeGPUGoodMatches()", "CameraGPU::addGPUKeyframe()",
                 "vi::VISystem::InitializePyramid(", "vi::VISystem::EstimatePoseFeatures(Frame*, Frame*)",
                 "vi::VISystem::WarpFunctionSE3(", "vi::VISystem::IdentityWeights(int)", "vi::VISystem::Track()",
                 "vi::VISystem::AddFrame(", "vi::VISystem::setGtRes(", "vi::VISystemGPU::AddFrameGPU(",
                 # dataset ingestion and angle helpers
                 "ImageReader::searchImages()", "ImageReader::getImageTime(int)", "ImageReader::getImage(int)",
                 "GroundTruth::getDataFromFile()", "GroundTruth::getGroundTruthData(int, int)",
                 "DataReader::setProperties(", "DataReader::UpdateDataReader(int, int)", "DataReader::UpdateImu(int, int)",
                 "rotationMatrix2RPY(", "RPY2rotationMatrix(", "toQuaternion(double, double, double)", "toRPY(Quaterniond const&)",
                 "vi::TrajectoryWriter::write(", "vi::imread_gray(",
                 # IMU prior
                 "Imu::initializate(", "Imu::estimate()", "Imu::estimateOrientation()", "Imu::computeAcceleration()",
                 "Imu::calibrateAng(int)", "ImuFilterNode::UpdatePublisher(", "ImuFilterNode::UpdateSubscriber()",
                 "vi::MadgwickFilter::update("]:
        assert name in syms, name
    # the class mirrors reach the device only through the C ABI: no CUDA runtime symbols of their own
    undefined = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True, check=True).stdout
    other = [ln.split()[-1] for ln in undefined.splitlines() if ln.strip() and not ln.split()[-1].startswith("vsb_")]
    assert not [s for s in other if s.startswith("cuda") or s.startswith("cu") and s[2:3].isupper()], other
    assert "vsb_gn_solve" in undefined and "vsb_knn2_hamming" in undefined


def test_host_classes_fail_loudly_without_device():
    _build()
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([os.path.join(PKG, "host_runner"), "nodevice"], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


CAL_XML = """<?xml version="1.0"?>
<!-- synthetic pinhole camera, EuRoC-shaped -->
<opencv_storage>
<in_width  type_id="integer"> 752 </in_width>
<in_height type_id="integer"> 480 </in_height>
<out_width  type_id="integer"> 752 </out_width>
<out_height type_id="integer"> 480 </out_height>
<calibration_values type_id="opencv-matrix">
  <rows>1</rows>
  <cols>4</cols>
  <dt>f</dt>
  <data>
    458.654	 457.296  367.215  248.375 </data></calibration_values>
<rectification type_id="opencv-matrix">
  <rows>1</rows>
  <cols>4</cols>
  <dt>f</dt>
  <data>
    %s </data></rectification>
<imu2cam0Transformation type_id="opencv-matrix">
  <rows>4</rows>
  <cols>4</cols>
  <dt>f</dt>
  <data>
     0.0148655429818 -0.999880929698 0.00414029679422 -0.0216401454975
         0.999557249008 0.0149672133247 0.025715529948 -0.064676986768
        -0.0257744366974 0.00375618835797 0.999660727178 0.00981073058949
         0.0 0.0 0.0 1.0 </data></imu2cam0Transformation>
<camera_frecuency  type_id="float"> 20 </camera_frecuency>
<imu_frecuency type_id="float"> 200 </imu_frecuency>
<min_features type_id="integer"> 20</min_features>
<num_max_keyframes type_id="integer"> 10</num_max_keyframes>
<start_index  type_id="integer"> 0 </start_index>
<use_gt type_id="integer">1</use_gt>
<use_ros type_id="integer">0</use_ros>
<num_cells type_id="integer"> 49</num_cells>
<length_patch type_id="integer"> 3</length_patch>
<detector type_id="integer">2</detector>
<matcher type_id="integer">4</matcher>
</opencv_storage>
"""


def test_camera_model_reads_the_reference_calibration_format(tmp_path):
    """vi::CameraModel::GetCameraModel on an OpenCV-FileStorage XML with the reference's keys (calibration/calibrationEUROC.xml
    layout): every value lands in the field the reference fills (src/CameraModel.cpp:25-76); a file with distortion
    coefficients is refused loudly (undistortion is outside this library)."""
    import ctypes as C
    import numpy as np
    _build()
    L = C.CDLL(os.path.join(PKG, "libvislam_host.so"))
    L.vih_last_error.restype = C.c_char_p
    good = tmp_path / "cal.xml"
    good.write_text(CAL_XML % "0 0 0 0")
    out = (C.c_double * 35)()
    assert L.vih_camera_model(str(good).encode(), out) == 0, L.vih_last_error()
    v = np.array(out[:])
    np.testing.assert_array_equal(v[:4], np.array([458.654, 457.296, 367.215, 248.375], np.float32).astype(np.float64))
    assert list(v[4:8]) == [752, 480, 752, 480] and list(v[8:10]) == [20, 200]
    assert list(v[10:19]) == [20, 10, 0, 1, 0, 49, 3, 2, 4]
    assert v[19] == np.float32(0.0148655429818) and v[22] == np.float32(-0.0216401454975) and v[34] == 1.0
    bad = tmp_path / "cal_dist.xml"
    bad.write_text(CAL_XML % "-0.28340811 0.07395907 0.00019359 1.76187114e-05")
    assert L.vih_camera_model(str(bad).encode(), out) == -1
    assert b"undistortion" in L.vih_last_error()
    assert L.vih_camera_model(str(tmp_path / "missing.xml").encode(), out) == -1


def test_reference_gpu_main_compiles_against_the_forwarding_headers():
    """Boundary proof by compilation: host/tests/main_vi_slamGPU_caller.cpp makes the calls of the reference's GPU
    executable (src/main_vi_slamGPU.cpp:58-65, 118-152) in the same form, includes the reference's header NAMES
    (host/include/refnames) and links against the class mirrors."""
    _build()
    exe = os.path.join(PKG, "ref_main_gpu")
    assert os.path.exists(exe)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "imagesPath imuFile gtFile calibrationFile outputFile" in out.stdout
    src = open(os.path.join(HOST, "tests", "main_vi_slamGPU_caller.cpp")).read()
    for call in ['#include "DataReader.hpp"', '#include "VISystemGPU.hpp"', "DataReader Data(imagesPath, imuFile, gtFile, separator);",
                 "visystem.InitializeSystemGPU(calibrationFile, Data.gtPosition[0], Data.gtLinearVelocity[0], Data.gtRPY[0], Data.image1);",
                 "visystem.AddFrameGPU(Data.image2, Data.imuAngularVelocity, Data.imuAcceleration);",
                 "rotationMatrix2RPY(visystem.imu2camRotation * RPY2rotationMatrix(toRPY(Data.gtQuaternion.back())))"]:
        assert call in src, call
    names = set(os.listdir(os.path.join(HOST, "include", "refnames")))
    assert {"MatcherGPU.hpp", "CameraGPU.hpp", "VISystemGPU.hpp", "DataReader.hpp", "Plus.hpp", "Imu.hpp", "Matcher.hpp",
            "Camera.hpp", "VISystem.hpp", "CameraModel.hpp"} <= names

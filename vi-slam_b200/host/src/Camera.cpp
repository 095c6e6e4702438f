// Camera.cpp — class mirror of the reference's Frame / Camera / CameraGPU (src/Camera.cpp, src/CameraGPU.cpp).
//   Update                              Camera.cpp:63-72     -> vsb_pyramid_build   (cv::resize x4)
//   computeGradient                     Camera.cpp:167-184   -> vsb_gradient_build  (cv::Scharr x10, addWeighted)
//   computeGoodMatches                  Camera.cpp:146-157   -> Matcher mirror
//   ObtainPatchesPointsPreviousFrame    Camera.cpp:358-409   -> vsb_candidates_build
//   addKeyframe / addGPUKeyframe        Camera.cpp:197-258 / CameraGPU.cpp:138-200: same sequencing
// The device results stay in the Frame for VISystem::EstimatePoseFeatures; with mirror_host they are also
// copied into the reference's public cv::Mat members.  No CPU fallback: vi::DeviceError without a CUDA device.
#include "vislam/Camera.hpp"

#include <chrono>
#include <cmath>
#include <iomanip>
#include <iostream>

using cv::KeyPoint;
using cv::Mat;
using std::vector;

namespace {
typedef std::chrono::steady_clock Clock;
double secs(const Clock::time_point& a, const Clock::time_point& b) { return std::chrono::duration<double>(b - a).count(); }
}  // namespace

Frame::Frame()
    : idFrame(0), imageTime(0.0), obtainedGradients(false), obtainedGoodMatches(false), isKeyFrame(false), cand_cap(0),
      pyr_on_device(false), grad_on_device(false), cand_on_device(false) {
    std::memset(&layout, 0, sizeof(layout));
    for (int l = 0; l < VSB_MAX_LEVELS; l++) n_cand[l] = 0;
}

Frame::~Frame() {   // Camera.cpp:14-22
    grayImage.clear();
    gradientX.clear();
    gradientY.clear();
    gradient.clear();
    candidatePoints.clear();
}

Camera::Camera()
    : currentFrame(nullptr), w_residual(0), h_residual(0), detectorType(0), matcherType(0), nPointsDetect(0),
      nBestMatches(0), n_cells(0), w_patch(0), h_patch(0), elapsed_detect(0), elapsed_descriptors(0),
      elapsed_computeGoodMatches(0), elapsed_computeGradient(0), elapsed_computePatches(0), elapsed_detect_mean(0),
      elapsed_descriptors_mean(0), elapsed_computeGoodMatches_mean(0), elapsed_computeGradient_mean(0),
      elapsed_computePatches_mean(0), nPointsDetect_mean(0), nBestMatches_mean(0), num_images(0), elapsed_detect_sum(0),
      elapsed_descriptors_sum(0), elapsed_computeGoodMatches_sum(0), elapsed_computeGradient_sum(0),
      elapsed_computePatches_sum(0), nPointsDetect_sum(0), nBestMatches_sum(0), mirror_host(true), orb_nfeatures(200), verbose(false) {}

Camera::Camera(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path) : Camera() {
    initializate(_detector, _matcher, _w_size, _h_size, _num_cells, _length_path);
}

void Camera::initializate(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_path) {
    w_size[0] = _w_size;   // Camera.cpp:42-47
    h_size[0] = _h_size;
    for (int lvl = 1; lvl < 5; lvl++) {
        w_size[lvl] = _w_size >> lvl;
        h_size[lvl] = _h_size >> lvl;
    }
    setDetector(_detector);
    setMatcher(_matcher);
    n_cells = _num_cells;
    w_patch = h_patch = _length_path;
    elapsed_detect_sum = elapsed_descriptors_sum = elapsed_computeGoodMatches_sum = 0.0;
    elapsed_computeGradient_sum = elapsed_computePatches_sum = 0.0;
    nPointsDetect_sum = nBestMatches_sum = 0.0;
    num_images = 0;
}

void Camera::Update(Mat _grayImage) {   // Camera.cpp:63-72
    if (_grayImage.empty() || _grayImage.type() != CV_8U)
        throw std::invalid_argument("Camera::Update: expects a non-empty CV_8UC1 image");
    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    currentFrame = new Frame();
    Frame* f = currentFrame;
    elapsed_computeGoodMatches = elapsed_computeGradient = elapsed_descriptors = elapsed_detect = 0.0;
    dev.check(vsb_pyr_layout(_grayImage.cols, _grayImage.rows, VSB_MAX_LEVELS, &f->layout), "vsb_pyr_layout");
    uint8_t* pyr = static_cast<uint8_t*>(f->d_pyr.reserve((size_t)f->layout.frame_stride));
    // level 0 goes straight into the packed pyramid; the kernel then cascades levels 1..4 in place
    dev.check(vsb_upload_2d(dev.ctx(), pyr, (size_t)_grayImage.cols, _grayImage.data, _grayImage.step, (size_t)_grayImage.cols,
                            (size_t)_grayImage.rows, st), "image upload");
    dev.check(vsb_pyramid_build(dev.ctx(), nullptr, 0, _grayImage.cols, 1, &f->layout, pyr, st), "vsb_pyramid_build");
    f->pyr_on_device = true;
    _grayImage.copyTo(f->grayImage[0]);
    if (mirror_host) {
        for (int l = 1; l < 5; l++) {
            f->grayImage[l].create(f->layout.h[l], f->layout.w[l], CV_8U);
            dev.check(vsb_download(dev.ctx(), f->grayImage[l].data, pyr + f->layout.offset[l],
                                   (size_t)f->layout.w[l] * f->layout.h[l], st), "pyramid download");
        }
    }
    dev.sync();
}

void Camera::setDetector(int _detector) { detectorType = _detector; }   // Camera.cpp:97-142 (detector objects are upstream of the path)

void Camera::setFeatures(const vector<KeyPoint>& keypoints, const Mat& descriptors) {
    if (!currentFrame) throw std::logic_error("Camera::setFeatures: call Update first");
    currentFrame->keypoints = keypoints;
    currentFrame->descriptors = descriptors;
}

// cv::ORB::create(orb_nfeatures)->detect / detectAndCompute on the device (vsb_orb_detect_compute_pyr): the level-0 image is
// already in the frame's packed pyramid; results come back as cv::KeyPoint (pt, size = 31 * scale, angle, response, octave)
// and a CV_8U n x 32 descriptor matrix, level by level in row-major order
struct CameraOrbBuffers { vi::DevBuf xy, oct, resp, ang, desc, n; };

int Camera::detectOrbOnDevice(bool describe) {
    Frame* f = currentFrame;
    if (!f || !f->pyr_on_device) throw std::logic_error("Camera::detectOrbOnDevice: call Update first");
    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    const int w = f->layout.w[0], h = f->layout.h[0];
    f->keypoints.clear();
    f->descriptors = Mat();
    if (w <= 62 || h <= 62 || orb_nfeatures <= 0) return 0;        // the 31-pixel border filter leaves nothing
    const int cap = 2 * orb_nfeatures + 256;                         // ties at the selection thresholds can exceed n
    if (!orb_bufs) orb_bufs = std::make_shared<CameraOrbBuffers>();
    vi::DevBuf &d_xy = orb_bufs->xy, &d_oct = orb_bufs->oct, &d_resp = orb_bufs->resp, &d_ang = orb_bufs->ang, &d_desc = orb_bufs->desc,
               &d_n = orb_bufs->n;
    float* xy = static_cast<float*>(d_xy.reserve((size_t)cap * 8));
    int32_t* oct = static_cast<int32_t*>(d_oct.reserve((size_t)cap * 4));
    float* resp = static_cast<float*>(d_resp.reserve((size_t)cap * 4));
    float* ang = static_cast<float*>(d_ang.reserve((size_t)cap * 4));
    uint8_t* desc = describe ? static_cast<uint8_t*>(d_desc.reserve((size_t)cap * 32)) : nullptr;
    int32_t* n = static_cast<int32_t*>(d_n.reserve(4));
    const uint8_t* img = f->d_pyr.as<uint8_t>() + f->layout.offset[0];
    dev.check(vsb_orb_detect_compute_pyr(dev.ctx(), img, (int64_t)w * h, w, w, h, 1, orb_nfeatures, 1.2f, 8, 20, cap, xy, oct, resp,
                                         ang, desc, n, st), "vsb_orb_detect_compute_pyr");
    int32_t found = 0;
    dev.check(vsb_download(dev.ctx(), &found, n, 4, st), "key-point count download");
    dev.sync();
    const int m = found < cap ? found : cap;
    if (m <= 0) return 0;
    std::vector<float> hxy(2 * (size_t)m), hresp(m), hang(m);
    std::vector<int32_t> hoct(m);
    dev.check(vsb_download(dev.ctx(), hxy.data(), xy, (size_t)m * 8, st), "key points download");
    dev.check(vsb_download(dev.ctx(), hoct.data(), oct, (size_t)m * 4, st), "octaves download");
    dev.check(vsb_download(dev.ctx(), hresp.data(), resp, (size_t)m * 4, st), "responses download");
    dev.check(vsb_download(dev.ctx(), hang.data(), ang, (size_t)m * 4, st), "angles download");
    if (describe) {
        f->descriptors.create(m, 32, CV_8U);
        dev.check(vsb_download(dev.ctx(), f->descriptors.data, desc, (size_t)m * 32, st), "descriptors download");
    }
    dev.sync();
    f->keypoints.resize(m);
    for (int i = 0; i < m; i++) {
        KeyPoint& k = f->keypoints[i];
        k.pt.x = hxy[2 * i]; k.pt.y = hxy[2 * i + 1];
        k.octave = hoct[i];
        k.size = 31.f * (float)std::pow((double)1.2f, (double)hoct[i]);     // patchSize * layer scale (orb.cpp)
        k.angle = hang[i];
        k.response = hresp[i];
        k.class_id = -1;
    }
    return m;
}

int Camera::detectFeatures() {   // Camera.cpp:74-82
    const Clock::time_point t0 = Clock::now();
    if (featureProvider && currentFrame && currentFrame->keypoints.empty()) {
        Mat unused;
        featureProvider(currentFrame->grayImage[0], currentFrame->keypoints, unused);
    } else if (detectorType == USE_ORB && currentFrame && currentFrame->keypoints.empty()) {
        detectOrbOnDevice(false);
    }
    elapsed_detect = secs(t0, Clock::now());
    return currentFrame ? (int)currentFrame->keypoints.size() : 0;
}

int Camera::detectAndComputeFeatures() {   // Camera.cpp:84-93
    const Clock::time_point t0 = Clock::now();
    if (featureProvider && currentFrame && currentFrame->keypoints.empty())
        featureProvider(currentFrame->grayImage[0], currentFrame->keypoints, currentFrame->descriptors);
    else if (detectorType == USE_ORB && currentFrame && currentFrame->keypoints.empty())
        detectOrbOnDevice(true);
    elapsed_detect = secs(t0, Clock::now());
    return currentFrame ? (int)currentFrame->keypoints.size() : 0;
}

void Camera::computeDescriptors() {}   // Camera.cpp:144-148: descriptors arrive with the key points

void Camera::match_with(Matcher& m, bool gpu_entry) {   // Camera.cpp:150-161 / CameraGPU.cpp:125-136
    if (frameList.empty() || !currentFrame) throw std::logic_error("Camera::computeGoodMatches: needs a saved frame and a current frame");
    Frame* prev = frameList[frameList.size() - 1];
    m.clear();
    m.setKeypoints(prev->keypoints, currentFrame->keypoints);
    m.setDescriptors(prev->descriptors, currentFrame->descriptors);
    if (gpu_entry) static_cast<MatcherGPU&>(m).computeGPUMatches();
    else m.computeMatches();
    m.computeBestMatches(n_cells);
    m.getGoodMatches(prev->nextGoodMatches, currentFrame->prevGoodMatches);
    currentFrame->obtainedGoodMatches = true;
}

void Camera::computeGoodMatches() { match_with(matcher, false); }

void Camera::setMatcher(int _matcher) {   // Camera.cpp:163-167
    matcherType = _matcher;
    matcher.setMatcher(_matcher);
    matcher.setImageDimensions(w_size[0], h_size[0]);
}

void Camera::computeGradient() {   // Camera.cpp:171-188
    if (!currentFrame || !currentFrame->pyr_on_device) throw std::logic_error("Camera::computeGradient: call Update first");
    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    Frame* f = currentFrame;
    const size_t px = (size_t)f->layout.frame_stride;
    int16_t* gx = static_cast<int16_t*>(f->d_gx.reserve(px * sizeof(int16_t)));
    int16_t* gy = static_cast<int16_t*>(f->d_gy.reserve(px * sizeof(int16_t)));
    uint8_t* gm = mirror_host ? static_cast<uint8_t*>(f->d_gmag.reserve(px)) : nullptr;
    dev.check(vsb_gradient_build(dev.ctx(), f->d_pyr.as<uint8_t>(), 1, &f->layout, gx, gy, gm, st), "vsb_gradient_build");
    f->grad_on_device = true;
    if (mirror_host) {
        for (int l = 0; l < 5; l++) {
            const size_t n = (size_t)f->layout.w[l] * f->layout.h[l];
            f->gradientX[l].create(f->layout.h[l], f->layout.w[l], CV_16S);
            f->gradientY[l].create(f->layout.h[l], f->layout.w[l], CV_16S);
            f->gradient[l].create(f->layout.h[l], f->layout.w[l], CV_8U);
            dev.check(vsb_download(dev.ctx(), f->gradientX[l].data, gx + f->layout.offset[l], n * 2, st), "gradient download");
            dev.check(vsb_download(dev.ctx(), f->gradientY[l].data, gy + f->layout.offset[l], n * 2, st), "gradient download");
            dev.check(vsb_download(dev.ctx(), f->gradient[l].data, gm + f->layout.offset[l], n, st), "gradient download");
        }
    }
    dev.sync();
    f->obtainedGradients = true;
}

void Camera::saveFrame() {   // Camera.cpp:192-197
    currentFrame->isKeyFrame = true;
    frameList.push_back(currentFrame);
}

void Camera::ObtainPatchesPointsPreviousFrame() {   // Camera.cpp:358-409
    if (frameList.empty()) throw std::logic_error("Camera::ObtainPatchesPointsPreviousFrame: no saved frame");
    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    Frame* f = frameList[frameList.size() - 1];
    const vector<KeyPoint>& good = f->nextGoodMatches;
    const int nf_all = (int)good.size();
    const int nf = nf_all < VSB_MAX_GN_FEATURES ? nf_all : VSB_MAX_GN_FEATURES;   // Camera.cpp:382
    const int good_cap = nf > 0 ? nf : 1;
    const int cand_cap = 121 * good_cap;
    vector<float> xy((size_t)2 * good_cap, 0.f);
    for (int i = 0; i < nf; i++) { xy[2 * i] = good[i].pt.x; xy[2 * i + 1] = good[i].pt.y; }
    vi::DevBuf d_xy;
    float* dxy = static_cast<float*>(d_xy.reserve(sizeof(float) * 2 * good_cap));
    int32_t* d_nc = static_cast<int32_t*>(f->d_ncand.reserve(sizeof(int32_t) * (VSB_MAX_LEVELS + 1)));
    float* d_cand = static_cast<float*>(f->d_cand.reserve(sizeof(float) * 4 * (size_t)cand_cap * VSB_MAX_LEVELS));
    const int32_t nf32 = nf;
    dev.check(vsb_upload(dev.ctx(), dxy, xy.data(), sizeof(float) * 2 * good_cap, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_nc + VSB_MAX_LEVELS, &nf32, sizeof(int32_t), st), "upload");
    int lw[VSB_MAX_LEVELS], lh[VSB_MAX_LEVELS];
    for (int l = 0; l < VSB_MAX_LEVELS; l++) { lw[l] = w_size[l]; lh[l] = h_size[l]; }
    dev.check(vsb_candidates_build(dev.ctx(), dxy, good_cap, d_nc + VSB_MAX_LEVELS, 1, VSB_MAX_LEVELS, lw, lh, d_cand,
                                   cand_cap, d_nc, st), "vsb_candidates_build");
    int32_t nc[VSB_MAX_LEVELS];
    dev.check(vsb_download(dev.ctx(), nc, d_nc, sizeof(nc), st), "download");
    dev.sync();
    f->cand_cap = cand_cap;
    f->cand_on_device = true;
    for (int l = 0; l < VSB_MAX_LEVELS; l++) f->n_cand[l] = nc[l];
    if (mirror_host) {
        // the reference appends rows to candidatePoints[lvl] (Mat::push_back); a frame gets them once
        for (int l = 0; l < VSB_MAX_LEVELS; l++) {
            if (nc[l] == 0) { f->candidatePoints[l].release(); continue; }
            f->candidatePoints[l].create(nc[l], 4, CV_32F);
            dev.check(vsb_download(dev.ctx(), f->candidatePoints[l].data, d_cand + (size_t)4 * cand_cap * l,
                                   sizeof(float) * 4 * nc[l], st), "candidate download");
        }
        dev.sync();
    }
}

void Camera::ObtainDebugPointsPreviousFrame() {}   // Camera.cpp:411-460: drawing aid, not on the path

void Camera::stats_accumulate() {   // Camera.cpp:259-296 / CameraGPU.cpp:176-196
    if (currentFrame && currentFrame->isKeyFrame && frameList.size() > 1) {
        elapsed_detect_sum += elapsed_detect;
        elapsed_descriptors_sum += elapsed_descriptors;
        elapsed_computeGoodMatches_sum += elapsed_computeGoodMatches;
        elapsed_computeGradient_sum += elapsed_computeGradient;
        elapsed_computePatches_sum += elapsed_computePatches;
        nPointsDetect_sum += nPointsDetect;
        nBestMatches_sum += nBestMatches;
        const double n = num_images > 0 ? (double)num_images : 1.0;
        elapsed_detect_mean = elapsed_detect_sum / n;
        elapsed_descriptors_mean = elapsed_descriptors_sum / n;
        elapsed_computeGoodMatches_mean = elapsed_computeGoodMatches_sum / n;
        elapsed_computeGradient_mean = elapsed_computeGradient_sum / n;
        elapsed_computePatches_mean = elapsed_computePatches_sum / n;
        nPointsDetect_mean = nPointsDetect_sum / n;
        nBestMatches_mean = nBestMatches_sum / n;
    }
}

bool Camera::addKeyframe() {   // Camera.cpp:201-299 (the CPU front end matches only; gradients/patches are commented out there)
    const Clock::time_point t0 = Clock::now();
    nPointsDetect = detectAndComputeFeatures();
    const Clock::time_point t1 = Clock::now();
    Clock::time_point t2 = t1;
    if (nPointsDetect > 10 && frameList.size() != 0) {
        num_images = num_images + 1;
        computeGoodMatches();
        t2 = Clock::now();
        saveFrame();
        nBestMatches = (int)matcher.goodMatches.size();
    } else if (nPointsDetect > 1 && frameList.size() == 0) {
        saveFrame();
        if (verbose) std::cout << "First Image detected" << "list = " << frameList.size() << std::endl;
    }
    elapsed_detect = secs(t0, t1);
    elapsed_computeGoodMatches = secs(t1, t2);
    stats_accumulate();
    return currentFrame->isKeyFrame;
}

void Camera::printStatistics() {   // Camera.cpp:301-356
    std::cout << "\nESTADISTICAS CAMARA" << std::fixed << std::setprecision(3)
              << "\nTiempo de matching: " << elapsed_computeGoodMatches * 1000 << " ms"
              << "\tTiempo de gradiente: " << elapsed_computeGradient * 1000 << " ms"
              << "\tTiempo de parches: " << elapsed_computePatches * 1000 << " ms"
              << "\nPuntos detectados: " << nPointsDetect << "\tMatches finales: " << nBestMatches << std::endl;
}

// ---- CameraGPU (src/CameraGPU.cpp) ----------------------------------------------------------------------------
CameraGPU::CameraGPU() : Camera(), useGPU(true) {}

CameraGPU::CameraGPU(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_patch)
    : Camera(), useGPU(true) {
    initializateCameraGPU(_detector, _matcher, _w_size, _h_size, _num_cells, _length_patch);
}

void CameraGPU::initializateCameraGPU(int _detector, int _matcher, int _w_size, int _h_size, int _num_cells, int _length_patch) {
    initializate(_detector, _matcher, _w_size, _h_size, _num_cells, _length_patch);   // CameraGPU.cpp:20-45
    setGPUDetector(_detector);
    setGPUMatcher(_matcher);
}

void CameraGPU::setGPUDetector(int _detector) { detectorType = _detector; orb_nfeatures = 1000; }   // CameraGPU.cpp:99 cuda::ORB::create(1000)
void CameraGPU::detectGPUFeatures() { nPointsDetect = detectFeatures(); }
int CameraGPU::detectAndComputeGPUFeatures() { return detectAndComputeFeatures(); }

void CameraGPU::setGPUMatcher(int _matcher) {   // CameraGPU.cpp:118-123
    matcherGPU.setGPUMatcher(_matcher);
    matcherGPU.setImageDimensions(w_size[0], h_size[0]);
}

void CameraGPU::computeGPUGoodMatches() { match_with(matcherGPU, true); }   // CameraGPU.cpp:125-136

bool CameraGPU::addGPUKeyframe() {   // CameraGPU.cpp:138-200
    const Clock::time_point t0 = Clock::now();
    nPointsDetect = detectAndComputeGPUFeatures();
    const Clock::time_point t1 = Clock::now();
    Clock::time_point t2 = t1, t3 = t1, t4 = t1;
    if (nPointsDetect > 1 && frameList.size() != 0) {
        num_images = num_images + 1;
        computeGPUGoodMatches();
        t2 = Clock::now();
        computeGradient();
        t3 = Clock::now();
        ObtainPatchesPointsPreviousFrame();
        ObtainDebugPointsPreviousFrame();
        t4 = Clock::now();
        saveFrame();
        nBestMatches = (int)matcherGPU.goodMatches.size();
    } else if (nPointsDetect > 1 && frameList.size() == 0) {
        computeGradient();
        t3 = t4 = Clock::now();
        saveFrame();
        if (verbose) std::cout << "First Image detected" << "list = " << frameList.size() << std::endl;
    }
    elapsed_detect = secs(t0, t1);
    elapsed_computeGoodMatches = secs(t1, t2);
    elapsed_computeGradient = secs(t2, t3);
    elapsed_computePatches = secs(t3, t4);
    stats_accumulate();
    return currentFrame->isKeyFrame;
}

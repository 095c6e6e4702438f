// Matcher.cpp — class mirror of the reference's Matcher / MatcherGPU (src/Matcher.cpp, src/MatcherGPU.cpp).
// Every stage is a kernel of libvislam_b200 reached through the C ABI; this file only moves the public
// std::vector / cv::Mat state of the class to the device and back.  Stage by stage:
//   computeMatches      Matcher.cpp:83-94 / MatcherGPU.cpp:44-66   -> vsb_knn2_hamming | vsb_knn2_l2
//   nnFilter            Matcher.cpp:148-169                        -> vsb_nn_filter
//   computeSymMatches   Matcher.cpp:96-144                         -> vsb_nn_filter x2 + vsb_sym_matches
//   sortMatches         Matcher.cpp:329-352                        -> vsb_sort_keys
//   bestMatchesFilter   Matcher.cpp:171-244                        -> vsb_grid_best
// There is no CPU fallback: without a CUDA device these methods throw vi::DeviceError.
#include "vislam/Matcher.hpp"

#include <chrono>
#include <cmath>
#include <iomanip>
#include <iostream>

using cv::DMatch;
using cv::KeyPoint;
using cv::Mat;
using std::vector;

namespace {

double seconds_since(const std::chrono::steady_clock::time_point& t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// rows of a knnMatch result as flat arrays; a missing neighbour is idx -1 / dist 0 (the C ABI's convention)
void serialize_live(const vector<vector<DMatch> >& m, vector<int32_t>& idx, vector<float>& dist) {
    const size_t n = m.size();
    idx.assign(2 * n, -1);
    dist.assign(2 * n, 0.f);
    for (size_t i = 0; i < n; i++)
        for (size_t k = 0; k < 2 && k < m[i].size(); k++) {
            idx[2 * i + k] = m[i][k].trainIdx;
            dist[2 * i + k] = m[i][k].distance;
        }
}

}  // namespace

Matcher::Matcher()
    : h_size(0), w_size(0), nSymMatches(0), nBestMatches(0), elapsed_detect1(0), elapsed_detect2(0), elapsed_knn1(0),
      elapsed_knn2(0), elapsed_symMatches(0), elapsed_sortMatches(0), elapsed_bestMatches(0), matchPercentage(0),
      norm_type(0), sym_mode(0), verbose(false), n1_(0), n2_(0), knn_valid_(false) {
    setMatcher(0);
}

Matcher::Matcher(int _matcher)
    : h_size(0), w_size(0), nSymMatches(0), nBestMatches(0), elapsed_detect1(0), elapsed_detect2(0), elapsed_knn1(0),
      elapsed_knn2(0), elapsed_symMatches(0), elapsed_sortMatches(0), elapsed_bestMatches(0), matchPercentage(0),
      norm_type(0), sym_mode(0), verbose(false), n1_(0), n2_(0), knn_valid_(false) {
    setMatcher(_matcher);
}

void Matcher::clear() {   // Matcher.cpp:18-29
    keypoints_1.clear();
    keypoints_2.clear();
    descriptors_1.release();
    descriptors_2.release();
    aux_matches1.clear();
    aux_matches2.clear();
    matches.clear();
    goodMatches.clear();
    sortedMatches.clear();
    knn_valid_ = false;
}

void Matcher::setImageDimensions(int w, int h) {
    w_size = w;
    h_size = h;
}

void Matcher::setKeypoints(vector<KeyPoint> _keypoints_1, vector<KeyPoint> _keypoints_2) {
    keypoints_1 = _keypoints_1;
    keypoints_2 = _keypoints_2;
}

void Matcher::setDescriptors(Mat _descriptors_1, Mat _descriptors_2) {
    descriptors_1 = _descriptors_1;
    descriptors_2 = _descriptors_2;
}

void Matcher::setMatcher(int _matcher) {   // Matcher.cpp:49-78
    switch (_matcher) {
        case USE_BRUTE_FORCE:
        case USE_BRUTE_FORCE_GPU:
            norm_type = 0;   // BFMatcher::create() == NORM_L2
            if (verbose) std::cout << "Using Brute Force B200 Matcher (L2)" << std::endl;
            break;
        case USE_BRUTE_FORCE_HAMMING:
        case USE_BRUTE_FORCE_GPU_HAMMING:
            norm_type = 1;
            if (verbose) std::cout << "Using Brute Force -Hamming B200 Matcher" << std::endl;
            break;
        default:
            // FLANN (approximate L2 search over float descriptors) is replaced by the exact L2 search.
            norm_type = 0;
            if (verbose) std::cout << "Using exact L2 B200 Matcher in place of FLANN" << std::endl;
            break;
    }
}

void Matcher::run_knn() {
    vi::Device& dev = vi::Device::get();
    const Mat& a = descriptors_1;
    const Mat& b = descriptors_2;
    n1_ = a.rows;
    n2_ = b.rows;
    aux_matches1.assign((size_t)n1_, vector<DMatch>());
    aux_matches2.assign((size_t)n2_, vector<DMatch>());
    h_idx12_.assign((size_t)2 * n1_, -1);
    h_dist12_.assign((size_t)2 * n1_, 0.f);
    h_idx21_.assign((size_t)2 * n2_, -1);
    h_dist21_.assign((size_t)2 * n2_, 0.f);
    knn_valid_ = true;
    if (n1_ == 0 || n2_ == 0) return;   // knnMatch against an empty train set returns empty rows
    if (a.type() != b.type() || a.cols != b.cols)
        throw std::invalid_argument("Matcher::computeMatches: descriptor sets differ in type or length");
    const bool hamming = (norm_type == 1);
    if (hamming && (a.type() != CV_8U || a.cols != 32))
        throw std::invalid_argument("Matcher::computeMatches: the Hamming matcher takes 32-byte CV_8U descriptors (ORB)");
    if (!hamming && a.type() != CV_32F)
        throw std::invalid_argument("Matcher::computeMatches: the L2 matcher takes CV_32F descriptors");
    const size_t row_bytes = (size_t)a.cols * a.elemSize();
    void* st = dev.stream();
    void* d1 = d_desc1_.reserve((size_t)n1_ * row_bytes);
    void* d2 = d_desc2_.reserve((size_t)n2_ * row_bytes);
    dev.check(vsb_upload_2d(dev.ctx(), d1, row_bytes, a.data, a.step, row_bytes, (size_t)n1_, st), "descriptor upload");
    dev.check(vsb_upload_2d(dev.ctx(), d2, row_bytes, b.data, b.step, row_bytes, (size_t)n2_, st), "descriptor upload");
    int32_t* i12 = static_cast<int32_t*>(d_idx12_.reserve(sizeof(int32_t) * 2 * n1_));
    int32_t* i21 = static_cast<int32_t*>(d_idx21_.reserve(sizeof(int32_t) * 2 * n2_));
    float* s12 = static_cast<float*>(d_dist12_.reserve(sizeof(float) * 2 * n1_));
    float* s21 = static_cast<float*>(d_dist21_.reserve(sizeof(float) * 2 * n2_));
    if (hamming)
        dev.check(vsb_knn2_hamming(dev.ctx(), static_cast<const uint8_t*>(d1), n1_, nullptr, static_cast<const uint8_t*>(d2),
                                   n2_, nullptr, 1, i12, s12, i21, s21, st), "vsb_knn2_hamming");
    else
        dev.check(vsb_knn2_l2(dev.ctx(), static_cast<const float*>(d1), n1_, nullptr, static_cast<const float*>(d2), n2_,
                              nullptr, a.cols, 1, i12, s12, i21, s21, st), "vsb_knn2_l2");
    dev.check(vsb_download(dev.ctx(), h_idx12_.data(), i12, sizeof(int32_t) * 2 * n1_, st), "download");
    dev.check(vsb_download(dev.ctx(), h_dist12_.data(), s12, sizeof(float) * 2 * n1_, st), "download");
    dev.check(vsb_download(dev.ctx(), h_idx21_.data(), i21, sizeof(int32_t) * 2 * n2_, st), "download");
    dev.check(vsb_download(dev.ctx(), h_dist21_.data(), s21, sizeof(float) * 2 * n2_, st), "download");
    dev.sync();
    // vector<vector<DMatch>> as BFMatcher::knnMatch fills it: queryIdx = row, imgIdx = 0, shorter rows when the
    // train set has fewer than 2 descriptors
    for (int i = 0; i < n1_; i++)
        for (int k = 0; k < 2; k++)
            if (h_idx12_[2 * i + k] >= 0) aux_matches1[i].push_back(DMatch(i, h_idx12_[2 * i + k], 0, h_dist12_[2 * i + k]));
    for (int j = 0; j < n2_; j++)
        for (int k = 0; k < 2; k++)
            if (h_idx21_[2 * j + k] >= 0) aux_matches2[j].push_back(DMatch(j, h_idx21_[2 * j + k], 0, h_dist21_[2 * j + k]));
}

void Matcher::computeMatches() {   // Matcher.cpp:83-94: both knnMatch calls, here from ONE pass over the distance matrix
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    run_knn();
    elapsed_knn1 = seconds_since(t0);
    elapsed_knn2 = 0.0;   // the second direction comes out of the same kernel
}

int Matcher::nnFilter(vector<vector<DMatch> >& m, double nn_ratio) {   // Matcher.cpp:148-169
    const int n = (int)m.size();
    if (n == 0) return 0;
    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    vector<int32_t> idx;
    vector<float> dist;
    serialize_live(m, idx, dist);
    int32_t* d_idx = static_cast<int32_t*>(d_list_q_.reserve(sizeof(int32_t) * 2 * n));
    float* d_dist = static_cast<float*>(d_list_d_.reserve(sizeof(float) * 2 * n));
    uint8_t* d_keep = static_cast<uint8_t*>(d_keys_.reserve((size_t)n));
    dev.check(vsb_upload(dev.ctx(), d_idx, idx.data(), sizeof(int32_t) * 2 * n, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_dist, dist.data(), sizeof(float) * 2 * n, st), "upload");
    dev.check(vsb_nn_filter(dev.ctx(), d_idx, d_dist, n, nullptr, 1, nn_ratio, d_keep, st), "vsb_nn_filter");
    vector<uint8_t> keep((size_t)n);
    dev.check(vsb_download(dev.ctx(), keep.data(), d_keep, (size_t)n, st), "download");
    dev.sync();
    int removed = 0;   // the reference returns an uninitialised counter; this one is exact
    for (int i = 0; i < n; i++)
        if (!keep[i]) { m[i].clear(); removed++; }
    return removed;
}

void Matcher::computeSymMatches() {   // Matcher.cpp:96-144
    const double nn_match_ratio = 0.8f;   // float literal widened to double, Matcher.cpp:103
    const int n1 = (int)aux_matches1.size(), n2 = (int)aux_matches2.size();
    // What the reference's symmetry loop can still read after nnFilter cleared rows of aux_matches2: the old
    // first element (std::vector::clear keeps the storage).  Rows already empty before this call read the
    // kNN result this object last computed, when there is one.
    vector<int32_t> mem_idx21;
    vector<float> mem_dist21;
    serialize_live(aux_matches2, mem_idx21, mem_dist21);
    if (knn_valid_ && n2 == n2_)
        for (int j = 0; j < n2; j++)
            if (aux_matches2[j].empty()) { mem_idx21[2 * j] = h_idx21_[2 * j]; mem_idx21[2 * j + 1] = h_idx21_[2 * j + 1]; }
    vector<int32_t> idx12;
    vector<float> dist12;
    serialize_live(aux_matches1, idx12, dist12);

    nnFilter(aux_matches1, nn_match_ratio);
    nnFilter(aux_matches2, nn_match_ratio);
    if (n1 == 0) { nSymMatches = (int)matches.size(); return; }

    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    vector<uint8_t> keep12((size_t)n1), keep21((size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n1; i++) keep12[i] = aux_matches1[i].size() >= 2 ? 1 : 0;   // Matcher.cpp:116
    for (int j = 0; j < n2; j++) keep21[j] = aux_matches2[j].size() >= 2 ? 1 : 0;
    int32_t* d_i12 = static_cast<int32_t*>(d_idx12_.reserve(sizeof(int32_t) * 2 * n1));
    float* d_s12 = static_cast<float*>(d_dist12_.reserve(sizeof(float) * 2 * n1));
    int32_t* d_i21 = static_cast<int32_t*>(d_idx21_.reserve(sizeof(int32_t) * 2 * (n2 > 0 ? n2 : 1)));
    uint8_t* d_k12 = static_cast<uint8_t*>(d_keep12_.reserve((size_t)n1));
    uint8_t* d_k21 = static_cast<uint8_t*>(d_keep21_.reserve((size_t)(n2 > 0 ? n2 : 1)));
    int32_t* d_q = static_cast<int32_t*>(d_list_q_.reserve(sizeof(int32_t) * n1));
    int32_t* d_t = static_cast<int32_t*>(d_list_t_.reserve(sizeof(int32_t) * n1));
    float* d_d = static_cast<float*>(d_list_d_.reserve(sizeof(float) * n1));
    int32_t* d_n = static_cast<int32_t*>(d_cnt_.reserve(sizeof(int32_t) * 4));
    dev.check(vsb_upload(dev.ctx(), d_i12, idx12.data(), sizeof(int32_t) * 2 * n1, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_s12, dist12.data(), sizeof(float) * 2 * n1, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_i21, mem_idx21.data(), sizeof(int32_t) * 2 * n2, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_k12, keep12.data(), (size_t)n1, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_k21, keep21.data(), (size_t)n2, st), "upload");
    dev.check(vsb_sym_matches(dev.ctx(), d_i12, d_s12, d_k12, n1, nullptr, d_i21, d_k21, n2, nullptr, 1, sym_mode, d_q,
                              d_t, d_d, d_n, st), "vsb_sym_matches");
    int32_t n_sym = 0;
    dev.check(vsb_download(dev.ctx(), &n_sym, d_n, sizeof(int32_t), st), "download");
    dev.sync();
    if (n_sym > 0) {
        vector<int32_t> q((size_t)n_sym), t((size_t)n_sym);
        vector<float> d((size_t)n_sym);
        dev.check(vsb_download(dev.ctx(), q.data(), d_q, sizeof(int32_t) * n_sym, st), "download");
        dev.check(vsb_download(dev.ctx(), t.data(), d_t, sizeof(int32_t) * n_sym, st), "download");
        dev.check(vsb_download(dev.ctx(), d.data(), d_d, sizeof(float) * n_sym, st), "download");
        dev.sync();
        for (int k = 0; k < n_sym; k++) matches.push_back(DMatch(q[k], t[k], d[k]));   // Matcher.cpp:129-131
    }
    nSymMatches = (int)matches.size();
}

void Matcher::sortMatches() {   // Matcher.cpp:329-352
    const int n = (int)matches.size();
    if (n == 0) return;
    vi::Device& dev = vi::Device::get();
    void* st = dev.stream();
    vector<float> y((size_t)n);
    for (int i = 0; i < n; i++) y[i] = keypoints_1.at((size_t)matches[i].queryIdx).pt.y;
    float* d_y = static_cast<float*>(d_keys_.reserve(sizeof(float) * n));
    int32_t* d_o = static_cast<int32_t*>(d_order_.reserve(sizeof(int32_t) * n));
    int32_t* d_n = static_cast<int32_t*>(d_cnt_.reserve(sizeof(int32_t) * 4));
    const int32_t nn = n;
    dev.check(vsb_upload(dev.ctx(), d_y, y.data(), sizeof(float) * n, st), "upload");
    dev.check(vsb_upload(dev.ctx(), d_n, &nn, sizeof(int32_t), st), "upload");
    dev.check(vsb_sort_keys(dev.ctx(), d_y, n, d_n, 1, d_o, st), "vsb_sort_keys");
    vector<int32_t> order((size_t)n);
    dev.check(vsb_download(dev.ctx(), order.data(), d_o, sizeof(int32_t) * n, st), "download");
    dev.sync();
    for (int i = 0; i < n; i++) sortedMatches.push_back(matches[(size_t)order[i]]);
}

int Matcher::bestMatchesFilter(int n_features) {   // Matcher.cpp:171-244
    const int n = (int)sortedMatches.size();
    const int n1 = (int)keypoints_1.size();
    if (n > 0 && n_features >= 1) {   // the reference dereferences begin() of an empty vector (App. B-3): empty in, empty out
        vi::Device& dev = vi::Device::get();
        void* st = dev.stream();
        const int root = (int)std::floor(std::sqrt((double)n_features));
        const int cap = root * root;
        vector<int32_t> q((size_t)n), t((size_t)n);
        vector<float> d((size_t)n), xy((size_t)2 * n1);
        for (int i = 0; i < n; i++) { q[i] = sortedMatches[i].queryIdx; t[i] = sortedMatches[i].trainIdx; d[i] = sortedMatches[i].distance; }
        for (int i = 0; i < n1; i++) { xy[2 * i] = keypoints_1[i].pt.x; xy[2 * i + 1] = keypoints_1[i].pt.y; }
        int32_t* d_q = static_cast<int32_t*>(d_list_q_.reserve(sizeof(int32_t) * n));
        int32_t* d_t = static_cast<int32_t*>(d_list_t_.reserve(sizeof(int32_t) * n));
        float* d_d = static_cast<float*>(d_list_d_.reserve(sizeof(float) * n));
        float* d_xy = static_cast<float*>(d_kp1_.reserve(sizeof(float) * 2 * n1));
        int32_t* d_n = static_cast<int32_t*>(d_cnt_.reserve(sizeof(int32_t) * 4));
        int32_t* g_q = static_cast<int32_t*>(d_good_q_.reserve(sizeof(int32_t) * 2 * cap));
        int32_t* g_t = static_cast<int32_t*>(d_good_t_.reserve(sizeof(int32_t) * cap));
        float* g_d = static_cast<float*>(d_good_d_.reserve(sizeof(float) * cap));
        int32_t* g_p = g_q + cap;
        const int32_t nn = n;
        dev.check(vsb_upload(dev.ctx(), d_q, q.data(), sizeof(int32_t) * n, st), "upload");
        dev.check(vsb_upload(dev.ctx(), d_t, t.data(), sizeof(int32_t) * n, st), "upload");
        dev.check(vsb_upload(dev.ctx(), d_d, d.data(), sizeof(float) * n, st), "upload");
        dev.check(vsb_upload(dev.ctx(), d_xy, xy.data(), sizeof(float) * 2 * n1, st), "upload");
        dev.check(vsb_upload(dev.ctx(), d_n, &nn, sizeof(int32_t), st), "upload");
        dev.check(vsb_grid_best(dev.ctx(), d_q, d_t, d_d, n, d_n, d_xy, n1, 1, w_size, h_size, n_features, g_q, g_t, g_d,
                                g_p, cap, d_n + 1, st), "vsb_grid_best");
        int32_t n_good = 0;
        vector<int32_t> pos((size_t)cap);
        dev.check(vsb_download(dev.ctx(), &n_good, d_n + 1, sizeof(int32_t), st), "download");
        dev.check(vsb_download(dev.ctx(), pos.data(), g_p, sizeof(int32_t) * cap, st), "download");
        dev.sync();
        for (int k = 0; k < n_good; k++) goodMatches.push_back(sortedMatches[(size_t)pos[k]]);   // Matcher.cpp:318-326
    }
    nBestMatches = (int)goodMatches.size();
    return (int)goodMatches.size();
}

void Matcher::computeBestMatches(int n_cells) {   // Matcher.cpp:353-367
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    computeSymMatches();
    const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    sortMatches();
    const std::chrono::steady_clock::time_point t2 = std::chrono::steady_clock::now();
    bestMatchesFilter(n_cells);
    elapsed_symMatches = std::chrono::duration<double>(t1 - t0).count();
    elapsed_sortMatches = std::chrono::duration<double>(t2 - t1).count();
    elapsed_bestMatches = seconds_since(t2);
}

void Matcher::resetVectorMatches(vector<DMatch>& v) {   // Matcher.cpp:309-315
    for (size_t i = 0; i < v.size(); i++) v[i].distance = 100000.0f;
}

void Matcher::pushBackVectorMatches(vector<DMatch>& v) {   // Matcher.cpp:317-326
    for (size_t i = 0; i < v.size(); i++)
        if (v[i].distance != 100000.0f) goodMatches.push_back(v[i]);
}

void Matcher::getGrid(int n_features, vector<KeyPoint>& grid_points) {   // Matcher.cpp:246-284
    const float winW = (float)(w_size / std::floor(std::sqrt((double)n_features)));
    const float winH = (float)(h_size / std::floor(std::sqrt((double)n_features)));
    const int root_n = (int)std::floor(std::sqrt((double)n_features));
    float h_final = winH;
    for (int j = 0; j < root_n; j++) {
        float w_final = winW;
        for (int i = 0; i < root_n; i++) {
            KeyPoint p;
            p.pt.x = w_final - winW / 2;
            p.pt.y = h_final - winH / 2;
            grid_points.push_back(p);
            w_final = w_final + winW;
        }
        h_final = h_final + winH;
    }
}

void Matcher::getMatches(vector<KeyPoint>& _matched1, vector<KeyPoint>& _matched2) {   // Matcher.cpp:286-292
    for (size_t i = 0; i < matches.size(); i++) {
        _matched1.push_back(keypoints_1.at((size_t)matches[i].queryIdx));
        _matched2.push_back(keypoints_2.at((size_t)matches[i].trainIdx));
    }
}

void Matcher::getGoodMatches(vector<KeyPoint>& _matched1, vector<KeyPoint>& _matched2) {   // Matcher.cpp:295-303
    _matched1.clear();
    _matched2.clear();
    for (size_t i = 0; i < goodMatches.size(); i++) {
        _matched1.push_back(keypoints_1.at((size_t)goodMatches[i].queryIdx));
        _matched2.push_back(keypoints_2.at((size_t)goodMatches[i].trainIdx));
    }
}

double Matcher::getMatchPercentage() { return 0.0; }   // Matcher.cpp:305-307

void Matcher::printStatistics() {   // Matcher.cpp:369-382
    std::cout << "\nESTADISTICAS"
              << "\nNumero de matches simetricos: " << nSymMatches << "\tNumero de matches finales: " << nBestMatches
              << "\nTiempo de knn I1+I2 (un kernel): " << std::fixed << std::setprecision(3) << elapsed_knn1 * 1000 << " ms"
              << "\nTiempo de symMatches: " << elapsed_symMatches * 1000 << " ms"
              << "\tTiempo de sortMatches " << elapsed_sortMatches * 1000 << " ms"
              << "\nTiempo de bestMatches " << elapsed_bestMatches * 1000 << " ms" << std::endl;
}

// ---- MatcherGPU (src/MatcherGPU.cpp) ------------------------------------------------------------------------
MatcherGPU::MatcherGPU() : Matcher(), useGPU(true), matcherType(0) { setGPUMatcher(0); }

MatcherGPU::MatcherGPU(int _matcher) : Matcher(_matcher), useGPU(true), matcherType(_matcher) { setGPUMatcher(_matcher); }

void MatcherGPU::setGPUMatcher(int _matcher) {   // MatcherGPU.cpp:17-42
    matcherType = _matcher;
    useGPU = true;   // every matcher kind runs on the device here
    setMatcher(_matcher);
}

void MatcherGPU::setGPUFrames(cv::Mat, cv::Mat) {}   // declared but never defined upstream (MatcherGPU.hpp:22)

void MatcherGPU::computeGPUMatches() { computeMatches(); }   // MatcherGPU.cpp:44-66

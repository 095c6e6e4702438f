// oracle/cvshim/ref_matcher_capi.cpp — C entry around the reference's own Matcher class (TEST INFRASTRUCTURE).
// Built by `make -C oracle ref` together with /root/reference/src/Matcher.cpp (unmodified, compiled where it
// lies) into oracle/_ref/libref_matcher.so.  Drives it exactly as Camera::computeGoodMatches does
// (reference src/Camera.cpp:146-157).
#include "Matcher.hpp"
#include <sstream>
#include <iostream>

extern "C" int ref_matcher_run(const void* d1, int n1, const void* d2, int n2, int dim, int norm,
                               const float* kp1_xy, const float* kp2_xy, int w, int h, int n_cells,
                               int* good_q, int* good_t, float* good_d,
                               int* sym_q, int* sym_t, int* n_sym, int* sorted_q,
                               float* prev_xy, float* cur_xy) {
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());            // the class prints which matcher it uses
    Matcher m(norm == 1 ? USE_BRUTE_FORCE_HAMMING : USE_BRUTE_FORCE);
    std::cout.rdbuf(old);
    const int type = norm == 1 ? CV_8U : CV_32F;
    cv::Mat a(n1, dim, type), b(n2, dim, type);
    if (n1) memcpy(a.data(), d1, (size_t)n1 * dim * cv::Mat::esz(type));
    if (n2) memcpy(b.data(), d2, (size_t)n2 * dim * cv::Mat::esz(type));
    std::vector<cv::KeyPoint> k1(n1), k2(n2);
    for (int i = 0; i < n1; i++) { k1[i].pt.x = kp1_xy[2 * i]; k1[i].pt.y = kp1_xy[2 * i + 1]; }
    for (int i = 0; i < n2; i++) { k2[i].pt.x = kp2_xy[2 * i]; k2[i].pt.y = kp2_xy[2 * i + 1]; }
    m.clear();
    m.setKeypoints(k1, k2);
    m.setDescriptors(a, b);
    m.setImageDimensions(w, h);
    m.computeMatches();
    m.computeSymMatches();
    *n_sym = (int)m.matches.size();
    for (size_t i = 0; i < m.matches.size(); i++) { sym_q[i] = m.matches[i].queryIdx; sym_t[i] = m.matches[i].trainIdx; }
    if (m.matches.empty()) return 0;          // bestMatchesFilter dereferences begin() of an empty vector (App. B-3)
    m.sortMatches();
    for (size_t i = 0; i < m.sortedMatches.size(); i++) sorted_q[i] = m.sortedMatches[i].queryIdx;
    m.bestMatchesFilter(n_cells);
    std::vector<cv::KeyPoint> p1, p2;
    m.getGoodMatches(p1, p2);
    for (size_t i = 0; i < m.goodMatches.size(); i++) {
        good_q[i] = m.goodMatches[i].queryIdx;
        good_t[i] = m.goodMatches[i].trainIdx;
        good_d[i] = m.goodMatches[i].distance;
        prev_xy[2 * i] = p1[i].pt.x; prev_xy[2 * i + 1] = p1[i].pt.y;
        cur_xy[2 * i] = p2[i].pt.x; cur_xy[2 * i + 1] = p2[i].pt.y;
    }
    return (int)m.goodMatches.size();
}

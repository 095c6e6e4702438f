// Matcher.hpp — drop-in mirror of the reference's Matcher / MatcherGPU classes (include/Matcher.hpp:28-67,
// include/MatcherGPU.hpp:17-32): same public methods and public state, every computation on the B200 through
// the C ABI of include/vislam_b200.h.  OpenCV's BFMatcher / cuda::DescriptorMatcher are gone: the matcher kind
// is an enum value and the kNN runs in vsb_knn2_hamming / vsb_knn2_l2.
#ifndef VISLAM_MATCHER_HPP_
#define VISLAM_MATCHER_HPP_

#include <vector>

#include "vislam/compat.hpp"
#include "vislam/device.hpp"

// include/Matcher.hpp:13-21
enum matcherType {
    USE_BRUTE_FORCE,
    USE_BRUTE_FORCE_HAMMING,
    USE_FLANN,
    USE_BRUTE_FORCE_GPU,
    USE_BRUTE_FORCE_GPU_HAMMING
};

class Matcher {
public:
    Matcher();
    explicit Matcher(int _matcher);
    virtual ~Matcher() {}
    void setKeypoints(std::vector<cv::KeyPoint> _keypoints_1, std::vector<cv::KeyPoint> _keypoints_2);
    void setDescriptors(cv::Mat _descriptors_1, cv::Mat _descriptors_2);
    void setMatcher(int _matcher);
    void setImageDimensions(int w, int h);
    void computeMatches();
    void computeSymMatches();
    int bestMatchesFilter(int n_features);
    int nnFilter(std::vector<std::vector<cv::DMatch> >& matches, double nn_ratio);
    void resetVectorMatches(std::vector<cv::DMatch>& matches);
    void pushBackVectorMatches(std::vector<cv::DMatch>& matches);
    void getMatches(std::vector<cv::KeyPoint>& _matched1, std::vector<cv::KeyPoint>& _matched2);
    void getGoodMatches(std::vector<cv::KeyPoint>& _matched1, std::vector<cv::KeyPoint>& _matched2);
    void sortMatches();
    double getMatchPercentage();
    void computeBestMatches(int n_features);
    void getGrid(int n_features, std::vector<cv::KeyPoint>& grid_point);
    void printStatistics();
    void clear();

    std::vector<std::vector<cv::DMatch> > aux_matches1;
    std::vector<std::vector<cv::DMatch> > aux_matches2;
    std::vector<cv::DMatch> matches;
    std::vector<cv::DMatch> sortedMatches;
    std::vector<cv::DMatch> goodMatches;
    std::vector<cv::KeyPoint> keypoints_1, keypoints_2;
    cv::Mat descriptors_1, descriptors_2;

    int h_size, w_size;
    int nSymMatches;
    int nBestMatches;

    double elapsed_detect1, elapsed_detect2, elapsed_knn1, elapsed_knn2;
    double elapsed_symMatches, elapsed_sortMatches, elapsed_bestMatches;
    double matchPercentage;

    // ---- additions (not in the reference) ----
    int norm_type;   // 1 = Hamming (ORB), 0 = L2 (float descriptors); FLANN requests run the exact L2 search
    int sym_mode;    // 0 = the reference's de-facto symmetry test, 1 = intended
    bool verbose;    // the reference prints the matcher kind from setMatcher; default off here

protected:
    void run_knn();
    // device scratch (grown on demand)
    vi::DevBuf d_desc1_, d_desc2_, d_idx12_, d_idx21_, d_dist12_, d_dist21_, d_keep12_, d_keep21_;
    vi::DevBuf d_list_q_, d_list_t_, d_list_d_, d_cnt_, d_keys_, d_order_, d_kp1_, d_good_q_, d_good_t_, d_good_d_;
    // results of the last computeMatches() as the device produced them (rows the nnFilter later clears on the
    // host stay readable here, which is what the reference's symmetry test does with its cleared vectors)
    std::vector<int32_t> h_idx12_, h_idx21_;
    std::vector<float> h_dist12_, h_dist21_;
    int n1_, n2_;
    bool knn_valid_;
};

// include/MatcherGPU.hpp:17-32 — in this implementation every Matcher runs on the device; the GPU-named
// entry points are kept so code written against the reference's GPU front end compiles unchanged.
class MatcherGPU : public Matcher {
public:
    MatcherGPU();
    explicit MatcherGPU(int _matcher);
    void setGPUFrames(cv::Mat _frame1, cv::Mat _frame2);
    void computeGPUMatches();
    void setGPUMatcher(int _matcher);
    bool useGPU;
    int matcherType;
};

#endif

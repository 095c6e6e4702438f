// oracle/cvshim/opencv2/core.hpp — MINIMAL OpenCV type shim (TEST INFRASTRUCTURE ONLY).
// Just enough of cv:: for the reference's src/Matcher.cpp + include/Matcher.hpp to compile UNMODIFIED where
// they lie under /root/reference, so the oracle's restatement of the Matcher filter chain (nnFilter,
// computeSymMatches, sortMatches, bestMatchesFilter, getGoodMatches) can be checked against the
// reference's own code.  The only arithmetic supplied here is what OpenCV itself would supply:
// BFMatcher::knnMatch (exhaustive, sorted by distance then train index — checked against cv2 4.13 in
// tests/test_oracle_cv2.py) and cv::sortIdx (stable ascending).  Nothing in the product links this.
#ifndef VSO_CVSHIM_CORE_HPP
#define VSO_CVSHIM_CORE_HPP
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <ostream>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F
#define CV_SORT_EVERY_ROW 0
#define CV_SORT_ASCENDING 0

namespace cv {

enum { NORM_L2 = 4, NORM_HAMMING = 6 };

struct Point2f { float x = 0, y = 0; };

// cv::Point3_ / cv::Matx33f as the reference's Plus / DataReader / Imu sources use them (containers + the element-wise
// operators OpenCV defines; Matx product and Matx * Point3f accumulate in float like cv::Matx does)
template <typename T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename U> Point3_(const Point3_<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
};
template <typename T> Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a) { return Point3_<T>(-a.x, -a.y, -a.z); }
template <typename T> Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }
template <typename T> Point3_<T> operator/(const Point3_<T>& a, int s) { return Point3_<T>((T)(a.x / s), (T)(a.y / s), (T)(a.z / s)); }
template <typename T> Point3_<T> operator/(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x / s), (T)(a.y / s), (T)(a.z / s)); }
template <typename T> std::ostream& operator<<(std::ostream& o, const Point3_<T>& p) { return o << "[" << p.x << ", " << p.y << ", " << p.z << "]"; }
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

struct Matx33f {
    float val[9];
    Matx33f() { for (int i = 0; i < 9; i++) val[i] = 0.f; }
    float& operator()(int r, int c) { return val[3 * r + c]; }
    const float& operator()(int r, int c) const { return val[3 * r + c]; }
    Matx33f t() const {
        Matx33f m;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) m(r, c) = (*this)(c, r);
        return m;
    }
};
inline Matx33f operator*(const Matx33f& a, const Matx33f& b) {   // cv::Matx product: float accumulator, k ascending
    Matx33f m;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            float s = 0;
            for (int k = 0; k < 3; k++) s += a(r, k) * b(k, c);
            m(r, c) = s;
        }
    return m;
}
inline Point3f operator*(const Matx33f& a, const Point3f& p) {
    const float v[3] = {p.x, p.y, p.z};
    float o[3];
    for (int r = 0; r < 3; r++) {
        float s = 0;
        for (int k = 0; k < 3; k++) s += a(r, k) * v[k];
        o[r] = s;
    }
    return Point3f(o[0], o[1], o[2]);
}

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

struct DMatch {
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38F) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    int queryIdx, trainIdx, imgIdx;
    float distance;
};

class Mat {
public:
    int rows = 0, cols = 0, type_ = CV_8U;
    std::shared_ptr<std::vector<uint8_t>> buf;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    static int esz(int t) { return t == CV_8U ? 1 : (t == CV_64F ? 8 : 4); }
    void create(int r, int c, int t) {
        rows = r; cols = c; type_ = t;
        buf = std::make_shared<std::vector<uint8_t>>((size_t)r * c * esz(t), 0);
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    void release() { buf.reset(); rows = cols = 0; }
    int type() const { return type_; }
    bool empty() const { return !buf || rows == 0 || cols == 0; }
    uint8_t* data() const { return buf ? buf->data() : nullptr; }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(buf->data())[(size_t)r * cols + c]; }
    template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(buf->data())[(size_t)r * cols + c]; }
};

template <typename T>
class Ptr : public std::shared_ptr<T> {
public:
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
    template <typename U> Ptr(const std::shared_ptr<U>& o) : std::shared_ptr<T>(o) {}
};

class DescriptorMatcher {
public:
    virtual ~DescriptorMatcher() {}
    // exhaustive k-NN, rows sorted by (distance ascending, train index ascending); fewer than k train rows
    // gives shorter lists; imgIdx = 0 — cv::BFMatcher::knnMatch semantics
    virtual void knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& out, int k) {
        out.clear();
        out.resize(q.rows);
        for (int i = 0; i < q.rows; i++) {
            std::vector<std::pair<float, int>> d(t.rows);
            for (int j = 0; j < t.rows; j++) d[j] = std::make_pair(dist(q, i, t, j), j);
            int kk = std::min(k, t.rows);
            std::partial_sort(d.begin(), d.begin() + kk, d.end());
            for (int m = 0; m < kk; m++) {
                DMatch dm(i, d[m].second, d[m].first);
                dm.imgIdx = 0;
                out[i].push_back(dm);
            }
        }
    }
protected:
    int norm_ = NORM_L2;
    float dist(const Mat& a, int i, const Mat& b, int j) const {
        if (norm_ == NORM_HAMMING) {
            const uint8_t* x = a.data() + (size_t)i * a.cols;
            const uint8_t* y = b.data() + (size_t)j * b.cols;
            int s = 0;
            for (int c = 0; c < a.cols; c++) s += __builtin_popcount((unsigned)(x[c] ^ y[c]));
            return (float)s;
        }
        const float* x = reinterpret_cast<const float*>(a.data()) + (size_t)i * a.cols;
        const float* y = reinterpret_cast<const float*>(b.data()) + (size_t)j * b.cols;
        double s = 0;
        for (int c = 0; c < a.cols; c++) { float d = x[c] - y[c]; s += (double)d * (double)d; }
        return std::sqrt((float)s);
    }
};

class BFMatcher : public DescriptorMatcher {
public:
    explicit BFMatcher(int norm = NORM_L2) { norm_ = norm; }
    static Ptr<BFMatcher> create(int norm = NORM_L2) { return Ptr<BFMatcher>(new BFMatcher(norm)); }
};

class FlannBasedMatcher : public DescriptorMatcher {   // exhaustive stand-in; the hot path never selects FLANN
public:
    static Ptr<FlannBasedMatcher> create() { return Ptr<FlannBasedMatcher>(new FlannBasedMatcher()); }
};

// cv::sortIdx(src, dst, CV_SORT_EVERY_ROW + CV_SORT_ASCENDING) for CV_32F rows; stable (decision for ties)
inline void sortIdx(const Mat& src, Mat& dst, int /*flags*/) {
    dst.create(src.rows, src.cols, CV_32S);
    for (int r = 0; r < src.rows; r++) {
        std::vector<int> idx(src.cols);
        for (int c = 0; c < src.cols; c++) idx[c] = c;
        const float* p = reinterpret_cast<const float*>(src.data()) + (size_t)r * src.cols;
        std::stable_sort(idx.begin(), idx.end(), [p](int a, int b) { return p[a] < p[b]; });
        for (int c = 0; c < src.cols; c++) dst.at<int>(r, c) = idx[c];
    }
}

namespace xfeatures2d {}

}  // namespace cv
#endif

// oracle/refshim/opencv2/core.hpp — FUNCTIONAL OpenCV stand-in (TEST INFRASTRUCTURE ONLY, never linked by the product).
//
// Purpose: let the reference's own src/Camera.cpp and src/VISystem.cpp (plus Matcher/Plus/Imu) compile UNMODIFIED, where
// they lie under /root/reference, into oracle/_ref/libref_visystem.so, so that the oracle's restatement of the
// Gauss-Newton pose solve can be checked against the reference's own source text being executed.  OpenCV 3.2 itself is not
// in this image, so this header supplies what OpenCV would supply on that path, following OpenCV 3.2's own evaluation
// rules (modules/core/src/matop.cpp, matmul.cpp, convert.cpp, arithm.cpp, lapack.cpp — restated, not copied):
//   * cv::Mat with row/col/ROI views that alias the parent's storage, create() that keeps storage of the right shape;
//   * cv::MatExpr LAZY expression folding exactly as matop.cpp does it, because it changes the arithmetic:
//       (A - s) * k        -> one convertTo(alpha = k, beta = -s*k [double]) : fl(fl(a*k_f) + beta_f), not fl(fl(a - s)*k);
//       A.t() * B, a*A.t()*B, -A.t() * B -> one gemm with transposition flags / alpha, accumulated in double, rounded once;
//       A.inv() * b        -> cv::solve(A, b, DECOMP_LU) (MatOp_Invert::matmul), NOT invert-then-multiply;
//   * cv::gemm on CV_32F: double accumulator, T(s * alpha) (GEMMSingleMul<float,double>);
//   * cv::solve / cv::invert DECOMP_LU for n > 3: hal::LU32f with partial pivoting, eps = FLT_EPSILON*10, back substitution
//     s / pivot (the form pinned bit-exactly against cv2 4.13; OpenCV 3.2 multiplied by the stored reciprocal instead —
//     define VSO_LU_RECIPROCAL to get that form);
//   * cv::resize(.., 0.5, 0.5) INTER_LINEAR 8-bit: exact-halving goes to the 2x2 area mean (a+b+c+d+2)>>2, otherwise the
//     11-bit fixed-point bilinear of resizeGeneric_/HResizeLinear/VResizeLinear; cv::Scharr 8U->16S with scale;
//   * BFMatcher::knnMatch, sortIdx as in oracle/cvshim.
// Each of those primitives is separately pinned against the real OpenCV (cv2 4.13) by tests/test_oracle_cv2.py /
// tests/test_ref_visystem.py.  Everything else the two translation units mention (GUI, FileStorage, calib3d, sfm, feature
// detectors) only has to compile: it is declared as an inline stub that aborts if it is ever reached.
#ifndef VSO_REFSHIM_CORE_HPP
#define VSO_REFSHIM_CORE_HPP
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <ostream>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_SORT_EVERY_ROW 0
#define CV_SORT_ASCENDING 0
#define CV_SORT_DESCENDING 16
#define CV_PI 3.1415926535897932384626433832795
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_Assert(x) do { if (!(x)) cv::shim_fail("CV_Assert(" #x ")"); } while (0)

namespace cv {

struct Exception : std::runtime_error { explicit Exception(const std::string& w) : std::runtime_error(w) {} };
[[noreturn]] inline void shim_fail(const char* what) { throw Exception(std::string("refshim: ") + what); }
// a value of any type for the stubs that only have to compile
struct Unreached {
    template <typename T> operator T() const { shim_fail("unreached stub value used"); }
};
#define VSO_STUB(name) template <typename... A> inline Unreached name(A&&...) { shim_fail("stub " #name " called"); }

typedef std::string String;
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_CHOLESKY = 3 };
enum { GEMM_1_T = 1, GEMM_2_T = 2, GEMM_3_T = 4 };
enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { FONT_HERSHEY_SIMPLEX = 0, LINE_AA = 16 };
enum { RANSAC = 8, LMEDS = 4 };
enum { COLOR_GRAY2BGR = 8, COLOR_BGR2GRAY = 6, CV_GRAY2BGR = 8, CV_BGR2GRAY = 6 };
enum { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1 };

inline int cvRound(double v) { return (int)lrint(v); }   // round-half-even in the default rounding mode
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline short saturate_cast<short>(double v) { int i = cvRound(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }
template <> inline ushort saturate_cast<ushort>(double v) { int i = cvRound(v); return (ushort)(i < 0 ? 0 : i > 65535 ? 65535 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

// ---------------------------------------------------------------------------------------------- small value types
template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
template <typename T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> Point_<T> operator*(const Point_<T>& a, double s) { return Point_<T>((T)(a.x * s), (T)(a.y * s)); }
template <typename T> std::ostream& operator<<(std::ostream& o, const Point_<T>& p) { return o << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename U> Point3_(const Point3_<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    T dot(const Point3_& o) const { return (T)(x * o.x + y * o.y + z * o.z); }
    Point3_ cross(const Point3_& o) const { return Point3_(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
};
template <typename T> Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> Point3_<T> operator-(const Point3_<T>& a) { return Point3_<T>(-a.x, -a.y, -a.z); }
template <typename T> Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }
template <typename T> Point3_<T> operator*(double s, const Point3_<T>& a) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }
template <typename T> Point3_<T> operator/(const Point3_<T>& a, int s) { return Point3_<T>((T)(a.x / s), (T)(a.y / s), (T)(a.z / s)); }
template <typename T> Point3_<T> operator/(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x / s), (T)(a.y / s), (T)(a.z / s)); }
template <typename T> Point3_<T>& operator+=(Point3_<T>& a, const Point3_<T>& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
template <typename T> std::ostream& operator<<(std::ostream& o, const Point3_<T>& p) { return o << "[" << p.x << ", " << p.y << ", " << p.z << "]"; }
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
typedef Point3_<int> Point3i;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;
typedef Size_<float> Size2f;
template <typename T> std::ostream& operator<<(std::ostream& o, const Size_<T>& s) { return o << "[" << s.width << " x " << s.height << "]"; }

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
    Rect_(const Point_<T>& p, const Point_<T>& q) : x(std::min(p.x, q.x)), y(std::min(p.y, q.y)), width(std::max(p.x, q.x) - std::min(p.x, q.x)), height(std::max(p.y, q.y) - std::min(p.y, q.y)) {}
};
typedef Rect_<int> Rect;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double a) { val[0] = a; val[1] = val[2] = val[3] = 0; }
    Scalar(double a, double b, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
    bool isReal() const { return val[1] == 0 && val[2] == 0 && val[3] == 0; }
    bool isZero() const { return val[0] == 0 && isReal(); }
};
inline Scalar operator*(const Scalar& a, double s) { return Scalar(a[0] * s, a[1] * s, a[2] * s, a[3] * s); }
inline Scalar operator-(const Scalar& a) { return Scalar(-a[0], -a[1], -a[2], -a[3]); }
inline Scalar operator+(const Scalar& a, const Scalar& b) { return Scalar(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]); }
inline Scalar operator-(const Scalar& a, const Scalar& b) { return Scalar(a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]); }

// cv::Matx: products accumulate in the element type, k ascending (matx.hpp Matx_MatMulOp)
template <typename T, int M, int N>
struct Matx {
    T val[M * N];
    Matx() { for (int i = 0; i < M * N; i++) val[i] = T(0); }
    template <typename U> Matx(const Matx<U, M, N>& o) { for (int i = 0; i < M * N; i++) val[i] = (T)o.val[i]; }
    Matx(T a, T b, T c) { static_assert(M * N == 3, ""); val[0] = a; val[1] = b; val[2] = c; }
    Matx(T a, T b, T c, T d) { static_assert(M * N == 4, ""); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    Matx(T a, T b, T c, T d, T e, T f, T g, T h, T i) {
        static_assert(M * N == 9, "");
        val[0] = a; val[1] = b; val[2] = c; val[3] = d; val[4] = e; val[5] = f; val[6] = g; val[7] = h; val[8] = i;
    }
    static Matx eye() { Matx m; for (int i = 0; i < (M < N ? M : N); i++) m(i, i) = T(1); return m; }
    static Matx zeros() { return Matx(); }
    T& operator()(int r, int c) { return val[N * r + c]; }
    const T& operator()(int r, int c) const { return val[N * r + c]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    Matx<T, N, M> t() const {
        Matx<T, N, M> m;
        for (int r = 0; r < M; r++) for (int c = 0; c < N; c++) m(c, r) = (*this)(r, c);
        return m;
    }
    Matx inv() const { shim_fail("Matx::inv not provided"); }
};
template <typename T, int M, int K, int N>
Matx<T, M, N> operator*(const Matx<T, M, K>& a, const Matx<T, K, N>& b) {
    Matx<T, M, N> m;
    for (int r = 0; r < M; r++)
        for (int c = 0; c < N; c++) {
            T s = 0;
            for (int k = 0; k < K; k++) s += a(r, k) * b(k, c);
            m(r, c) = s;
        }
    return m;
}
template <typename T, int M, int N> Matx<T, M, N> operator+(const Matx<T, M, N>& a, const Matx<T, M, N>& b) { Matx<T, M, N> m; for (int i = 0; i < M * N; i++) m.val[i] = a.val[i] + b.val[i]; return m; }
template <typename T, int M, int N> Matx<T, M, N> operator-(const Matx<T, M, N>& a, const Matx<T, M, N>& b) { Matx<T, M, N> m; for (int i = 0; i < M * N; i++) m.val[i] = a.val[i] - b.val[i]; return m; }
template <typename T, int M, int N> Matx<T, M, N> operator-(const Matx<T, M, N>& a) { Matx<T, M, N> m; for (int i = 0; i < M * N; i++) m.val[i] = -a.val[i]; return m; }
template <typename T, int M, int N> Matx<T, M, N> operator*(const Matx<T, M, N>& a, double s) { Matx<T, M, N> m; for (int i = 0; i < M * N; i++) m.val[i] = (T)(a.val[i] * s); return m; }
template <typename T, int M, int N> Matx<T, M, N> operator*(double s, const Matx<T, M, N>& a) { return a * s; }
template <typename T> Point3_<T> operator*(const Matx<T, 3, 3>& a, const Point3_<T>& p) {
    Matx<T, 3, 1> v(p.x, p.y, p.z);
    Matx<T, 3, 1> o = a * v;
    return Point3_<T>(o(0), o(1), o(2));
}
template <typename T, int M, int N> std::ostream& operator<<(std::ostream& o, const Matx<T, M, N>& m) {
    o << "[";
    for (int r = 0; r < M; r++) { for (int c = 0; c < N; c++) o << m(r, c) << (c + 1 < N ? ", " : ""); o << (r + 1 < M ? ";\n " : ""); }
    return o << "]";
}
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<double, 3, 3> Matx33d;
typedef Matx<float, 3, 1> Matx31f;
typedef Matx<double, 3, 1> Matx31d;
typedef Matx<float, 4, 4> Matx44f;
typedef Matx<float, 3, 4> Matx34f;
template <typename T, int N> struct Vec : Matx<T, N, 1> {
    Vec() {}
    Vec(T a, T b, T c) { static_assert(N == 3, ""); this->val[0] = a; this->val[1] = b; this->val[2] = c; }
    T& operator[](int i) { return this->val[i]; }
    const T& operator[](int i) const { return this->val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<float, 3> Vec3f;
typedef Vec<double, 3> Vec3d;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
    static void convert(const std::vector<KeyPoint>& k, std::vector<Point2f>& p, const std::vector<int>& = std::vector<int>()) {
        p.resize(k.size());
        for (size_t i = 0; i < k.size(); i++) p[i] = k[i].pt;
    }
    static void convert(const std::vector<Point2f>& p, std::vector<KeyPoint>& k, float size = 1, float response = 1, int octave = 0, int class_id = -1) {
        k.resize(p.size());
        for (size_t i = 0; i < p.size(); i++) k[i] = KeyPoint(p[i], size, -1, response, octave, class_id);
    }
};

struct DMatch {
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    int queryIdx, trainIdx, imgIdx;
    float distance;
};

template <typename T>
class Ptr : public std::shared_ptr<T> {
public:
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    Ptr(const Unreached&) {}
    template <typename U> Ptr(const Ptr<U>& o) : std::shared_ptr<T>(o) {}
    template <typename U> Ptr(const std::shared_ptr<U>& o) : std::shared_ptr<T>(o) {}
    bool empty() const { return !this->get(); }
};

// ---------------------------------------------------------------------------------------------- Mat
class MatExpr;
class Mat {
public:
    int flags = CV_8U;     // the type
    int dims = 2, rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;       // bytes between rows
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(Size s, int t) { create(s.height, s.width, t); }
    Mat(int r, int c, int t, const Scalar& s) { create(r, c, t); setTo(s); }
    Mat(Size sz, int t, const Scalar& s) { create(sz.height, sz.width, t); setTo(s); }
    Mat(int r, int c, int t, void* d, size_t st = 0) : flags(t), rows(r), cols(c), data((uchar*)d) { step = st ? st : (size_t)c * esz(t); }
    Mat(const Mat& m, const Range& rr, const Range& cr) { *this = m.sub(rr, cr); }
    Mat(const Mat& m, const Rect& r) { *this = m.sub(Range(r.y, r.y + r.height), Range(r.x, r.x + r.width)); }
    Mat(const Unreached&) {}
    template <typename T, int M, int N> explicit Mat(const Matx<T, M, N>& m, bool = true) {
        create(M, N, depth_of((T*)0));
        for (int r = 0; r < M; r++) for (int c = 0; c < N; c++) at<T>(r, c) = m(r, c);
    }
    template <typename T> explicit Mat(const std::vector<T>& v, bool = false) {
        create((int)v.size(), 1, type_of((T*)0));
        if (!v.empty()) memcpy(data, v.data(), v.size() * sizeof(T));
    }
    template <typename T> explicit Mat(const Point3_<T>& p, bool = true) {
        create(3, 1, depth_of((T*)0));
        at<T>(0, 0) = p.x; at<T>(1, 0) = p.y; at<T>(2, 0) = p.z;
    }
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Unreached&) { shim_fail("unreached stub value assigned"); }
    Mat& operator=(const Scalar& s) { setTo(s); return *this; }
    const Mat& operator=(const Scalar& s) const { const_cast<Mat*>(this)->setTo(s); return *this; }

    static int depth_of(uchar*) { return CV_8U; }
    static int depth_of(short*) { return CV_16S; }
    static int depth_of(ushort*) { return CV_16U; }
    static int depth_of(int*) { return CV_32S; }
    static int depth_of(float*) { return CV_32F; }
    static int depth_of(double*) { return CV_64F; }
    template <typename T> static int type_of(T* p) { return depth_of(p); }
    template <typename T> static int type_of(Point_<T>*) { return CV_MAKETYPE(depth_of((T*)0), 2); }
    template <typename T> static int type_of(Point3_<T>*) { return CV_MAKETYPE(depth_of((T*)0), 3); }
    static int esz1(int t) { static const int s[8] = {1, 1, 2, 2, 4, 4, 8, 0}; return s[CV_MAT_DEPTH(t)]; }
    static int esz(int t) { return esz1(t) * CV_MAT_CN(t); }

    void create(int r, int c, int t) {
        if (data && rows == r && cols == c && flags == t) return;       // Mat::create keeps matching storage (views!)
        rows = r; cols = c; flags = t;
        step = (size_t)c * esz(t);
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, 0);
        data = buf->data();
    }
    void create(Size s, int t) { create(s.height, s.width, t); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return esz(flags); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * esz(flags) || rows <= 1; }
    Size size() const { return Size(cols, rows); }

    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    // Mat::at does not check bounds in release builds of OpenCV; the reference relies on that in one place
    // (VISystem.cpp:1321, SURVEY App. B-4: round(y2) can equal rows).  An access past the last element of the matrix is
    // counted and served from a zero cell, so that such runs are detectable instead of reading the heap; running past the
    // end of a ROW of a continuous matrix lands in the next row, as it does in OpenCV.
    template <typename T> T& at(int r, int c) {
        const long long off = (long long)r * (long long)step + (long long)c * (long long)sizeof(T);
        const long long end = rows > 0 ? (long long)(rows - 1) * (long long)step + (long long)cols * esz(flags) : 0;
        if (off < 0 || off + (long long)sizeof(T) > end) { oob_reads()++; static thread_local double zero; zero = 0; return *(T*)&zero; }
        return *(T*)(data + off);
    }
    template <typename T> const T& at(int r, int c) const { return const_cast<Mat*>(this)->at<T>(r, c); }
    static long long& oob_reads() { static long long n = 0; return n; }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : (cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols)); }
    template <typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }

    Mat sub(Range rr, Range cr) const {
        if (rr.start == INT_MIN) rr = Range(0, rows);
        if (cr.start == INT_MIN) cr = Range(0, cols);
        if (!(0 <= rr.start && rr.start <= rr.end && rr.end <= rows && 0 <= cr.start && cr.start <= cr.end && cr.end <= cols))
            shim_fail("Mat range out of bounds (OpenCV would throw)");
        Mat m;
        m.flags = flags; m.rows = rr.end - rr.start; m.cols = cr.end - cr.start; m.step = step; m.buf = buf;
        m.data = data + (size_t)rr.start * step + (size_t)cr.start * esz(flags);
        return m;
    }
    Mat row(int r) const { return sub(Range(r, r + 1), Range::all()); }
    Mat col(int c) const { return sub(Range::all(), Range(c, c + 1)); }
    Mat rowRange(int a, int b) const { return sub(Range(a, b), Range::all()); }
    Mat colRange(int a, int b) const { return sub(Range::all(), Range(a, b)); }
    Mat operator()(const Rect& r) const { return Mat(*this, r); }
    Mat operator()(Range rr, Range cr) const { return sub(rr, cr); }

    void copyTo(Mat& m) const {
        if (empty()) { m.release(); return; }
        if (m.data == data && m.rows == rows && m.cols == cols) return;
        Mat keep = *this;                           // m may alias this
        m.create(keep.rows, keep.cols, keep.flags);
        size_t w = (size_t)keep.cols * esz(keep.flags);
        for (int r = 0; r < keep.rows; r++) memmove(m.ptr(r), keep.ptr(r), w);
    }
    void copyTo(Mat& m, const Mat& /*mask*/) const { shim_fail("masked copyTo not provided"); }
    Mat clone() const { Mat m; Mat src = *this; m.buf.reset(); m.data = nullptr; src.copyTo(m); return m; }
    double getd(int r, int c) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(r, c);
            case CV_8S: return at<signed char>(r, c);
            case CV_16U: return at<ushort>(r, c);
            case CV_16S: return at<short>(r, c);
            case CV_32S: return at<int>(r, c);
            case CV_32F: return at<float>(r, c);
            default: return at<double>(r, c);
        }
    }
    void setd(int r, int c, double v) {
        switch (depth()) {
            case CV_8U: at<uchar>(r, c) = saturate_cast<uchar>(v); break;
            case CV_8S: at<signed char>(r, c) = (signed char)std::max(-128, std::min(127, cvRound(v))); break;
            case CV_16U: at<ushort>(r, c) = saturate_cast<ushort>(v); break;
            case CV_16S: at<short>(r, c) = saturate_cast<short>(v); break;
            case CV_32S: at<int>(r, c) = cvRound(v); break;
            case CV_32F: at<float>(r, c) = (float)v; break;
            default: at<double>(r, c) = v;
        }
    }
    Mat& setTo(const Scalar& s) {
        if (channels() != 1) { if (s.isZero()) { for (int r = 0; r < rows; r++) memset(ptr(r), 0, (size_t)cols * esz(flags)); return *this; } shim_fail("multi-channel setTo"); }
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) setd(r, c, s[0]);
        return *this;
    }
    // convertTo with scale: cvtScale_<..., float> for everything but CV_64F work types — src*alpha_f + beta_f evaluated in
    // float without contraction (OpenCV 3.2 cvtScale32f; SSE2 mul + add), double when source or destination is CV_64F
    void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const {
        Mat src = *this;
        int dt = rtype < 0 ? src.flags : CV_MAKETYPE(CV_MAT_DEPTH(rtype), src.channels());
        if (src.channels() != 1) shim_fail("multi-channel convertTo");
        Mat dst;
        if (m.data && m.rows == src.rows && m.cols == src.cols && m.flags == dt) dst = m; else dst.create(src.rows, src.cols, dt);
        bool noscale = std::fabs(alpha - 1) < DBL_EPSILON && std::fabs(beta) < DBL_EPSILON;
        bool wide = src.depth() == CV_64F || CV_MAT_DEPTH(dt) == CV_64F || src.depth() == CV_32S;
        for (int r = 0; r < src.rows; r++)
            for (int c = 0; c < src.cols; c++) {
                if (noscale) dst.setd(r, c, src.getd(r, c));
                else if (wide) { volatile double p = src.getd(r, c) * alpha; dst.setd(r, c, p + beta); }
                else { volatile float p = (float)src.getd(r, c) * (float)alpha; volatile float q = p + (float)beta; dst.setd(r, c, q); }
            }
        m = dst;
    }
    void push_back(const Mat& m) {
        if (m.empty()) return;
        if (empty()) { Mat c = m.clone(); *this = c; return; }
        if (m.cols != cols || m.flags != flags) shim_fail("push_back shape/type mismatch (OpenCV would throw)");
        const size_t w = (size_t)cols * esz(flags);
        if (buf && buf.use_count() == 1 && data == buf->data() && step == w && m.buf != buf) {
            // sole owner of a tight buffer: grow it in place (amortised, like Mat::reserve); values are unchanged
            const size_t need = (size_t)(rows + m.rows) * w + 16;
            if (buf->size() < need) { buf->resize(std::max(need, 2 * buf->size())); data = buf->data(); }
            for (int r = 0; r < m.rows; r++) memcpy(data + (size_t)(rows + r) * w, m.ptr(r), w);
            rows += m.rows;
            return;
        }
        Mat old = *this;
        Mat n;
        n.create(old.rows + m.rows, cols, flags);
        for (int r = 0; r < old.rows; r++) memcpy(n.ptr(r), old.ptr(r), w);
        for (int r = 0; r < m.rows; r++) memcpy(n.ptr(old.rows + r), m.ptr(r), w);
        *this = n;
    }
    void push_back(const MatExpr& e);
    template <typename T> void push_back(const T& v) {
        Mat one(1, 1, type_of((T*)0));
        memcpy(one.data, &v, sizeof(T));
        if (!empty() && cols != 1) shim_fail("push_back(value) on a multi-column Mat");
        push_back(one);
    }
    void pop_back(size_t n = 1) { rows -= (int)n; }
    Mat reshape(int cn, int r = 0) const {
        if (!isContinuous()) shim_fail("reshape on a non-continuous Mat");
        Mat m = *this;
        size_t totalscalars = (size_t)rows * cols * channels();
        if (cn == 0) cn = channels();
        m.flags = CV_MAKETYPE(depth(), cn);
        m.rows = r ? r : rows;
        m.cols = (int)(totalscalars / cn / m.rows);
        m.step = (size_t)m.cols * esz(m.flags);
        return m;
    }
    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;
    MatExpr mul(const Mat& m, double scale = 1) const;
    MatExpr mul(const MatExpr& m, double scale = 1) const;
    MatExpr mul(double s, double scale = 1) const;
    double dot(const Mat& m) const { double s = 0; for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) s += getd(r, c) * m.getd(r, c); return s; }
    Mat cross(const Mat&) const { shim_fail("Mat::cross not provided"); }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    static Mat zeros(Size s, int t) { return Mat(s, t); }
    static Mat ones(int r, int c, int t) { return Mat(r, c, t, Scalar(1)); }
    static Mat ones(Size s, int t) { return Mat(s, t, Scalar(1)); }
    static Mat eye(int r, int c, int t) { Mat m(r, c, t); for (int i = 0; i < std::min(r, c); i++) m.setd(i, i, 1); return m; }
    static Mat eye(Size s, int t) { return eye(s.height, s.width, t); }
    template <typename T, int M, int N> operator Matx<T, M, N>() const {
        if (rows != M || cols != N) shim_fail("Mat -> Matx size mismatch");
        Matx<T, M, N> m;
        for (int r = 0; r < M; r++) for (int c = 0; c < N; c++) m(r, c) = (T)getd(r, c);
        return m;
    }
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, Mat::depth_of((T*)0)) {}
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    // Mat_<T>(r, c) << a, b, c ... comma initialiser
    struct Init {
        Mat_* m; int i;
        Init& operator,(T v) { m->template at<T>(i / m->cols, i % m->cols) = v; i++; return *this; }
        operator Mat_() const { return *m; }
        operator Mat() const { return *m; }
    };
    Init operator<<(T v) { Init in = {this, 0}; in, v; return in; }
};
typedef Mat_<float> Mat1f;
typedef Mat_<double> Mat1d;

typedef const Mat& InputArray;
typedef Mat& OutputArray;
typedef Mat& InputOutputArray;
typedef const std::vector<Mat>& InputArrayOfArrays;
typedef std::vector<Mat>& OutputArrayOfArrays;
inline Mat& noArray() { static Mat m; m.release(); return m; }

inline std::ostream& operator<<(std::ostream& o, const Mat& m) {
    o << "[";
    for (int r = 0; r < m.rows; r++) {
        for (int c = 0; c < m.cols * m.channels(); c++) {
            Mat one = m;
            one.flags = m.depth();
            o << one.getd(r, c) << (c + 1 < m.cols * m.channels() ? ", " : "");
        }
        o << (r + 1 < m.rows ? ";\n " : "");
    }
    return o << "]";
}

// ---------------------------------------------------------------------------------------------- array primitives
namespace shim {
inline void check_same(const Mat& a, const Mat& b) {
    if (a.rows != b.rows || a.cols != b.cols || a.flags != b.flags) shim_fail("array size/type mismatch (OpenCV would throw)");
}
// element-wise binary operation in the arrays' own precision (float arrays: float arithmetic; others: double + saturate)
template <typename F> inline void binary(const Mat& a0, const Mat& b0, Mat& dst, F f) {
    Mat a = a0, b = b0;
    check_same(a, b);
    Mat out;
    if (dst.data && dst.rows == a.rows && dst.cols == a.cols && dst.flags == a.flags) out = dst; else out.create(a.rows, a.cols, a.flags);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < a.cols * a.channels(); c++) {
            Mat a1 = a, b1 = b, o1 = out;
            a1.flags = b1.flags = o1.flags = a.depth();
            if (a.depth() == CV_32F) { volatile float v = f(a1.at<float>(r, c), b1.at<float>(r, c)); o1.at<float>(r, c) = v; }
            else { volatile double v = f(a1.getd(r, c), b1.getd(r, c)); o1.setd(r, c, v); }
        }
    dst = out;
}
struct OpAdd { template <typename T> T operator()(T x, T y) const { return x + y; } };
struct OpSub { template <typename T> T operator()(T x, T y) const { return x - y; } };
struct OpMul { template <typename T> T operator()(T x, T y) const { return x * y; } };
struct OpDiv { template <typename T> T operator()(T x, T y) const { return x / y; } };
}  // namespace shim

inline void add(const Mat& a, const Mat& b, Mat& dst) { shim::binary(a, b, dst, shim::OpAdd()); }
inline void subtract(const Mat& a, const Mat& b, Mat& dst) { shim::binary(a, b, dst, shim::OpSub()); }
// array (+|-) scalar: the scalar is converted to the array's working type first (arithm_op, convertAndUnrollScalar)
inline void add(const Mat& a0, const Scalar& s, Mat& dst) {
    Mat a = a0, sm(a0.rows, a0.cols, a0.flags, Scalar(s[0]));
    if (a.channels() != 1) shim_fail("multi-channel scalar add");
    shim::binary(a, sm, dst, shim::OpAdd());
}
inline void subtract(const Scalar& s, const Mat& a0, Mat& dst) {
    Mat a = a0, sm(a0.rows, a0.cols, a0.flags, Scalar(s[0]));
    shim::binary(sm, a, dst, shim::OpSub());
}
inline void subtract(const Mat& a0, const Scalar& s, Mat& dst) {
    Mat a = a0, sm(a0.rows, a0.cols, a0.flags, Scalar(s[0]));
    shim::binary(a, sm, dst, shim::OpSub());
}
// cv::multiply / cv::divide: scale == 1 is a plain product / quotient; otherwise float: a*scale_f*b (mul_) resp. a*scale_f/b
inline void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1) {
    if (scale == 1) { shim::binary(a, b, dst, shim::OpMul()); return; }
    const float sf = (float)scale;
    shim::binary(a, b, dst, [sf, scale](double x, double y) -> double { (void)scale; volatile float p = sf * (float)x; volatile float q = p * (float)y; return q; });
}
inline void divide(const Mat& a, const Mat& b, Mat& dst, double scale = 1) {
    if (scale == 1) { shim::binary(a, b, dst, shim::OpDiv()); return; }
    const float sf = (float)scale;
    shim::binary(a, b, dst, [sf](double x, double y) -> double { volatile float p = (float)x * sf; volatile float q = p / (float)y; return q; });
}
inline void divide(double scale, const Mat& b0, Mat& dst) {
    Mat b = b0, sm(b0.rows, b0.cols, b0.flags, Scalar(scale));
    shim::binary(sm, b, dst, shim::OpDiv());
}
inline void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst) {
    if (a.depth() == CV_32F) { const float af = (float)alpha; shim::binary(a, b, dst, [af](double x, double y) -> double { volatile float p = (float)x * af; volatile float q = p + (float)y; return q; }); }
    else shim::binary(a, b, dst, [alpha](double x, double y) -> double { volatile double p = x * alpha; return p + y; });
}
inline void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst, int = -1) {
    // addWeighted_<T, float/double>: saturate(a*alpha + b*beta + gamma); 8-bit goes through float tables in 3.2 with the same value
    if (a.depth() == CV_64F) shim::binary(a, b, dst, [=](double x, double y) -> double { return x * alpha + y * beta + gamma; });
    else { const float fa = (float)alpha, fb = (float)beta, fg = (float)gamma; shim::binary(a, b, dst, [=](double x, double y) -> double { volatile float p = (float)x * fa; volatile float q = (float)y * fb; volatile float s = p + q; volatile float t = s + fg; return t; }); }
}
inline void convertScaleAbs(const Mat& src0, Mat& dst, double alpha = 1, double beta = 0) {
    Mat src = src0, out(src0.rows, src0.cols, CV_MAKETYPE(CV_8U, src0.channels()));
    for (int r = 0; r < src.rows; r++) for (int c = 0; c < src.cols; c++) {
        volatile float v = (float)src.getd(r, c) * (float)alpha; volatile float w = v + (float)beta;
        out.at<uchar>(r, c) = saturate_cast<uchar>(std::fabs((double)w));
    }
    dst = out;
}
inline void transpose(const Mat& a0, Mat& dst) {
    Mat a = a0, out(a0.cols, a0.rows, a0.flags);
    size_t e = Mat::esz(a.flags);
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) memcpy(out.ptr(c) + e * r, a.ptr(r) + e * c, e);
    dst = out;
}
// cv::gemm for CV_32F / CV_64F: dst = alpha*op(A)*op(B) + beta*op(C), accumulated in double, one rounding (GEMMSingleMul /
// GEMMBlockMul + GEMMStore with WT = double)
inline void gemm(const Mat& A0, const Mat& B0, double alpha, const Mat& C0, double beta, Mat& dst, int flags = 0) {
    Mat A = A0, B = B0, C = C0;
    if (A.flags != B.flags || (A.flags != CV_32F && A.flags != CV_64F)) shim_fail("gemm type mismatch (OpenCV would throw)");
    const bool ta = flags & GEMM_1_T, tb = flags & GEMM_2_T, tc = flags & GEMM_3_T;
    const int M = ta ? A.cols : A.rows, K = ta ? A.rows : A.cols, Kb = tb ? B.cols : B.rows, N = tb ? B.rows : B.cols;
    if (K != Kb) shim_fail("gemm inner dimension mismatch (OpenCV would throw)");
    Mat out(M, N, A.flags);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int k = 0; k < K; k++) {
                double a = ta ? A.getd(k, i) : A.getd(i, k), b = tb ? B.getd(j, k) : B.getd(k, j);
                volatile double p = a * b;
                s += p;
            }
            volatile double v = s * alpha;
            if (!C.empty() && beta != 0) { volatile double cc = (tc ? C.getd(j, i) : C.getd(i, j)) * beta; v = v + cc; }
            out.setd(i, j, v);
        }
    dst = out;
}
// hal::LU32f / LU64f (lapack.cpp LUImpl): in-place LU with partial pivoting on [A | b]; returns 0 when singular
template <typename T> inline int LUImpl(T* A, int astep, int m, T* b, int bstep, int n, T eps) {
    int p = 1;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++) if (std::abs(A[j * astep + i]) > std::abs(A[k * astep + i])) k = j;
        if (std::abs(A[k * astep + i]) < eps) return 0;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * astep + j], A[k * astep + j]);
            if (b) for (int j = 0; j < n; j++) std::swap(b[i * bstep + j], b[k * bstep + j]);
            p = -p;
        }
        volatile T d = -1 / A[i * astep + i];
        for (int j = i + 1; j < m; j++) {
            volatile T alpha = A[j * astep + i] * d;
            for (k = i + 1; k < m; k++) { volatile T t = alpha * A[i * astep + k]; A[j * astep + k] += t; }
            if (b) for (k = 0; k < n; k++) { volatile T t = alpha * b[i * bstep + k]; b[j * bstep + k] += t; }
        }
#ifdef VSO_LU_RECIPROCAL
        A[i * astep + i] = -d;
#endif
    }
    if (b) {
        for (int i = m - 1; i >= 0; i--)
            for (int j = 0; j < n; j++) {
                T s = b[i * bstep + j];
                for (int k = i + 1; k < m; k++) { volatile T t = A[i * astep + k] * b[k * bstep + j]; s -= t; }
#ifdef VSO_LU_RECIPROCAL
                b[i * bstep + j] = s * A[i * astep + i];
#else
                b[i * bstep + j] = s / A[i * astep + i];
#endif
            }
    }
    return p;
}
// cv::solve(A, B, X, DECOMP_LU), n > 3 general path (n <= 3 has closed forms in OpenCV which this path does not need)
inline bool solve(const Mat& A0, const Mat& B0, Mat& dst, int method = DECOMP_LU) {
    Mat A = A0.clone(), X = B0.clone();
    if (method != DECOMP_LU || A.rows != A.cols || A.rows <= 3) shim_fail("solve: only DECOMP_LU with n > 3 is provided");
    int ok;
    if (A.flags == CV_32F) ok = LUImpl<float>(A.ptr<float>(), A.cols, A.rows, X.ptr<float>(), X.cols, X.cols, FLT_EPSILON * 10);
    else if (A.flags == CV_64F) ok = LUImpl<double>(A.ptr<double>(), A.cols, A.rows, X.ptr<double>(), X.cols, X.cols, DBL_EPSILON * 100);
    else shim_fail("solve type");
    if (!ok) X.setTo(Scalar(0));
    dst = X;
    return ok != 0;
}
inline double invert(const Mat& A0, Mat& dst, int method = DECOMP_LU) {
    if (A0.rows != A0.cols || A0.rows <= 3) shim_fail("invert: only n > 3 is provided");
    Mat I = Mat::eye(A0.rows, A0.cols, A0.flags);
    Mat X;
    bool ok = solve(A0, I, X, method);
    dst = X;
    return ok ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------- MatExpr (matop.cpp)
class MatExpr {
public:
    enum Op { IDENT, ADDEX, BIN, T, GEMM, INV, SOLVE };
    Op op = IDENT;
    int flags = 0;
    Mat a, b, c;
    double alpha = 1, beta = 0;
    Scalar s;
    MatExpr() {}
    MatExpr(const Mat& m) : a(m) {}
    MatExpr(const Unreached&) {}
    static MatExpr make(Op op, int flags, const Mat& a, const Mat& b = Mat(), const Mat& c = Mat(), double alpha = 1, double beta = 0, const Scalar& s = Scalar()) {
        MatExpr e; e.op = op; e.flags = flags; e.a = a; e.b = b; e.c = c; e.alpha = alpha; e.beta = beta; e.s = s; return e;
    }
    bool isAddEx() const { return op == ADDEX; }
    bool isScaled() const { return op == ADDEX && (!b.data || beta == 0) && s.isZero() && s.isReal(); }
    bool isT() const { return op == T; }
    bool isBin(char ch) const { return op == BIN && flags == ch; }
    bool isReciprocal() const { return isBin('/') && (!b.data || beta == 0); }

    void assign(Mat& m, int type = -1) const {
        switch (op) {
            case IDENT:
                if (type == -1 || type == a.type()) m = a; else a.convertTo(m, type);
                break;
            case ADDEX: assignAddEx(m, type); break;
            case BIN:
                if (flags == '*') cv::multiply(a, b, m, alpha);
                else if (flags == '/' && b.data) cv::divide(a, b, m, alpha);
                else if (flags == '/' && !b.data) cv::divide(alpha, a, m);
                else shim_fail("MatExpr binary op not provided");
                break;
            case T: {
                Mat tmp;
                cv::transpose(a, tmp);
                if (alpha != 1) tmp.convertTo(tmp, -1, alpha);
                assignInto(m, tmp);
                break; }
            case GEMM: { Mat tmp; cv::gemm(a, b, alpha, c, beta, tmp, flags); assignInto(m, tmp); break; }
            case INV: { Mat tmp; cv::invert(a, tmp, flags); assignInto(m, tmp); break; }
            case SOLVE: { Mat tmp; cv::solve(a, b, tmp, flags); assignInto(m, tmp); break; }
        }
    }
    // dst.create(size, type) + write: keeps m's storage when it already has the right shape (so views are written through)
    static void assignInto(Mat& m, const Mat& v) {
        if (m.data && m.rows == v.rows && m.cols == v.cols && m.flags == v.flags && m.data != v.data) v.copyTo(m); else m = v;
    }
    void assignAddEx(Mat& m, int type) const {
        if (type != -1 && type != a.type()) shim_fail("typed MatExpr assignment not provided");
        if (b.data) {
            if (s.isZero() || !s.isReal()) {
                if (alpha == 1) {
                    if (beta == 1) cv::add(a, b, m);
                    else if (beta == -1) cv::subtract(a, b, m);
                    else cv::scaleAdd(b, beta, a, m);
                } else if (beta == 1) {
                    if (alpha == -1) cv::subtract(b, a, m);
                    else cv::scaleAdd(a, alpha, b, m);
                } else cv::addWeighted(a, alpha, b, beta, 0, m);
                if (!s.isReal()) shim_fail("non-real scalar");
            } else cv::addWeighted(a, alpha, b, beta, s[0], m);
        } else if (s.isReal() && std::fabs(alpha) != 1) {
            a.convertTo(m, type, alpha, s[0]);
        } else if (alpha == 1) cv::add(a, s, m);
        else if (alpha == -1) cv::subtract(s, a, m);
        else { Mat tmp; a.convertTo(tmp, a.type(), alpha); cv::add(tmp, s, m); }
    }
    template <typename TT, int M, int N> operator Matx<TT, M, N>() const { Mat m; assign(m); return (Matx<TT, M, N>)m; }
    Size size() const { Mat m = *this; return m.size(); }
    int type() const { return a.type(); }
    MatExpr t() const;
    MatExpr inv(int method = DECOMP_LU) const;
    MatExpr mul(const MatExpr& e, double scale = 1) const;
    Mat row(int r) const { Mat m = *this; return m.row(r); }
    Mat col(int cc) const { Mat m = *this; return m.col(cc); }
    double dot(const Mat& m) const { Mat x = *this; return x.dot(m); }
    template <typename TT> TT at(int r, int cc) const { Mat m = *this; return m.at<TT>(r, cc); }
};

inline Mat::Mat(const MatExpr& e) { e.assign(*this); }
inline Mat& Mat::operator=(const MatExpr& e) { e.assign(*this); return *this; }
inline void Mat::push_back(const MatExpr& e) { Mat m = e; push_back(m); }

namespace shim {
typedef MatExpr E;
inline E addex(const Mat& a, const Mat& b, double alpha, double beta, const Scalar& s = Scalar()) { return E::make(E::ADDEX, 0, a, b, Mat(), alpha, beta, s); }
inline Mat eval(const E& e) { Mat m; e.assign(m); return m; }
// MatOp::add / subtract (generic, this == e2.op after dispatch)
inline E add(const E& e1, const E& e2, double sign) {
    if (e1.op == E::GEMM || e2.op == E::GEMM) {
        // MatOp_GEMM::add/subtract fold a matrix addend into the gemm's C operand
        if (e1.op == E::GEMM && !e1.c.data && e2.op == E::IDENT) return E::make(E::GEMM, e1.flags, e1.a, e1.b, e2.a, e1.alpha, sign);
        if (e1.op == E::GEMM && !e1.c.data && e2.isT() ) return E::make(E::GEMM, e1.flags | GEMM_3_T, e1.a, e1.b, e2.a, e1.alpha, sign * e2.alpha);
        if (sign == 1 && e2.op == E::GEMM && !e2.c.data && e1.op == E::IDENT) return E::make(E::GEMM, e2.flags, e2.a, e2.b, e1.a, e2.alpha, 1);
    }
    double alpha = 1, beta = sign;
    Scalar s;
    Mat m1, m2;
    if (e1.isAddEx() && (!e1.b.data || e1.beta == 0)) { m1 = e1.a; alpha = e1.alpha; s = e1.s; } else m1 = eval(e1);
    if (e2.isAddEx() && (!e2.b.data || e2.beta == 0)) { m2 = e2.a; beta = sign * e2.alpha; s = sign == 1 ? s + e2.s : s - e2.s; } else m2 = eval(e2);
    return addex(m1, m2, alpha, beta, s);
}
inline E add(const E& e, const Scalar& s) {
    if (e.isAddEx()) { E r = e; r.s = r.s + s; return r; }
    return addex(eval(e), Mat(), 1, 0, s);
}
inline E rsub(const Scalar& s, const E& e) {   // s - e
    if (e.isAddEx()) { E r = e; r.alpha = -r.alpha; r.beta = -r.beta; r.s = s - r.s; return r; }
    return addex(eval(e), Mat(), -1, 0, s);
}
inline E scale(const E& e, double k) {
    if (e.isAddEx()) { E r = e; r.alpha *= k; r.beta *= k; r.s = r.s * k; return r; }
    if (e.op == E::BIN || e.op == E::T || e.op == E::GEMM) { E r = e; r.alpha *= k; if (e.op == E::GEMM) r.beta *= k; return r; }
    return addex(eval(e), Mat(), k, 0);
}
inline E rdiv(double k, const E& e) {          // k / e
    if (e.isBin('/') && !e.b.data) return addex(e.a, Mat(), k / e.alpha, 0);
    return E::make(E::BIN, '/', eval(e), Mat(), Mat(), k);
}
inline E mulel(const E& e1, const E& e2, double sc, char op) {   // MatOp::multiply / divide (element-wise)
    Mat m1, m2;
    if (e1.isScaled()) { m1 = e1.a; sc *= e1.alpha; } else m1 = eval(e1);
    if (e2.isScaled()) { m2 = e2.a; if (op == '*') sc *= e2.alpha; else sc /= e2.alpha; }
    else if (e2.isReciprocal()) { m2 = e2.a; sc /= e2.alpha; op = op == '*' ? '/' : '*'; }
    else m2 = eval(e2);
    return E::make(E::BIN, op, m1, m2, Mat(), sc);
}
inline E transpose(const E& e) {
    if (e.isT()) return e.alpha == 1 ? E(e.a) : addex(e.a, Mat(), e.alpha, 0);
    if (e.isScaled()) return E::make(E::T, 0, e.a, Mat(), Mat(), e.alpha);          // MatOp_AddEx::transpose
    if (e.op == E::GEMM) { E r = e; r.flags = (!(e.flags & GEMM_1_T) ? GEMM_2_T : 0) | (!(e.flags & GEMM_2_T) ? GEMM_1_T : 0) | (!(e.flags & GEMM_3_T) ? GEMM_3_T : 0); std::swap(r.a, r.b); return r; }
    return E::make(E::T, 0, eval(e), Mat(), Mat(), 1);
}
inline E matmul(const E& e1, const E& e2) {
    if (e1.op == E::INV && e2.op == E::IDENT) return E::make(E::SOLVE, e1.flags, e1.a, e2.a);   // MatOp_Invert::matmul
    double sc = 1;
    int flags = 0;
    Mat m1, m2;
    if (e1.isT()) { flags = GEMM_1_T; sc = e1.alpha; m1 = e1.a; }
    else if (e1.isScaled()) { sc = e1.alpha; m1 = e1.a; }
    else m1 = eval(e1);
    if (e2.isT()) { flags |= GEMM_2_T; sc *= e2.alpha; m2 = e2.a; }
    else if (e2.isScaled()) { sc *= e2.alpha; m2 = e2.a; }
    else m2 = eval(e2);
    return E::make(E::GEMM, flags, m1, m2, Mat(), sc);
}
inline E invert(const E& e, int method) { return E::make(E::INV, method, eval(e)); }
}  // namespace shim

inline MatExpr Mat::t() const { return shim::transpose(MatExpr(*this)); }
inline MatExpr Mat::inv(int method) const { return shim::invert(MatExpr(*this), method); }
inline MatExpr Mat::mul(const Mat& m, double scale) const { return shim::mulel(MatExpr(*this), MatExpr(m), scale, '*'); }
inline MatExpr Mat::mul(const MatExpr& m, double scale) const { return shim::mulel(MatExpr(*this), m, scale, '*'); }
// Mat::mul(InputArray(1)): the 1x1 double array is taken as a scalar by arithm_op; x * 1 is exact
inline MatExpr Mat::mul(double sc, double scale) const { return shim::addex(*this, Mat(), sc * scale, 0); }
inline MatExpr MatExpr::t() const { return shim::transpose(*this); }
inline MatExpr MatExpr::inv(int method) const { return shim::invert(*this, method); }
inline MatExpr MatExpr::mul(const MatExpr& e, double scale) const { return shim::mulel(*this, e, scale, '*'); }

inline MatExpr operator+(const MatExpr& a, const MatExpr& b) { return shim::add(a, b, 1); }
inline MatExpr operator-(const MatExpr& a, const MatExpr& b) { return shim::add(a, b, -1); }
inline MatExpr operator+(const MatExpr& a, const Scalar& s) { return shim::add(a, s); }
inline MatExpr operator+(const Scalar& s, const MatExpr& a) { return shim::add(a, s); }
inline MatExpr operator-(const MatExpr& a, const Scalar& s) { return shim::add(a, -s); }
inline MatExpr operator-(const Scalar& s, const MatExpr& a) { return shim::rsub(s, a); }
inline MatExpr operator+(const MatExpr& a, double s) { return shim::add(a, Scalar(s)); }
inline MatExpr operator+(double s, const MatExpr& a) { return shim::add(a, Scalar(s)); }
inline MatExpr operator-(const MatExpr& a, double s) { return shim::add(a, Scalar(-s)); }
inline MatExpr operator-(double s, const MatExpr& a) { return shim::rsub(Scalar(s), a); }
inline MatExpr operator-(const MatExpr& a) { return shim::rsub(Scalar(0), a); }
inline MatExpr operator*(const MatExpr& a, double s) { return shim::scale(a, s); }
inline MatExpr operator*(double s, const MatExpr& a) { return shim::scale(a, s); }
inline MatExpr operator/(const MatExpr& a, double s) { return shim::scale(a, 1 / s); }
inline MatExpr operator/(double s, const MatExpr& a) { return shim::rdiv(s, a); }
inline MatExpr operator*(const MatExpr& a, const MatExpr& b) { return shim::matmul(a, b); }
inline MatExpr operator/(const MatExpr& a, const MatExpr& b) { return shim::mulel(a, b, 1, '/'); }
// Mat-typed overloads so that two plain Mats (each one user conversion away from MatExpr) resolve without ambiguity
inline MatExpr operator+(const Mat& a, const Mat& b) { return shim::add(MatExpr(a), MatExpr(b), 1); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return shim::add(MatExpr(a), MatExpr(b), -1); }
inline MatExpr operator*(const Mat& a, const Mat& b) { return shim::matmul(MatExpr(a), MatExpr(b)); }
inline MatExpr operator/(const Mat& a, const Mat& b) { return shim::mulel(MatExpr(a), MatExpr(b), 1, '/'); }
inline MatExpr operator-(const Mat& a) { return shim::rsub(Scalar(0), MatExpr(a)); }
inline MatExpr operator*(const Mat& a, double s) { return shim::scale(MatExpr(a), s); }
inline MatExpr operator*(double s, const Mat& a) { return shim::scale(MatExpr(a), s); }
inline MatExpr operator/(const Mat& a, double s) { return shim::scale(MatExpr(a), 1 / s); }
inline MatExpr operator+(const Mat& a, double s) { return shim::add(MatExpr(a), Scalar(s)); }
inline MatExpr operator-(const Mat& a, double s) { return shim::add(MatExpr(a), Scalar(-s)); }
inline MatExpr operator+(const Mat& a, const Scalar& s) { return shim::add(MatExpr(a), s); }
inline MatExpr operator-(const Mat& a, const Scalar& s) { return shim::add(MatExpr(a), -s); }
inline MatExpr operator*(const Mat& a, const MatExpr& b) { return shim::matmul(MatExpr(a), b); }
inline MatExpr operator*(const MatExpr& a, const Mat& b) { return shim::matmul(a, MatExpr(b)); }
inline MatExpr operator+(const Mat& a, const MatExpr& b) { return shim::add(MatExpr(a), b, 1); }
inline MatExpr operator+(const MatExpr& a, const Mat& b) { return shim::add(a, MatExpr(b), 1); }
inline MatExpr operator-(const Mat& a, const MatExpr& b) { return shim::add(MatExpr(a), b, -1); }
inline MatExpr operator-(const MatExpr& a, const Mat& b) { return shim::add(a, MatExpr(b), -1); }
template <typename T, int M, int N> inline MatExpr operator*(const Mat& a, const Matx<T, M, N>& b) { return shim::matmul(MatExpr(a), MatExpr(Mat(b))); }
template <typename T, int M, int N> inline MatExpr operator*(const Matx<T, M, N>& a, const Mat& b) { return shim::matmul(MatExpr(Mat(a)), MatExpr(b)); }

// compound assignment: OpenCV declares both Mat& and const Mat& forms so that they apply to row()/col() temporaries
inline const Mat& operator+=(const Mat& a, const Mat& b) { cv::add(a, b, const_cast<Mat&>(a)); return a; }
inline const Mat& operator+=(const Mat& a, const Scalar& s) { cv::add(a, s, const_cast<Mat&>(a)); return a; }
inline const Mat& operator+=(const Mat& a, const MatExpr& e) { Mat m = e; cv::add(a, m, const_cast<Mat&>(a)); return a; }
inline const Mat& operator-=(const Mat& a, const Mat& b) { cv::subtract(a, b, const_cast<Mat&>(a)); return a; }
inline const Mat& operator-=(const Mat& a, const Scalar& s) { cv::subtract(a, s, const_cast<Mat&>(a)); return a; }
inline const Mat& operator-=(const Mat& a, const MatExpr& e) { Mat m = e; cv::subtract(a, m, const_cast<Mat&>(a)); return a; }
inline const Mat& operator*=(const Mat& a, double s) { a.convertTo(const_cast<Mat&>(a), -1, s); return a; }
inline const Mat& operator*=(const Mat& a, const Mat& b) { Mat m; cv::gemm(a, b, 1, Mat(), 0, m, 0); const_cast<Mat&>(a) = m; return a; }
inline const Mat& operator/=(const Mat& a, double s) { a.convertTo(const_cast<Mat&>(a), -1, 1. / s); return a; }
inline const Mat& operator/=(const Mat& a, const Mat& b) { cv::divide(a, b, const_cast<Mat&>(a)); return a; }

inline double norm(const Mat& a, int type = NORM_L2) {
    double s = 0;
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) { double v = a.getd(r, c); s = type == NORM_L2 ? s + v * v : (type == NORM_L1 ? s + std::fabs(v) : std::max(s, std::fabs(v))); }
    return type == NORM_L2 ? std::sqrt(s) : s;
}
inline double norm(const MatExpr& e, int type = NORM_L2) { Mat m = e; return norm(m, type); }
inline double norm(const Mat& a, const Mat& b, int type = NORM_L2) { Mat d = a - b; return norm(d, type); }
template <typename T> inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }
template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
inline Scalar mean(const Mat& a) { double s = 0; for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) s += a.getd(r, c); return Scalar(a.total() ? s / a.total() : 0); }
inline Scalar sum(const Mat& a) { double s = 0; for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) s += a.getd(r, c); return Scalar(s); }
inline void hconcat(const Mat& a0, const Mat& b0, Mat& dst) {
    Mat a = a0, b = b0, out(a0.rows, a0.cols + b0.cols, a0.flags);
    if (a.rows != b.rows || a.flags != b.flags) shim_fail("hconcat mismatch");
    size_t e = Mat::esz(a.flags);
    for (int r = 0; r < a.rows; r++) { memcpy(out.ptr(r), a.ptr(r), e * a.cols); memcpy(out.ptr(r) + e * a.cols, b.ptr(r), e * b.cols); }
    dst = out;
}
inline void vconcat(const Mat& a, const Mat& b, Mat& dst) { Mat out = a.clone(); out.push_back(b); dst = out; }
inline Mat abs(const Mat& a0) { Mat a = a0.clone(); for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) a.setd(r, c, std::fabs(a.getd(r, c))); return a; }
inline Mat abs(const MatExpr& e) { Mat m = e; return abs(m); }

// cv::sortIdx(src, dst, CV_SORT_EVERY_ROW + CV_SORT_ASCENDING) for CV_32F rows; stable (decision for ties, App. B)
inline void sortIdx(const Mat& src, Mat& dst, int /*flags*/) {
    Mat out(src.rows, src.cols, CV_32S);
    for (int r = 0; r < src.rows; r++) {
        std::vector<int> idx(src.cols);
        for (int c = 0; c < src.cols; c++) idx[c] = c;
        const float* p = src.ptr<float>(r);
        std::stable_sort(idx.begin(), idx.end(), [p](int a, int b) { return p[a] < p[b]; });
        for (int c = 0; c < src.cols; c++) out.at<int>(r, c) = idx[c];
    }
    dst = out;
}
inline void sort(const Mat& src, Mat& dst, int) {
    Mat out = src.clone();
    if (src.flags != CV_32F) shim_fail("sort type");
    for (int r = 0; r < out.rows; r++) std::sort(out.ptr<float>(r), out.ptr<float>(r) + out.cols);
    dst = out;
}

// ---------------------------------------------------------------------------------------------- imgproc
// cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR) for CV_8UC1 (imgproc/src/imgwarp.cpp):
// dsize = saturate_cast<int>(src.size * f) while the scale stays 1/f, so f = 0.5 is ALWAYS handed to the area-fast path (2x2 mean,
// +2 >> 2; clipped blocks at odd edges average what exists), even when the source size is odd;
// everything else is the 11-bit fixed-point bilinear (INTER_RESIZE_COEF_BITS = 11, VResizeLinear's >>4, >>16, +2 >>2).
inline void resize(const Mat& src0, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interp = INTER_LINEAR) {
    Mat src = src0;
    if (src.empty()) return;                              // (OpenCV asserts; the GUI paths that reach this are disabled)
    if (src.flags != CV_8UC1 || interp != INTER_LINEAR) shim_fail("resize: only CV_8UC1 INTER_LINEAR is provided");
    if (dsize.area() == 0) dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
    else { fx = (double)dsize.width / src.cols; fy = (double)dsize.height / src.rows; }
    const double sx = 1. / fx, sy = 1. / fy;
    const int dw = dsize.width, dh = dsize.height, sw = src.cols, sh = src.rows;
    Mat out(dh, dw, CV_8UC1);
    const int isx = saturate_cast<int>(sx), isy = saturate_cast<int>(sy);
    const bool area_fast = std::abs(sx - isx) < DBL_EPSILON && std::abs(sy - isy) < DBL_EPSILON;
    if (area_fast && isx == 2 && isy == 2) {
        // resizeAreaFast_: whole 2x2 blocks (+2 >> 2); destination columns / rows past ssize/2 (odd source sizes) average
        // the source pixels that exist, in float, with saturate_cast rounding
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                const int sx0 = 2 * x, sy0 = 2 * y;
                if (x < sw / 2 && y < sh / 2) {
                    const uchar* r0 = src.ptr(sy0) + sx0;
                    const uchar* r1 = src.ptr(sy0 + 1) + sx0;
                    out.at<uchar>(y, x) = (uchar)((r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2);
                    continue;
                }
                int sum = 0, count = 0;
                for (int yy = sy0; yy < sy0 + 2 && yy < sh; yy++)
                    for (int xx = sx0; xx < sx0 + 2 && xx < sw; xx++) { sum += src.at<uchar>(yy, xx); count++; }
                out.at<uchar>(y, x) = count ? saturate_cast<uchar>((float)sum / count) : 0;
            }
        dst = out;
        return;
    }
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> xa(2 * dw), ya(2 * dh);
    auto coeffs = [](int d, int n, double scale, std::vector<int>& ofs, std::vector<short>& ab, int dn) {
        (void)dn;
        for (int i = 0; i < d; i++) {
            float f = (float)((i + 0.5) * scale - 0.5);
            int s = cvFloor(f);
            f -= s;
            if (s < 0) { f = 0; s = 0; }
            if (s >= n - 1) { f = 0; s = n - 1; }
            ofs[i] = s;
            ab[2 * i] = saturate_cast<short>((1.f - f) * 2048);
            ab[2 * i + 1] = saturate_cast<short>(f * 2048);
        }
    };
    coeffs(dw, sw, sx, xofs, xa, dw);
    coeffs(dh, sh, sy, yofs, ya, dh);
    std::vector<int> row0(dw), row1(dw);
    for (int y = 0; y < dh; y++) {
        const uchar* s0 = src.ptr(yofs[y]);
        const uchar* s1 = src.ptr(std::min(yofs[y] + 1, sh - 1));
        for (int x = 0; x < dw; x++) {
            int x0 = xofs[x], x1 = std::min(x0 + 1, sw - 1);
            row0[x] = s0[x0] * xa[2 * x] + s0[x1] * xa[2 * x + 1];
            row1[x] = s1[x0] * xa[2 * x] + s1[x1] * xa[2 * x + 1];
        }
        const short b0 = ya[2 * y], b1 = ya[2 * y + 1];
        for (int x = 0; x < dw; x++)
            out.at<uchar>(y, x) = (uchar)((((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2);
    }
    dst = out;
}
// cv::Scharr(src 8U, dst, CV_16S, dx, dy, scale, delta = 0, BORDER_REFLECT_101): separable [-1 0 1] x [3 10 3], the scale is
// folded into the kernel; all integer, exact
inline void Scharr(const Mat& src0, Mat& dst, int ddepth, int dx, int dy, double scale = 1, double delta = 0, int border = BORDER_DEFAULT) {
    Mat src = src0;
    if (src.flags != CV_8UC1 || ddepth != CV_16S || dx + dy != 1 || delta != 0 || border != BORDER_REFLECT_101) shim_fail("Scharr: configuration not provided");
    const int w = src.cols, h = src.rows;
    Mat out(h, w, CV_16SC1);
    auto refl = [](int p, int n) { if (n == 1) return 0; if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; };
    for (int y = 0; y < h; y++) {
        const uchar* rm = src.ptr(refl(y - 1, h));
        const uchar* r0 = src.ptr(y);
        const uchar* rp = src.ptr(refl(y + 1, h));
        for (int x = 0; x < w; x++) {
            int xm = refl(x - 1, w), xp = refl(x + 1, w);
            int v;
            if (dx == 1) v = 3 * (rm[xp] - rm[xm]) + 10 * (r0[xp] - r0[xm]) + 3 * (rp[xp] - rp[xm]);
            else v = 3 * (rp[xm] - rm[xm]) + 10 * (rp[x] - rm[x]) + 3 * (rp[xp] - rm[xp]);
            out.at<short>(y, x) = saturate_cast<short>(v * scale);
        }
    }
    dst = out;
}

// cv::calcHist for the one configuration the reference uses (VISystem.cpp:1860: one CV_8UC1 image, no mask, 1-D, uniform
// bins over [lo, hi)): CV_32F counts, histSize x 1
inline void calcHist(const Mat* images, int nimages, const int* channels, const Mat& mask, Mat& hist, int dims,
                     const int* histSize, const float** ranges, bool uniform = true, bool accumulate = false) {
    if (nimages != 1 || dims != 1 || !uniform || accumulate || !mask.empty() || (channels && channels[0] != 0) || images[0].flags != CV_8UC1)
        shim_fail("calcHist: configuration not provided");
    const int n = histSize[0];
    const double lo = ranges[0][0], hi = ranges[0][1];
    Mat out(n, 1, CV_32FC1);
    const Mat img = images[0];
    for (int r = 0; r < img.rows; r++)
        for (int c = 0; c < img.cols; c++) {
            const double v = img.at<uchar>(r, c);
            if (v < lo || v >= hi) continue;
            const int bin = cvFloor((v - lo) * n / (hi - lo));
            if (bin >= 0 && bin < n) out.at<float>(bin, 0) += 1.f;
        }
    hist = out;
}

// ---------------------------------------------------------------------------------------------- features2d
class Feature2D {
public:
    virtual ~Feature2D() {}
    virtual void detect(const Mat&, std::vector<KeyPoint>&, const Mat& = Mat()) { shim_fail("Feature2D::detect: no detector in the shim (key points are supplied)"); }
    virtual void compute(const Mat&, std::vector<KeyPoint>&, Mat&) { shim_fail("Feature2D::compute"); }
    virtual void compute(const std::vector<Mat>&, std::vector<KeyPoint>&, Mat&) { shim_fail("Feature2D::compute"); }
    virtual void detectAndCompute(const Mat&, const Mat&, std::vector<KeyPoint>&, Mat&, bool = false) { shim_fail("Feature2D::detectAndCompute"); }
};
#define VSO_DETECTOR(name) struct name : Feature2D { enum { DESCRIPTOR_MLDB = 5, DESCRIPTOR_KAZE = 3 }; template <typename... A> static Ptr<name> create(A&&...) { return Ptr<name>(new name()); } };
VSO_DETECTOR(KAZE) VSO_DETECTOR(AKAZE) VSO_DETECTOR(ORB) VSO_DETECTOR(BRISK) VSO_DETECTOR(FastFeatureDetector)
namespace xfeatures2d { VSO_DETECTOR(SIFT) VSO_DETECTOR(SURF) }

class DescriptorMatcher {
public:
    virtual ~DescriptorMatcher() {}
    // exhaustive k-NN, rows sorted by (distance ascending, train index ascending); fewer than k train rows gives shorter
    // lists; imgIdx = 0 — cv::BFMatcher::knnMatch semantics (checked against cv2 4.13 in tests/test_oracle_cv2.py)
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
    virtual void match(const Mat& q, const Mat& t, std::vector<DMatch>& out) {
        std::vector<std::vector<DMatch>> kn;
        knnMatch(q, t, kn, 1);
        out.clear();
        for (auto& v : kn) if (!v.empty()) out.push_back(v[0]);
    }
protected:
    int norm_ = NORM_L2;
    float dist(const Mat& a, int i, const Mat& b, int j) const {
        if (norm_ == NORM_HAMMING) {
            const uchar* x = a.ptr(i);
            const uchar* y = b.ptr(j);
            int s = 0;
            for (int c = 0; c < a.cols; c++) s += __builtin_popcount((unsigned)(x[c] ^ y[c]));
            return (float)s;
        }
        const float* x = a.ptr<float>(i);
        const float* y = b.ptr<float>(j);
        double s = 0;
        for (int c = 0; c < a.cols; c++) { float d = x[c] - y[c]; s += (double)d * (double)d; }
        return std::sqrt((float)s);
    }
};
class BFMatcher : public DescriptorMatcher {
public:
    explicit BFMatcher(int norm = NORM_L2, bool = false) { norm_ = norm; }
    static Ptr<BFMatcher> create(int norm = NORM_L2, bool = false) { return Ptr<BFMatcher>(new BFMatcher(norm)); }
};
class FlannBasedMatcher : public DescriptorMatcher {   // exhaustive stand-in; the hot path never selects FLANN
public:
    static Ptr<FlannBasedMatcher> create() { return Ptr<FlannBasedMatcher>(new FlannBasedMatcher()); }
};
struct DrawMatchesFlags { enum { DEFAULT = 0, DRAW_OVER_OUTIMG = 1, NOT_DRAW_SINGLE_POINTS = 2, DRAW_RICH_KEYPOINTS = 4 }; };

// ---------------------------------------------------------------------------------------------- GUI / drawing: no-ops
// (the reference draws and shows a debug image inside every solver iteration, VISystem.cpp:1239-1268; excluded)
template <typename... A> inline void drawKeypoints(A&&...) {}
template <typename... A> inline void drawMatches(A&&...) {}
template <typename... A> inline void putText(A&&...) {}
template <typename... A> inline void imshow(A&&...) {}
template <typename... A> inline void namedWindow(A&&...) {}
template <typename... A> inline void line(A&&...) {}
template <typename... A> inline void circle(A&&...) {}
template <typename... A> inline void rectangle(A&&...) {}
template <typename... A> inline void arrowedLine(A&&...) {}
template <typename... A> inline void destroyAllWindows(A&&...) {}
template <typename... A> inline void resizeWindow(A&&...) {}
template <typename... A> inline void moveWindow(A&&...) {}
inline int waitKey(int = 0) { return -1; }
inline Mat imread(const std::string& name, int /*flags*/ = 0) {     // binary PGM (P5, maxval 255) only, as oracle/cvshim
    Mat m;
    FILE* f = fopen(name.c_str(), "rb");
    if (!f) return m;
    int w = 0, h = 0, maxv = 0;
    char magic[3] = {0, 0, 0};
    if (fscanf(f, "%2s %d %d %d", magic, &w, &h, &maxv) == 4 && magic[0] == 'P' && magic[1] == '5' && maxv == 255) {
        fgetc(f);
        m.create(h, w, CV_8U);
        if (fread(m.data, 1, (size_t)w * h, f) != (size_t)w * h) m.release();
    }
    fclose(f);
    return m;
}
VSO_STUB(imwrite)

// ---------------------------------------------------------------------------------------------- compile-only stubs
VSO_STUB(cvtColor) VSO_STUB(remap) VSO_STUB(undistort) VSO_STUB(undistortPoints) VSO_STUB(initUndistortRectifyMap)
VSO_STUB(getOptimalNewCameraMatrix) VSO_STUB(findEssentialMat) VSO_STUB(findFundamentalMat) VSO_STUB(recoverPose)
VSO_STUB(triangulatePoints) VSO_STUB(decomposeEssentialMat) VSO_STUB(Rodrigues) VSO_STUB(convertPointsFromHomogeneous)
VSO_STUB(convertPointsToHomogeneous) VSO_STUB(projectPoints) VSO_STUB(solvePnP) VSO_STUB(solvePnPRansac) VSO_STUB(findHomography)
VSO_STUB(perspectiveTransform) VSO_STUB(warpPerspective) VSO_STUB(warpAffine) VSO_STUB(GaussianBlur) VSO_STUB(Sobel)
VSO_STUB(minMaxLoc) VSO_STUB(normalize) VSO_STUB(threshold) VSO_STUB(applyColorMap) VSO_STUB(computeCorrespondEpilines)
VSO_STUB(correctMatches) VSO_STUB(SVDecomp) VSO_STUB(eigen) VSO_STUB(determinant) VSO_STUB(calcOpticalFlowPyrLK) VSO_STUB(goodFeaturesToTrack)
VSO_STUB(meanStdDev) VSO_STUB(countNonZero) VSO_STUB(merge) VSO_STUB(split)
namespace sfm { VSO_STUB(triangulatePoints) VSO_STUB(reconstruct) VSO_STUB(projectionFromKRt) VSO_STUB(KRtFromProjection) VSO_STUB(essentialFromRt) VSO_STUB(motionFromEssential) }

struct FileNode {
    template <typename T> void operator>>(T&) const { shim_fail("FileStorage is not provided"); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    template <typename T> operator T() const { shim_fail("FileStorage is not provided"); }
    bool empty() const { return true; }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) { shim_fail("FileStorage is not provided (intrinsics are passed in)"); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};

}  // namespace cv
#endif

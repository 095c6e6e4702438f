// compat.hpp — the type vocabulary the reference's Matcher / Camera / VISystem interfaces are written in
// (cv::Mat, cv::KeyPoint, cv::DMatch, cv::Point3d, cv::Matx33f, Sophus::SE3f), so the class mirrors keep the
// reference's signatures (include/Matcher.hpp:28-67, include/Camera.hpp:32-118, include/VISystem.hpp:40-148)
// in a tree that has neither OpenCV nor Eigen/Sophus.  Build with -DVISLAM_WITH_OPENCV to use the real
// OpenCV types instead of the cv-lite ones below (the SE3 type stays ours: it is 7 floats in Sophus'
// storage order and every group operation goes through the C ABI so host and device agree bit for bit).
//
// Nothing here computes on the hot path: these are containers.
#ifndef VISLAM_COMPAT_HPP_
#define VISLAM_COMPAT_HPP_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "vislam_b200.h"

#ifdef VISLAM_WITH_OPENCV
#include <opencv2/core.hpp>
#else

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_16SC1 CV_16S
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

namespace cv {

typedef unsigned char uchar;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point;

template <typename T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    template <typename U> explicit Point3_(const Point3_<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <typename U> Point3_& operator=(const Point3_<U>& o) { x = (T)o.x; y = (T)o.y; z = (T)o.z; return *this; }
    Point3_ operator/(int s) const { return Point3_((T)(x / s), (T)(y / s), (T)(z / s)); }
    Point3_ operator/(double s) const { return Point3_((T)(x / s), (T)(y / s), (T)(z / s)); }
    Point3_ operator+(const Point3_& o) const { return Point3_(x + o.x, y + o.y, z + o.z); }
    Point3_ operator-(const Point3_& o) const { return Point3_(x - o.x, y - o.y, z - o.z); }
    Point3_ operator-() const { return Point3_(-x, -y, -z); }
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s = 0.f) : pt(x, y), size(s), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38F) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

// Row-major 3x3 float matrix (cv::Matx33f): val[3*r + c].
struct Matx33f {
    float val[9];
    Matx33f() { for (int i = 0; i < 9; i++) val[i] = 0.f; }
    Matx33f(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
        val[0] = a; val[1] = b; val[2] = c; val[3] = d; val[4] = e; val[5] = f; val[6] = g; val[7] = h; val[8] = i;
    }
    static Matx33f eye() { return Matx33f(1, 0, 0, 0, 1, 0, 0, 0, 1); }
    float& operator()(int r, int c) { return val[3 * r + c]; }
    const float& operator()(int r, int c) const { return val[3 * r + c]; }
    Matx33f t() const { return Matx33f(val[0], val[3], val[6], val[1], val[4], val[7], val[2], val[5], val[8]); }
};
// cv::Matx products accumulate in the element type (float), k ascending from a zero accumulator
inline Matx33f operator*(const Matx33f& a, const Matx33f& b) {
    Matx33f m;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            float s = 0.f;
            for (int k = 0; k < 3; k++) s += a(r, k) * b(k, c);
            m(r, c) = s;
        }
    return m;
}
inline Point3f operator*(const Matx33f& a, const Point3f& p) {
    const float v[3] = {p.x, p.y, p.z};
    float o[3];
    for (int r = 0; r < 3; r++) {
        float s = 0.f;
        for (int k = 0; k < 3; k++) s += a(r, k) * v[k];
        o[r] = s;
    }
    return Point3f(o[0], o[1], o[2]);
}

// Reference-counted dense 2-D array, single channel (the subset of cv::Mat the path touches).
class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;   // bytes per row

    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8U) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), type_(CV_8U) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(nullptr), step(0), type_(CV_8U) { create(s.height, s.width, type); }
    // non-owning view over caller memory (cv::Mat(rows, cols, type, data, step))
    Mat(int r, int c, int type, void* ext, size_t step_bytes = 0)
        : rows(r), cols(c), data(static_cast<uchar*>(ext)),
          step(step_bytes ? step_bytes : (size_t)c * elem_size(type)), type_(type) {}

    static size_t elem_size(int type) {
        switch (type) {
            case CV_8U: case CV_8S: return 1;
            case CV_16U: case CV_16S: return 2;
            case CV_32S: case CV_32F: return 4;
            case CV_64F: return 8;
            default: throw std::invalid_argument("cv-lite Mat: unsupported type");
        }
    }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && own_ && isContinuous()) return;
        const size_t bytes = (size_t)r * c * elem_size(type);
        own_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>(bytes ? bytes : 1, 0));
        rows = r; cols = c; type_ = type;
        data = own_->data();
        step = (size_t)c * elem_size(type);
    }
    static Mat zeros(int r, int c, int type) { Mat m; m.create(r, c, type); std::memset(m.data, 0, m.total() * m.elemSize()); return m; }
    static Mat ones(int r, int c, int type) {
        Mat m = zeros(r, c, type);
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) m.set_scalar(i, j, 1.0);
        return m;
    }
    void release() { own_.reset(); rows = cols = 0; data = nullptr; step = 0; }
    int type() const { return type_; }
    int depth() const { return type_; }
    int channels() const { return 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t elemSize() const { return elem_size(type_); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        Mat out;
        out.create(rows, cols, type_);
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) std::memcpy(out.data + (size_t)r * out.step, data + (size_t)r * step, rb);
        dst = out;
    }
    Mat row(int r) const { Mat m(1, cols, type_, data + (size_t)r * step, step); m.own_ = own_; return m; }

private:
    void set_scalar(int r, int c, double v) {
        switch (type_) {
            case CV_8U: at<uchar>(r, c) = (uchar)v; break;
            case CV_16S: at<int16_t>(r, c) = (int16_t)v; break;
            case CV_32S: at<int32_t>(r, c) = (int32_t)v; break;
            case CV_32F: at<float>(r, c) = (float)v; break;
            case CV_64F: at<double>(r, c) = v; break;
            default: break;
        }
    }
    int type_;
    std::shared_ptr<std::vector<uchar>> own_;
};

}  // namespace cv
#endif  // VISLAM_WITH_OPENCV

namespace vi {

// Sophus::SE3f as the reference uses it (Options.hpp:53 `typedef Sophus::SE3f SE3`): unit quaternion +
// translation, float.  Storage {qx, qy, qz, qw, tx, ty, tz} = Sophus' own order.  operator*, exp and matrix
// are evaluated by libvislam_b200 (vsb_se3_mul / vsb_se3_exp / vsb_se3_matrix: se3.hpp:285-321, 723-742,
// 253-268) — the same code the device runs.
class SE3 {
public:
    struct Quat {
        float qx, qy, qz, qw;
        float x() const { return qx; }
        float y() const { return qy; }
        float z() const { return qz; }
        float w() const { return qw; }
    };
    struct Vec3 {
        float v[3];
        float operator()(int i) const { return v[i]; }
        float& operator()(int i) { return v[i]; }
        float x() const { return v[0]; }
        float y() const { return v[1]; }
        float z() const { return v[2]; }
    };
    static const int DoF = 6;

    SE3() { p_[0] = p_[1] = p_[2] = 0.f; p_[3] = 1.f; p_[4] = p_[5] = p_[6] = 0.f; }
    explicit SE3(const float pose7[7]) { for (int i = 0; i < 7; i++) p_[i] = pose7[i]; }
    // SE3(quaternion(w, x, y, z), translation) — the argument order of Eigen::Quaternion's constructor
    SE3(float qw, float qx, float qy, float qz, float tx, float ty, float tz) {
        p_[0] = qx; p_[1] = qy; p_[2] = qz; p_[3] = qw; p_[4] = tx; p_[5] = ty; p_[6] = tz;
    }
    Quat unit_quaternion() const { Quat q = {p_[0], p_[1], p_[2], p_[3]}; return q; }
    Vec3 translation() const { Vec3 t = {{p_[4], p_[5], p_[6]}}; return t; }
    const float* data() const { return p_; }
    float* data() { return p_; }
    SE3 operator*(const SE3& o) const { SE3 r; vsb_se3_mul(p_, o.p_, r.p_); return r; }
    SE3& operator*=(const SE3& o) { float t[7]; vsb_se3_mul(p_, o.p_, t); for (int i = 0; i < 7; i++) p_[i] = t[i]; return *this; }
    static SE3 exp(const float delta[6]) { SE3 r; vsb_se3_exp(delta, r.p_); return r; }
    // row-major 4x4, as Eigen's matrix() read row by row
    void matrix(float m[16]) const { vsb_se3_matrix(p_, m); }

private:
    float p_[7];
};

}  // namespace vi

#endif

// oracle/refshim/opencv2/core/eigen.hpp — cv2eigen / eigen2cv element copies (TEST INFRASTRUCTURE ONLY)
#ifndef VSO_REFSHIM_CV_EIGEN_HPP
#define VSO_REFSHIM_CV_EIGEN_HPP
#include "../core.hpp"
#include "Eigen/Core"
namespace cv {
template <typename T, int R, int C> inline void cv2eigen(const Matx<T, R, C>& src, Eigen::Matrix<T, R, C>& dst) {
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) dst(r, c) = src(r, c);
}
template <typename T, int R, int C> inline void cv2eigen(const Mat& src, Eigen::Matrix<T, R, C>& dst) {
    if (src.rows != R || src.cols != C) shim_fail("cv2eigen size mismatch");
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) dst(r, c) = (T)src.getd(r, c);
}
template <typename T, int R, int C> inline void eigen2cv(const Eigen::Matrix<T, R, C>& src, Mat& dst) {
    Mat out(R, C, Mat::depth_of((T*)0));
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) out.at<T>(r, c) = src(r, c);
    dst = out;                                       // eigen2cv copies into a newly created matrix
}
}  // namespace cv
#endif

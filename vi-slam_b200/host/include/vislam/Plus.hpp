// Plus.hpp — angle / pose conversion helpers with the reference's names and conventions (include/Plus.hpp:11-36,
// src/Plus.cpp): ZYX Euler angles packed as Point3d(roll, pitch, yaw), quaternion {w, x, y, z} in double, rotation
// matrices in float (Matx33f / 4x4 CV_32F Mat).  Host-side only; the GN initial pose (VISystem.cpp:1135-1168) is formed
// from rotationMatrix2RPY / RPY2rotationMatrix.
#ifndef VISLAM_PLUS_HPP_
#define VISLAM_PLUS_HPP_
#include "compat.hpp"

struct Quaterniond {                       // include/Plus.hpp:10-16
    double w, x, y, z;
    Quaterniond() : w(1), x(0), y(0), z(0) {}
};

Quaterniond toQuaternion(double roll, double pitch, double yaw);
cv::Point3d toRPY(const Quaterniond& q);
cv::Point3d toRPY360(cv::Point3d angles);
double computeDiff(double gt_angle, double gt_est);
cv::Mat point2MatPlusOne(cv::Point3d point);
cv::Mat point2Mat(cv::Point3d point);
cv::Point3d Mat2point(cv::Mat position);
cv::Point3d transformationMatrix2RPY(cv::Mat transformationMatrix);
cv::Point3d transformationMatrix2position(cv::Mat transformationMatrix);
cv::Point3d rotationMatrix2RPY(cv::Matx33f rotationMatrix);
cv::Matx33f RPY2rotationMatrix(cv::Point3d rpy);
cv::Mat RPYAndPosition2transformationMatrix(cv::Point3d rpy, cv::Point3d position);
cv::Mat transformationMatrix2rotationMatrix(cv::Mat transformationMatrix);
cv::Mat RPYWorld2ResidualAngImu(cv::Point3d rpy);

#endif

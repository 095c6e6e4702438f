// Stand-in for the OpenCV header of this name when OpenCV is not installed: the cv:: types the frame-tracking classes use
// (Mat, KeyPoint, DMatch, Point3d, Matx33f, ...) come from vislam/compat.hpp.  With a real OpenCV on the include path, drop
// this directory from it and define VISLAM_WITH_OPENCV.
#include "vislam/compat.hpp"

// GroundTruth.hpp — forwarding header with the reference's file name (include/GroundTruth.hpp); the class lives in vislam/DataReader.hpp.
#ifndef VISLAM_REFNAMES_GroundTruth_HPP_
#define VISLAM_REFNAMES_GroundTruth_HPP_
#include "vislam/DataReader.hpp"
using namespace cv;
using namespace std;
#endif

// ImageReader.hpp — forwarding header with the reference's file name (include/ImageReader.hpp); the class lives in vislam/DataReader.hpp.
#ifndef VISLAM_REFNAMES_ImageReader_HPP_
#define VISLAM_REFNAMES_ImageReader_HPP_
#include "vislam/DataReader.hpp"
using namespace cv;
using namespace std;
#endif

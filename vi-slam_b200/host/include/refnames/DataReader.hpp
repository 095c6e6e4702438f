// DataReader.hpp — forwarding header with the reference's file name (include/DataReader.hpp): a caller written against the reference
// includes "DataReader.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_DataReader_HPP_
#define VISLAM_REFNAMES_DataReader_HPP_
#include "vislam/DataReader.hpp"
using namespace cv;
using namespace std;
#endif

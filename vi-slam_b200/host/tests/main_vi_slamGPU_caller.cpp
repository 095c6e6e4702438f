// main_vi_slamGPU_caller.cpp — boundary proof by compilation: the calls the reference's GPU executable makes on this path
// (/root/reference/src/main_vi_slamGPU.cpp:58-65 set-up, :118-152 loop), written against the header NAMES and the class /
// function names of the reference and compiled against host/include/refnames + libvislam_host.so.  What is left out are the
// ROS and visualiser lines (:12-18, :70-75, :133-136) and cv::CommandLineParser (plain argv here).  Every statement that
// touches DataReader, VISystemGPU or the Plus helpers has the same form as upstream, so if this file compiles a maintainer
// of the reference can swap the classes in without touching the call sites.
//   usage: ref_main_gpu <imagesPath> <imuFile> <gtFile> <calibrationFile> <outputFile> [first index = 1]
#include <fstream>
#include <iostream>

#include "DataReader.hpp"
#include "VISystemGPU.hpp"
#include "opencv2/cudafeatures2d.hpp"
#include "opencv2/xfeatures2d/cuda.hpp"
#include "opencv2/core.hpp"
#include "opencv2/highgui.hpp"
#include "opencv2/calib3d.hpp"
#include <ctime>

using namespace cv;
using namespace std;
using namespace vi;

int main(int argc, char** argv) {
    if (argc < 6) {
        cout << "usage: " << argv[0] << " imagesPath imuFile gtFile calibrationFile outputFile [first index]" << endl;
        return 2;
    }
    string imagesPath = argv[1];
    string imuFile = argv[2];
    string gtFile = argv[3];
    string calibrationFile = argv[4];
    string outputFile = argv[5];
    char separator = ',';

    DataReader Data(imagesPath, imuFile, gtFile, separator);                              // main_vi_slamGPU.cpp:58

    int j = argc > 6 ? atoi(argv[6]) : 1;                                                 // upstream: 210, a hard-coded start
    Data.UpdateDataReader(j - 1, j);                                                      // :62
    VISystemGPU visystem(argc, argv);                                                     // :63
    visystem.InitializeSystemGPU(calibrationFile, Data.gtPosition[0], Data.gtLinearVelocity[0], Data.gtRPY[0], Data.image1);   // :65
    cout << "Initializate System" << endl;
    Quaterniond qinit = toQuaternion(Data.gtRPY[0].x, Data.gtRPY[0].y, Data.gtRPY[0].z);  // :67
    (void)qinit;

    // ground truth expressed for the camera (:101-104)
    Quaterniond qOrientationCamGT;
    Point3d RPYOrientationCamGT;
    Point3d positionCamGT;

    std::ofstream outputFilecsv;
    outputFilecsv.open(outputFile.c_str(), std::ofstream::out | std::ofstream::trunc);    // upstream: a hard-coded path (:116)
    while (j < Data.indexLastData) {                                                      // :118
        Data.UpdateDataReader(j, j + 1);                                                  // :121
        j = j + 1;
        visystem.AddFrameGPU(Data.image2, Data.imuAngularVelocity, Data.imuAcceleration); // :123

        positionCamGT = Data.gtPosition.back() + visystem.imu2camTranslation;             // :125
        RPYOrientationCamGT = rotationMatrix2RPY(visystem.imu2camRotation * RPY2rotationMatrix(toRPY(Data.gtQuaternion.back())));   // :126
        qOrientationCamGT = toQuaternion(RPYOrientationCamGT.x, RPYOrientationCamGT.y, RPYOrientationCamGT.z);   // :127

        cout << " Current time = " << Data.currentTimeMs << " ms " << endl;               // :138

        outputFilecsv << visystem.positionCam.x << ","                                    // :140-152
                      << visystem.positionCam.y << ","
                      << visystem.positionCam.z << ","
                      << visystem.qOrientationCam.x << ","
                      << visystem.qOrientationCam.y << ","
                      << visystem.qOrientationCam.z << ","
                      << visystem.qOrientationCam.w << ","
                      << positionCamGT.x << ","
                      << positionCamGT.y << ","
                      << positionCamGT.z << ","
                      << qOrientationCamGT.x << ","
                      << qOrientationCamGT.y << ","
                      << qOrientationCamGT.z << ","
                      << qOrientationCamGT.w
                      << endl;
    }
    return 0;
}

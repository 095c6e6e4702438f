// DataReader.hpp — dataset ingestion with the reference's class interfaces:
//   ImageReader  (include/ImageReader.hpp:13-35, src/ImageReader.cpp)  — directory of <timestamp_ns>.<ext> images
//   GroundTruth  (include/GroundTruth.hpp:5-27,  src/GroundTruth.cpp)  — numeric CSV (EuRoC imu0/data.csv and
//                                                                        state_groundtruth_estimate0/data.csv)
//   DataReader   (include/DataReader.hpp:6-56,   src/DataReader.cpp)   — camera / IMU / ground-truth synchronisation
//   TrajectoryWriter — the per-frame CSV row of src/main_vi_slam.cpp:183-210
// plus cv::imread's job for this path (8-bit PGM and non-interlaced 8-bit PNG, decoded to grayscale).
// Host-side only, no device work; on-disk formats: SURVEY.md Appendix C.
#ifndef VISLAM_DATAREADER_HPP_
#define VISLAM_DATAREADER_HPP_
#include <fstream>
#include <string>
#include <vector>

#include "Plus.hpp"
#include "compat.hpp"

namespace vi {
// cv::imread(name, CV_LOAD_IMAGE_GRAYSCALE): empty Mat when the file is missing or not a supported format
cv::Mat imread_gray(const std::string& file);
}

class ImageReader {
public:
    ImageReader();
    explicit ImageReader(std::string directory);
    void setPath(std::string directory);
    std::string getImageName(int index);
    long int getImageTime(int index);
    void searchImages();
    cv::Mat getImage(int index);
    size_t splitStrings(const std::string& txt, std::vector<std::string>& strs, char separator);
    size_t getSize();
    void computeTimeStep();
    double TimeStep;

private:
    std::string path;
    std::vector<std::string> file_names;
};

class GroundTruth {
public:
    GroundTruth();
    GroundTruth(std::string file, char separator);
    void setFileProperties(std::string file, char separator);
    void getDataFromFile();
    int getLines();
    int getRows();
    int getCols();
    size_t splitStrings(const std::string& txt, std::vector<std::string>& strs, char separator);
    std::string getFileName();
    char getCharSeparator();
    double getGroundTruthData(int line, int colData);
    void computeTimeStep();
    double TimeStep;

private:
    char charSeparator;
    std::string fileName;
    cv::Mat data;
    int cols;
    int rows;
};

class DataReader {
public:
    DataReader();
    DataReader(std::string image_path, std::string imu_path, std::string gt_path, char separator);
    void setProperties(std::string image_path, std::string imu_path, std::string gt_path, char separator);
    void UpdateDataReader(int index, int index2);
    void UpdateImu(int index, int n_measures);

    std::vector<cv::Point3d> imuAngularVelocity;
    std::vector<cv::Point3d> imuAcceleration;
    cv::Mat image1;
    cv::Mat image2;
    std::vector<cv::Point3d> gtPosition, gtLinearVelocity;
    std::vector<Quaterniond> gtQuaternion;
    std::vector<cv::Point3d> gtRPY;
    std::vector<cv::Point3d> accBias;
    cv::Point3d angBias;
    double currentTimeMs;
    double initialTime;
    double lastTime;
    int imageIndex0;
    int gtIndex0;
    int imuIndex0;
    double timeStepImu;
    double timeStepCamara;
    double timeStepGt;
    int indexLastData;

private:
    ImageReader imageReader;
    GroundTruth gtReader;
    GroundTruth imuReader;
};

namespace vi {
// One row per tracked frame, 27 comma-separated columns (src/main_vi_slam.cpp:183-210):
// t_s, p_est(3), v_est(3), a_est(3), q_est(x,y,z,w), p_gt(3), v_gt(3), q_gt(x,y,z,w), w_filtered(3)
class TrajectoryWriter {
public:
    explicit TrajectoryWriter(const std::string& file);
    bool ok() const { return out_.is_open(); }
    void write(double time_s, const cv::Point3d& p, const cv::Point3d& v, const cv::Point3d& a, const Quaterniond& q,
               const cv::Point3d& p_gt, const cv::Point3d& v_gt, const Quaterniond& q_gt, const cv::Point3d& w_filtered);

private:
    std::ofstream out_;
};
}  // namespace vi

#endif

// DataReader.cpp — see DataReader.hpp.  The index arithmetic follows the reference step by step (file:line cited per
// block) because the synchronisation result IS the behaviour: which IMU rows and ground-truth rows belong to a frame pair.
#include "vislam/DataReader.hpp"

#include <dirent.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <stdexcept>

// ------------------------------------------------------------------------------------------------ ImageReader
ImageReader::ImageReader() : TimeStep(0.0) { setPath(""); }

ImageReader::ImageReader(std::string directory) : TimeStep(0.0) {
    setPath(directory);
    searchImages();
    computeTimeStep();
}

void ImageReader::setPath(std::string directory) { path = directory; }

size_t ImageReader::splitStrings(const std::string& txt, std::vector<std::string>& strs, char separator) {
    strs.clear();
    size_t from = 0;
    for (size_t pos = txt.find(separator); pos != std::string::npos; pos = txt.find(separator, from)) {
        strs.push_back(txt.substr(from, pos - from));
        from = pos + 1;
    }
    strs.push_back(txt.substr(from));
    return strs.size();
}

// "<dir>/<timestamp>.<ext>" -> "<timestamp>" (ImageReader.cpp:23-40): last path component, text before its first '.'
std::string ImageReader::getImageName(int index) {
    std::vector<std::string> parts;
    splitStrings(file_names[index], parts, '/');
    std::string name = parts.back();
    if (name.find(".") != 0) {            // the reference's `if (imageName.find("."))`: skipped only for a leading dot
        splitStrings(name, parts, '.');
        name = parts[0];
    }
    return name;
}

long int ImageReader::getImageTime(int index) { return std::atol(getImageName(index).c_str()); }   // :42-48

// Every directory entry, sorted; the first two sorted entries are dropped unseen — upstream that removes "." and ".."
// (ImageReader.cpp:50-78) — and fewer than 15 remaining files is a fatal configuration error there (exit(0)).
void ImageReader::searchImages() {
    file_names.clear();
    DIR* dir = opendir(path.c_str());
    if (!dir) throw std::runtime_error("ImageReader: could not open directory of images: " + path);
    while (struct dirent* ent = readdir(dir)) file_names.push_back(path + std::string(ent->d_name));
    closedir(dir);
    std::sort(file_names.begin(), file_names.end());
    file_names.erase(file_names.begin(), file_names.begin() + std::min<size_t>(2, file_names.size()));
    if (file_names.size() < 15)
        throw std::runtime_error("ImageReader: insufficient number of images found (need >= 15): " + path);
}

cv::Mat ImageReader::getImage(int index) { return vi::imread_gray(file_names[index]); }   // :80-82

size_t ImageReader::getSize() { return file_names.size(); }

void ImageReader::computeTimeStep() {                                                      // :103-108
    const double t0 = (double)getImageTime(0), t1 = (double)getImageTime(1);
    TimeStep = t1 - t0;
}

// ------------------------------------------------------------------------------------------------ GroundTruth
GroundTruth::GroundTruth() : TimeStep(0), charSeparator(','), cols(0), rows(0) { data = cv::Mat::zeros(1, 1, CV_64F); }

GroundTruth::GroundTruth(std::string file, char separator) : TimeStep(0), cols(0), rows(0) {
    setFileProperties(file, separator);
    getDataFromFile();
    computeTimeStep();
}

void GroundTruth::setFileProperties(std::string file, char separator) {
    fileName = file;
    charSeparator = separator;
    data = cv::Mat::zeros(1, 1, CV_64F);
}

int GroundTruth::getLines() {                                                              // GroundTruth.cpp:29-45
    std::ifstream in(fileName.c_str());
    if (!in) throw std::runtime_error("GroundTruth: unable to open " + fileName);
    int n = 0;
    std::string line;
    while (std::getline(in, line)) n++;
    return n;
}

int GroundTruth::getRows() { return rows; }
int GroundTruth::getCols() { return cols; }

size_t GroundTruth::splitStrings(const std::string& txt, std::vector<std::string>& strs, char separator) {
    strs.clear();
    size_t from = 0;
    for (size_t pos = txt.find(separator); pos != std::string::npos; pos = txt.find(separator, from)) {
        strs.push_back(txt.substr(from, pos - from));
        from = pos + 1;
    }
    strs.push_back(txt.substr(from));
    return strs.size();
}

// GroundTruth.cpp:57-106.  Leading lines containing '#' are skipped; the first data line fixes the column count; the
// matrix gets one row per LINE OF THE FILE (comment lines included: `rows = getLines()`), so rows past the data stay
// zero — DataReader relies on that (`getRows() - 2` is the last data row of a file with one header line).  Fields are
// split on ',' whatever separator was configured (upstream passes the literal) and converted with atof.  A line with
// fewer fields than the first one leaves the rest 0 (upstream reads past the end of its vector there).
void GroundTruth::getDataFromFile() {
    std::ifstream in(fileName.c_str());
    if (!in) throw std::runtime_error("GroundTruth: unable to open " + fileName);
    std::string line;
    std::vector<std::string> fields;
    do {
        if (!std::getline(in, line)) { line.clear(); break; }
    } while (line.find("#") != std::string::npos);
    splitStrings(line, fields, ',');
    rows = getLines();
    cols = (int)fields.size();
    if (line.empty() || cols == 0 || rows == 0) throw std::runtime_error("GroundTruth: empty file " + fileName);
    data = cv::Mat::zeros(rows, cols, CV_64F);
    int r = 0;
    for (;;) {
        for (int c = 0; c < cols && c < (int)fields.size(); c++) data.at<double>(r, c) = std::atof(fields[c].c_str());
        if (!std::getline(in, line) || ++r >= rows) break;
        splitStrings(line, fields, ',');
    }
}

std::string GroundTruth::getFileName() { return fileName; }
char GroundTruth::getCharSeparator() { return charSeparator; }
double GroundTruth::getGroundTruthData(int line, int colData) { return data.at<double>(line, colData); }

void GroundTruth::computeTimeStep() {                                                      // :143-148
    TimeStep = data.at<double>(1, 0) - data.at<double>(0, 0);
}

// ------------------------------------------------------------------------------------------------ DataReader
DataReader::DataReader()
    : currentTimeMs(0), initialTime(0), lastTime(0), imageIndex0(0), gtIndex0(0), imuIndex0(0), timeStepImu(0),
      timeStepCamara(0), timeStepGt(0), indexLastData(0) {}

DataReader::DataReader(std::string image_path, std::string imu_path, std::string gt_path, char separator) {
    setProperties(image_path, imu_path, gt_path, separator);
}

void DataReader::setProperties(std::string image_path, std::string imu_path, std::string gt_path, char separator) {
    imageReader.setPath(image_path);
    imageReader.searchImages();
    imageReader.computeTimeStep();
    imuReader.setFileProperties(imu_path, separator);
    imuReader.getDataFromFile();
    imuReader.computeTimeStep();
    gtReader.setFileProperties(gt_path, separator);
    gtReader.getDataFromFile();
    gtReader.computeTimeStep();

    // DataReader.cpp:26-65: first image that is not older than the first IMU and the first GT sample, then the first
    // GT / IMU sample that is not older than that image
    imageIndex0 = imuIndex0 = gtIndex0 = 0;
    double t_img = (double)imageReader.getImageTime(imageIndex0);
    double t_imu = imuReader.getGroundTruthData(imuIndex0, 0);
    double t_gt = gtReader.getGroundTruthData(gtIndex0, 0);
    while (t_img - t_gt < 0.0 || t_img - t_imu < 0.0) t_img = (double)imageReader.getImageTime(++imageIndex0);
    while (t_gt - t_img < 0.0) t_gt = gtReader.getGroundTruthData(++gtIndex0, 0);
    while (t_imu - t_img < 0.0) t_imu = imuReader.getGroundTruthData(++imuIndex0, 0);

    timeStepCamara = imageReader.TimeStep;
    timeStepImu = imuReader.TimeStep;
    timeStepGt = gtReader.TimeStep;

    // :99-103: last image index that still has ground truth
    const double t_last_gt = gtReader.getGroundTruthData(gtReader.getRows() - 2, 0);
    initialTime = t_img;
    indexLastData = static_cast<int>(std::floor((t_last_gt - initialTime) / timeStepCamara) - imageIndex0);
    lastTime = (indexLastData * timeStepCamara) / 1000000;
}

void DataReader::UpdateDataReader(int index, int index2) {                                 // DataReader.cpp:110-240
    image1 = imageReader.getImage(imageIndex0 + index);
    image2 = imageReader.getImage(imageIndex0 + index2);
    // first guess from the nominal rates, then walk forward to the sample at / after the first image
    int i_gt = static_cast<int>(std::floor((imageReader.TimeStep / gtReader.TimeStep) * index)) + gtIndex0;
    int i_imu = static_cast<int>(std::floor((imageReader.TimeStep / imuReader.TimeStep) * index)) + imuIndex0;
    const double t_img1 = (double)imageReader.getImageTime(imageIndex0 + index);
    const double t_img2 = static_cast<double>(imageReader.getImageTime(imageIndex0 + index2));
    while (gtReader.getGroundTruthData(i_gt, 0) - t_img1 < 0.0) i_gt++;          // GT sample not older than image 1
    while (imuReader.getGroundTruthData(i_imu, 0) - t_img1 <= 0.0) i_imu++;      // IMU sample strictly after image 1

    imuAngularVelocity.clear();
    imuAcceleration.clear();
    gtLinearVelocity.clear();
    gtPosition.clear();
    gtQuaternion.clear();
    gtRPY.clear();
    accBias.clear();
    // every IMU sample up to and including image 2 (:171-187)
    for (; imuReader.getGroundTruthData(i_imu, 0) <= t_img2; i_imu++) {
        imuAngularVelocity.push_back(cv::Point3d(imuReader.getGroundTruthData(i_imu, 1), imuReader.getGroundTruthData(i_imu, 2),
                                                 imuReader.getGroundTruthData(i_imu, 3)));
        imuAcceleration.push_back(cv::Point3d(imuReader.getGroundTruthData(i_imu, 4), imuReader.getGroundTruthData(i_imu, 5),
                                              imuReader.getGroundTruthData(i_imu, 6)));
    }
    // gyro bias of the GT row at image 1 (:189-192), then every LATER GT row up to image 2 (:195-233)
    angBias = cv::Point3d(gtReader.getGroundTruthData(i_gt, 11), gtReader.getGroundTruthData(i_gt, 12),
                          gtReader.getGroundTruthData(i_gt, 13));
    for (i_gt++; gtReader.getGroundTruthData(i_gt, 0) <= t_img2; i_gt++) {
        Quaterniond q;
        q.w = gtReader.getGroundTruthData(i_gt, 4);
        q.x = gtReader.getGroundTruthData(i_gt, 5);
        q.y = gtReader.getGroundTruthData(i_gt, 6);
        q.z = gtReader.getGroundTruthData(i_gt, 7);
        gtPosition.push_back(cv::Point3d(gtReader.getGroundTruthData(i_gt, 1), gtReader.getGroundTruthData(i_gt, 2),
                                         gtReader.getGroundTruthData(i_gt, 3)));
        gtQuaternion.push_back(q);
        gtRPY.push_back(toRPY(q));
        gtLinearVelocity.push_back(cv::Point3d(gtReader.getGroundTruthData(i_gt, 8), gtReader.getGroundTruthData(i_gt, 9),
                                               gtReader.getGroundTruthData(i_gt, 10)));
        accBias.push_back(cv::Point3d(gtReader.getGroundTruthData(i_gt, 14), gtReader.getGroundTruthData(i_gt, 15),
                                      gtReader.getGroundTruthData(i_gt, 16)));
    }
    currentTimeMs = (imageReader.getImageTime(imageIndex0 + index2) - initialTime) / 1000000.0;   // :238
}

void DataReader::UpdateImu(int index, int n_measures) {                                    // :243-266
    imuAngularVelocity.clear();
    imuAcceleration.clear();
    for (int i = index; i < index + n_measures; i++) {
        imuAngularVelocity.push_back(cv::Point3d(imuReader.getGroundTruthData(i, 1), imuReader.getGroundTruthData(i, 2),
                                                 imuReader.getGroundTruthData(i, 3)));
        imuAcceleration.push_back(cv::Point3d(imuReader.getGroundTruthData(i, 4), imuReader.getGroundTruthData(i, 5),
                                              imuReader.getGroundTruthData(i, 6)));
    }
}

// ------------------------------------------------------------------------------------------------ TrajectoryWriter
vi::TrajectoryWriter::TrajectoryWriter(const std::string& file) : out_(file.c_str()) {}

void vi::TrajectoryWriter::write(double time_s, const cv::Point3d& p, const cv::Point3d& v, const cv::Point3d& a,
                                 const Quaterniond& q, const cv::Point3d& p_gt, const cv::Point3d& v_gt,
                                 const Quaterniond& q_gt, const cv::Point3d& w) {
    out_ << time_s << "," << p.x << "," << p.y << "," << p.z << "," << v.x << "," << v.y << "," << v.z << "," << a.x << ","
         << a.y << "," << a.z << "," << q.x << "," << q.y << "," << q.z << "," << q.w << "," << p_gt.x << "," << p_gt.y << ","
         << p_gt.z << "," << v_gt.x << "," << v_gt.y << "," << v_gt.z << "," << q_gt.x << "," << q_gt.y << "," << q_gt.z << ","
         << q_gt.w << "," << w.x << "," << w.y << "," << w.z << std::endl;
}

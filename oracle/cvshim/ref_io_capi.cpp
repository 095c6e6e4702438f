// oracle/cvshim/ref_io_capi.cpp — C entry around the reference's own Plus / GroundTruth / ImageReader / DataReader
// sources (TEST INFRASTRUCTURE).  Built by `make -C oracle ref` together with /root/reference/src/{Plus,GroundTruth,
// ImageReader,DataReader}.cpp (unmodified, compiled where they lie) into oracle/_ref/libref_io.so.
#include "DataReader.hpp"
#include <sstream>
#include <iostream>

namespace {
struct Quiet {   // the reference classes narrate on std::cout
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf()) { std::cout.rdbuf(sink.rdbuf()); }
    ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// ---- Plus.cpp --------------------------------------------------------------------------------------------
void ref_toQuaternion(double roll, double pitch, double yaw, double out_wxyz[4]) {
    Quaterniond q = toQuaternion(roll, pitch, yaw);
    out_wxyz[0] = q.w; out_wxyz[1] = q.x; out_wxyz[2] = q.y; out_wxyz[3] = q.z;
}
void ref_toRPY(const double wxyz[4], double out[3]) {
    Quaterniond q; q.w = wxyz[0]; q.x = wxyz[1]; q.y = wxyz[2]; q.z = wxyz[3];
    Point3d a = toRPY(q);
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
}
void ref_toRPY360(const double in[3], double out[3]) {
    Point3d a = toRPY360(Point3d(in[0], in[1], in[2]));
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
}
double ref_computeDiff(double a, double b) { return computeDiff(a, b); }
void ref_rotationMatrix2RPY(const float m[9], double out[3]) {
    Matx33f r;
    for (int i = 0; i < 9; i++) r.val[i] = m[i];
    Point3d a = rotationMatrix2RPY(r);
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
}
void ref_RPY2rotationMatrix(const double rpy[3], float out[9]) {
    Matx33f r = RPY2rotationMatrix(Point3d(rpy[0], rpy[1], rpy[2]));
    for (int i = 0; i < 9; i++) out[i] = r.val[i];
}
void ref_RPYAndPosition2transformationMatrix(const double rpy[3], const double pos[3], float out[16]) {
    Mat t = RPYAndPosition2transformationMatrix(Point3d(rpy[0], rpy[1], rpy[2]), Point3d(pos[0], pos[1], pos[2]));
    for (int i = 0; i < 16; i++) out[i] = t.at<float>(i / 4, i % 4);
}
void ref_transformationMatrix2RPY_position(const float m[16], double rpy[3], double pos[3]) {
    Mat t = Mat::zeros(4, 4, CV_32FC1);
    for (int i = 0; i < 16; i++) t.at<float>(i / 4, i % 4) = m[i];
    Point3d a = transformationMatrix2RPY(t), p = transformationMatrix2position(t);
    rpy[0] = a.x; rpy[1] = a.y; rpy[2] = a.z;
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
}

// ---- GroundTruth.cpp -------------------------------------------------------------------------------------
// Parses `file`; returns rows (= line count, as the class defines it), writes cols, time step and up to cap doubles.
int ref_groundtruth_read(const char* file, char sep, int* cols, double* timestep, double* data, int cap) {
    Quiet q;
    GroundTruth g(file, sep);
    *cols = g.getCols();
    *timestep = g.TimeStep;
    const int rows = g.getRows();
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < g.getCols(); c++)
            if (r * g.getCols() + c < cap) data[r * g.getCols() + c] = g.getGroundTruthData(r, c);
    return rows;
}

// ---- ImageReader.cpp -------------------------------------------------------------------------------------
int ref_imagereader_list(const char* dir, long* times, int cap, double* timestep) {
    Quiet q;
    ImageReader r(dir);
    *timestep = r.TimeStep;
    const int n = (int)r.getSize();
    for (int i = 0; i < n && i < cap; i++) times[i] = r.getImageTime(i);
    return n;
}

// ---- DataReader.cpp --------------------------------------------------------------------------------------
struct RefReader { DataReader d; };
void* ref_datareader_open(const char* image_dir, const char* imu_csv, const char* gt_csv, char sep, int out_idx[4],
                          double out_t[6]) {
    Quiet q;
    RefReader* r = new RefReader();
    r->d.setProperties(image_dir, imu_csv, gt_csv, sep);
    out_idx[0] = r->d.imageIndex0; out_idx[1] = r->d.imuIndex0; out_idx[2] = r->d.gtIndex0; out_idx[3] = r->d.indexLastData;
    out_t[0] = r->d.timeStepCamara; out_t[1] = r->d.timeStepImu; out_t[2] = r->d.timeStepGt;
    out_t[3] = r->d.initialTime; out_t[4] = r->d.lastTime; out_t[5] = 0;
    return r;
}
// One UpdateDataReader(index, index2) step.  counts = {n_imu, n_gt, image1 rows, image1 cols, image2 rows, image2 cols};
// imu = n_imu x 6 (w, a); gt = n_gt x 16 (p3, q wxyz, v3, rpy3, accBias3); misc = {angBias3, currentTimeMs};
// checksums = sum of pixels of image1 / image2.
void ref_datareader_update(void* h, int index, int index2, int counts[6], double* imu, double* gt, double misc[4],
                           long checksums[2], int cap) {
    Quiet q;
    DataReader& d = static_cast<RefReader*>(h)->d;
    d.UpdateDataReader(index, index2);
    counts[0] = (int)d.imuAngularVelocity.size();
    counts[1] = (int)d.gtPosition.size();
    counts[2] = d.image1.rows; counts[3] = d.image1.cols; counts[4] = d.image2.rows; counts[5] = d.image2.cols;
    for (int i = 0; i < counts[0] && i < cap; i++) {
        imu[6 * i] = d.imuAngularVelocity[i].x; imu[6 * i + 1] = d.imuAngularVelocity[i].y; imu[6 * i + 2] = d.imuAngularVelocity[i].z;
        imu[6 * i + 3] = d.imuAcceleration[i].x; imu[6 * i + 4] = d.imuAcceleration[i].y; imu[6 * i + 5] = d.imuAcceleration[i].z;
    }
    for (int i = 0; i < counts[1] && i < cap; i++) {
        double* g = gt + 16 * i;
        g[0] = d.gtPosition[i].x; g[1] = d.gtPosition[i].y; g[2] = d.gtPosition[i].z;
        g[3] = d.gtQuaternion[i].w; g[4] = d.gtQuaternion[i].x; g[5] = d.gtQuaternion[i].y; g[6] = d.gtQuaternion[i].z;
        g[7] = d.gtLinearVelocity[i].x; g[8] = d.gtLinearVelocity[i].y; g[9] = d.gtLinearVelocity[i].z;
        g[10] = d.gtRPY[i].x; g[11] = d.gtRPY[i].y; g[12] = d.gtRPY[i].z;
        g[13] = d.accBias[i].x; g[14] = d.accBias[i].y; g[15] = d.accBias[i].z;
    }
    misc[0] = d.angBias.x; misc[1] = d.angBias.y; misc[2] = d.angBias.z; misc[3] = d.currentTimeMs;
    for (int k = 0; k < 2; k++) {
        const Mat& m = k ? d.image2 : d.image1;
        long s = 0;
        for (int i = 0; i < m.rows * m.cols; i++) s += m.data()[i];
        checksums[k] = s;
    }
}
void ref_datareader_close(void* h) { delete static_cast<RefReader*>(h); }

}  // extern "C"

// host_capi.cpp — plain-C entry points over the host-side class mirrors (Plus helpers, ImageReader, GroundTruth,
// DataReader, image decode, Imu), so that language bindings and the parity tests (ctypes) can drive them without a
// C++ toolchain.  Same shapes as the test harness around the reference's own sources (oracle/cvshim/ref_io_capi.cpp),
// which lets a test feed identical inputs to both and compare bit for bit.  Exceptions become negative status codes.
#include <cstring>
#include <exception>
#include <string>

#include "vislam/CameraModel.hpp"
#include "vislam/DataReader.hpp"
#include "vislam/Imu.hpp"
#include "vislam/Plus.hpp"

namespace {
thread_local std::string g_last_error;
int fail(const std::exception& e) { g_last_error = e.what(); return -1; }
}  // namespace

extern "C" {

const char* vih_last_error(void) { return g_last_error.c_str(); }

// ---- Plus ------------------------------------------------------------------------------------------------
void vih_toQuaternion(double roll, double pitch, double yaw, double out_wxyz[4]) {
    const Quaterniond q = toQuaternion(roll, pitch, yaw);
    out_wxyz[0] = q.w; out_wxyz[1] = q.x; out_wxyz[2] = q.y; out_wxyz[3] = q.z;
}
void vih_toRPY(const double wxyz[4], double out[3]) {
    Quaterniond q; q.w = wxyz[0]; q.x = wxyz[1]; q.y = wxyz[2]; q.z = wxyz[3];
    const cv::Point3d a = toRPY(q);
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
}
void vih_toRPY360(const double in[3], double out[3]) {
    const cv::Point3d a = toRPY360(cv::Point3d(in[0], in[1], in[2]));
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
}
double vih_computeDiff(double a, double b) { return computeDiff(a, b); }
void vih_rotationMatrix2RPY(const float m[9], double out[3]) {
    cv::Matx33f r;
    for (int i = 0; i < 9; i++) r.val[i] = m[i];
    const cv::Point3d a = rotationMatrix2RPY(r);
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
}
void vih_RPY2rotationMatrix(const double rpy[3], float out[9]) {
    const cv::Matx33f r = RPY2rotationMatrix(cv::Point3d(rpy[0], rpy[1], rpy[2]));
    for (int i = 0; i < 9; i++) out[i] = r.val[i];
}
void vih_RPYAndPosition2transformationMatrix(const double rpy[3], const double pos[3], float out[16]) {
    cv::Mat t = RPYAndPosition2transformationMatrix(cv::Point3d(rpy[0], rpy[1], rpy[2]), cv::Point3d(pos[0], pos[1], pos[2]));
    for (int i = 0; i < 16; i++) out[i] = t.at<float>(i / 4, i % 4);
}
void vih_transformationMatrix2RPY_position(const float m[16], double rpy[3], double pos[3]) {
    cv::Mat t = cv::Mat::zeros(4, 4, CV_32FC1);
    for (int i = 0; i < 16; i++) t.at<float>(i / 4, i % 4) = m[i];
    const cv::Point3d a = transformationMatrix2RPY(t), p = transformationMatrix2position(t);
    rpy[0] = a.x; rpy[1] = a.y; rpy[2] = a.z;
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
}

// ---- image decode ----------------------------------------------------------------------------------------
// Returns 0 and the size; with pixels != NULL and cap >= rows * cols also the pixels (row-major u8).
int vih_imread_gray(const char* file, int* rows, int* cols, unsigned char* pixels, long cap) {
    try {
        cv::Mat m = vi::imread_gray(file);
        *rows = m.rows; *cols = m.cols;
        if (m.empty()) { g_last_error = std::string("cannot decode ") + file; return -1; }
        if (pixels && cap >= (long)m.rows * m.cols)
            for (int r = 0; r < m.rows; r++) std::memcpy(pixels + (size_t)r * m.cols, m.ptr<unsigned char>(r), (size_t)m.cols);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- GroundTruth -----------------------------------------------------------------------------------------
int vih_groundtruth_read(const char* file, char sep, int* cols, double* timestep, double* data, int cap) {
    try {
        GroundTruth g(file, sep);
        *cols = g.getCols();
        *timestep = g.TimeStep;
        const int rows = g.getRows();
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < g.getCols(); c++)
                if (r * g.getCols() + c < cap) data[r * g.getCols() + c] = g.getGroundTruthData(r, c);
        return rows;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- ImageReader -----------------------------------------------------------------------------------------
int vih_imagereader_list(const char* dir, long* times, int cap, double* timestep) {
    try {
        ImageReader r(dir);
        *timestep = r.TimeStep;
        const int n = (int)r.getSize();
        for (int i = 0; i < n && i < cap; i++) times[i] = r.getImageTime(i);
        return n;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- DataReader ------------------------------------------------------------------------------------------
void* vih_datareader_open(const char* image_dir, const char* imu_csv, const char* gt_csv, char sep, int out_idx[4],
                          double out_t[6]) {
    try {
        DataReader* d = new DataReader(image_dir, imu_csv, gt_csv, sep);
        out_idx[0] = d->imageIndex0; out_idx[1] = d->imuIndex0; out_idx[2] = d->gtIndex0; out_idx[3] = d->indexLastData;
        out_t[0] = d->timeStepCamara; out_t[1] = d->timeStepImu; out_t[2] = d->timeStepGt;
        out_t[3] = d->initialTime; out_t[4] = d->lastTime; out_t[5] = 0;
        return d;
    } catch (const std::exception& e) { fail(e); return nullptr; }
}
// counts = {n_imu, n_gt, image1 rows, image1 cols, image2 rows, image2 cols}; imu = n_imu x 6 (w, a);
// gt = n_gt x 16 (p3, q wxyz, v3, rpy3, accBias3); misc = {angBias3, currentTimeMs}; checksums = pixel sums.
int vih_datareader_update(void* h, int index, int index2, int counts[6], double* imu, double* gt, double misc[4],
                          long checksums[2], int cap) {
    try {
        DataReader& d = *static_cast<DataReader*>(h);
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
            const cv::Mat& m = k ? d.image2 : d.image1;
            long s = 0;
            for (int r = 0; r < m.rows; r++)
                for (int c = 0; c < m.cols; c++) s += m.at<unsigned char>(r, c);
            checksums[k] = s;
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}
void vih_datareader_close(void* h) { delete static_cast<DataReader*>(h); }

// ---- TrajectoryWriter ------------------------------------------------------------------------------------
// rows: n x 27 doubles in the column order of the CSV (q as x, y, z, w)
int vih_trajectory_write(const char* file, const double* rows, int n) {
    try {
        vi::TrajectoryWriter w(file);
        if (!w.ok()) { g_last_error = std::string("cannot open ") + file; return -1; }
        for (int i = 0; i < n; i++) {
            const double* r = rows + 27 * i;
            Quaterniond q, qg;
            q.x = r[10]; q.y = r[11]; q.z = r[12]; q.w = r[13];
            qg.x = r[20]; qg.y = r[21]; qg.z = r[22]; qg.w = r[23];
            w.write(r[0], cv::Point3d(r[1], r[2], r[3]), cv::Point3d(r[4], r[5], r[6]), cv::Point3d(r[7], r[8], r[9]), q,
                    cv::Point3d(r[14], r[15], r[16]), cv::Point3d(r[17], r[18], r[19]), qg, cv::Point3d(r[24], r[25], r[26]));
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// ---- Imu -------------------------------------------------------------------------------------------------
// The filter alone: n samples (w, a: n x 3) -> n orientations (w, x, y, z).
void vih_madgwick_run(double gain, double dt, const double* w, const double* a, int n, double* q_wxyz) {
    vi::MadgwickFilter f(gain, dt);
    for (int i = 0; i < n; i++) {
        f.update(w[3 * i], w[3 * i + 1], w[3 * i + 2], a[3 * i], a[3 * i + 1], a[3 * i + 2]);
        f.getOrientation(q_wxyz[4 * i], q_wxyz[4 * i + 1], q_wxyz[4 * i + 2], q_wxyz[4 * i + 3]);
    }
}

namespace {
struct ImuTable { const double* in; double* out; int n, next; vi::MadgwickFilter filter; };
void imu_table_source(void* user, const vi::ImuMsg& raw, vi::ImuMsg& fused) {
    ImuTable* t = static_cast<ImuTable*>(user);
    double q[4];
    if (t->in) {
        for (int k = 0; k < 4; k++) q[k] = t->next < t->n ? t->in[4 * t->next + k] : (k == 0 ? 1.0 : 0.0);
    } else {
        t->filter.update(raw.angular_velocity.x, raw.angular_velocity.y, raw.angular_velocity.z, raw.linear_acceleration.x,
                         raw.linear_acceleration.y, raw.linear_acceleration.z);
        t->filter.getOrientation(q[0], q[1], q[2], q[3]);
    }
    if (t->out && t->next < t->n) for (int k = 0; k < 4; k++) t->out[4 * t->next + k] = q[k];
    t->next++;
    fused.orientation.w = q[0]; fused.orientation.x = q[1]; fused.orientation.y = q[2]; fused.orientation.z = q[3];
}
}  // namespace

// One Imu life: initializate(gt_yaw, gt_velocity, first n0 samples), then `steps` x { setImuData(next n_per samples);
// estimate() }.  q_in != NULL: the orientation answered for each published sample (w, x, y, z), otherwise the built-in
// Madgwick filter; q_out (optional) records them.  out: 60 doubles after initializate and after every step, in the
// layout documented in oracle/cvshim/ref_imu_capi.cpp (the reference's Imu runs behind the same signature there).
int vih_imu_run(double timestep, double gt_yaw, const double gt_vel[3], const double* w, const double* a,
                const double* q_in, double* q_out, int n0, int n_per, int steps, double* out) {
    try {
        ImuTable table = {q_in, q_out, n0 + steps * n_per, 0, vi::MadgwickFilter(0.1, timestep)};
        Imu imu(timestep);
        imu.setOrientationSource(imu_table_source, &table);
        auto pts = [](const double* v, int n) {
            std::vector<cv::Point3d> o;
            for (int i = 0; i < n; i++) o.push_back(cv::Point3d(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
            return o;
        };
        std::vector<cv::Point3d> wv = pts(w, n0), av = pts(a, n0);
        imu.initializate(gt_yaw, cv::Point3d(gt_vel[0], gt_vel[1], gt_vel[2]), wv, av);
        auto dump = [&](double* o) {
            const cv::Point3d* p[] = {&imu.residualRPY, &imu.residualPosition, &imu.residualVelocity, &imu.velocity,
                                      &imu.position, &imu.angBias, &imu.accBias, &imu.initialVelocity};
            for (int k = 0; k < 8; k++) { o[3 * k] = p[k]->x; o[3 * k + 1] = p[k]->y; o[3 * k + 2] = p[k]->z; }
            for (int i = 0; i < 9; i++) {
                o[24 + i] = imu.init_rotationMatrix.val[i];
                o[33 + i] = imu.final_rotationMatrix.val[i];
                o[42 + i] = imu.residual_rotationMatrix.val[i];
            }
            const cv::Point3d r = imu.rpyAnglesWorld.empty() ? cv::Point3d() : imu.rpyAnglesWorld.back();
            const cv::Point3d aw = imu.accelerationWorld.empty() ? cv::Point3d() : imu.accelerationWorld.back();
            o[51] = r.x; o[52] = r.y; o[53] = r.z;
            o[54] = aw.x; o[55] = aw.y; o[56] = aw.z;
            o[57] = imu.angularVelocity.x; o[58] = imu.angularVelocity.y; o[59] = imu.angularVelocity.z;
        };
        dump(out);
        for (int s = 0; s < steps; s++) {
            wv = pts(w + 3 * (n0 + s * n_per), n_per);
            av = pts(a + 3 * (n0 + s * n_per), n_per);
            imu.setImuData(wv, av);
            imu.estimate();
            dump(out + 60 * (s + 1));
        }
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

// vi::CameraModel::GetCameraModel (src/CameraModel.cpp:16-101) on a calibration XML: out[0..3] = fx fy cx cy, [4..7] = in / out
// sizes, [8..9] = camera / imu rate, [10..18] = min_features num_max_keyframes start_index use_gt use_ros num_cells length_patch
// detector matcher, [19..34] = imu2cam0Transformation row-major.  Returns -1 (message in vih_last_error) where the reference exits.
int vih_camera_model(const char* path, double* out) {
    try {
        vi::CameraModel m;
        m.GetCameraModel(path);
        const cv::Mat& K = m.GetK();
        out[0] = K.at<float>(0, 0); out[1] = K.at<float>(1, 1); out[2] = K.at<float>(0, 2); out[3] = K.at<float>(1, 2);
        out[4] = m.GetInputWidth(); out[5] = m.GetInputHeight(); out[6] = m.GetOutputWidth(); out[7] = m.GetOutputHeight();
        out[8] = m.camera_frecuency; out[9] = m.imu_frecuency;
        const int v[9] = {m.min_features, m.num_max_keyframes, m.start_index, m.use_gt, m.use_ros, m.num_cells, m.length_patch,
                          m.detector, m.matcher};
        for (int i = 0; i < 9; i++) out[10 + i] = v[i];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) out[19 + 4 * i + j] = m.imu2cam0Transformation.at<float>(i, j);
        return 0;
    } catch (const std::exception& e) { return fail(e); }
}

}  // extern "C"

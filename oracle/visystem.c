/*
 * oracle/visystem.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see vso.h) for the Gauss-Newton
 * photometric pose solve: VISystem::EstimatePoseFeatures (src/VISystem.cpp:1113-1448),
 * WarpFunctionSE3 (:1495-1558), InitializePyramid (:1451-1493), IdentityWeights (:1561-1565),
 * TukeyFunctionWeights / MedianAbsoluteDeviation / MedianMat (:1797-1870), the Sophus SE3f
 * arithmetic it uses (thirdparty/sophus/se3.hpp:253-321,723-742; so3.hpp:338-353,534-568;
 * common.hpp:154-158) and Plus.cpp RPY helpers (:56-83,182-220).
 *
 * Arithmetic conventions (what "the reference's result" means where third-party code is involved):
 *  - float arithmetic is IEEE single, evaluated left to right exactly as the C++ source spells it,
 *    with NO fused multiply-add (build with -ffp-contract=off);
 *  - cv::gemm on CV_32F accumulates in double and rounds once to float (OpenCV GEMMSingleMul<float,double>);
 *  - cv::Mat::inv() is the float LU with partial pivoting of OpenCV hal::LU32f (eps = 10*FLT_EPSILON),
 *    zeros when singular;
 *  - sinf/cosf of Sophus are taken as the correctly rounded float of the double function
 *    ((float)sin((double)x)), which is what glibc's sinf/cosf return in all but vanishingly rare cases
 *    and is reproducible on the device.
 */
#include "vso.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- intrinsics ---- */
void vso_init_pyramid(int w, int h, float fx, float fy, float cx, float cy, vso_intr_t out[VSO_MAX_LEVELS]) {
    out[0].w = w; out[0].h = h;
    out[0].fx = fx; out[0].fy = fy; out[0].cx = cx; out[0].cy = cy;
    out[0].invfx = 1 / out[0].fx;
    out[0].invfy = 1 / out[0].fy;
    for (int lvl = 1; lvl < VSO_MAX_LEVELS; lvl++) {
        out[lvl].w = w >> lvl;                                   /* VISystem.cpp:1468 */
        out[lvl].h = h >> lvl;
        out[lvl].fx = (float)(out[lvl - 1].fx * 0.5);            /* :1470 */
        out[lvl].fy = (float)(out[lvl - 1].fy * 0.5);
        out[lvl].cx = (float)((out[0].cx + 0.5) / ((int)1 << lvl) - 0.5); /* :1472 */
        out[lvl].cy = (float)((out[0].cy + 0.5) / ((int)1 << lvl) - 0.5);
        out[lvl].invfx = 1 / out[lvl].fx;                        /* :1481 */
        out[lvl].invfy = 1 / out[lvl].fy;
    }
}

/* ---------------------------------------------------------------- SE3 (float) ---- */
static const float kSophusEps = 1e-5f; /* common.hpp:154-158 */

static float sinf_cr(float x) { return (float)sin((double)x); }
static float cosf_cr(float x) { return (float)cos((double)x); }

static void quat_to_rot(const float q[4], float R[9]) {
    /* Eigen::QuaternionBase::toRotationMatrix */
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float tx = 2 * x, ty = 2 * y, tz = 2 * z;
    float twx = tx * w, twy = ty * w, twz = tz * w;
    float txx = tx * x, txy = ty * x, txz = tz * x;
    float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void vso_se3_matrix(const float pose[7], float m[16]) {
    float R[9];
    quat_to_rot(pose, R);
    m[0] = R[0]; m[1] = R[1]; m[2] = R[2];  m[3] = pose[4];
    m[4] = R[3]; m[5] = R[4]; m[6] = R[5];  m[7] = pose[5];
    m[8] = R[6]; m[9] = R[7]; m[10] = R[8]; m[11] = pose[6];
    m[12] = 0; m[13] = 0; m[14] = 0; m[15] = 1;
}

void vso_se3_exp(const float d[6], float pose[7]) {
    float ox = d[3], oy = d[4], oz = d[5];
    /* SO3::expAndTheta, so3.hpp:534-568 */
    float theta_sq = (ox * ox + oy * oy) + oz * oz;
    float theta = sqrtf(theta_sq);
    float half_theta = 0.5f * theta;
    float imag_factor, real_factor;
    if (theta < kSophusEps) {
        float theta_po4 = theta_sq * theta_sq;
        imag_factor = 0.5f - (float)(1.0 / 48.0) * theta_sq + (float)(1.0 / 3840.0) * theta_po4;
        real_factor = 1.0f - (float)(1.0 / 8.0) * theta_sq + (float)(1.0 / 384.0) * theta_po4;
    } else {
        float sin_half_theta = sinf_cr(half_theta);
        imag_factor = sin_half_theta / theta;
        real_factor = cosf_cr(half_theta);
    }
    float q[4] = {imag_factor * ox, imag_factor * oy, imag_factor * oz, real_factor};
    /* SE3::exp, se3.hpp:723-742 */
    float Om[9] = {0, -oz, oy, oz, 0, -ox, -oy, ox, 0};
    float Om2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Om2[3 * i + j] = (Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j]) + Om[3 * i + 2] * Om[6 + j];
    float V[9];
    if (theta < kSophusEps) {
        quat_to_rot(q, V);
    } else {
        float th2 = theta * theta;
        float c1 = (1.0f - cosf_cr(theta)) / th2;
        float c2 = (theta - sinf_cr(theta)) / (th2 * theta);
        for (int i = 0; i < 9; i++) {
            float id = (i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f;
            V[i] = (id + c1 * Om[i]) + c2 * Om2[i];
        }
    }
    pose[0] = q[0]; pose[1] = q[1]; pose[2] = q[2]; pose[3] = q[3];
    for (int i = 0; i < 3; i++)
        pose[4 + i] = (V[3 * i] * d[0] + V[3 * i + 1] * d[1]) + V[3 * i + 2] * d[2];
}

void vso_se3_mul(const float a[7], const float b[7], float out[7]) {
    /* translation() += so3() * other.translation()  (se3.hpp:317; Eigen _transformVector) */
    float ax = a[0], ay = a[1], az = a[2], aw = a[3];
    float vx = b[4], vy = b[5], vz = b[6];
    float ux = ay * vz - az * vy, uy = az * vx - ax * vz, uz = ax * vy - ay * vx;
    ux = ux + ux; uy = uy + uy; uz = uz + uz;
    float cx = ay * uz - az * uy, cy = az * ux - ax * uz, cz = ax * uy - ay * ux;
    float rx = (vx + aw * ux) + cx, ry = (vy + aw * uy) + cy, rz = (vz + aw * uz) + cz;
    float tx = a[4] + rx, ty = a[5] + ry, tz = a[6] + rz;
    /* unit_quaternion *= other (so3.hpp:338-353) */
    float bx = b[0], by = b[1], bz = b[2], bw = b[3];
    float qw = ((aw * bw - ax * bx) - ay * by) - az * bz;
    float qx = ((aw * bx + ax * bw) + ay * bz) - az * by;
    float qy = ((aw * by + ay * bw) + az * bx) - ax * bz;
    float qz = ((aw * bz + az * bw) + ax * by) - ay * bx;
    float sn = ((qx * qx + qy * qy) + qz * qz) + qw * qw;
    if (sn != 1.0f) {
        float s = 2.0f / (1.0f + sn);
        qx *= s; qy *= s; qz *= s; qw *= s;
    }
    out[0] = qx; out[1] = qy; out[2] = qz; out[3] = qw;
    out[4] = tx; out[5] = ty; out[6] = tz;
}

void vso_rpy_to_rot(const double rpy[3], float r[9]) {
    double c1 = cos(rpy[0]), s1 = sin(rpy[0]);
    double c2 = cos(rpy[1]), s2 = sin(rpy[1]);
    double c3 = cos(rpy[2]), s3 = sin(rpy[2]);
    r[0] = (float)(c3 * c2); r[1] = (float)(c3 * s2 * s1 - s3 * c1); r[2] = (float)(c3 * s2 * c1 + s3 * s1);
    r[3] = (float)(s3 * c2); r[4] = (float)(s3 * s2 * s1 + c3 * c1); r[5] = (float)(s3 * s2 * c1 - c3 * s1);
    r[6] = (float)(-s2);     r[7] = (float)(c2 * s1);                r[8] = (float)(c2 * c1);
}

void vso_rot_to_rpy(const float r[9], double rpy[3]) {
    double r11 = r[0], r21 = r[3], r31 = r[6], r32 = r[7], r33 = r[8];
    rpy[2] = atan2(r21, r11);
    rpy[1] = atan2(-r31, sqrt(r32 * r32 + r33 * r33));
    rpy[0] = atan2(r32, r33);
}

void vso_rot_to_quat(const float m[9], float q[4]) {
    /* Eigen quaternionbase_assign_impl<Matrix3> */
    float t = (m[0] + m[4]) + m[8];
    if (t > 0.0f) {
        t = sqrtf(t + 1.0f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (m[7] - m[5]) * t;
        q[1] = (m[2] - m[6]) * t;
        q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrtf(((m[4 * i] - m[4 * j]) - m[4 * k]) + 1.0f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

/* Sophus::SE3f(rotation matrix, translation) (se3.hpp:438-440 -> SO3(R), so3.hpp:422-427): the quaternion of R, t as given. */
int vso_se3_from_rt(const float r[9], const float t[3], float pose[7]) {
    vso_rot_to_quat(r, pose);
    pose[4] = t[0]; pose[5] = t[1]; pose[6] = t[2];
    return 0;
}

static void mat33_mul(const float a[9], const float b[9], float c[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            c[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}

void vso_initial_pose(const float imu2cam[9], const float r_imu_res[9], const float t_res[3], float pose[7]) {
    /* VISystem.cpp:1135: rotationMatrix2RPY(imu2cam.t() * R_imu_res * imu2cam); :1146 RPY2rotationMatrix(-rpy);
     * :1162 SE3(rotationEigen, Point(-sx,-sy,-sz)) */
    float it[9] = {imu2cam[0], imu2cam[3], imu2cam[6], imu2cam[1], imu2cam[4], imu2cam[7],
                   imu2cam[2], imu2cam[5], imu2cam[8]};
    float tmp[9], rc[9], r0[9];
    mat33_mul(it, r_imu_res, tmp);
    mat33_mul(tmp, imu2cam, rc);
    double rpy[3];
    vso_rot_to_rpy(rc, rpy);
    rpy[0] = -rpy[0]; rpy[1] = -rpy[1]; rpy[2] = -rpy[2];
    vso_rpy_to_rot(rpy, r0);
    vso_rot_to_quat(r0, pose);
    pose[4] = -t_res[0]; pose[5] = -t_res[1]; pose[6] = -t_res[2];
}

/* ---------------------------------------------------------------- warp ---- */
static inline void warp_point(const float* p, const float m[16], const vso_intr_t* K, float* o) {
    /* VISystem.cpp:1519-1524: col = (col - cx) * invfx; col = col.mul(z).  cv::MatExpr folds "(A - s) * k" into ONE
     * scaled conversion (matop.cpp MatOp_AddEx::multiply + assign -> A.convertTo(dst, type, alpha = k, beta = -s*k)),
     * beta formed in double and rounded to float, the conversion evaluated in float as a*alpha + beta (cvtScale32f,
     * multiply then add, not fused) — i.e. x*invfx - cx*invfx, NOT (x - cx)*invfx. */
    float bx = (float)(-(double)K->cx * (double)K->invfx);
    float by = (float)(-(double)K->cy * (double)K->invfy);
    float X = (p[0] * K->invfx + bx) * p[2];
    float Y = (p[1] * K->invfy + by) * p[2];
    float Z = p[2], W = p[3];
    /* :1536 rigid * pts.t()  — cv::gemm, double accumulation, one rounding to float */
    float r[4];
    for (int i = 0; i < 4; i++) {
        double s = 0.0;
        s += (double)m[4 * i + 0] * (double)X;
        s += (double)m[4 * i + 1] * (double)Y;
        s += (double)m[4 * i + 2] * (double)Z;
        s += (double)m[4 * i + 3] * (double)W;
        r[i] = (float)s;
    }
    /* :1540-1547  row *= f; row /= Z'; row += c;  :1552-1553 row = row.mul(W') */
    float x2 = r[0] * K->fx; x2 = x2 / r[2]; x2 = x2 + K->cx;
    float y2 = r[1] * K->fy; y2 = y2 / r[2]; y2 = y2 + K->cy;
    o[0] = x2 * r[3];
    o[1] = y2 * r[3];
    o[2] = r[2];
    o[3] = r[3];
}

void vso_warp(const float* pts, int n, const float pose[7], const vso_intr_t* K, float* out) {
    float m[16];
    vso_se3_matrix(pose, m);
    for (int i = 0; i < n; i++) warp_point(pts + 4 * i, m, K, out + 4 * i);
}

/* ---------------------------------------------------------------- 6x6 inverse ---- */
/* hal::LU32f (modules/core/src/lapack.cpp, LUImpl<float>): in-place LU with partial pivoting on [A | b], b has n columns.
 * Back-substitution divides by the pivot (s / A[ii]): bit-exact with cv2 4.13 (the only OpenCV that can be executed
 * here; tests/golden/inv6_cv2.npz, solve6_cv2.npz); OpenCV 3.2 multiplied by a stored reciprocal instead, which differs
 * by <= 1 ulp per element — far inside the 1e-5 pose tolerance. */
static int lu32f_6(float A[36], float* b, int n) {
    const int m = 6;
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (fabsf(A[j * m + i]) > fabsf(A[k * m + i])) k = j;
        if (fabsf(A[k * m + i]) < eps) return 0;
        if (k != i) {
            for (int j = i; j < m; j++) { float t = A[i * m + j]; A[i * m + j] = A[k * m + j]; A[k * m + j] = t; }
            for (int j = 0; j < n; j++) { float t = b[i * n + j]; b[i * n + j] = b[k * n + j]; b[k * n + j] = t; }
        }
        float d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            float alpha = A[j * m + i] * d;
            for (k = i + 1; k < m; k++) A[j * m + k] += alpha * A[i * m + k];
            for (k = 0; k < n; k++) b[j * n + k] += alpha * b[i * n + k];
        }
    }
    for (int i = m - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
            float s = b[i * n + j];
            for (int k = i + 1; k < m; k++) s -= A[i * m + k] * b[k * n + j];
            b[i * n + j] = s / A[i * m + i];
        }
    return 1;
}

int vso_inv6(const float a[36], float out[36]) {
    /* cv::invert(DECOMP_LU) n>3: copy, identity RHS, hal::LU32f; failure => zeros. */
    float A[36], b[36];
    memcpy(A, a, sizeof(A));
    for (int i = 0; i < 36; i++) b[i] = 0.f;
    for (int i = 0; i < 6; i++) b[i * 6 + i] = 1.f;
    if (!lu32f_6(A, b, 6)) { memset(out, 0, sizeof(float) * 36); return 0; }
    memcpy(out, b, sizeof(b));
    return 1;
}

int vso_solve6(const float a[36], const float rhs[6], float out[6]) {
    /* cv::solve(A, b, x, DECOMP_LU) n>3: copy A, x = b, hal::LU32f on [A | x]; failure => zeros.
     * This is what `A.inv() * b` evaluates to (VISystem.cpp:1412): cv::MatExpr turns inverse-times-matrix into a solve
     * (matop.cpp MatOp_Invert::matmul -> MatOp_Solve), it never forms the inverse. */
    float A[36], b[6];
    memcpy(A, a, sizeof(A));
    memcpy(b, rhs, sizeof(b));
    if (!lu32f_6(A, b, 1)) { memset(out, 0, sizeof(float) * 6); return 0; }
    memcpy(out, b, sizeof(b));
    return 1;
}

/* ---------------------------------------------------------------- weights ---- */
static float median_hist_u8(const float* v, int n) {
    /* VISystem::MedianMat (:1846-1870): convertTo CV_8U (saturating cvRound), 256-bin histogram */
    int hist[256];
    memset(hist, 0, sizeof(hist));
    for (int i = 0; i < n; i++) {
        long r = lrintf(v[i]);
        if (r < 0) r = 0;
        if (r > 255) r = 255;
        hist[r]++;
    }
    float m = (float)(n / 2);
    int bin = 0;
    float med = -1.0f;
    for (int i = 0; i < 256 && med < 0.0f; ++i) {
        bin += hist[i];
        if ((float)bin > m && med < 0.0f) med = (float)i;
    }
    return med;
}

static void tukey_weights(const float* r, int n, float* w) {
    /* VISystem::TukeyFunctionWeights (:1797-1826), MedianAbsoluteDeviation (:1829-1842) */
    float b = 4.6851f;
    float* dev = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float median = median_hist_u8(r, n);
    for (int i = 0; i < n; i++) dev[i] = fabsf(r[i] - median);
    float MAD = 1.4826f * median_hist_u8(dev, n);
    free(dev);
    if (MAD == 0) MAD = 1;
    float inv_MAD = (float)(1.0 / MAD);
    float inv_b2 = (float)(1.0 / (b * b));
    for (int i = 0; i < n; i++) {
        float x = r[i] * inv_MAD;
        if (fabsf(x) <= b) {
            float tukey = (float)(1.0 - (x * x) * inv_b2);
            w[i] = tukey * tukey;
        } else {
            w[i] = 0.0f;
        }
    }
}

void vso_tukey_weights(const float* r, int n, float* w) { tukey_weights(r, n, w); }

/* ---------------------------------------------------------------- GN ---- */
static inline int round_half_away_pos(float v) {
    float f = floorf(v);
    return (int)f + ((v - f) >= 0.5f ? 1 : 0); /* == round(v) for v > 0 (VISystem.cpp:1321) */
}

int vso_gn_solve(const vso_gn_frames_t* f, const vso_intr_t K[VSO_MAX_LEVELS], const float pose_in[7],
                 const vso_gn_opts_t* o, float pose_out[7], vso_gn_trace_t* trace, int trace_cap) {
    float pose[7];
    memcpy(pose, pose_in, sizeof(pose));
    int ntrace = 0;
    int maxn = 1;
    for (int l = 0; l < VSO_MAX_LEVELS; l++) if (f->n_cand[l] > maxn) maxn = f->n_cand[l];
    float* J = (float*)malloc(sizeof(float) * 6 * (size_t)maxn);
    float* R = (float*)malloc(sizeof(float) * (size_t)maxn);
    float* Wt = (float*)malloc(sizeof(float) * (size_t)maxn);

    for (int lvl = o->first_lvl; lvl >= o->last_lvl; lvl--) {              /* :1181 */
        float last_error = 50000.0f;                                        /* :1185 */
        const uint8_t* image1 = f->prev_img[lvl];
        const uint8_t* image2 = f->cur_img[lvl];
        const int16_t* gx1 = f->prev_gx[lvl];
        const int16_t* gy1 = f->prev_gy[lvl];
        const float* cand = f->cand[lvl];
        int ncand = f->n_cand[lvl];
        int cols = f->img_w[lvl], rows = f->img_h[lvl];
        float fx = K[lvl].fx, fy = K[lvl].fy;
        float z_factor = o->z_factor;
        for (int k = 0; k < o->max_iterations; k++) {                       /* :1214 */
            float m[16];
            vso_se3_matrix(pose, m);
            int nv = 0;
            for (int i = 0; i < ncand; i++) {                               /* :1281-1338 */
                float wp[4];
                warp_point(cand + 4 * i, m, &K[lvl], wp);
                float x1 = cand[4 * i], y1 = cand[4 * i + 1];
                float x2 = wp[0], y2 = wp[1], z2 = wp[2];
                float inv_z2 = 1 / z2;
                if (!(y2 > 0 && y2 < (float)rows && x2 > 0 && x2 < (float)cols)) continue;
                if (!(z2 != 0)) continue;
                if (inv_z2 < 0) inv_z2 = 0;
                float i2;
                if (o->sample_mode == 0) {
                    long lin = (long)round_half_away_pos(y2) * cols + round_half_away_pos(x2);
                    if (lin >= (long)rows * cols) continue;                 /* App. B-4 */
                    i2 = (float)image2[lin];
                } else {
                    float x0 = floorf(x2), y0 = floorf(y2);
                    int ix = (int)x0, iy = (int)y0;
                    if (ix + 1 >= cols || iy + 1 >= rows) continue;
                    float ax = x2 - x0, ay = y2 - y0;
                    const uint8_t* p0 = image2 + (size_t)iy * cols + ix;
                    float i00 = p0[0], i01 = p0[1], i10 = p0[cols], i11 = p0[cols + 1];
                    float top = i00 + ax * (i01 - i00);
                    float bot = i10 + ax * (i11 - i10);
                    i2 = top + ay * (bot - top);
                }
                float Jw[2][6];
                Jw[0][0] = fx * inv_z2;
                Jw[0][1] = 0.0f;
                Jw[0][2] = -(fx * x2 * inv_z2 * inv_z2) * z_factor;
                Jw[0][3] = -(fx * x2 * y2 * inv_z2 * inv_z2);
                Jw[0][4] = (fx * (1 + x2 * x2 * inv_z2 * inv_z2));
                Jw[0][5] = -fx * y2 * inv_z2;
                Jw[1][0] = 0.0f;
                Jw[1][1] = fy * inv_z2;
                Jw[1][2] = -(fy * y2 * inv_z2 * inv_z2) * z_factor;
                Jw[1][3] = -(fy * (1 + y2 * y2 * inv_z2 * inv_z2));
                Jw[1][4] = fy * x2 * y2 * inv_z2 * inv_z2;
                Jw[1][5] = -fy * x2 * inv_z2;
                size_t src = (size_t)((int)y1) * cols + (int)x1;
                float i1 = (float)image1[src];
                R[nv] = i2 - i1;                                            /* :1323 */
                float jl0 = (float)gx1[src], jl1 = (float)gy1[src];         /* :1324-1325 */
                for (int c = 0; c < 6; c++) {                               /* :1327 gemm 1x2 * 2x6 */
                    double s = 0.0;
                    s += (double)jl0 * (double)Jw[0][c];
                    s += (double)jl1 * (double)Jw[1][c];
                    J[6 * nv + c] = (float)s;
                }
                nv++;
            }
            vso_gn_trace_t tr;
            memset(&tr, 0, sizeof(tr));
            tr.lvl = lvl; tr.iter = k; tr.n_valid = nv;
            if (nv == 0) {                                                  /* App. B-12 */
                memcpy(tr.pose, pose, sizeof(pose));
                tr.error = 0.f;
                if (trace && ntrace < trace_cap) trace[ntrace] = tr;
                ntrace++;
                break;
            }
            /* weights :1343-1344 */
            if (o->weight_mode == 1) {
                tukey_weights(R, nv, Wt);
            } else if (o->weight_mode == 2) {
                for (int i = 0; i < nv; i++) {
                    float a = fabsf(R[i]);
                    Wt[i] = (a <= o->huber_k) ? 1.0f : o->huber_k / a;
                }
            } else {
                for (int i = 0; i < nv; i++) Wt[i] = 1.0f;
            }
            /* error :1347-1350 : inv_n * R^T * (R .* W) */
            float inv_num_residuals = (float)(1.0 / nv);
            double es = 0.0;
            for (int i = 0; i < nv; i++) es += (double)R[i] * (double)(R[i] * Wt[i]);
            float error = (float)((double)inv_num_residuals * es);
            tr.error = error;
            if (error >= last_error || k == o->max_iterations - 1 || fabsf(error - last_error) < o->epsilon) { /* :1357 */
                memcpy(tr.pose, pose, sizeof(pose));
                if (trace && ntrace < trace_cap) trace[ntrace] = tr;
                ntrace++;
                break;
            }
            last_error = error;
            /* :1402-1409 */
            double Ad[36], bd[6];
            memset(Ad, 0, sizeof(Ad));
            memset(bd, 0, sizeof(bd));
            for (int i = 0; i < nv; i++) {
                float jw[6];
                for (int c = 0; c < 6; c++) jw[c] = Wt[i] * J[6 * i + c];   /* :1404-1405 */
                float rw = R[i] * Wt[i];
                for (int a = 0; a < 6; a++) {
                    for (int c = 0; c < 6; c++) Ad[6 * a + c] += (double)jw[a] * (double)jw[c];
                    bd[a] += (double)jw[a] * (double)rw;
                }
            }
            float A[36], b[6], delta[6];
            for (int i = 0; i < 36; i++) A[i] = (float)Ad[i];
            for (int i = 0; i < 6; i++) b[i] = (float)(-1.0 * bd[i]);
            vso_solve6(A, b, delta);                                        /* :1412  A.inv() * b == cv::solve(A, b) */
            float e[7], np[7];
            vso_se3_exp(delta, e);
            vso_se3_mul(pose, e, np);                                       /* :1421 */
            memcpy(pose, np, sizeof(pose));
            tr.updated = 1;
            memcpy(tr.pose, pose, sizeof(pose));
            memcpy(tr.delta, delta, sizeof(delta));
            if (trace && ntrace < trace_cap) trace[ntrace] = tr;
            ntrace++;
        }
    }
    memcpy(pose_out, pose, sizeof(pose));                                   /* :1445 */
    free(J); free(R); free(Wt);
    return ntrace;
}

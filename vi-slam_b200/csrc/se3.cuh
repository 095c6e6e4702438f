// se3.cuh — Sophus::SE3f arithmetic as the reference uses it, written once for host and device.
// Restates thirdparty/sophus/se3.hpp:253-321 (matrix, operator*=), :723-742 (exp), so3.hpp:338-353
// (quaternion product + first-order renormalisation), :534-568 (expAndTheta), common.hpp:154-158 (eps),
// Eigen's Quaternion::toRotationMatrix / _transformVector / Quaternion(Matrix3), and Plus.cpp:56-83,182-220.
//
// Every float operation is spelled with an explicitly rounded primitive (no FMA contraction on either
// side), and sin/cos are the float rounding of the double function, so host, device and the CPU oracle
// produce the same bits.
#pragma once
#include <math.h>

#if defined(__CUDA_ARCH__)
#define VSB_HD __host__ __device__ __forceinline__
#define F_MUL(a, b) __fmul_rn((a), (b))
#define F_ADD(a, b) __fadd_rn((a), (b))
#define F_SUB(a, b) __fsub_rn((a), (b))
#define F_DIV(a, b) __fdiv_rn((a), (b))
#define F_SQRT(a) __fsqrt_rn((a))
#else
#if defined(__CUDACC__)
#define VSB_HD __host__ __device__ inline
#else
#define VSB_HD inline
#endif
// host: plain IEEE single ops; the translation unit is built with -ffp-contract=off
static inline float vsb_f_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float vsb_f_add(float a, float b) { volatile float r = a + b; return r; }
static inline float vsb_f_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float vsb_f_div(float a, float b) { volatile float r = a / b; return r; }
#define F_MUL(a, b) vsb_f_mul((a), (b))
#define F_ADD(a, b) vsb_f_add((a), (b))
#define F_SUB(a, b) vsb_f_sub((a), (b))
#define F_DIV(a, b) vsb_f_div((a), (b))
#define F_SQRT(a) sqrtf((a))
#endif

namespace vsb {

constexpr float kSophusEps = 1e-5f;

#if defined(__CUDA_ARCH__)
// Device: for |x| <= 0.5 (every Gauss-Newton update is far inside) the double sine / cosine is a Taylor polynomial in Horner
// form with fused multiply-adds — 10 dependent FP64 operations instead of libdevice's argument reduction + kernel (~5x the
// latency, and the pose update is the serial part of every iteration).  Truncation error < 2e-23, evaluation error
// 2.3e-16 relative; checked on the CPU (same operation sequence, fma()) for EVERY float in [0, 0.5] — 1 056 964 609 values —
// against (float) sin((double) x) / (float) cos((double) x) of glibc: no result differs, so on this range the device
// returns the host's bits by construction (tests/test_gpu_se3.py repeats the sweep on the device).  Odd / even symmetry of
// the polynomials covers negative arguments exactly.  Larger arguments take libdevice's double functions as before.
VSB_HD double sin_small(double x) {
    const double z = __dmul_rn(x, x);
    double p = -1.0 / 355687428096000.0;                 // -1/17!
    p = __fma_rn(p, z, 1.0 / 1307674368000.0);           //  1/15!
    p = __fma_rn(p, z, -1.0 / 6227020800.0);             // -1/13!
    p = __fma_rn(p, z, 1.0 / 39916800.0);                //  1/11!
    p = __fma_rn(p, z, -1.0 / 362880.0);                 // -1/9!
    p = __fma_rn(p, z, 1.0 / 5040.0);
    p = __fma_rn(p, z, -1.0 / 120.0);
    p = __fma_rn(p, z, 1.0 / 6.0);
    return __fma_rn(__dmul_rn(x, z), -p, x);
}
VSB_HD double cos_small(double x) {
    const double z = __dmul_rn(x, x);
    double p = 1.0 / 6402373705728000.0;                 //  1/18!
    p = __fma_rn(p, z, -1.0 / 20922789888000.0);         // -1/16!
    p = __fma_rn(p, z, 1.0 / 87178291200.0);             //  1/14!
    p = __fma_rn(p, z, -1.0 / 479001600.0);              // -1/12!
    p = __fma_rn(p, z, 1.0 / 3628800.0);                 //  1/10!
    p = __fma_rn(p, z, -1.0 / 40320.0);
    p = __fma_rn(p, z, 1.0 / 720.0);
    p = __fma_rn(p, z, -1.0 / 24.0);
    p = __fma_rn(p, z, 0.5);
    return __fma_rn(-z, p, 1.0);
}
VSB_HD float sin_cr(float x) { return fabsf(x) <= 0.5f ? (float)sin_small((double)x) : (float)sin((double)x); }
VSB_HD float cos_cr(float x) { return fabsf(x) <= 0.5f ? (float)cos_small((double)x) : (float)cos((double)x); }
#else
VSB_HD float sin_cr(float x) { return (float)sin((double)x); }
VSB_HD float cos_cr(float x) { return (float)cos((double)x); }
#endif

// q = {x, y, z, w}
VSB_HD void quat_to_rot(const float* q, float* R) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = F_MUL(2.f, x), ty = F_MUL(2.f, y), tz = F_MUL(2.f, z);
    const float twx = F_MUL(tx, w), twy = F_MUL(ty, w), twz = F_MUL(tz, w);
    const float txx = F_MUL(tx, x), txy = F_MUL(ty, x), txz = F_MUL(tz, x);
    const float tyy = F_MUL(ty, y), tyz = F_MUL(tz, y), tzz = F_MUL(tz, z);
    R[0] = F_SUB(1.f, F_ADD(tyy, tzz)); R[1] = F_SUB(txy, twz);             R[2] = F_ADD(txz, twy);
    R[3] = F_ADD(txy, twz);             R[4] = F_SUB(1.f, F_ADD(txx, tzz)); R[5] = F_SUB(tyz, twx);
    R[6] = F_SUB(txz, twy);             R[7] = F_ADD(tyz, twx);             R[8] = F_SUB(1.f, F_ADD(txx, tyy));
}

// row-major 3x4 [R | t] of pose {qx,qy,qz,qw,tx,ty,tz}  (SE3::matrix, se3.hpp:253-268)
VSB_HD void se3_matrix34(const float* pose, float* m) {
    float R[9];
    quat_to_rot(pose, R);
    m[0] = R[0]; m[1] = R[1]; m[2] = R[2];  m[3] = pose[4];
    m[4] = R[3]; m[5] = R[4]; m[6] = R[5];  m[7] = pose[5];
    m[8] = R[6]; m[9] = R[7]; m[10] = R[8]; m[11] = pose[6];
}

VSB_HD float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return F_ADD(F_ADD(F_MUL(a0, b0), F_MUL(a1, b1)), F_MUL(a2, b2));
}

VSB_HD void se3_exp(const float* d, float* pose) {
    const float ox = d[3], oy = d[4], oz = d[5];
    const float theta_sq = F_ADD(F_ADD(F_MUL(ox, ox), F_MUL(oy, oy)), F_MUL(oz, oz));
    const float theta = F_SQRT(theta_sq);
    const float half_theta = F_MUL(0.5f, theta);
    float imag_factor, real_factor;
    if (theta < kSophusEps) {
        const float theta_po4 = F_MUL(theta_sq, theta_sq);
        imag_factor = F_ADD(F_SUB(0.5f, F_MUL((float)(1.0 / 48.0), theta_sq)), F_MUL((float)(1.0 / 3840.0), theta_po4));
        real_factor = F_ADD(F_SUB(1.0f, F_MUL((float)(1.0 / 8.0), theta_sq)), F_MUL((float)(1.0 / 384.0), theta_po4));
    } else {
        const float sin_half_theta = sin_cr(half_theta);
        imag_factor = F_DIV(sin_half_theta, theta);
        real_factor = cos_cr(half_theta);
    }
    float q[4] = {F_MUL(imag_factor, ox), F_MUL(imag_factor, oy), F_MUL(imag_factor, oz), real_factor};
    const float Om[9] = {0.f, -oz, oy, oz, 0.f, -ox, -oy, ox, 0.f};
    float Om2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            Om2[3 * i + j] = dot3(Om[3 * i], Om[j], Om[3 * i + 1], Om[3 + j], Om[3 * i + 2], Om[6 + j]);
    float V[9];
    if (theta < kSophusEps) {
        quat_to_rot(q, V);
    } else {
        const float th2 = F_MUL(theta, theta);
        const float c1 = F_DIV(F_SUB(1.0f, cos_cr(theta)), th2);
        const float c2 = F_DIV(F_SUB(theta, sin_cr(theta)), F_MUL(th2, theta));
        for (int i = 0; i < 9; i++) {
            const float id = (i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f;
            V[i] = F_ADD(F_ADD(id, F_MUL(c1, Om[i])), F_MUL(c2, Om2[i]));
        }
    }
    pose[0] = q[0]; pose[1] = q[1]; pose[2] = q[2]; pose[3] = q[3];
    for (int i = 0; i < 3; i++) pose[4 + i] = dot3(V[3 * i], d[0], V[3 * i + 1], d[1], V[3 * i + 2], d[2]);
}

VSB_HD void se3_mul(const float* a, const float* b, float* out) {
    const float ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const float vx = b[4], vy = b[5], vz = b[6];
    float ux = F_SUB(F_MUL(ay, vz), F_MUL(az, vy));
    float uy = F_SUB(F_MUL(az, vx), F_MUL(ax, vz));
    float uz = F_SUB(F_MUL(ax, vy), F_MUL(ay, vx));
    ux = F_ADD(ux, ux); uy = F_ADD(uy, uy); uz = F_ADD(uz, uz);
    const float cx = F_SUB(F_MUL(ay, uz), F_MUL(az, uy));
    const float cy = F_SUB(F_MUL(az, ux), F_MUL(ax, uz));
    const float cz = F_SUB(F_MUL(ax, uy), F_MUL(ay, ux));
    const float rx = F_ADD(F_ADD(vx, F_MUL(aw, ux)), cx);
    const float ry = F_ADD(F_ADD(vy, F_MUL(aw, uy)), cy);
    const float rz = F_ADD(F_ADD(vz, F_MUL(aw, uz)), cz);
    const float tx = F_ADD(a[4], rx), ty = F_ADD(a[5], ry), tz = F_ADD(a[6], rz);
    const float bx = b[0], by = b[1], bz = b[2], bw = b[3];
    float qw = F_SUB(F_SUB(F_SUB(F_MUL(aw, bw), F_MUL(ax, bx)), F_MUL(ay, by)), F_MUL(az, bz));
    float qx = F_SUB(F_ADD(F_ADD(F_MUL(aw, bx), F_MUL(ax, bw)), F_MUL(ay, bz)), F_MUL(az, by));
    float qy = F_SUB(F_ADD(F_ADD(F_MUL(aw, by), F_MUL(ay, bw)), F_MUL(az, bx)), F_MUL(ax, bz));
    float qz = F_SUB(F_ADD(F_ADD(F_MUL(aw, bz), F_MUL(az, bw)), F_MUL(ax, by)), F_MUL(ay, bx));
    const float sn = F_ADD(F_ADD(F_ADD(F_MUL(qx, qx), F_MUL(qy, qy)), F_MUL(qz, qz)), F_MUL(qw, qw));
    if (sn != 1.0f) {
        const float s = F_DIV(2.0f, F_ADD(1.0f, sn));
        qx = F_MUL(qx, s); qy = F_MUL(qy, s); qz = F_MUL(qz, s); qw = F_MUL(qw, s);
    }
    out[0] = qx; out[1] = qy; out[2] = qz; out[3] = qw;
    out[4] = tx; out[5] = ty; out[6] = tz;
}

// VISystem.cpp:1519-1524 writes col = (col - c) * invf; cv::MatExpr folds "(A - s) * k" into ONE scaled conversion
// A.convertTo(dst, type, alpha = k, beta = -s*k): beta is formed in double and rounded to float, the conversion is
// a*alpha + beta in float (multiply, then add; not fused).  So X = x*invf + backproj_offset(c, invf), not (x - c)*invf.
VSB_HD float backproj_offset(float c, float invf) { return (float)(-(double)c * (double)invf); }

}  // namespace vsb

/*
 * oracle/orb.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see vso.h) for the feature front end the reference runs before the
 * tracked path: cv::ORB::detectAndCompute (Camera.cpp:79-86, 124-129: ORB::create(n); CameraGPU.cpp:99-104: cuda::ORB), one
 * pyramid level (nlevels = 1): FAST-9/16 -> border filter -> retainBest(2n) on the FAST score -> Harris response ->
 * retainBest(n) -> intensity-centroid orientation -> 7x7 Gaussian blur -> steered rBRIEF (256 tests, WTA_K = 2).
 *
 * OpenCV is a third-party dependency of the reference (3.2 by prose, un-vendored); this restates the published algorithm of
 * features2d/src/orb.cpp and is PINNED against the OpenCV that can be executed here, cv2 4.13 (tests/test_oracle_orb.py,
 * tests/golden/orb_cv2.npz): key-point sets, Harris responses and angles bit for bit, descriptors bit for bit.
 * Conventions that were established against cv2 (they are not visible in the API):
 *  - Harris: integer sums a, b, c of the 3x3 Sobel-like gradients over 7x7; the response is evaluated in float in source
 *    order, without fused multiply-add;
 *  - orientation: integer moments over the circular patch, cv::fastAtan2's degree-7 polynomial in float, without FMA;
 *  - the blur inside ORB is NOT cv::GaussianBlur's bit-exact 8-bit path (that path is skipped for a sub-matrix, and ORB
 *    blurs a region of its bordered pyramid buffer): it is the generic separable filter in float — row pass
 *    s = x0*k0, then s = fma(x_i, k_i, s); column pass s = r3*k3, then s = fma(r[3+i] + r[3-i], k[3+i], s) — rounded to the
 *    nearest byte.  cv2's own float intermediate differs in the last bit in the SIMD tail columns; the rounded bytes agreed
 *    on every pixel tested (~4e5) under all summation orders tried;
 *  - retainBest keeps every point whose score is >= the n-th best (ties kept, as cv::KeyPointsFilter::retainBest does); the
 *    ORDER std::nth_element leaves them in is an artefact of libstdc++ and is not reproduced: output is row-major (y, x).
 */
#include "vso.h"
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

static const signed char kPattern[256 * 4] = {
#include "orb_pattern.inc"
};

static int cv_round_f(float v) { return (int)lrintf(v); }

/* HarrisResponses(img, pts, blockSize = 7, k = 0.04): orb.cpp */
void vso_orb_harris(const uint8_t* img, int w, int h, int pitch, const int32_t* xy, int n, float* resp) {
    (void)w; (void)h;
    const int r = 3;
    float scale = 1.f / ((1 << 2) * 7 * 255.f);
    float scale_sq_sq = scale * scale * scale * scale;
    for (int p = 0; p < n; p++) {
        const uint8_t* ptr0 = img + (size_t)(xy[2 * p + 1] - r) * pitch + (xy[2 * p] - r);
        int a = 0, b = 0, c = 0;
        for (int i = 0; i < 7; i++)
            for (int j = 0; j < 7; j++) {
                const uint8_t* q = ptr0 + (size_t)i * pitch + j;
                int Ix = (q[1] - q[-1]) * 2 + (q[-pitch + 1] - q[-pitch - 1]) + (q[pitch + 1] - q[pitch - 1]);
                int Iy = (q[pitch] - q[-pitch]) * 2 + (q[pitch - 1] - q[-pitch - 1]) + (q[pitch + 1] - q[-pitch + 1]);
                a += Ix * Ix;
                b += Iy * Iy;
                c += Ix * Iy;
            }
        float fa = (float)a, fb = (float)b, fc = (float)c;
        float t1 = fa * fb;
        float t2 = fc * fc;
        float s = fa + fb;
        float t3 = 0.04f * s;
        t3 = t3 * s;
        float d = t1 - t2;
        d = d - t3;
        resp[p] = d * scale_sq_sq;
    }
}

/* cv::fastAtan2 (degrees), scalar form */
static float fast_atan2f_deg(float y, float x) {
    const float c = (float)(180 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * c, p3 = -0.3258083974640975f * c, p5 = 0.1555786518463281f * c,
                p7 = -0.04432655554792128f * c;
    float ax = fabsf(x), ay = fabsf(y), a, q, q2;
    if (ax >= ay) {
        q = ay / (ax + (float)DBL_EPSILON);
        q2 = q * q;
        a = p7 * q2; a = a + p5; a = a * q2; a = a + p3; a = a * q2; a = a + p1; a = a * q;
    } else {
        q = ax / (ay + (float)DBL_EPSILON);
        q2 = q * q;
        a = p7 * q2; a = a + p5; a = a * q2; a = a + p3; a = a * q2; a = a + p1; a = a * q;
        a = 90.f - a;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

static void umax_table(int half, int* umax) {
    /* end of each row of the circular patch (orb.cpp computeKeyPoints) */
    int vmax = (int)floor(half * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(half * sqrtf(2.f) / 2);
    for (int v = 0; v <= vmax; v++) umax[v] = (int)lrint(sqrt((double)half * half - (double)v * v));
    for (int v = half, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

/* ICAngles(img, pts, umax, half_k = 15): orb.cpp */
void vso_orb_ic_angle(const uint8_t* img, int w, int h, int pitch, const int32_t* xy, int n, float* angle) {
    (void)w; (void)h;
    const int half = 15;
    int umax[18];
    memset(umax, 0, sizeof(umax));
    umax_table(half, umax);
    for (int p = 0; p < n; p++) {
        const uint8_t* center = img + (size_t)xy[2 * p + 1] * pitch + xy[2 * p];
        int m_01 = 0, m_10 = 0;
        for (int u = -half; u <= half; ++u) m_10 += u * center[u];
        for (int v = 1; v <= half; ++v) {
            int v_sum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * pitch], val_minus = center[u - v * pitch];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        angle[p] = fast_atan2f_deg((float)m_01, (float)m_10);
    }
}

/* 7x7 sigma-2 Gaussian as ORB applies it (generic float separable filter, BORDER_REFLECT_101), see the header */
static const uint32_t kGaussBits[4] = {1032826801u, 1040595070u, 1044597305u, 1046301408u};   /* cv::getGaussianKernel(7, 2, CV_32F)[0..3] */
void vso_orb_gauss_kernel(float k[7]) {
    for (int i = 0; i < 4; i++) { float v; memcpy(&v, &kGaussBits[i], 4); k[i] = v; k[6 - i] = v; }
}
static inline int refl101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; }
    return p;
}
void vso_orb_blur(const uint8_t* img, int w, int h, int pitch, uint8_t* out) {
    float k[7];
    vso_orb_gauss_kernel(k);
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = img + (size_t)y * pitch;
        for (int x = 0; x < w; x++) {
            float acc = (float)s[refl101(x - 3, w)] * k[0];
            for (int i = 1; i < 7; i++) acc = fmaf((float)s[refl101(x - 3 + i, w)], k[i], acc);
            rows[(size_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float acc = rows[(size_t)y * w + x] * k[3];
            for (int i = 1; i <= 3; i++) {
                float t = rows[(size_t)refl101(y + i, h) * w + x] + rows[(size_t)refl101(y - i, h) * w + x];
                acc = fmaf(t, k[3 + i], acc);
            }
            int v = cv_round_f(acc);
            out[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(rows);
}

/* computeOrbDescriptors (WTA_K = 2) on the blurred image: 32 bytes per key point */
void vso_orb_describe(const uint8_t* blurred, int w, int h, int pitch, const int32_t* xy, const float* angle_deg, int n,
                      uint8_t* desc) {
    (void)w; (void)h;
    for (int p = 0; p < n; p++) {
        float angle = angle_deg[p];
        angle *= (float)(3.14159265358979323846 / 180.f);
        float a = (float)cos((double)angle), b = (float)sin((double)angle);
        const uint8_t* center = blurred + (size_t)xy[2 * p + 1] * pitch + xy[2 * p];
        for (int j = 0; j < 32; j++) {
            int val = 0;
            for (int i = 0; i < 8; i++) {
                const signed char* t = kPattern + 4 * (8 * j + i);
                int v[2];
                for (int q = 0; q < 2; q++) {
                    float px = (float)t[2 * q], py = (float)t[2 * q + 1];
                    float m0 = px * a, m1 = py * b, m2 = px * b, m3 = py * a;
                    float x = m0 - m1, y = m2 + m3;
                    v[q] = center[(ptrdiff_t)cv_round_f(y) * pitch + cv_round_f(x)];
                }
                val |= (v[0] < v[1]) << i;
            }
            desc[(size_t)p * 32 + j] = (uint8_t)val;
        }
    }
}

/* KeyPointsFilter::retainBest on (score, index) pairs: keeps every entry with score >= the n-th largest; order preserved */
static int cmp_desc(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return x < y ? 1 : x > y ? -1 : 0;
}
static int retain_best(const float* score, int count, int n, uint8_t* keep) {
    if (n < 0 || count <= n) { memset(keep, 1, (size_t)count); return count; }
    if (n == 0) { memset(keep, 0, (size_t)count); return 0; }
    float* s = (float*)malloc(sizeof(float) * (size_t)count);
    memcpy(s, score, sizeof(float) * (size_t)count);
    qsort(s, (size_t)count, sizeof(float), cmp_desc);
    float thr = s[n - 1];
    free(s);
    int kept = 0;
    for (int i = 0; i < count; i++) { keep[i] = score[i] >= thr; kept += keep[i]; }
    return kept;
}

/* cv::ORB::create(nfeatures, 1.2f, nlevels = 1, edgeThreshold = 31, 0, 2, HARRIS_SCORE, 31, fast_threshold)->detectAndCompute.
 * Outputs (row-major order): xy, Harris response, angle (degrees), 32-byte descriptors; returns the number of key points
 * (<= cap entries are written; ties at the selection thresholds can make it exceed nfeatures, as in OpenCV). */
int vso_orb_detect_compute(const uint8_t* img, int w, int h, int pitch, int nfeatures, int fast_threshold, int32_t* out_xy,
                           float* out_resp, float* out_angle, uint8_t* out_desc, int cap) {
    const int edge = 31;
    int fcap = w * h / 4 + 16;
    int32_t* fxy = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)fcap);
    int32_t* fsc = (int32_t*)malloc(sizeof(int32_t) * (size_t)fcap);
    int nf = vso_fast9(img, w, h, pitch, fast_threshold, 1, fxy, fsc, fcap);
    if (nf > fcap) nf = fcap;
    /* KeyPointsFilter::runByImageBorder(edgeThreshold): Rect(edge, edge, w - 2 edge, h - 2 edge).contains(pt) */
    int m = 0;
    for (int i = 0; i < nf; i++) {
        int x = fxy[2 * i], y = fxy[2 * i + 1];
        if (x >= edge && x < w - edge && y >= edge && y < h - edge) { fxy[2 * m] = x; fxy[2 * m + 1] = y; fsc[m] = fsc[i]; m++; }
    }
    float* sc = (float*)malloc(sizeof(float) * (size_t)(m + 1));
    uint8_t* keep = (uint8_t*)malloc((size_t)m + 1);
    for (int i = 0; i < m; i++) sc[i] = (float)fsc[i];
    retain_best(sc, m, 2 * nfeatures, keep);                        /* on the FAST score */
    int m2 = 0;
    for (int i = 0; i < m; i++) if (keep[i]) { fxy[2 * m2] = fxy[2 * i]; fxy[2 * m2 + 1] = fxy[2 * i + 1]; m2++; }
    vso_orb_harris(img, w, h, pitch, fxy, m2, sc);
    retain_best(sc, m2, nfeatures, keep);                           /* on the Harris response */
    int m3 = 0;
    for (int i = 0; i < m2; i++) if (keep[i]) { fxy[2 * m3] = fxy[2 * i]; fxy[2 * m3 + 1] = fxy[2 * i + 1]; sc[m3] = sc[i]; m3++; }
    int nout = m3 < cap ? m3 : cap;
    float* ang = (float*)malloc(sizeof(float) * (size_t)(m3 + 1));
    vso_orb_ic_angle(img, w, h, pitch, fxy, m3, ang);
    if (out_xy) memcpy(out_xy, fxy, sizeof(int32_t) * 2 * (size_t)nout);
    if (out_resp) memcpy(out_resp, sc, sizeof(float) * (size_t)nout);
    if (out_angle) memcpy(out_angle, ang, sizeof(float) * (size_t)nout);
    if (out_desc) {
        uint8_t* blurred = (uint8_t*)malloc((size_t)w * h);
        vso_orb_blur(img, w, h, pitch, blurred);
        vso_orb_describe(blurred, w, h, w, fxy, ang, nout, out_desc);
        free(blurred);
    }
    free(ang); free(keep); free(sc); free(fsc); free(fxy);
    return m3;
}

/* ---------------------------------------------------------------------------------------------- the scale pyramid ---- */
/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR_EXACT) for 8-bit images (imgproc/src/resize.cpp, resize_bitExact with
 * interpolationLinear<uchar>): coefficients in 8.8 fixed point from scale = 1 / (dsize / ssize) evaluated in double,
 * horizontal pass into 8.8 values, vertical pass into 16.16 and rounding; positions that fall before the first / after the
 * last source sample take that sample.  Bit-exact against cv2 4.13 (tests/test_oracle_orb.py). */
static void lin_exact_coeff(int v, int ssize, int dsize, int* ofs, int* c1, int* edge) {
    double inv = (double)dsize / (double)ssize;
    double scale = 1.0 / inv;
    double fval = scale * ((double)v + 0.5) - 0.5;
    int ival = (int)floor(fval);
    *edge = 0; *ofs = 0; *c1 = 0;
    if (ival >= 0 && ssize > 1) {
        if (ival < ssize - 1) { *ofs = ival; *c1 = (int)lrint((fval - (double)ival) * 256.0); }
        else { *ofs = ssize - 1; *edge = 1; }
    } else {
        *edge = -1;
    }
}
void vso_resize_linear_exact(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh) {
    int* ox = (int*)malloc(sizeof(int) * 3 * (size_t)dw);
    int *ax = ox + dw, *ex = ox + 2 * dw;
    for (int x = 0; x < dw; x++) lin_exact_coeff(x, sw, dw, &ox[x], &ax[x], &ex[x]);
    uint32_t* h0 = (uint32_t*)malloc(sizeof(uint32_t) * 2 * (size_t)dw);
    uint32_t* h1 = h0 + dw;
    for (int y = 0; y < dh; y++) {
        int oy, ay, ey;
        lin_exact_coeff(y, sh, dh, &oy, &ay, &ey);
        const uint8_t* r0 = src + (size_t)(ey < 0 ? 0 : oy) * spitch;
        const uint8_t* r1 = src + (size_t)(ey != 0 ? (ey < 0 ? 0 : oy) : oy + 1) * spitch;
        for (int x = 0; x < dw; x++) {
            if (ex[x] < 0) { h0[x] = (uint32_t)r0[0] << 8; h1[x] = (uint32_t)r1[0] << 8; }
            else if (ex[x] > 0) { h0[x] = (uint32_t)r0[sw - 1] << 8; h1[x] = (uint32_t)r1[sw - 1] << 8; }
            else {
                h0[x] = (uint32_t)(256 - ax[x]) * r0[ox[x]] + (uint32_t)ax[x] * r0[ox[x] + 1];
                h1[x] = (uint32_t)(256 - ax[x]) * r1[ox[x]] + (uint32_t)ax[x] * r1[ox[x] + 1];
            }
        }
        for (int x = 0; x < dw; x++) {
            uint32_t v = ey != 0 ? h0[x] << 8 : (uint32_t)(256 - ay) * h0[x] + (uint32_t)ay * h1[x];
            v = (v + (1u << 15)) >> 16;
            dst[(size_t)y * dw + x] = (uint8_t)(v > 255 ? 255 : v);
        }
    }
    free(h0); free(ox);
}

/* layer scale and per-level feature budget of cv::ORB (orb.cpp getScale / computeKeyPoints).  ORB::create takes the scale
 * factor as a FLOAT and stores it in a double, so every level scale is (float)pow((double)1.2f, level). */
float vso_orb_level_scale(float scale_factor, int level) { return (float)pow((double)scale_factor, (double)level); }
void vso_orb_level_budget(int nfeatures, float scale_factor, int nlevels, int* per_level) {
    float factor = (float)(1.0 / (double)scale_factor);
    float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        per_level[l] = (int)lrintf(nd);
        sum += per_level[l];
        nd *= factor;
    }
    per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

/* cv::ORB::create(nfeatures, scale_factor, nlevels, 31, 0, 2, HARRIS_SCORE, 31, fast_threshold)->detectAndCompute:
 * every level is the one-level pipeline above on the level image (level l = INTER_LINEAR_EXACT resize of level l-1 to
 * cvRound(size / scale_l)) with that level's budget; key points are reported level by level (row-major inside a level) with
 * pt = level coordinates * scale_l, octave = level.  Returns the number of key points (<= cap are written). */
int vso_orb_detect_compute_pyr(const uint8_t* img, int w, int h, int pitch, int nfeatures, float scale_factor, int nlevels,
                               int fast_threshold, float* out_xy, int32_t* out_octave, float* out_resp, float* out_angle,
                               uint8_t* out_desc, int cap) {
    int budget[32];
    if (nlevels < 1) nlevels = 1;
    if (nlevels > 32) nlevels = 32;
    vso_orb_level_budget(nfeatures, scale_factor, nlevels, budget);
    uint8_t* prev = NULL;
    int pw = w, ph = h, total = 0;
    for (int l = 0; l < nlevels; l++) {
        float scale = vso_orb_level_scale(scale_factor, l);
        const uint8_t* cur = img;
        int cw = w, ch = h, cp = pitch;
        uint8_t* mine = NULL;
        if (l > 0) {
            cw = cv_round_f((float)w / scale);
            ch = cv_round_f((float)h / scale);
            if (cw < 1 || ch < 1) break;
            mine = (uint8_t*)malloc((size_t)cw * ch);
            vso_resize_linear_exact(prev ? prev : img, pw, ph, prev ? pw : pitch, mine, cw, ch);
            cur = mine; cp = cw;
        }
        if (cw > 62 && ch > 62) {
            int lcap = cw * ch / 4 + 16;
            int32_t* xy = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)lcap);
            float* resp = (float*)malloc(sizeof(float) * (size_t)lcap);
            float* ang = (float*)malloc(sizeof(float) * (size_t)lcap);
            uint8_t* desc = out_desc ? (uint8_t*)malloc((size_t)lcap * 32) : NULL;
            int n = vso_orb_detect_compute(cur, cw, ch, cp, budget[l], fast_threshold, xy, resp, ang, desc, lcap);
            if (n > lcap) n = lcap;
            for (int i = 0; i < n; i++, total++) {
                if (total >= cap) continue;
                if (out_xy) { out_xy[2 * total] = (float)xy[2 * i] * scale; out_xy[2 * total + 1] = (float)xy[2 * i + 1] * scale; }
                if (out_octave) out_octave[total] = l;
                if (out_resp) out_resp[total] = resp[i];
                if (out_angle) out_angle[total] = ang[i];
                if (out_desc) memcpy(out_desc + (size_t)total * 32, desc + (size_t)i * 32, 32);
            }
            free(desc); free(ang); free(resp); free(xy);
        }
        free(prev);
        prev = mine;
        if (l > 0) { pw = cw; ph = ch; }
    }
    free(prev);
    return total;
}

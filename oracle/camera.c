/*
 * oracle/camera.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see vso.h) for src/Camera.cpp:
 * Camera::Update pyramid (:63-72), computeGradient (:167-184), ObtainPatchesPointsPreviousFrame
 * (:358-409).  cv::resize / cv::Scharr / cv::addWeighted semantics are restated from OpenCV and
 * checked bit-exactly against cv2 4.13 in tests/test_oracle_cv2.py.
 */
#include "vso.h"
#include <math.h>
#include <stdlib.h>

static int cv_round(double v) { return (int)lrint(v); } /* round-half-even, as cvRound on SSE2 */

void vso_pyr_size(int w, int h, int* dw, int* dh) {
    /* resize(..., Size(), 0.5, 0.5): dsize = saturate_cast<int>(ssize * 0.5) = cvRound */
    *dw = cv_round(w * 0.5);
    *dh = cv_round(h * 0.5);
}

void vso_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst) {
    /* fx = fy = 0.5 exactly => scale 2 => OpenCV swaps INTER_LINEAR for the INTER_AREA fast path:
     * full 2x2 blocks give (a+b+c+d+2)>>2; blocks clipped by an odd edge average what exists in float. */
    int dw, dh;
    vso_pyr_size(w, h, &dw, &dh);
    for (int dy = 0; dy < dh; dy++) {
        for (int dx = 0; dx < dw; dx++) {
            int sx0 = 2 * dx, sy0 = 2 * dy;
            if (sx0 + 1 < w && sy0 + 1 < h) {
                const uint8_t* r0 = src + (size_t)sy0 * w + sx0;
                const uint8_t* r1 = r0 + w;
                dst[(size_t)dy * dw + dx] = (uint8_t)((r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2);
            } else {
                int sum = 0, count = 0;
                if (sx0 < w) {
                    for (int sy = 0; sy < 2; sy++) {
                        if (sy0 + sy >= h) break;
                        for (int sx = 0; sx < 2; sx++) {
                            if (sx0 + sx >= w) break;
                            sum += src[(size_t)(sy0 + sy) * w + sx0 + sx];
                            count++;
                        }
                    }
                }
                int v = count ? cv_round((float)sum / (float)count) : 0;
                dst[(size_t)dy * dw + dx] = (uint8_t)(v > 255 ? 255 : v);
            }
        }
    }
}

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    if (p < 0) return -p;
    if (p >= len) return 2 * len - 2 - p;
    return p;
}

void vso_scharr3(const uint8_t* src, int w, int h, int16_t* gx, int16_t* gy) {
    /* Scharr(..., CV_16S, 1,0, scale=3) : deriv [-1 0 1] along x, smooth [3 10 3] along y, times 3.
     * (Camera.cpp:171-172 — the literal 3 is the scale argument, App. B-8.)  Integer-exact. */
    for (int y = 0; y < h; y++) {
        int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
        const uint8_t* r0 = src + (size_t)ym * w;
        const uint8_t* r1 = src + (size_t)y * w;
        const uint8_t* r2 = src + (size_t)yp * w;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int dx = 3 * (r0[xp] - r0[xm]) + 10 * (r1[xp] - r1[xm]) + 3 * (r2[xp] - r2[xm]);
            int dy = 3 * (r2[xm] - r0[xm]) + 10 * (r2[x] - r0[x]) + 3 * (r2[xp] - r0[xp]);
            gx[(size_t)y * w + x] = (int16_t)(3 * dx);
            gy[(size_t)y * w + x] = (int16_t)(3 * dy);
        }
    }
}

void vso_grad_mag(const int16_t* gx, const int16_t* gy, int n, uint8_t* g) {
    /* convertScaleAbs (u8 saturating |v|) then addWeighted(0.5, 0.5) with cvRound (Camera.cpp:174-180) */
    for (int i = 0; i < n; i++) {
        int ax = abs((int)gx[i]), ay = abs((int)gy[i]);
        if (ax > 255) ax = 255;
        if (ay > 255) ay = 255;
        float t = (float)ax * 0.5f + (float)ay * 0.5f;
        int v = cv_round(t);
        g[i] = (uint8_t)(v > 255 ? 255 : v);
    }
}

int vso_candidates(const float* good_xy, int nf, int lvl, int lw, int lh, float* out) {
    static const int patch_size[VSO_MAX_LEVELS] = {5, 3, 2, 5, 5}; /* Camera.cpp:369-373 */
    float factor_lvl = (float)(1.0 / pow(2, lvl));                  /* Camera.cpp:379 */
    int start_point = patch_size[lvl] - 1 / 2;                      /* == patch_size (int division), :381 */
    int nmax = nf < 200 ? nf : 200;                                 /* Camera.cpp:382 */
    int n = 0;
    for (int k = 0; k < nmax; k++) {
        float x = (float)(((double)good_xy[2 * k] + 0.5) * (double)factor_lvl - 0.5);     /* :384 */
        float y = (float)(((double)good_xy[2 * k + 1] + 0.5) * (double)factor_lvl - 0.5); /* :385 */
        float xhi = x + (float)start_point, yhi = y + (float)start_point;
        for (int i = (int)(x - (float)start_point); (float)i <= xhi; i++) {         /* :391 */
            for (int j = (int)(y - (float)start_point); (float)j <= yhi; j++) {     /* :392 */
                if (i > 0 && i < lw && j > 0 && j < lh) {                           /* :393 */
                    out[4 * n + 0] = (float)i;
                    out[4 * n + 1] = (float)j;
                    out[4 * n + 2] = 1.0f;
                    out[4 * n + 3] = 1.0f;
                    n++;
                }
            }
        }
    }
    return n;
}

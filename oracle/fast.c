/* oracle/fast.c — CPU restatement of the FAST-9/16 corner detector (TEST INFRASTRUCTURE ONLY: imported by tests/ and by
 * bench legs that measure the CPU baseline, never by the product).
 *
 * Where it sits in the reference: the first stage of feature detection — cv::ORB (Camera.cpp:124-129, ORB::create(200))
 * and cv::cuda::ORB (CameraGPU.cpp:99-104, cuda::ORB::create(1000)) both find their key points with FAST-9/16, threshold
 * 20, with non-maximum suppression — SURVEY.md 8f row N-4.  The algorithm itself lives in OpenCV (pinned by prose to 3.2,
 * README.md:15; not vendored), so this file restates the published algorithm (Rosten & Drummond; cv::FAST, TYPE_9_16) and
 * is pinned by golden vectors produced with cv2 4.13 in this container (tests/golden/fast_cv2.npz) and by a live cv2
 * comparison (tests/test_oracle_cv2.py):
 *   - a pixel p (3 <= x < w-3, 3 <= y < h-3) is a corner iff 9 contiguous pixels of the 16-pixel Bresenham ring of
 *     radius 3 are all > p + t or all < p - t;
 *   - its score is the largest threshold for which it would still be a corner, minus one
 *     (max over the 9-arcs of the smallest |difference| on the arc, both signs) — always >= t for a corner;
 *   - with non-maximum suppression a corner is kept iff its score is strictly greater than the scores of its 8
 *     neighbours (non-corners count as 0);
 *   - corners are reported in row-major order with response = score (0 without suppression). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int RING_X[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_Y[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* 0 = not a corner, otherwise the corner score (>= threshold >= 0; a score of 0 is reported as 0 too, like cv::FAST
 * whose suppression treats "no corner" and "score 0" alike) plus one, so that callers can tell the two apart. */
static int corner_score_plus1(const uint8_t* p, int pitch, int t) {
    int d[16];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - (int)p[RING_Y[k] * pitch + RING_X[k]];
    int best = -1;                                    /* max over arcs of min |d| with one sign */
    for (int s = 0; s < 16; s++) {
        int lo = 1 << 30, hi = 1 << 30;               /* min of d (centre brighter), min of -d (centre darker) */
        for (int k = 0; k < 9; k++) {
            const int x = d[(s + k) & 15];
            if (x < lo) lo = x;
            if (-x < hi) hi = -x;
        }
        if (lo > best) best = lo;
        if (hi > best) best = hi;
    }
    if (best <= t) return 0;                          /* corner iff some arc has every |d| > t */
    return best;                                      /* score = best - 1; returned + 1 */
}

/* img: h rows of `pitch` bytes.  Writes up to cap corners (x, y, score) in row-major order; returns the TOTAL number found. */
int vso_fast9(const uint8_t* img, int w, int h, int pitch, int threshold, int nonmax, int32_t* out_xy, int32_t* out_score,
              int cap) {
    if (w < 7 || h < 7) return 0;
    int* score = (int*)calloc((size_t)w * h, sizeof(int));      /* score + 1, 0 = no corner */
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) score[(size_t)y * w + x] = corner_score_plus1(img + (size_t)y * pitch + x, pitch, threshold);
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const int s1 = score[(size_t)y * w + x];
            if (!s1) continue;
            const int s = s1 - 1;
            int keep = 1;
            if (nonmax) {
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        const int o1 = score[(size_t)(y + dy) * w + (x + dx)];
                        const int o = o1 ? o1 - 1 : 0;
                        if (!(s > o)) { keep = 0; break; }
                    }
            }
            if (!keep) continue;
            if (n < cap) {
                out_xy[2 * n] = x; out_xy[2 * n + 1] = y;
                out_score[n] = nonmax ? s : 0;
            }
            n++;
        }
    free(score);
    return n;
}

/*
 * oracle/matcher.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see vso.h) for src/Matcher.cpp.
 * Restates: computeMatches (:83-94), nnFilter (:148-169), computeSymMatches (:96-144),
 * sortMatches (:329-352), bestMatchesFilter (:171-244), getGoodMatches (:295-303),
 * computeBestMatches (:353-367).  The O(N*M*D) kernels live in OpenCV (BFMatcher::knnMatch);
 * their semantics (exhaustive, sorted by distance then train index) are checked against cv2 4.13.
 */
#include "vso.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline void top2_insert(int32_t* bi, float* bd, int* cnt, int j, float d) {
    /* strict '<' keeps the lower train index first on exact ties (j ascends) */
    if (*cnt < 1 || d < bd[0]) {
        if (*cnt >= 1) { bd[1] = bd[0]; bi[1] = bi[0]; }
        bd[0] = d; bi[0] = j;
        if (*cnt < 2) (*cnt)++;
    } else if (*cnt < 2 || d < bd[1]) {
        bd[1] = d; bi[1] = j;
        if (*cnt < 2) (*cnt)++;
    }
}

void vso_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                      int32_t* idx, float* dist) {
    for (int i = 0; i < nq; i++) {
        int32_t bi[2] = {-1, -1};
        float bd[2] = {0.f, 0.f};
        int cnt = 0;
        const uint8_t* a = q + (size_t)i * nbytes;
        for (int j = 0; j < nt; j++) {
            const uint8_t* b = t + (size_t)j * nbytes;
            int d = 0, k = 0;
            for (; k + 8 <= nbytes; k += 8) {
                uint64_t x, y;
                memcpy(&x, a + k, 8);
                memcpy(&y, b + k, 8);
                d += __builtin_popcountll(x ^ y);
            }
            for (; k < nbytes; k++) d += __builtin_popcount((unsigned)(a[k] ^ b[k]));
            top2_insert(bi, bd, &cnt, j, (float)d);
        }
        idx[2 * i] = bi[0]; idx[2 * i + 1] = bi[1];
        dist[2 * i] = bd[0]; dist[2 * i + 1] = bd[1];
    }
}

void vso_knn2_l2(const float* q, int nq, const float* t, int nt, int dim, int32_t* idx, float* dist) {
    for (int i = 0; i < nq; i++) {
        int32_t bi[2] = {-1, -1};
        float bd[2] = {0.f, 0.f};
        int cnt = 0;
        const float* a = q + (size_t)i * dim;
        for (int j = 0; j < nt; j++) {
            const float* b = t + (size_t)j * dim;
            double s = 0.0;
            for (int k = 0; k < dim; k++) {
                float d = a[k] - b[k];
                s += (double)d * (double)d;
            }
            top2_insert(bi, bd, &cnt, j, sqrtf((float)s));
        }
        idx[2 * i] = bi[0]; idx[2 * i + 1] = bi[1];
        dist[2 * i] = bd[0]; dist[2 * i + 1] = bd[1];
    }
}

void vso_nn_filter(const int32_t* idx, const float* dist, int n, double ratio, uint8_t* keep) {
    for (int i = 0; i < n; i++) {
        if (idx[2 * i] >= 0 && idx[2 * i + 1] >= 0) {
            /* Matcher.cpp:156: float > double*float, evaluated in double */
            keep[i] = !((double)dist[2 * i] > ratio * (double)dist[2 * i + 1]);
        } else {
            keep[i] = 0; /* fewer than 2 neighbours, Matcher.cpp:162-165 */
        }
    }
}

int vso_sym_matches(const int32_t* idx1, const float* dist1, int n1,
                    const int32_t* idx2, const float* dist2, int n2,
                    double ratio, int mode, int32_t* mq, int32_t* mt, float* md) {
    uint8_t* keep1 = (uint8_t*)malloc((size_t)(n1 > 0 ? n1 : 1));
    uint8_t* keep2 = (uint8_t*)malloc((size_t)(n2 > 0 ? n2 : 1));
    vso_nn_filter(idx1, dist1, n1, ratio, keep1);
    vso_nn_filter(idx2, dist2, n2, ratio, keep2);
    int cnt = 0;
    for (int i = 0; i < n1; i++) {
        if (!keep1[i]) continue;           /* Matcher.cpp:116 */
        int j = idx1[2 * i];               /* aux2[j][0].queryIdx == j, so the scan :119-139 finds j or nothing */
        if (j < 0 || j >= n2) continue;
        if (idx2[2 * j] < 0) continue;     /* row never had an element: nothing to read, even stale */
        if (mode == 1 && !keep2[j]) continue; /* intended semantics; de-facto reads the cleared row (:122-125) */
        if (idx2[2 * j] == i) {
            mq[cnt] = i; mt[cnt] = j; md[cnt] = dist1[2 * i];
            cnt++;
        }
    }
    free(keep1); free(keep2);
    return cnt;
}

typedef struct { float y; int32_t pos; } sort_key_t;
static int cmp_key(const void* a, const void* b) {
    const sort_key_t* x = (const sort_key_t*)a;
    const sort_key_t* y = (const sort_key_t*)b;
    if (x->y < y->y) return -1;
    if (x->y > y->y) return 1;
    return (x->pos > y->pos) - (x->pos < y->pos); /* stable (decision for cv::sortIdx ties) */
}

void vso_sort_matches(const int32_t* mq, int n, const float* kp1_xy, int32_t* order) {
    sort_key_t* k = (sort_key_t*)malloc(sizeof(sort_key_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { k[i].y = kp1_xy[2 * mq[i] + 1]; k[i].pos = i; }
    qsort(k, (size_t)n, sizeof(sort_key_t), cmp_key);
    for (int i = 0; i < n; i++) order[i] = k[i].pos;
    free(k);
}

int vso_grid_filter(const int32_t* mq, const int32_t* mt, const float* md, const int32_t* order, int n,
                    const float* kp1_xy, int w, int h, int n_cells,
                    int32_t* gq, int32_t* gt, float* gd) {
    if (n <= 0 || n_cells < 1) return 0;     /* App. B-3: empty in => empty out */
    float winW = (float)((double)w / floor(sqrt((double)n_cells)));   /* Matcher.cpp:177 */
    float winH = (float)((double)h / floor(sqrt((double)n_cells)));   /* Matcher.cpp:178 */
    int root_n = (int)floor(sqrt((double)n_cells));                   /* Matcher.cpp:191 */
    float* cd = (float*)malloc(sizeof(float) * (size_t)root_n);
    int32_t* cq = (int32_t*)malloc(sizeof(int32_t) * (size_t)root_n);
    int32_t* ct = (int32_t*)malloc(sizeof(int32_t) * (size_t)root_n);
    for (int i = 0; i < root_n; i++) cd[i] = 100000.0f;
    int it = 0, out = 0;
    float h_final = winH;
    for (int j = 0; j < root_n; j++) {
        while (kp1_xy[2 * mq[order[it]] + 1] <= h_final) {
            float x = kp1_xy[2 * mq[order[it]]];
            float w_final = winW;
            int i = 0;
            while (x > w_final) { w_final = w_final + winW; i++; }
            if (i >= root_n) i = root_n - 1;  /* App. B-13: reference would index out of bounds */
            if (md[order[it]] < cd[i]) {
                cd[i] = md[order[it]]; cq[i] = mq[order[it]]; ct[i] = mt[order[it]];
            }
            ++it;
            if (it == n) break;
        }
        for (int i = 0; i < root_n; i++) {
            if (cd[i] != 100000.0f) { gq[out] = cq[i]; gt[out] = ct[i]; gd[out] = cd[i]; out++; }
            cd[i] = 100000.0f;
        }
        h_final = h_final + winH;
        if (it == n) break;
    }
    free(cd); free(cq); free(ct);
    return out;
}

int vso_match_pipeline(const void* d1, int n1, const void* d2, int n2, int dim, int norm,
                       const float* kp1_xy, int w, int h, int n_cells, double ratio, int mode,
                       int32_t* gq, int32_t* gt, float* gd, int* n_sym) {
    size_t c1 = (size_t)(n1 > 0 ? n1 : 1), c2 = (size_t)(n2 > 0 ? n2 : 1);
    int32_t* idx1 = (int32_t*)malloc(sizeof(int32_t) * 2 * c1);
    int32_t* idx2 = (int32_t*)malloc(sizeof(int32_t) * 2 * c2);
    float* dist1 = (float*)malloc(sizeof(float) * 2 * c1);
    float* dist2 = (float*)malloc(sizeof(float) * 2 * c2);
    int32_t* mq = (int32_t*)malloc(sizeof(int32_t) * c1);
    int32_t* mt = (int32_t*)malloc(sizeof(int32_t) * c1);
    float* md = (float*)malloc(sizeof(float) * c1);
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * c1);
    if (norm == 1) {
        vso_knn2_hamming((const uint8_t*)d1, n1, (const uint8_t*)d2, n2, dim, idx1, dist1);
        vso_knn2_hamming((const uint8_t*)d2, n2, (const uint8_t*)d1, n1, dim, idx2, dist2);
    } else {
        vso_knn2_l2((const float*)d1, n1, (const float*)d2, n2, dim, idx1, dist1);
        vso_knn2_l2((const float*)d2, n2, (const float*)d1, n1, dim, idx2, dist2);
    }
    int ns = vso_sym_matches(idx1, dist1, n1, idx2, dist2, n2, ratio, mode, mq, mt, md);
    if (n_sym) *n_sym = ns;
    vso_sort_matches(mq, ns, kp1_xy, order);
    int ng = vso_grid_filter(mq, mt, md, order, ns, kp1_xy, w, h, n_cells, gq, gt, gd);
    free(idx1); free(idx2); free(dist1); free(dist2); free(mq); free(mt); free(md); free(order);
    return ng;
}

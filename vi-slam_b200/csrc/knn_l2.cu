// knn_l2.cu — brute-force kNN (k = 2) over float descriptors, BFMatcher(NORM_L2)::knnMatch as called by
// Matcher::computeMatches (reference src/Matcher.cpp:55, 83-94), both directions from ONE pass over the
// distance matrix.
//
// This file holds the EXACT kernel: every distance is sqrtf((float) sum_k (double)(a_k - b_k)^2) with the
// difference taken in float and the sum accumulated in FP64 in k order — bit-identical to the CPU oracle's
// restatement of cv::normL2Sqr + sqrt (oracle/matcher.c), independent of tiling.  Squares of floats are exact in
// double, so each term is one DFMA on the FP64 pipe.  Candidates are ordered by the 64-bit key
// (float bits << 32 | index): non-negative floats order like their bit patterns, so (distance asc, index asc) —
// cv::BFMatcher's order, ties to the lowest train index — is an unsigned min.
//
// Layout: a CTA owns 64 rows of d1 and sweeps 64-column tiles of d2; both are staged k-chunk by k-chunk
// (32 floats) in shared memory, transposed to [k][row] so a thread's 4 rows are one LDS.128 and its 4 columns
// are conflict-free scalar loads.  A thread keeps 4 x 4 FP64 accumulators.
#include "common.cuh"
#include "knn_keys.cuh"

namespace {

constexpr int LQ = 64;          // rows of d1 per CTA
constexpr int LT = 64;          // columns of d2 per tile
constexpr int LK = 32;          // k chunk
constexpr int LSTR = 68;        // padded row stride of the transposed staging tiles (floats)
constexpr int LTHREADS = 256;   // 16 (tx: columns) x 16 (ty: rows)

__device__ __forceinline__ unsigned long long l2_key(double s, uint32_t idx) {
    const float d = __fsqrt_rn((float)s);
    return ((unsigned long long)__float_as_uint(d) << 32) | idx;
}

__global__ void __launch_bounds__(LTHREADS, 2)
knn2_l2_exact_kernel(const float* __restrict__ d1, int n1_max, const int32_t* __restrict__ n1_arr,
                     const float* __restrict__ d2, int n2_max, const int32_t* __restrict__ n2_arr, int dim,
                     int tiles_per_chunk, unsigned long long* __restrict__ key12,
                     unsigned long long* __restrict__ key21) {
    const int prob = blockIdx.z;
    const int n1 = n1_arr ? min(n1_arr[prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[prob], n2_max) : n2_max;
    const int row0 = blockIdx.x * LQ;
    if (row0 >= n1) return;
    const int ntiles = (n2 + LT - 1) / LT;
    const int tile_begin = blockIdx.y * tiles_per_chunk;
    const int tile_end = min(ntiles, tile_begin + tiles_per_chunk);
    if (tile_begin >= tile_end) return;

    const float* __restrict__ g1 = d1 + (size_t)prob * n1_max * dim;
    const float* __restrict__ g2 = d2 + (size_t)prob * n2_max * dim;
    unsigned long long* k12 = key12 + (size_t)prob * n1_max * 2;
    unsigned long long* k21 = key21 + (size_t)prob * n2_max * 2;

    __shared__ __align__(16) float s_q[LK][LSTR];
    __shared__ __align__(16) float s_t[LK][LSTR];
    __shared__ unsigned long long s_col[LTHREADS / 32][LT][2];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int warp = tid >> 5;
    const int nkc = (dim + LK - 1) / LK;

    unsigned long long rb0[4], rb1[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { rb0[r] = KEY64_INF; rb1[r] = KEY64_INF; }

    // staging: thread -> (row = tid / 4, 8 consecutive k = (tid % 4) * 8 ..)
    const int ld_row = tid >> 2, ld_k = (tid & 3) * 8;

    for (int tile = tile_begin; tile < tile_end; tile++) {
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = 0.0;

        for (int kc = 0; kc < nkc; kc++) {
            __syncthreads();
            {
                const int qrow = row0 + ld_row, tcol = tile * LT + ld_row;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int k = kc * LK + ld_k + e;
                    const bool kok = k < dim;
                    s_q[ld_k + e][ld_row] = (kok && qrow < n1) ? __ldg(g1 + (size_t)qrow * dim + k) : 0.f;
                    s_t[ld_k + e][ld_row] = (kok && tcol < n2) ? __ldg(g2 + (size_t)tcol * dim + k) : 0.f;
                }
            }
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < LK; k++) {
                const float4 a = *reinterpret_cast<const float4*>(&s_q[k][ty * 4]);
                float b[4];
#pragma unroll
                for (int c = 0; c < 4; c++) b[c] = s_t[k][tx + 16 * c];
                const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const double d = (double)__fsub_rn(av[r], b[c]);   // difference in float, as cv::normL2Sqr
                        acc[r][c] = __fma_rn(d, d, acc[r][c]);              // d*d is exact in double: one rounding
                    }
            }
        }
        // ---- top-2 bookkeeping for this tile ------------------------------------------------------------
        unsigned long long cb0[4], cb1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int col = tile * LT + tx + 16 * c;
            const bool cok = col < n2;
            unsigned long long c0 = KEY64_INF, c1 = KEY64_INF;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int row = row0 + ty * 4 + r;
                const bool rok = row < n1;
                const unsigned long long kr = (cok && rok) ? l2_key(acc[r][c], (uint32_t)col) : KEY64_INF;
                const unsigned long long kc2 = (cok && rok) ? l2_key(acc[r][c], (uint32_t)row) : KEY64_INF;
                top2_insert(rb0[r], rb1[r], kr);
                top2_insert(c0, c1, kc2);
            }
            cb0[c] = c0; cb1[c] = c1;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, cb0[c], 16);
            const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, cb1[c], 16);
            top2_merge(cb0[c], cb1[c], o0, o1);
            if ((tid & 16) == 0) {
                s_col[warp][tx + 16 * c][0] = cb0[c];
                s_col[warp][tx + 16 * c][1] = cb1[c];
            }
        }
        __syncthreads();
        if (tid < LT) {
            unsigned long long m0 = s_col[0][tid][0], m1 = s_col[0][tid][1];
#pragma unroll
            for (int wv = 1; wv < LTHREADS / 32; wv++) top2_merge(m0, m1, s_col[wv][tid][0], s_col[wv][tid][1]);
            const int col = tile * LT + tid;
            if (col < n2) top2_publish(k21 + (size_t)col * 2, m0, m1);
        }
    }
    // ---- row results: reduce over the 16 column groups, publish ------------------------------------------
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, rb0[r], off);
            const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, rb1[r], off);
            top2_merge(rb0[r], rb1[r], o0, o1);
        }
        const int row = row0 + ty * 4 + r;
        if (tx == 0 && row < n1) top2_publish(k12 + (size_t)row * 2, rb0[r], rb1[r]);
    }
}

__global__ void knn_unpack64_kernel(const unsigned long long* __restrict__ keys, int n_max,
                                    const int32_t* __restrict__ n_arr, int count, int32_t* __restrict__ idx,
                                    float* __restrict__ dist) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)count * n_max * 2;
    if (i >= total) return;
    const int prob = (int)(i / ((size_t)n_max * 2));
    const int row = (int)((i / 2) % n_max);
    const int n = n_arr ? min(n_arr[prob], n_max) : n_max;
    const unsigned long long k = keys[i];
    if (row >= n || k == KEY64_INF) { idx[i] = -1; dist[i] = 0.f; return; }
    idx[i] = (int32_t)(uint32_t)(k & 0xFFFFFFFFull);
    dist[i] = __uint_as_float((uint32_t)(k >> 32));
}

}  // namespace

int vsb_knn2_l2_tc(vsb_ctx* ctx, const float* d1, int n1_max, const int32_t* n1, const float* d2, int n2_max,
                   const int32_t* n2, int dim, int count, unsigned long long* key12, unsigned long long* key21,
                   cudaStream_t st);

// Internal entry (tracker too): leaves packed 64-bit keys (float bits << 32 | index) in key12 / key21.
int vsb_knn2_l2_keys(vsb_ctx* ctx, const float* d1, int n1_max, const int32_t* n1, const float* d2, int n2_max,
                     const int32_t* n2, int dim, int count, unsigned long long* key12, unsigned long long* key21,
                     cudaStream_t st) {
    if (!ctx || count < 0 || n1_max < 0 || n2_max < 0 || dim <= 0) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    if (n2_max > 0) VSB_CUDA(ctx, cudaMemsetAsync(key21, 0xFF, (size_t)count * n2_max * 2 * sizeof(unsigned long long), st));
    if (n1_max > 0) VSB_CUDA(ctx, cudaMemsetAsync(key12, 0xFF, (size_t)count * n1_max * 2 * sizeof(unsigned long long), st));
    if (n1_max == 0 || n2_max == 0) return VSB_OK;
    if (!d1 || !d2) return VSB_ERR_INVALID;
    if (ctx->knn_l2_impl == 1 && dim <= 64 && (dim & 7) == 0 && ((((uintptr_t)d1 | (uintptr_t)d2) & 15) == 0))
        return vsb_knn2_l2_tc(ctx, d1, n1_max, n1, d2, n2_max, n2, dim, count, key12, key21, st);
    const int row_tiles = vsb_div_up(n1_max, LQ);
    const int col_tiles = vsb_div_up(n2_max, LT);
    int chunks = 1;
    const long long want = 2LL * ctx->sm_count * 2;
    while ((long long)row_tiles * chunks * count < want && chunks < col_tiles) chunks *= 2;
    if (chunks > col_tiles) chunks = col_tiles;
    const int tiles_per_chunk = vsb_div_up(col_tiles, chunks);
    chunks = vsb_div_up(col_tiles, tiles_per_chunk);
    for (int z0 = 0; z0 < count; z0 += 65535) {
        const int zc = count - z0 < 65535 ? count - z0 : 65535;
        dim3 grid(row_tiles, chunks, zc);
        ProfScope ps(ctx, VSB_K_KNN_L2, st);
        knn2_l2_exact_kernel<<<grid, LTHREADS, 0, st>>>(
            d1 + (size_t)z0 * n1_max * dim, n1_max, n1 ? n1 + z0 : nullptr,
            d2 + (size_t)z0 * n2_max * dim, n2_max, n2 ? n2 + z0 : nullptr, dim, tiles_per_chunk,
            key12 + (size_t)z0 * n1_max * 2, key21 + (size_t)z0 * n2_max * 2);
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

int vsb_knn_unpack64(vsb_ctx* ctx, const unsigned long long* keys, int n_max, const int32_t* n, int count,
                     int32_t* idx, float* dist, cudaStream_t st) {
    const size_t total = (size_t)count * n_max * 2;
    if (total == 0) return VSB_OK;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    ProfScope ps(ctx, VSB_K_KNN_UNPACK, st);
    knn_unpack64_kernel<<<blocks, 256, 0, st>>>(keys, n_max, n, count, idx, dist);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_knn2_l2(vsb_ctx_t* ctx, const float* d1, int n1_max, const int32_t* n1, const float* d2,
                           int n2_max, const int32_t* n2, int dim, int count, int32_t* idx12, float* dist12,
                           int32_t* idx21, float* dist21, void* stream) {
    if (!ctx) return VSB_ERR_INVALID;
    if (count < 0 || n1_max < 0 || n2_max < 0 || dim <= 0) return VSB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nk = (size_t)count * ((size_t)n1_max + n2_max) * 2;
    void* scratch = nullptr;
    int rc = vsb_scratch_reserve(ctx, nk * sizeof(unsigned long long) + 256, &scratch);
    if (rc) return rc;
    unsigned long long* key12 = (unsigned long long*)scratch;
    unsigned long long* key21 = key12 + (size_t)count * n1_max * 2;
    rc = vsb_knn2_l2_keys(ctx, d1, n1_max, n1, d2, n2_max, n2, dim, count, key12, key21, st);
    if (rc) return rc;
    rc = vsb_knn_unpack64(ctx, key12, n1_max, n1, count, idx12, dist12, st);
    if (rc) return rc;
    return vsb_knn_unpack64(ctx, key21, n2_max, n2, count, idx21, dist21, st);
}

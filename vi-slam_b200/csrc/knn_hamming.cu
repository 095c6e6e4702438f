// knn_hamming.cu — brute-force kNN (k = 2) over 256-bit binary descriptors, both directions from ONE
// pass over the distance matrix.  Replaces the two BFMatcher(NORM_HAMMING)::knnMatch(...,2) calls of
// Matcher::computeMatches (reference src/Matcher.cpp:83-94) / MatcherGPU::computeGPUMatches
// (src/MatcherGPU.cpp:44-66).
//
// Layout: one CTA owns a 64-row tile of d1 and sweeps a chunk of 128-column tiles of d2 staged in shared
// memory as two 16-byte planes (conflict-free LDS.128).  Each thread keeps 4 rows of d1 in registers
// (8 x u32 each) and evaluates 4 x 8 distances per tile: XOR on the ALU pipe, __popc on the integer pipe.
// Candidates are ordered by the packed key (distance << 23 | index), so "distance ascending, then lowest
// index" — cv::BFMatcher's order — is a plain unsigned min.  Row results live in registers until the sweep
// ends; column results are merged per tile (warp shuffle, shared memory) and published with the
// two-atomicMin top-2 protocol on a global key array.
#include "common.cuh"
#include "knn_keys.cuh"

namespace {

constexpr int TQ = 64;        // rows of d1 per CTA
constexpr int TT = 128;       // columns of d2 per tile
constexpr int RQ = 4;         // rows per thread
constexpr int RT = 8;         // columns per thread
constexpr int NTHREADS = 256; // 16 (tx) x 16 (ty)

__device__ __forceinline__ int popc256(const uint32_t (&a)[8], const uint4& b0, const uint4& b1) {
    int s0 = __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y);
    int s1 = __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w);
    int s2 = __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y);
    int s3 = __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
    return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(NTHREADS, 2)
knn2_hamming_kernel(const uint8_t* __restrict__ d1, int n1_max, const int32_t* __restrict__ n1_arr,
                    const uint8_t* __restrict__ d2, int n2_max, const int32_t* __restrict__ n2_arr,
                    int tiles_per_chunk, uint32_t* __restrict__ key12, uint32_t* __restrict__ key21) {
    const int prob = blockIdx.z;
    const int n1 = n1_arr ? min(n1_arr[prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[prob], n2_max) : n2_max;
    const int row0 = blockIdx.x * TQ;
    if (row0 >= n1) return;
    const int ntiles = (n2 + TT - 1) / TT;
    const int tile_begin = blockIdx.y * tiles_per_chunk;
    const int tile_end = min(ntiles, tile_begin + tiles_per_chunk);
    if (tile_begin >= tile_end) return;

    const uint4* __restrict__ g1 = reinterpret_cast<const uint4*>(d1 + (size_t)prob * n1_max * 32);
    const uint4* __restrict__ g2 = reinterpret_cast<const uint4*>(d2 + (size_t)prob * n2_max * 32);
    uint32_t* k12 = key12 + (size_t)prob * n1_max * 2;
    uint32_t* k21 = key21 + (size_t)prob * n2_max * 2;

    __shared__ uint4 s_lo[2][TT];              // first 16 bytes of each staged d2 descriptor
    __shared__ uint4 s_hi[2][TT];              // last 16 bytes
    __shared__ uint32_t s_col[NTHREADS / 32][TT][2];

    const int tid = threadIdx.x;
    const int tx = tid & 15;   // column group
    const int ty = tid >> 4;   // row group
    const int warp = tid >> 5;

    // ---- rows of d1 into registers -------------------------------------------------------------
    uint32_t q[RQ][8];
    uint32_t rowc[RQ];         // row index, or all-ones when the row does not exist
#pragma unroll
    for (int r = 0; r < RQ; r++) {
        int row = row0 + ty * RQ + r;
        bool ok = row < n1;
        uint4 a = ok ? __ldg(g1 + (size_t)row * 2) : make_uint4(0, 0, 0, 0);
        uint4 b = ok ? __ldg(g1 + (size_t)row * 2 + 1) : make_uint4(0, 0, 0, 0);
        q[r][0] = a.x; q[r][1] = a.y; q[r][2] = a.z; q[r][3] = a.w;
        q[r][4] = b.x; q[r][5] = b.y; q[r][6] = b.z; q[r][7] = b.w;
        rowc[r] = ok ? (uint32_t)row : KEY_INF;
    }
    uint32_t rb0[RQ], rb1[RQ];
#pragma unroll
    for (int r = 0; r < RQ; r++) { rb0[r] = KEY_INF; rb1[r] = KEY_INF; }

    // ---- stage the first tile ------------------------------------------------------------------
    auto load_tile = [&](int tile, uint4& v) {
        int col = tile * TT + (tid >> 1);
        v = (col < n2) ? __ldg(g2 + (size_t)col * 2 + (tid & 1)) : make_uint4(0, 0, 0, 0);
    };
    auto store_tile = [&](int buf, const uint4& v) {
        if (tid & 1) s_hi[buf][tid >> 1] = v; else s_lo[buf][tid >> 1] = v;
    };
    uint4 stage;
    load_tile(tile_begin, stage);
    store_tile(0, stage);
    __syncthreads();

    for (int tile = tile_begin; tile < tile_end; tile++) {
        const int buf = (tile - tile_begin) & 1;
        if (tile + 1 < tile_end) load_tile(tile + 1, stage);   // prefetch into registers

        uint32_t cb0[RT], cb1[RT];
#pragma unroll
        for (int c = 0; c < RT; c++) {
            const int lc = tx + 16 * c;
            const int col = tile * TT + lc;
            const uint32_t colc = (col < n2) ? (uint32_t)col : KEY_INF;
            const uint4 b0 = s_lo[buf][lc];
            const uint4 b1 = s_hi[buf][lc];
            uint32_t c0 = KEY_INF, c1 = KEY_INF;
#pragma unroll
            for (int r = 0; r < RQ; r++) {
                const uint32_t d = (uint32_t)popc256(q[r], b0, b1) << KEY_SHIFT;
                const uint32_t kr = d | colc;     // candidate for row r  (all-ones when the column is padding)
                const uint32_t kc = d | rowc[r];  // candidate for column c
                top2_insert(rb0[r], rb1[r], kr);
                top2_insert(c0, c1, kc);
            }
            cb0[c] = c0; cb1[c] = c1;
        }
        // ---- column results: the two ty-groups of a warp, then the 8 warps via shared memory ----
#pragma unroll
        for (int c = 0; c < RT; c++) {
            uint32_t o0 = __shfl_xor_sync(0xffffffffu, cb0[c], 16);
            uint32_t o1 = __shfl_xor_sync(0xffffffffu, cb1[c], 16);
            top2_merge(cb0[c], cb1[c], o0, o1);
            if ((tid & 16) == 0) {
                s_col[warp][tx + 16 * c][0] = cb0[c];
                s_col[warp][tx + 16 * c][1] = cb1[c];
            }
        }
        __syncthreads();
        if (tid < TT) {
            uint32_t m0 = s_col[0][tid][0], m1 = s_col[0][tid][1];
#pragma unroll
            for (int wv = 1; wv < NTHREADS / 32; wv++) top2_merge(m0, m1, s_col[wv][tid][0], s_col[wv][tid][1]);
            const int col = tile * TT + tid;
            if (col < n2) top2_publish(k21 + (size_t)col * 2, m0, m1);
        }
        if (tile + 1 < tile_end) store_tile(buf ^ 1, stage);
        __syncthreads();
    }

    // ---- row results: reduce over the 16 column groups (lanes sharing ty), publish ------------------
#pragma unroll
    for (int r = 0; r < RQ; r++) {
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            uint32_t o0 = __shfl_xor_sync(0xffffffffu, rb0[r], off);
            uint32_t o1 = __shfl_xor_sync(0xffffffffu, rb1[r], off);
            top2_merge(rb0[r], rb1[r], o0, o1);
        }
        const int row = row0 + ty * RQ + r;
        if (tx == 0 && row < n1) {
            if (gridDim.y == 1) {   // this CTA saw every column: plain stores
                k12[(size_t)row * 2] = rb0[r];
                k12[(size_t)row * 2 + 1] = rb1[r];
            } else {
                top2_publish(k12 + (size_t)row * 2, rb0[r], rb1[r]);
            }
        }
    }
}

__global__ void knn_unpack_kernel(const uint32_t* __restrict__ keys, int n_max, const int32_t* __restrict__ n_arr,
                                  int count, int32_t* __restrict__ idx, float* __restrict__ dist, int is_float_key) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)count * n_max * 2;
    if (i >= total) return;
    int prob = (int)(i / ((size_t)n_max * 2));
    int row = (int)((i / 2) % n_max);
    int n = n_arr ? min(n_arr[prob], n_max) : n_max;
    uint32_t k = keys[i];
    if (row >= n || k == KEY_INF) { idx[i] = -1; dist[i] = 0.f; return; }
    idx[i] = (int32_t)(k & KEY_IDX_MASK);
    dist[i] = (float)(k >> KEY_SHIFT);
}

}  // namespace

int vsb_knn2_hamming_tc(vsb_ctx* ctx, const uint8_t* d1, int n1_max, const int32_t* n1, const uint8_t* d2,
                        int n2_max, const int32_t* n2, int count, uint32_t* key12, uint32_t* key21, int pack16,
                        int32_t* dump, int dump_ld, cudaStream_t st);

int vsb_knn2_hamming_mx(vsb_ctx* ctx, const uint8_t* d1, int n1_max, const int32_t* n1, const uint8_t* d2, int n2_max,
                        const int32_t* n2, int count, uint32_t* key12, uint32_t* key21, int pre, cudaStream_t st);

// Internal entry used by the tracker too: leaves packed keys (distance << 23 | index) in key12 / key21.
int vsb_knn2_hamming_keys(vsb_ctx* ctx, const uint8_t* d1, int n1_max, const int32_t* n1, const uint8_t* d2,
                          int n2_max, const int32_t* n2, int count, uint32_t* key12, uint32_t* key21,
                          cudaStream_t st) {
    if (!ctx || count < 0 || n1_max < 0 || n2_max < 0) return VSB_ERR_INVALID;
    if (n1_max > (int)KEY_IDX_MASK || n2_max > (int)KEY_IDX_MASK) return VSB_ERR_CAPACITY;
    if (count == 0) return VSB_OK;
    int impl = ctx->knn_impl;
    const int n_big = n1_max > n2_max ? n1_max : n2_max;
    if (impl == 6) impl = (n_big >= 768 && n_big <= 24576) ? 5 : 2;      // auto: the 4-bit persistent kernel pays from ~768 rows (its fold holds the tile index in a byte: <= 256 tiles of 96)
    if (impl == 5 && n_big > 24576) return VSB_ERR_CAPACITY;
    if (impl >= 3) return vsb_knn2_hamming_mx(ctx, d1, n1_max, n1, d2, n2_max, n2, count, key12, key21, impl - 3, st);
    if (impl != 0)   // tensor-core path: every valid row is written by exactly one CTA, no memset, no atomics
        return vsb_knn2_hamming_tc(ctx, d1, n1_max, n1, d2, n2_max, n2, count, key12, key21, impl == 2,
                                   nullptr, 0, st);
    if (n2_max > 0)
        VSB_CUDA(ctx, cudaMemsetAsync(key21, 0xFF, (size_t)count * n2_max * 2 * sizeof(uint32_t), st));
    if (n1_max > 0)
        VSB_CUDA(ctx, cudaMemsetAsync(key12, 0xFF, (size_t)count * n1_max * 2 * sizeof(uint32_t), st));
    if (n1_max == 0 || n2_max == 0) return VSB_OK;
    const int row_tiles = vsb_div_up(n1_max, TQ);
    const int col_tiles = vsb_div_up(n2_max, TT);
    // split the column sweep only when the batch alone cannot fill the machine (>= 2 waves of CTAs)
    int chunks = 1;
    const long long want = 2LL * ctx->sm_count * 4;
    while ((long long)row_tiles * chunks * count < want && chunks < col_tiles) chunks *= 2;
    if (chunks > col_tiles) chunks = col_tiles;
    const int tiles_per_chunk = vsb_div_up(col_tiles, chunks);
    chunks = vsb_div_up(col_tiles, tiles_per_chunk);
    for (int z0 = 0; z0 < count; z0 += 65535) {
        int zc = count - z0 < 65535 ? count - z0 : 65535;
        dim3 grid(row_tiles, chunks, zc);
        ProfScope ps(ctx, VSB_K_KNN_HAMMING, st);
        knn2_hamming_kernel<<<grid, NTHREADS, 0, st>>>(
            d1 + (size_t)z0 * n1_max * 32, n1_max, n1 ? n1 + z0 : nullptr,
            d2 + (size_t)z0 * n2_max * 32, n2_max, n2 ? n2 + z0 : nullptr, tiles_per_chunk,
            key12 + (size_t)z0 * n1_max * 2, key21 + (size_t)z0 * n2_max * 2);
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

int vsb_knn_unpack(vsb_ctx* ctx, const uint32_t* keys, int n_max, const int32_t* n, int count, int32_t* idx,
                   float* dist, cudaStream_t st) {
    size_t total = (size_t)count * n_max * 2;
    if (total == 0) return VSB_OK;
    unsigned blocks = (unsigned)((total + 255) / 256);
    ProfScope ps(ctx, VSB_K_KNN_UNPACK, st);
    knn_unpack_kernel<<<blocks, 256, 0, st>>>(keys, n_max, n, count, idx, dist, 0);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_knn2_hamming(vsb_ctx_t* ctx, const uint8_t* d1, int n1_max, const int32_t* n1,
                                const uint8_t* d2, int n2_max, const int32_t* n2, int count,
                                int32_t* idx12, float* dist12, int32_t* idx21, float* dist21, void* stream) {
    if (!ctx) return VSB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    size_t nk = (size_t)count * ((size_t)n1_max + n2_max) * 2;
    void* scratch = nullptr;
    int rc = vsb_scratch_reserve(ctx, nk * sizeof(uint32_t) + 256, &scratch);
    if (rc) return rc;
    uint32_t* key12 = (uint32_t*)scratch;
    uint32_t* key21 = key12 + (size_t)count * n1_max * 2;
    rc = vsb_knn2_hamming_keys(ctx, d1, n1_max, n1, d2, n2_max, n2, count, key12, key21, st);
    if (rc) return rc;
    rc = vsb_knn_unpack(ctx, key12, n1_max, n1, count, idx12, dist12, st);
    if (rc) return rc;
    return vsb_knn_unpack(ctx, key21, n2_max, n2, count, idx21, dist21, st);
}

// ---- INT-pipe ceiling: independent POPC chains on every SM (the denominator of the matcher's roofline) -----
namespace {
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t seed, int iters, uint32_t* out) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = seed * (threadIdx.x + 1) + k * 0x9E3779B9u + blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = __popc(v[k]) + seed;   // 1 POPC + 1 IADD per link
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= v[k];
    if (acc == 0xDEADBEEFu) out[0] = acc;   // keeps the chains alive
}
}  // namespace

extern "C" int vsb_popc_peak(vsb_ctx_t* ctx, double* popc_per_s, void* stream) {
    if (!ctx || !popc_per_s) return VSB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    void* scratch = nullptr;
    int rc = vsb_scratch_reserve(ctx, 256, &scratch);
    if (rc) return rc;
    const int blocks = ctx->sm_count * 8, iters = 8192;
    cudaEvent_t a, b;
    VSB_CUDA(ctx, cudaEventCreate(&a));
    VSB_CUDA(ctx, cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        VSB_CUDA(ctx, cudaEventRecord(a, st));
        popc_peak_kernel<<<blocks, 256, 0, st>>>(0x12345u + rep, iters, (uint32_t*)scratch);
        VSB_LAUNCHED(ctx);
        VSB_CUDA(ctx, cudaEventRecord(b, st));
        VSB_CUDA(ctx, cudaEventSynchronize(b));
        float ms = 0.f;
        VSB_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *popc_per_s = (double)blocks * 256.0 * iters * 8.0 / (best * 1e-3);
    return VSB_OK;
}

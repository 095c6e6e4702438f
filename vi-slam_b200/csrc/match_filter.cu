// match_filter.cu — ratio test, symmetry test, sort-by-y and grid-best selection on the device, one CTA per
// frame pair.  Replaces Matcher::computeBestMatches + getGoodMatches (reference src/Matcher.cpp:353-367,
// 96-169, 171-244, 295-303, 329-352) so descriptors -> GN never returns to the host.
//
// The reference's sequential band sweep over y-sorted matches is order-equivalent to: every symmetric match
// falls into cell (band, column) given by the float-accumulated band/column edges, and each cell keeps the
// match that is smallest under (distance, y, queryIdx) — strict '<' in the sweep keeps the first of equal
// distances, the sweep order is y ascending with ties in match (= queryIdx) order.  Cells are emitted
// band-major, column-minor.  That form needs no sort and is bit-identical to the sweep.
#include "common.cuh"
#include "knn_keys.cuh"

namespace {

constexpr int FT = 256;
constexpr int MAX_ROOT = 64;  // floor(sqrt(n_cells)) <= 64

__device__ __forceinline__ uint32_t float_order_bits(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// kNN results as the filter reads them.  KEYS == 0: the public form, idx / dist arrays (-1 = missing neighbour).
// KEYS == 1 / 2: the kNN kernels' packed keys read in place — 32-bit (hamming << 23 | index) or 64-bit
// (float bits << 32 | index), all-ones = missing — which is what knn_unpack would have expanded (tracker path).
template <int KEYS>
struct KnnView {
    const int32_t* idx; const float* dist; const void* keys;
    __device__ __forceinline__ int index(int e) const {
        if (KEYS == 0) return idx[e];
        if (KEYS == 1) { const uint32_t k = reinterpret_cast<const uint32_t*>(keys)[e]; return k == KEY_INF ? -1 : (int)(k & KEY_IDX_MASK); }
        const unsigned long long k = reinterpret_cast<const unsigned long long*>(keys)[e];
        return k == KEY64_INF ? -1 : (int)(uint32_t)(k & 0xFFFFFFFFull);
    }
    __device__ __forceinline__ float distance(int e) const {
        if (KEYS == 0) return dist[e];
        if (KEYS == 1) { const uint32_t k = reinterpret_cast<const uint32_t*>(keys)[e]; return k == KEY_INF ? 0.f : (float)(k >> KEY_SHIFT); }
        const unsigned long long k = reinterpret_cast<const unsigned long long*>(keys)[e];
        return k == KEY64_INF ? 0.f : __uint_as_float((uint32_t)(k >> 32));
    }
};

template <int KEYS>
__global__ void __launch_bounds__(FT)
match_filter_kernel(const int32_t* __restrict__ idx12, const float* __restrict__ dist12, const void* __restrict__ keys12,
                    int n1_max, const int32_t* __restrict__ n1_arr, const int32_t* __restrict__ idx21,
                    const float* __restrict__ dist21, const void* __restrict__ keys21, int n2_max,
                    const int32_t* __restrict__ n2_arr,
                    const float* __restrict__ kp1_xy, int w, int h, int n_cells, float ratio_f, int sym_mode,
                    int32_t* __restrict__ good_q, int32_t* __restrict__ good_t, float* __restrict__ good_d,
                    int good_cap, int32_t* __restrict__ n_good, int32_t* __restrict__ n_sym,
                    float* __restrict__ good_xy) {
    extern __shared__ unsigned char smem_raw[];
    const int prob = blockIdx.x;
    const int tid = threadIdx.x;
    const int n1 = n1_arr ? min(n1_arr[prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[prob], n2_max) : n2_max;
    const size_t o12 = (size_t)prob * n1_max * 2, o21 = (size_t)prob * n2_max * 2;
    const size_t kb = KEYS == 2 ? 8 : 4;
    const KnnView<KEYS> v12 = {KEYS == 0 ? idx12 + o12 : nullptr, KEYS == 0 ? dist12 + o12 : nullptr,
                               KEYS ? static_cast<const unsigned char*>(keys12) + o12 * kb : nullptr};
    const KnnView<KEYS> v21 = {KEYS == 0 ? idx21 + o21 : nullptr, KEYS == 0 ? dist21 + o21 : nullptr,
                               KEYS ? static_cast<const unsigned char*>(keys21) + o21 * kb : nullptr};
    const float* kp = kp1_xy + (size_t)prob * n1_max * 2;

    const int root = (int)floor(sqrt((double)n_cells));         // Matcher.cpp:191
    const int ncell = root * root;
    int16_t* s_cell = reinterpret_cast<int16_t*>(smem_raw);     // [n1_max] cell of row i, -1 = no symmetric match
    uint32_t* s_best = reinterpret_cast<uint32_t*>(smem_raw + (((size_t)n1_max * 2 + 15) & ~(size_t)15)); // [ncell]
    __shared__ float s_hf[MAX_ROOT], s_wf[MAX_ROOT];
    __shared__ int s_nsym, s_base;
    __shared__ int s_warp_cnt[FT / 32];

    if (tid == 0) {
        s_nsym = 0;
        s_base = 0;
        float winW = (float)((double)w / floor(sqrt((double)n_cells)));   // Matcher.cpp:177
        float winH = (float)((double)h / floor(sqrt((double)n_cells)));   // Matcher.cpp:178
        float hf = winH, wf = winW;
        for (int k = 0; k < root; k++) {
            s_hf[k] = hf; s_wf[k] = wf;
            hf = __fadd_rn(hf, winH);                                     // Matcher.cpp:236
            wf = __fadd_rn(wf, winW);                                     // Matcher.cpp:213
        }
    }
    for (int c = tid; c < ncell; c += FT) s_best[c] = 0xFFFFFFFFu;
    __syncthreads();

    const double ratio = (double)ratio_f;                                  // Matcher.cpp:103
    // ---- pass 0: ratio + symmetry, cell of every surviving row, per-cell min distance -----------------
    int my_sym = 0;
    for (int i = tid; i < n1; i += FT) {
        int cell = -1;
        // rows past the set's true size are "missing" in either form (knn_unpack wrote -1 there)
        const int j0 = v12.index(2 * i), j1 = v12.index(2 * i + 1);
        const float dd0 = v12.distance(2 * i), dd1 = v12.distance(2 * i + 1);
        bool keep = (j0 >= 0 && j1 >= 0) && !((double)dd0 > ratio * (double)dd1);   // Matcher.cpp:153-166
        if (keep && j0 < n2) {
            const int b0 = v21.index(2 * j0), b1 = v21.index(2 * j0 + 1);
            bool ok = (b0 >= 0);
            if (sym_mode == 1)   // intended: the 2->1 row must survive its own ratio test
                ok = ok && (b1 >= 0) && !((double)v21.distance(2 * j0) > ratio * (double)v21.distance(2 * j0 + 1));
            if (ok && b0 == i) { // Matcher.cpp:124-125
                my_sym++;
                const float x = kp[2 * i], y = kp[2 * i + 1];
                int band = -1;
                for (int k = 0; k < root; k++) if (y <= s_hf[k]) { band = k; break; }   // Matcher.cpp:205
                if (band >= 0) {
                    int col = 0;
                    while (col < root && x > s_wf[col]) col++;                          // Matcher.cpp:211-216
                    if (col >= root) col = root - 1;                                    // App. B-13
                    cell = band * root + col;
                    if (dd0 < 100000.0f) atomicMin(&s_best[cell], __float_as_uint(dd0)); // Matcher.cpp:218
                    else cell = -1;
                }
            }
        }
        s_cell[i] = (int16_t)cell;
    }
    if (my_sym) atomicAdd(&s_nsym, my_sym);
    __syncthreads();
    // ---- pass 1/2: among rows at the cell's min distance keep the smallest y, then the smallest index --
    // (s_best is reused: distance bits -> y order bits -> row index, each pass filtering on the previous)
    uint32_t* s_y = s_best + ncell;     // [ncell]
    uint32_t* s_i = s_y + ncell;        // [ncell]
    for (int c = tid; c < ncell; c += FT) { s_y[c] = 0xFFFFFFFFu; s_i[c] = 0xFFFFFFFFu; }
    __syncthreads();
    for (int i = tid; i < n1; i += FT) {
        int cell = s_cell[i];
        if (cell >= 0 && __float_as_uint(v12.distance(2 * i)) == s_best[cell])
            atomicMin(&s_y[cell], float_order_bits(kp[2 * i + 1]));
    }
    __syncthreads();
    for (int i = tid; i < n1; i += FT) {
        int cell = s_cell[i];
        if (cell >= 0 && __float_as_uint(v12.distance(2 * i)) == s_best[cell] &&
            float_order_bits(kp[2 * i + 1]) == s_y[cell])
            atomicMin(&s_i[cell], (uint32_t)i);
    }
    __syncthreads();
    // ---- emit non-empty cells band-major / column-minor (Matcher.cpp:232-233, 318-326) ---------------
    int32_t* gq = good_q + (size_t)prob * good_cap;
    int32_t* gt = good_t + (size_t)prob * good_cap;
    float* gd = good_d + (size_t)prob * good_cap;
    for (int c0 = 0; c0 < ncell; c0 += FT) {
        const int c = c0 + tid;
        const bool full = (c < ncell) && (s_i[c] != 0xFFFFFFFFu);
        const unsigned ballot = __ballot_sync(0xffffffffu, full);
        const int lane = tid & 31, wid = tid >> 5;
        if (lane == 0) s_warp_cnt[wid] = __popc(ballot);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < wid; k++) off += s_warp_cnt[k];
        off += __popc(ballot & ((1u << lane) - 1u));
        if (full && off < good_cap) {
            const int i = (int)s_i[c];
            gq[off] = i;
            gt[off] = v12.index(2 * i);
            gd[off] = v12.distance(2 * i);
            if (good_xy) {                                   // Matcher::getGoodMatches for the previous frame (:295-303)
                good_xy[((size_t)prob * good_cap + off) * 2] = kp[2 * i];
                good_xy[((size_t)prob * good_cap + off) * 2 + 1] = kp[2 * i + 1];
            }
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int k = 0; k < FT / 32; k++) tot += s_warp_cnt[k];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) {
        n_good[prob] = min(s_base, good_cap);
        if (n_sym) n_sym[prob] = s_nsym;
    }
}

__global__ void gather_keypoints_kernel(const float2* __restrict__ kp, int n_max, const int32_t* __restrict__ good,
                                        int good_cap, const int32_t* __restrict__ n_good, float2* __restrict__ out) {
    const int prob = blockIdx.x;
    const int m = blockIdx.y * blockDim.x + threadIdx.x;
    if (m >= good_cap) return;
    const int n = min(n_good[prob], good_cap);
    float2 v = make_float2(0.f, 0.f);
    if (m < n) v = kp[(size_t)prob * n_max + good[(size_t)prob * good_cap + m]];
    out[(size_t)prob * good_cap + m] = v;
}

}  // namespace

static int match_filter_launch(vsb_ctx_t* ctx, int key_mode, const int32_t* idx12, const float* dist12, const void* keys12,
                               int n1_max, const int32_t* n1, const int32_t* idx21, const float* dist21,
                               const void* keys21, int n2_max, const int32_t* n2, const float* kp1_xy, int count, int w,
                               int h, int n_cells, float ratio, int sym_mode, int32_t* good_q, int32_t* good_t,
                               float* good_d, int good_cap, int32_t* n_good, int32_t* n_sym, float* good_xy, void* stream) {
    if (!ctx || count < 0 || n_cells < 1 || n1_max < 0 || n2_max < 0 || !n_good) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    const int root = (int)floor(sqrt((double)n_cells));
    if (root > MAX_ROOT || n1_max > 32767 * 2) return VSB_ERR_CAPACITY;
    if (root * root > 32767) return VSB_ERR_CAPACITY;
    if (good_cap < root * root) return VSB_ERR_INVALID;
    size_t smem = (((size_t)n1_max * 2 + 15) & ~(size_t)15) + (size_t)root * root * 3 * sizeof(uint32_t);
    if (smem > 200 * 1024) return VSB_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 48 * 1024) {
        VSB_CUDA(ctx, cudaFuncSetAttribute(match_filter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VSB_CUDA(ctx, cudaFuncSetAttribute(match_filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VSB_CUDA(ctx, cudaFuncSetAttribute(match_filter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    ProfScope ps(ctx, VSB_K_MATCH_FILTER, st);
#define MF_ARGS idx12, dist12, keys12, n1_max, n1, idx21, dist21, keys21, n2_max, n2, kp1_xy, w, h, n_cells, ratio, sym_mode, \
                good_q, good_t, good_d, good_cap, n_good, n_sym, good_xy
    if (key_mode == 1) match_filter_kernel<1><<<count, FT, smem, st>>>(MF_ARGS);
    else if (key_mode == 2) match_filter_kernel<2><<<count, FT, smem, st>>>(MF_ARGS);
    else match_filter_kernel<0><<<count, FT, smem, st>>>(MF_ARGS);
#undef MF_ARGS
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_match_filter(vsb_ctx_t* ctx, const int32_t* idx12, const float* dist12, int n1_max,
                                const int32_t* n1, const int32_t* idx21, const float* dist21, int n2_max,
                                const int32_t* n2, const float* kp1_xy, int count, int w, int h, int n_cells,
                                float ratio, int sym_mode, int32_t* good_q, int32_t* good_t, float* good_d,
                                int good_cap, int32_t* n_good, int32_t* n_sym, void* stream) {
    return match_filter_launch(ctx, 0, idx12, dist12, nullptr, n1_max, n1, idx21, dist21, nullptr, n2_max, n2, kp1_xy, count,
                               w, h, n_cells, ratio, sym_mode, good_q, good_t, good_d, good_cap, n_good, n_sym, nullptr,
                               stream);
}

// Internal entry (tracker): reads the kNN kernels' packed keys in place (key_bytes 4: Hamming, 8: float) and also
// gathers the key points of the good matches (Matcher::getGoodMatches), so neither knn_unpack nor the gather runs.
int vsb_match_filter_keys(vsb_ctx_t* ctx, const void* keys12, const void* keys21, int key_bytes, int n1_max,
                          const int32_t* n1, int n2_max, const int32_t* n2, const float* kp1_xy, int count, int w, int h,
                          int n_cells, float ratio, int sym_mode, int32_t* good_q, int32_t* good_t, float* good_d,
                          int good_cap, int32_t* n_good, int32_t* n_sym, float* good_xy, void* stream) {
    if (!keys12 || !keys21 || (key_bytes != 4 && key_bytes != 8)) return VSB_ERR_INVALID;
    return match_filter_launch(ctx, key_bytes == 4 ? 1 : 2, nullptr, nullptr, keys12, n1_max, n1, nullptr, nullptr, keys21,
                               n2_max, n2, kp1_xy, count, w, h, n_cells, ratio, sym_mode, good_q, good_t, good_d, good_cap,
                               n_good, n_sym, good_xy, stream);
}

extern "C" int vsb_gather_keypoints(vsb_ctx_t* ctx, const float* kp_xy, int n_max, const int32_t* good_idx,
                                    int good_cap, const int32_t* n_good, int count, float* out_xy, void* stream) {
    if (!ctx || count < 0 || good_cap < 0) return VSB_ERR_INVALID;
    if (count == 0 || good_cap == 0) return VSB_OK;
    dim3 grid(count, vsb_div_up(good_cap, 128));
    ProfScope ps(ctx, VSB_K_GATHER, (cudaStream_t)stream);
    gather_keypoints_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(kp_xy), n_max, good_idx, good_cap, n_good, reinterpret_cast<float2*>(out_xy));
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

// =====================================================================================================
// Per-stage entries: the reference's Matcher exposes nnFilter / computeSymMatches / sortMatches /
// bestMatchesFilter as separate public methods over public vectors (Matcher.hpp:34-47), so the class mirror in
// vi-slam_b200/host/ needs each stage on its own.  One CTA per frame pair; all O(N) except the stable rank sort.
// =====================================================================================================
namespace {

// Matcher::nnFilter (Matcher.cpp:148-169): keep[i] = 0 when the row would be cleared.
__global__ void __launch_bounds__(FT)
nn_filter_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist, int n_max,
                 const int32_t* __restrict__ n_arr, double ratio, uint8_t* __restrict__ keep) {
    const int prob = blockIdx.y;
    const int i = blockIdx.x * FT + threadIdx.x;
    if (i >= n_max) return;
    const int n = n_arr ? min(n_arr[prob], n_max) : n_max;
    const size_t o = (size_t)prob * n_max + i;
    uint8_t k = 0;
    if (i < n) {
        const int j0 = idx[2 * o], j1 = idx[2 * o + 1];
        if (j0 >= 0 && j1 >= 0) k = !((double)dist[2 * o] > ratio * (double)dist[2 * o + 1]);   // :156
    }
    keep[o] = k;
}

// Ordered compaction of a per-thread flag over consecutive chunks of FT items (query-ascending output order).
__device__ __forceinline__ int block_ordered_slot(bool flag, int* s_warp_cnt, int* s_base) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp_cnt[wid] = __popc(ballot);
    __syncthreads();
    int off = *s_base;
    for (int k = 0; k < wid; k++) off += s_warp_cnt[k];
    off += __popc(ballot & ((1u << lane) - 1u));
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int k = 0; k < FT / 32; k++) tot += s_warp_cnt[k];
        *s_base += tot;
    }
    __syncthreads();
    return off;
}

// Matcher::computeSymMatches after its two nnFilter calls (Matcher.cpp:113-143): rows of aux_matches1 that
// survived (keep12) and whose best neighbour j names them back.  sym_mode 0 reads aux_matches2[j][0] even when
// row j was cleared (the reference's de-facto behaviour); sym_mode 1 requires keep21[j].
__global__ void __launch_bounds__(FT)
sym_matches_kernel(const int32_t* __restrict__ idx12, const float* __restrict__ dist12,
                   const uint8_t* __restrict__ keep12, int n1_max, const int32_t* __restrict__ n1_arr,
                   const int32_t* __restrict__ idx21, const uint8_t* __restrict__ keep21, int n2_max,
                   const int32_t* __restrict__ n2_arr, int sym_mode, int32_t* __restrict__ sym_q,
                   int32_t* __restrict__ sym_t, float* __restrict__ sym_d, int32_t* __restrict__ n_sym) {
    __shared__ int s_warp_cnt[FT / 32];
    __shared__ int s_base;
    const int prob = blockIdx.x, tid = threadIdx.x;
    const int n1 = n1_arr ? min(n1_arr[prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[prob], n2_max) : n2_max;
    const int32_t* i12 = idx12 + (size_t)prob * n1_max * 2;
    const float* d12 = dist12 + (size_t)prob * n1_max * 2;
    const uint8_t* k12 = keep12 + (size_t)prob * n1_max;
    const int32_t* i21 = idx21 + (size_t)prob * n2_max * 2;
    const uint8_t* k21 = keep21 ? keep21 + (size_t)prob * n2_max : nullptr;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n1; i0 += FT) {
        const int i = i0 + tid;
        bool sym = false;
        int j = -1;
        if (i < n1 && k12[i]) {
            j = i12[2 * i];
            if (j >= 0 && j < n2) {
                const int b0 = i21[2 * j];
                sym = (b0 >= 0) && (b0 == i);
                if (sym_mode == 1) sym = sym && k21 && k21[j];
            }
        }
        const int off = block_ordered_slot(sym, s_warp_cnt, &s_base);
        if (sym) {
            sym_q[(size_t)prob * n1_max + off] = i;
            sym_t[(size_t)prob * n1_max + off] = j;
            sym_d[(size_t)prob * n1_max + off] = d12[2 * i];
        }
    }
    if (tid == 0) n_sym[prob] = s_base;
}

// Matcher::sortMatches (Matcher.cpp:329-352): stable ascending order of float keys (cv::sortIdx; ties keep input
// order — the decision recorded in SURVEY App. A.1-4).  Rank sort: order[rank(k)] = k.
__global__ void __launch_bounds__(FT)
sort_keys_kernel(const float* __restrict__ keys, int cap, const int32_t* __restrict__ n_arr,
                 int32_t* __restrict__ order) {
    extern __shared__ unsigned char smem_raw[];
    float* s_key = reinterpret_cast<float*>(smem_raw);
    const int prob = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_arr[prob], cap);
    const float* kk = keys + (size_t)prob * cap;
    for (int i = tid; i < n; i += FT) s_key[i] = kk[i];
    __syncthreads();
    for (int i = tid; i < n; i += FT) {
        const float y = s_key[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const float v = s_key[j];
            rank += (v < y || (v == y && j < i)) ? 1 : 0;
        }
        order[(size_t)prob * cap + rank] = i;
    }
}

// Matcher::bestMatchesFilter (Matcher.cpp:171-244) over a y-sorted match list (Matcher::sortedMatches): cell of
// every match from the float-accumulated band / column edges, strict '<' on the distance keeps the first match
// of the sweep among equals => per-cell minimum of (distance bits, list position), one 64-bit atomicMin.
__global__ void __launch_bounds__(FT)
grid_best_kernel(const int32_t* __restrict__ list_q, const int32_t* __restrict__ list_t,
                 const float* __restrict__ list_d, int list_cap, const int32_t* __restrict__ n_list,
                 const float* __restrict__ kp1_xy, int n1_max, int w, int h, int n_cells,
                 int32_t* __restrict__ good_q, int32_t* __restrict__ good_t, float* __restrict__ good_d,
                 int32_t* __restrict__ good_pos, int good_cap, int32_t* __restrict__ n_good) {
    extern __shared__ unsigned char smem_raw[];
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem_raw);
    __shared__ float s_hf[MAX_ROOT], s_wf[MAX_ROOT];
    __shared__ int s_warp_cnt[FT / 32];
    __shared__ int s_base;
    const int prob = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_list[prob], list_cap);
    const int32_t* lq = list_q + (size_t)prob * list_cap;
    const int32_t* lt = list_t + (size_t)prob * list_cap;
    const float* ld = list_d + (size_t)prob * list_cap;
    const float* kp = kp1_xy + (size_t)prob * n1_max * 2;
    const int root = (int)floor(sqrt((double)n_cells));
    const int ncell = root * root;
    if (tid == 0) {
        s_base = 0;
        const float winW = (float)((double)w / floor(sqrt((double)n_cells)));   // Matcher.cpp:177
        const float winH = (float)((double)h / floor(sqrt((double)n_cells)));   // Matcher.cpp:178
        float hf = winH, wf = winW;
        for (int k = 0; k < root; k++) {
            s_hf[k] = hf; s_wf[k] = wf;
            hf = __fadd_rn(hf, winH);
            wf = __fadd_rn(wf, winW);
        }
    }
    for (int c = tid; c < ncell; c += FT) s_best[c] = 0xFFFFFFFFFFFFFFFFull;
    __syncthreads();
    for (int p = tid; p < n; p += FT) {
        const int q = lq[p];
        if (q < 0 || q >= n1_max) continue;
        const float x = kp[2 * q], y = kp[2 * q + 1], d = ld[p];
        int band = -1;
        for (int k = 0; k < root; k++) if (y <= s_hf[k]) { band = k; break; }
        if (band < 0) continue;
        int col = 0;
        while (col < root && x > s_wf[col]) col++;
        if (col >= root) col = root - 1;
        if (d < 100000.0f)   // distances are non-negative: float bits order like the values
            atomicMin(&s_best[band * root + col], ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)p);
    }
    __syncthreads();
    for (int c0 = 0; c0 < ncell; c0 += FT) {
        const int c = c0 + tid;
        const bool full = (c < ncell) && (s_best[c] != 0xFFFFFFFFFFFFFFFFull);
        const int off = block_ordered_slot(full, s_warp_cnt, &s_base);
        if (full && off < good_cap) {
            const int p = (int)(s_best[c] & 0xFFFFFFFFull);
            good_q[(size_t)prob * good_cap + off] = lq[p];
            good_t[(size_t)prob * good_cap + off] = lt[p];
            good_d[(size_t)prob * good_cap + off] = ld[p];
            if (good_pos) good_pos[(size_t)prob * good_cap + off] = p;
        }
    }
    if (tid == 0) n_good[prob] = min(s_base, good_cap);
}

}  // namespace

extern "C" int vsb_nn_filter(vsb_ctx_t* ctx, const int32_t* idx, const float* dist, int n_max, const int32_t* n,
                             int count, double ratio, uint8_t* keep, void* stream) {
    if (!ctx || count < 0 || n_max < 0 || (n_max > 0 && (!idx || !dist || !keep))) return VSB_ERR_INVALID;
    if (count == 0 || n_max == 0) return VSB_OK;
    dim3 grid(vsb_div_up(n_max, FT), count);
    ProfScope ps(ctx, VSB_K_MATCH_STAGE, (cudaStream_t)stream);
    nn_filter_kernel<<<grid, FT, 0, (cudaStream_t)stream>>>(idx, dist, n_max, n, ratio, keep);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_sym_matches(vsb_ctx_t* ctx, const int32_t* idx12, const float* dist12, const uint8_t* keep12,
                               int n1_max, const int32_t* n1, const int32_t* idx21, const uint8_t* keep21, int n2_max,
                               const int32_t* n2, int count, int sym_mode, int32_t* sym_q, int32_t* sym_t,
                               float* sym_d, int32_t* n_sym, void* stream) {
    if (!ctx || count < 0 || n1_max < 0 || n2_max < 0 || !n_sym) return VSB_ERR_INVALID;
    if (n1_max > 0 && (!idx12 || !dist12 || !keep12 || !sym_q || !sym_t || !sym_d)) return VSB_ERR_INVALID;
    if (n2_max > 0 && !idx21) return VSB_ERR_INVALID;
    if (sym_mode == 1 && n2_max > 0 && !keep21) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    ProfScope ps(ctx, VSB_K_MATCH_STAGE, (cudaStream_t)stream);
    sym_matches_kernel<<<count, FT, 0, (cudaStream_t)stream>>>(idx12, dist12, keep12, n1_max, n1, idx21, keep21,
                                                                 n2_max, n2, sym_mode, sym_q, sym_t, sym_d, n_sym);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_sort_keys(vsb_ctx_t* ctx, const float* keys, int cap, const int32_t* n, int count,
                             int32_t* order, void* stream) {
    if (!ctx || count < 0 || cap < 0 || !n || (cap > 0 && (!keys || !order))) return VSB_ERR_INVALID;
    if (count == 0 || cap == 0) return VSB_OK;
    const size_t smem = (size_t)cap * sizeof(float);
    if (smem > 200 * 1024) return VSB_ERR_CAPACITY;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 48 * 1024)
        VSB_CUDA(ctx, cudaFuncSetAttribute(sort_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, VSB_K_MATCH_STAGE, st);
    sort_keys_kernel<<<count, FT, smem, st>>>(keys, cap, n, order);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_grid_best(vsb_ctx_t* ctx, const int32_t* list_q, const int32_t* list_t, const float* list_d,
                             int list_cap, const int32_t* n_list, const float* kp1_xy, int n1_max, int count, int w,
                             int h, int n_cells, int32_t* good_q, int32_t* good_t, float* good_d, int32_t* good_pos,
                             int good_cap, int32_t* n_good, void* stream) {
    if (!ctx || count < 0 || n_cells < 1 || list_cap < 0 || !n_list || !n_good) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    const int root = (int)floor(sqrt((double)n_cells));
    if (root > MAX_ROOT) return VSB_ERR_CAPACITY;
    if (good_cap < root * root) return VSB_ERR_INVALID;
    const size_t smem = (size_t)root * root * sizeof(unsigned long long);
    ProfScope ps(ctx, VSB_K_MATCH_STAGE, (cudaStream_t)stream);
    grid_best_kernel<<<count, FT, smem, (cudaStream_t)stream>>>(list_q, list_t, list_d, list_cap, n_list, kp1_xy,
                                                                  n1_max, w, h, n_cells, good_q, good_t, good_d,
                                                                  good_pos, good_cap, n_good);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

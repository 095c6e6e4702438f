// knn_l2_tc.cu — float-descriptor kNN (k = 2) as a tensor-core distance GEMM followed by an exact re-check of the
// top candidates: BFMatcher(NORM_L2)::knnMatch as called by Matcher::computeMatches (reference src/Matcher.cpp:55,
// 83-94) / MatcherGPU::computeGPUMatches (src/MatcherGPU.cpp:23, 44-66), for SURF-like descriptors (dim <= 64,
// dim % 8 == 0).  Results are bit-identical to the exact kernel in knn_l2.cu (and therefore to the oracle): the tensor
// cores only SELECT candidates, every reported distance is recomputed with the reference arithmetic, and a row whose
// selection cannot be proven complete is recomputed exhaustively.
//
// Stage 0 (l2_prep_kernel): per descriptor, -|b|^2 / 2 split into three TF32-exact pieces, and the largest norm of
//   each set (for the error bound).
// Stage 1 (knn2_l2_tc_kernel): score(i, j) = a_i . b_j - |b_j|^2 / 2  (maximal score == minimal distance for a fixed
//   row) on tcgen05.mma kind::tf32 with FP32 accumulators in TMEM.  Operands are split a = hi + lo with hi = a with the
//   low 13 mantissa bits cleared (exact in TF32 whatever the hardware's conversion rounding) and lo = a - hi, and three
//   products hi.hi + hi.lo + lo.hi are accumulated (error <= 3 * 2^-20 |a||b|); the norm term rides along as one extra
//   K step (a constant row of ones times the three pieces), so the accumulator IS the score and masked columns
//   (piece = -1e30) lose by themselves.  Same warp-specialised pipeline as the Hamming kernel (knn_tc.cu): 4 epilogue
//   warps (TMEM lane == row), 4 producer warps (global -> split -> 128-byte-swizzled K-major tiles), 1 UMMA warp,
//   operand and accumulator stages double-buffered.  The epilogue keeps the 3 best scores per row with the column
//   packed into the low 7 mantissa bits (one LOP3 + five FMNMX per element, two interleaved lists).
// Stage 2 (l2_finalize_kernel): exact distances sqrtf((float) sum_k (double)(a_k - b_k)^2) of the 3 candidates,
//   top-2 by (distance, index), and the completeness proof: every column that was NOT kept has score <= s3 (the
//   third-best kept score), hence true squared distance >= |a|^2 - 2 (s3 + eps); if that lower bound does not clear
//   the second-best exact distance the row is flagged.
// Stage 3 (l2_fallback_kernel): flagged rows only (ties among >= 3 columns, near-duplicates), one warp per row over
//   all columns with the exact arithmetic.
#include "common.cuh"
#include "knn_keys.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128, TN = 128;
constexpr int L2TC_THREADS = 288;
constexpr int KB_BYTES = 128 * 128;              // one K block: 128 rows x 32 floats
constexpr int EXT_BYTES = 4096;                  // norm operand: 128 rows x 8 TF32, no swizzle
constexpr int TMEM_COLS = 2 * TN;
constexpr uint32_t IDESC_TF32 = umma::idesc(/*D f32*/ 1, /*A tf32*/ 2, /*B tf32*/ 2, TM, TN);
constexpr float MASKED_SCORE = -1.0e30f;
constexpr double EPS_REL = 6.0e-5;               // 2x the worst-case bound derived in DESIGN.md (split + FP32 accumulation + packing)

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(id), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// exact squared distance with the reference arithmetic (difference in float, squares summed in double in k order)
__device__ __forceinline__ double l2_exact_sq(const float* __restrict__ a, const float* __restrict__ b, int dim) {
    double s = 0.0;
    for (int k = 0; k < dim; k++) {
        const double d = (double)__fsub_rn(__ldg(a + k), __ldg(b + k));
        s = __fma_rn(d, d, s);
    }
    return s;
}
__device__ __forceinline__ unsigned long long l2_key(double s, uint32_t idx) {
    const float d = __fsqrt_rn((float)s);
    return ((unsigned long long)__float_as_uint(d) << 32) | idx;
}

// ---- stage 0 ---------------------------------------------------------------------------------------------
// prep[desc] = {h1, h2, h3, |b|} with h1 + h2 + h3 = -|b|^2 / 2 (each piece exact in TF32); nmax[prob*2 + set] = max |b|.
// Eight lanes per descriptor: a lane sums the squares of its float4 chunks (chunk c belongs to lane c % 8, so the eight
// lanes of a descriptor read 128 contiguous bytes per step), the partial sums meet in an xor tree.  The norm only feeds the
// candidate SELECTION and an upper bound — exactness comes from the re-check — so the summation order is free.
__global__ void __launch_bounds__(256)
l2_prep_kernel(const float* __restrict__ d1, int n1_max, const int32_t* __restrict__ n1_arr,
               const float* __restrict__ d2, int n2_max, const int32_t* __restrict__ n2_arr, int dim,
               float4* __restrict__ prep1, float4* __restrict__ prep2, uint32_t* __restrict__ nmax) {
    const int prob = blockIdx.z, set = blockIdx.y;
    const int n_max = set ? n2_max : n1_max;
    const int32_t* n_arr = set ? n2_arr : n1_arr;
    const int n = n_arr ? min(n_arr[prob], n_max) : n_max;
    const int i = blockIdx.x * 32 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    float norm = 0.f;
    double s = 0.0;
    if (i < n) {
        const float4* v = reinterpret_cast<const float4*>((set ? d2 : d1) + ((size_t)prob * n_max + i) * dim);
        for (int c = sub; c < (dim >> 2); c += 8) {
            const float4 x = __ldg(v + c);
            s = __fma_rn((double)x.x, (double)x.x, s); s = __fma_rn((double)x.y, (double)x.y, s);
            s = __fma_rn((double)x.z, (double)x.z, s); s = __fma_rn((double)x.w, (double)x.w, s);
        }
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (i < n && sub == 0) {
        const double h = -0.5 * s;
        const float h1 = tf32_trunc((float)h);
        const double r1 = h - (double)h1;
        const float h2 = tf32_trunc((float)r1);
        const float h3 = tf32_trunc((float)(r1 - (double)h2));
        norm = (float)sqrt(s) * 1.0000002f;                                   // rounded up: it feeds an upper bound
        (set ? prep2 : prep1)[(size_t)prob * n_max + i] = make_float4(h1, h2, h3, norm);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) norm = fmaxf(norm, __shfl_xor_sync(0xffffffffu, norm, off));
    if ((threadIdx.x & 31) == 0 && norm > 0.f) atomicMax(nmax + prob * 2 + set, __float_as_uint(norm));
}

// ---- stage 1 ---------------------------------------------------------------------------------------------
// Producer mapping.  A 128-row operand tile is 128 x (kblocks * 8) chunks of 16 bytes (4 floats).  Thread t of the 128
// producers owns chunk column c = t & 15 of rows (t >> 4) + 8 i, i = 0..15: a warp's load instruction reads two whole
// 256-byte descriptors (fully coalesced), and in the 128B-swizzled K-major tile the 16 chunks of a thread sit exactly
// 1024 bytes apart:  offset(i) = (c >> 3) * KB_BYTES + i * 1024 + (t >> 4) * 128 + (((c & 7) ^ (t >> 4)) << 4).
// hi = value with the low 13 mantissa bits cleared, lo = value - hi (both exact).
__device__ __forceinline__ void load_chunks(float4 (&v)[16], const float* __restrict__ g, int row_begin, int n_rows,
                                            int dim, int t, bool active) {
    const int c = t & 15;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int r = row_begin + (t >> 4) + 8 * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && r < n_rows) v[i] = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * dim) + c);
    }
}
__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, const float4 (&v)[16], int t) {
    const int c = t & 15, rr = t >> 4;
    const uint32_t base = (uint32_t)(c >> 3) * KB_BYTES + (uint32_t)rr * 128u + (uint32_t)(((c & 7) ^ rr) << 4);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float4 h = make_float4(tf32_trunc(v[i].x), tf32_trunc(v[i].y), tf32_trunc(v[i].z), tf32_trunc(v[i].w));
        const float4 l = make_float4(__fsub_rn(v[i].x, h.x), __fsub_rn(v[i].y, h.y), __fsub_rn(v[i].z, h.z),
                                     __fsub_rn(v[i].w, h.w));
        *reinterpret_cast<float4*>(hi_tile + base + i * 1024) = h;
        *reinterpret_cast<float4*>(lo_tile + base + i * 1024) = l;
    }
}

__device__ __forceinline__ void top3_insert(float& b0, float& b1, float& b2, float x) {
    const float m0 = fminf(b0, x), m1 = fminf(b1, x);
    b0 = fmaxf(b0, x);
    b1 = fmaxf(b1, m0);
    b2 = fmaxf(b2, m1);
}

// sorted insertion of (score, column) into the row's running top-3; strict '>' keeps the earlier column on equal scores
__device__ __forceinline__ void top3_insert_idx(float (&gs)[3], int (&gi)[3], float x, int c) {
    if (x > gs[2]) {
        if (x > gs[1]) {
            gs[2] = gs[1]; gi[2] = gi[1];
            if (x > gs[0]) { gs[1] = gs[0]; gi[1] = gi[0]; gs[0] = x; gi[0] = c; }
            else { gs[1] = x; gi[1] = c; }
        } else { gs[2] = x; gi[2] = c; }
    }
}

struct L2TcParams {
    const float* d1; const float* d2;
    const float4* prep1; const float4* prep2;
    int n1_max, n2_max;
    const int32_t* n1_arr; const int32_t* n2_arr;
    int dim, kblocks;
    int4* cand12; int4* cand21;          // {col0, col1, col2, bits(s3)} per row
};

__global__ void __launch_bounds__(L2TC_THREADS, 1)
knn2_l2_tc_kernel(const L2TcParams P) {
    const int prob = blockIdx.z, dir = blockIdx.y;
    const int n1 = P.n1_arr ? min(P.n1_arr[prob], P.n1_max) : P.n1_max;
    const int n2 = P.n2_arr ? min(P.n2_arr[prob], P.n2_max) : P.n2_max;
    const int n_rows = dir ? n2 : n1, n_cols = dir ? n1 : n2;
    const int row0 = blockIdx.x * TM;
    if (row0 >= n_rows) return;
    const int dim = P.dim, kblocks = P.kblocks, chunks = dim >> 2;
    const float* g_rows = dir ? P.d2 + (size_t)prob * P.n2_max * dim : P.d1 + (size_t)prob * P.n1_max * dim;
    const float* g_cols = dir ? P.d1 + (size_t)prob * P.n1_max * dim : P.d2 + (size_t)prob * P.n2_max * dim;
    const float4* col_prep = dir ? P.prep1 + (size_t)prob * P.n1_max : P.prep2 + (size_t)prob * P.n2_max;
    int4* cand_out = dir ? P.cand21 + (size_t)prob * P.n2_max : P.cand12 + (size_t)prob * P.n1_max;
    const int T = (n_cols + TN - 1) / TN;

    extern __shared__ __align__(1024) uint8_t smem[];
    const int op_bytes = kblocks * KB_BYTES;
    uint8_t* sAhi = smem;
    uint8_t* sAlo = sAhi + op_bytes;
    uint8_t* sB = sAlo + op_bytes;                       // [stage][hi, lo][op_bytes]
    uint8_t* sAext = sB + 4 * op_bytes;
    uint8_t* sBext = sAext + EXT_BYTES;                  // [stage][EXT_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBext + 2 * EXT_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t bar0 = umma::smem_u32(bars);
    auto BAR = [&](int id) { return bar0 + 8u * (uint32_t)id; };   // bfull 0-1, bempty 2-3, tfull 4-5, tempty 6-7

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        umma::mbar_init(BAR(0), 128); umma::mbar_init(BAR(1), 128);
        umma::mbar_init(BAR(2), 1);   umma::mbar_init(BAR(3), 1);
        umma::mbar_init(BAR(4), 1);   umma::mbar_init(BAR(5), 1);
        umma::mbar_init(BAR(6), 128); umma::mbar_init(BAR(7), 128);
        umma::fence_mbar_init();
    }
    if (warp == 8) umma::tmem_alloc<TMEM_COLS>(umma::smem_u32(tmem_slot));
    // constant row operand of the norm step (no swizzle, K-major): 8-row core matrices of 16-byte rows, 128 bytes apart
    // (SBO); K elements 4..7 in a second 16-byte column 2048 bytes further (LBO).  Row = (1, 1, 1, 0 | 0, 0, 0, 0).
    for (int ci = tid; ci < 256; ci += L2TC_THREADS)
        reinterpret_cast<float4*>(sAext)[ci] = ci < 128 ? make_float4(1.f, 1.f, 1.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (umma::smem_u32(smem) & 1023u) __trap();

    if (warp < 4) {
        // ===================================== epilogue =====================================================
        const int row = row0 + warp * 32 + lane;
        float gs[3] = {-INFINITY, -INFINITY, -INFINITY};
        int gi[3] = {-1, -1, -1};
        for (int j = 0; j < T; j++) {
            const int s = j & 1, n = j >> 1;
            umma::mbar_wait(BAR(4 + s), n & 1);
            umma::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * TN);
            const int col0 = j * TN;
            uint32_t va[32], vb[32];
            umma::tmem_ld32(taddr, va);
            {
                // Branch-free top-3 of the tile on packed scores, two interleaved lists.  (A threshold test against the
                // row's running third-best score needs fewer ALU instructions per element but serialises one
                // compare -> vote -> branch chain per pair in the single epilogue warp of each scheduler: measured 2x
                // slower.)
                float l0[3] = {-INFINITY, -INFINITY, -INFINITY}, l1[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int chunk = 0; chunk < 4; chunk++) {
                    uint32_t (&v)[32] = (chunk & 1) ? vb : va;
                    uint32_t (&nx)[32] = (chunk & 1) ? va : vb;
                    umma::tmem_wait_ld_dep(v);                                  // chunk's registers are valid from here on
                    if (chunk < 3) umma::tmem_ld32(taddr + (chunk + 1) * 32, nx);   // next chunk in flight during the selection
                    else { umma::fence_before_sync(); umma::mbar_arrive(BAR(6 + s)); }
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const uint32_t c = chunk * 32 + i;
                        // column-in-tile in the low 7 mantissa bits (a 2^-16 relative perturbation, inside the bound)
                        const float x0 = __uint_as_float((v[i] & 0xFFFFFF80u) | (127u - c));
                        const float x1 = __uint_as_float((v[i + 1] & 0xFFFFFF80u) | (126u - c));
                        top3_insert(l0[0], l0[1], l0[2], x0);
                        top3_insert(l1[0], l1[1], l1[2], x1);
                    }
                }
                // fold the tile's six survivors into the row's global top-3 (score, global column)
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    const float x = q < 3 ? l0[q] : l1[q - 3];
                    top3_insert_idx(gs, gi, x, col0 + 127 - (int)(__float_as_uint(x) & 127u));
                }
            }
        }
        if (row < n_rows) cand_out[row] = make_int4(gi[0], gi[1], gi[2], (int)__float_as_uint(gs[2]));
    } else if (warp < 8) {
        // ===================================== producers ====================================================
        const int p = tid - 128;
        const bool active = (p & 15) < chunks && ((p & 15) >> 3) < kblocks;   // chunk column inside the descriptor
        const bool stores = ((p & 15) >> 3) < kblocks;                        // chunk column inside the staged K blocks
        float4 v[16];
        load_chunks(v, g_rows, row0, n_rows, dim, p, active);
        if (stores) store_split(sAhi, sAlo, v, p);
        // norm operand slot of column p: core matrix p / 8, row p % 8, first 16-byte K column
        const uint32_t ext_off = (uint32_t)((p >> 3) * 128 + (p & 7) * 16);
        if (T > 0) load_chunks(v, g_cols, 0, n_cols, dim, p, active);
        for (int j = 0; j < T; j++) {
            const int s = j & 1, n = j >> 1;
            const int cn = j * TN + p;
            float4 e = make_float4(MASKED_SCORE, 0.f, 0.f, 0.f);
            if (cn < n_cols) { const float4 h = __ldg(col_prep + cn); e = make_float4(h.x, h.y, h.z, 0.f); }
            umma::mbar_wait(BAR(2 + s), (n & 1) ^ 1);                  // the UMMAs that read this stage have completed
            uint8_t* stage = sB + (size_t)s * 2 * op_bytes;
            if (stores) store_split(stage, stage + op_bytes, v, p);
            *reinterpret_cast<float4*>(sBext + s * EXT_BYTES + ext_off) = e;
            *reinterpret_cast<float4*>(sBext + s * EXT_BYTES + 2048 + ext_off) = make_float4(0.f, 0.f, 0.f, 0.f);
            umma::fence_proxy_async();
            umma::mbar_arrive(BAR(0 + s));
            if (j + 1 < T) load_chunks(v, g_cols, (j + 1) * TN, n_cols, dim, p, active);   // in flight during the next wait
        }
    } else {
        // ===================================== UMMA issuer ==================================================
        if (lane == 0) {
            const uint32_t aHi = umma::smem_u32(sAhi), aLo = umma::smem_u32(sAlo), aB = umma::smem_u32(sB);
            const uint64_t aext_desc = umma::smem_desc(umma::smem_u32(sAext), 2048, 128, umma::LAYOUT_NONE);
            const int ksteps = dim >> 3;                                 // 8 TF32 (32 bytes) per UMMA
            for (int j = 0; j < T; j++) {
                const int s = j & 1, n = j >> 1;
                umma::mbar_wait(BAR(0 + s), n & 1);
                umma::mbar_wait(BAR(6 + s), (n & 1) ^ 1);
                umma::fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * TN);
                const uint32_t bHi = aB + (uint32_t)(s * 2 * op_bytes), bLo = bHi + (uint32_t)op_bytes;
                uint32_t acc = 0u;
                // small terms first: lo.hi, hi.lo, then hi.hi, then the norm step
                for (int prod = 0; prod < 3; prod++) {
                    const uint32_t a0 = prod == 0 ? aLo : aHi;
                    const uint32_t b0 = prod == 1 ? bLo : bHi;
                    for (int k = 0; k < ksteps; k++) {
                        const uint32_t off = (uint32_t)((k >> 2) * KB_BYTES + (k & 3) * 32);
                        mma_tf32(d_tmem, umma::smem_desc(a0 + off, 16, 1024, umma::LAYOUT_SW128),
                                 umma::smem_desc(b0 + off, 16, 1024, umma::LAYOUT_SW128), IDESC_TF32, acc);
                        acc = 1u;
                    }
                }
                mma_tf32(d_tmem, aext_desc,
                         umma::smem_desc(umma::smem_u32(sBext) + (uint32_t)(s * EXT_BYTES), 2048, 128, umma::LAYOUT_NONE),
                         IDESC_TF32, acc);
                umma::commit(BAR(2 + s));
                umma::commit(BAR(4 + s));
            }
        }
        __syncwarp();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        umma::fence_after_sync();
        umma::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---- stage 2 ---------------------------------------------------------------------------------------------
// Four lanes per row: lanes 0..2 recompute one candidate each with the reference arithmetic, lane 3 forms |a|^2.
__global__ void __launch_bounds__(128)
l2_finalize_kernel(const L2TcParams P, int prob0, const uint32_t* __restrict__ nmax, unsigned long long* __restrict__ key12,
                   unsigned long long* __restrict__ key21, int2* __restrict__ flagged,
                   unsigned long long* __restrict__ n_flagged) {
    const int prob = blockIdx.z, dir = blockIdx.y;
    const int n1 = P.n1_arr ? min(P.n1_arr[prob], P.n1_max) : P.n1_max;
    const int n2 = P.n2_arr ? min(P.n2_arr[prob], P.n2_max) : P.n2_max;
    const int n_rows = dir ? n2 : n1, n_cols = dir ? n1 : n2;
    const int rows_max = dir ? P.n2_max : P.n1_max, cols_max = dir ? P.n1_max : P.n2_max;
    const int row = blockIdx.x * 32 + (threadIdx.x >> 2), q = threadIdx.x & 3;
    const bool live = row < n_rows && n_cols > 0;
    const int dim = P.dim;
    const int rrow = min(row, rows_max - 1);
    const float4* a = reinterpret_cast<const float4*>((dir ? P.d2 : P.d1) + ((size_t)prob * rows_max + rrow) * dim);
    const float* cols = (dir ? P.d1 : P.d2) + (size_t)prob * cols_max * dim;
    const int4 c = live ? (dir ? P.cand21 : P.cand12)[(size_t)prob * rows_max + row] : make_int4(-1, -1, -1, 0);
    const int cq = q == 0 ? c.x : q == 1 ? c.y : c.z;
    const bool valid = live && q < 3 && cq >= 0 && cq < n_cols;
    double s = 0.0;
    if (valid) {
        const float4* b = reinterpret_cast<const float4*>(cols + (size_t)cq * dim);
        for (int k = 0; k < (dim >> 2); k++) {
            const float4 x = __ldg(a + k), y = __ldg(b + k);
            double d = (double)__fsub_rn(x.x, y.x); s = __fma_rn(d, d, s);
            d = (double)__fsub_rn(x.y, y.y); s = __fma_rn(d, d, s);
            d = (double)__fsub_rn(x.z, y.z); s = __fma_rn(d, d, s);
            d = (double)__fsub_rn(x.w, y.w); s = __fma_rn(d, d, s);
        }
    } else if (live && q == 3) {
        for (int k = 0; k < (dim >> 2); k++) {
            const float4 x = __ldg(a + k);
            s = __fma_rn((double)x.x, (double)x.x, s); s = __fma_rn((double)x.y, (double)x.y, s);
            s = __fma_rn((double)x.z, (double)x.z, s); s = __fma_rn((double)x.w, (double)x.w, s);
        }
    }
    const unsigned long long kq = valid ? l2_key(s, (uint32_t)cq) : KEY64_INF;
    const int lane = threadIdx.x & 31, g0 = lane & ~3;
    double sv[4];
    unsigned long long kv[3];
#pragma unroll
    for (int t = 0; t < 4; t++) sv[t] = __shfl_sync(0xffffffffu, s, g0 + t);
#pragma unroll
    for (int t = 0; t < 3; t++) kv[t] = __shfl_sync(0xffffffffu, kq, g0 + t);
    if (q != 0 || row >= rows_max) return;
    unsigned long long* keys = (dir ? key21 : key12) + ((size_t)prob * rows_max + row) * 2;
    if (!live) { keys[0] = KEY64_INF; keys[1] = KEY64_INF; return; }
    unsigned long long k0 = KEY64_INF, k1 = KEY64_INF;
    double s_second = 0.0, s_first = 0.0;
    int n_valid = 0;
#pragma unroll
    for (int t = 0; t < 3; t++) {
        if (kv[t] == KEY64_INF) continue;
        n_valid++;
        if (kv[t] < k0) { k1 = k0; s_second = s_first; k0 = kv[t]; s_first = sv[t]; }
        else if (kv[t] < k1) { k1 = kv[t]; s_second = sv[t]; }
    }
    keys[0] = k0; keys[1] = k1;
    bool flag;
    if (n_valid < 3) {
        flag = n_valid < min(n_cols, 3);     // a masked third candidate means every valid column was kept
    } else {
        // completeness: a column that was not kept has packed score <= s3, so its true squared distance is at least
        // |a|^2 - 2 (s3 + eps); the second-best exact distance must be strictly below that
        const double na2 = sv[3];
        const double bmax = (double)__uint_as_float(nmax[prob * 2 + (dir ? 0 : 1)]);
        const double eps = EPS_REL * (sqrt(na2) * bmax + 0.5 * bmax * bmax) + 1e-30;
        const double s3 = (double)__uint_as_float((uint32_t)c.w);
        const double lower = na2 - 2.0 * (s3 + eps);
        flag = !(lower * (1.0 - 1e-6) > s_second * (1.0 + 1e-6));
    }
    if (flag) flagged[atomicAdd(n_flagged, 1ull)] = make_int2((prob0 + prob) * 2 + dir, row);
}

// ---- stage 3 ---------------------------------------------------------------------------------------------
// Flagged rows only, from the compacted list: one CTA per row, the row in shared memory, every thread walks its
// columns four at a time (independent FP64 chains), block-wide top-2 by (distance, index).
__global__ void __launch_bounds__(256)
l2_fallback_kernel(const L2TcParams P, unsigned long long* __restrict__ key12, unsigned long long* __restrict__ key21,
                   const int2* __restrict__ flagged, const unsigned long long* __restrict__ n_flagged) {
    __shared__ float s_a[64];
    __shared__ unsigned long long s_k[8][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dim = P.dim;
    const unsigned long long total = *n_flagged;
    for (unsigned long long e = blockIdx.x; e < total; e += gridDim.x) {
        const int2 f = flagged[e];
        const int prob = f.x >> 1, dir = f.x & 1, row = f.y;
        const int rows_max = dir ? P.n2_max : P.n1_max, cols_max = dir ? P.n1_max : P.n2_max;
        const int n1 = P.n1_arr ? min(P.n1_arr[prob], P.n1_max) : P.n1_max;
        const int n2 = P.n2_arr ? min(P.n2_arr[prob], P.n2_max) : P.n2_max;
        const int n_cols = dir ? n1 : n2;
        const float* a = (dir ? P.d2 : P.d1) + ((size_t)prob * rows_max + row) * dim;
        const float* cols = (dir ? P.d1 : P.d2) + (size_t)prob * cols_max * dim;
        __syncthreads();
        if (tid < dim) s_a[tid] = __ldg(a + tid);
        __syncthreads();
        unsigned long long k0 = KEY64_INF, k1 = KEY64_INF;
        for (int j0 = tid; j0 < n_cols; j0 += 4 * 256) {
            double s[4] = {0.0, 0.0, 0.0, 0.0};
            const float4* b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) b[u] = reinterpret_cast<const float4*>(cols + (size_t)min(j0 + u * 256, n_cols - 1) * dim);
            for (int k = 0; k < (dim >> 2); k++) {
                const float4 x = *reinterpret_cast<const float4*>(&s_a[4 * k]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float4 y = __ldg(b[u] + k);
                    double d = (double)__fsub_rn(x.x, y.x); s[u] = __fma_rn(d, d, s[u]);
                    d = (double)__fsub_rn(x.y, y.y); s[u] = __fma_rn(d, d, s[u]);
                    d = (double)__fsub_rn(x.z, y.z); s[u] = __fma_rn(d, d, s[u]);
                    d = (double)__fsub_rn(x.w, y.w); s[u] = __fma_rn(d, d, s[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (j0 + u * 256 < n_cols) top2_insert(k0, k1, l2_key(s[u], (uint32_t)(j0 + u * 256)));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, off);
            const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off);
            top2_merge(k0, k1, o0, o1);
        }
        if (lane == 0) { s_k[warp][0] = k0; s_k[warp][1] = k1; }
        __syncthreads();
        if (tid == 0) {
            unsigned long long m0 = s_k[0][0], m1 = s_k[0][1];
            for (int wv = 1; wv < 8; wv++) top2_merge(m0, m1, s_k[wv][0], s_k[wv][1]);
            unsigned long long* keys = (dir ? key21 : key12) + ((size_t)prob * rows_max + row) * 2;
            keys[0] = m0; keys[1] = m1;
        }
    }
}

}  // namespace

// Tensor-core implementation behind vsb_knn2_l2_keys (knn_l2.cu dispatches on ctx->knn_l2_impl and the dimension).
int vsb_knn2_l2_tc(vsb_ctx* ctx, const float* d1, int n1_max, const int32_t* n1, const float* d2, int n2_max,
                   const int32_t* n2, int dim, int count, unsigned long long* key12, unsigned long long* key21,
                   cudaStream_t st) {
    if (!ctx || !d1 || !d2 || count <= 0 || n1_max <= 0 || n2_max <= 0) return VSB_ERR_INVALID;
    if (dim <= 0 || dim > 64 || (dim & 7)) return VSB_ERR_UNSUPPORTED;
    if ((((uintptr_t)d1 | (uintptr_t)d2) & 15) != 0) return VSB_ERR_UNSUPPORTED;
    const size_t nd = (size_t)count * ((size_t)n1_max + n2_max);
    // workspace of the launching STREAM (the tracker's host entry has two chunks in flight on two streams, and a
    // context-wide area would be shared by both): prep float4 [nd] | cand int4 [nd] | flagged rows int2 [nd] |
    // nmax u32 [2 count] | flagged-row counter
    const size_t off_cand = nd * sizeof(float4);
    const size_t off_flag = off_cand + nd * sizeof(int4);
    const size_t off_nmax = (off_flag + nd * sizeof(int2) + 255) & ~(size_t)255;
    const size_t off_cnt = (off_nmax + (size_t)count * 2 * sizeof(uint32_t) + 7) & ~(size_t)7;
    void* scratch = nullptr;
    int rc = vsb_stream_ws_reserve(ctx, st, off_cnt + 64, &scratch);
    if (rc) return rc;
    uint8_t* base = static_cast<uint8_t*>(scratch);
    float4* prep1 = reinterpret_cast<float4*>(base);
    float4* prep2 = prep1 + (size_t)count * n1_max;
    int4* cand12 = reinterpret_cast<int4*>(base + off_cand);
    int4* cand21 = cand12 + (size_t)count * n1_max;
    int2* flagged = reinterpret_cast<int2*>(base + off_flag);
    uint32_t* nmax = reinterpret_cast<uint32_t*>(base + off_nmax);
    unsigned long long* n_fallback = reinterpret_cast<unsigned long long*>(base + off_cnt);
    VSB_CUDA(ctx, cudaMemsetAsync(nmax, 0, off_cnt + 8 - off_nmax, st));      // set maxima and the flagged-row counter

    L2TcParams P;
    P.d1 = d1; P.d2 = d2; P.prep1 = prep1; P.prep2 = prep2;
    P.n1_max = n1_max; P.n2_max = n2_max; P.n1_arr = n1; P.n2_arr = n2;
    P.dim = dim; P.kblocks = (dim + 31) / 32;
    P.cand12 = cand12; P.cand21 = cand21;
    const int smem = 6 * P.kblocks * KB_BYTES + 3 * EXT_BYTES + 128;
    if (smem > ctx->attr_l2_tc_smem) {
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_l2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ctx->attr_l2_tc_smem = smem;
    }
    const int n_big = n1_max > n2_max ? n1_max : n2_max;
    for (int z0 = 0; z0 < count; z0 += 65535) {
        const int zc = count - z0 < 65535 ? count - z0 : 65535;
        L2TcParams Q = P;
        Q.d1 += (size_t)z0 * n1_max * dim; Q.d2 += (size_t)z0 * n2_max * dim;
        Q.prep1 += (size_t)z0 * n1_max; Q.prep2 += (size_t)z0 * n2_max;
        Q.cand12 += (size_t)z0 * n1_max; Q.cand21 += (size_t)z0 * n2_max;
        if (n1) Q.n1_arr += z0;
        if (n2) Q.n2_arr += z0;
        {
            ProfScope ps(ctx, VSB_K_KNN_L2_PREP, st);
            l2_prep_kernel<<<dim3(vsb_div_up(n_big, 32), 2, zc), 256, 0, st>>>(
                Q.d1, n1_max, Q.n1_arr, Q.d2, n2_max, Q.n2_arr, dim, const_cast<float4*>(Q.prep1),
                const_cast<float4*>(Q.prep2), nmax + (size_t)z0 * 2);
            VSB_LAUNCHED(ctx);
        }
        {
            ProfScope ps(ctx, VSB_K_KNN_L2, st);
            knn2_l2_tc_kernel<<<dim3(vsb_div_up(n_big, TM), 2, zc), L2TC_THREADS, smem, st>>>(Q);
            VSB_LAUNCHED(ctx);
        }
        {
            ProfScope ps(ctx, VSB_K_KNN_L2_FINAL, st);
            l2_finalize_kernel<<<dim3(vsb_div_up(n_big, 32), 2, zc), 128, 0, st>>>(
                Q, z0, nmax + (size_t)z0 * 2, key12 + (size_t)z0 * n1_max * 2, key21 + (size_t)z0 * n2_max * 2, flagged,
                n_fallback);
            VSB_LAUNCHED(ctx);
        }
    }
    {
        ProfScope ps(ctx, VSB_K_KNN_L2_FINAL, st);
        l2_fallback_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(P, key12, key21, flagged, n_fallback);
        VSB_LAUNCHED(ctx);
    }
    ctx->l2_fallback_counter = n_fallback;
    return VSB_OK;
}

// Test / diagnostics: number of rows the last vsb_knn2_l2* call on this context had to recompute exhaustively.
extern "C" int vsb_debug_l2_fallback_rows(vsb_ctx_t* ctx, long long* out) {
    if (!ctx || !out) return VSB_ERR_INVALID;
    *out = 0;
    if (!ctx->l2_fallback_counter) return VSB_OK;
    unsigned long long v = 0;
    VSB_CUDA(ctx, cudaDeviceSynchronize());
    VSB_CUDA(ctx, cudaMemcpy(&v, ctx->l2_fallback_counter, sizeof(v), cudaMemcpyDeviceToHost));
    *out = (long long)v;
    return VSB_OK;
}

// knn_tc.cu — Hamming kNN (k = 2) on the 5th-generation tensor cores (tcgen05 / TMEM), the B200-native form
// of Matcher::computeMatches (reference src/Matcher.cpp:83-94) / MatcherGPU::computeGPUMatches
// (src/MatcherGPU.cpp:44-66) for 256-bit binary descriptors.
//
// Arithmetic.  Each descriptor bit becomes one signed byte: +-1 in the row operand, +-64 in the column operand
// (the column tiles are re-expanded per tile, and +-64 costs one shift and one LOP3 per four bytes), so
//     dot = sum_k a_k b_k = 64 * (256 - 2 * hamming)   (exact in the int32 accumulator of tcgen05.mma kind::i8).
// A CTA owns 128 rows and sweeps the other set in 128-column tiles: 8 UMMA instructions (M = 128, N = 128,
// K = 32 bytes each) fill one 128 x 128 int32 accumulator tile in tensor memory.  The epilogue warps read it
// back with tcgen05.ld — TMEM lane == row, so a thread owns one row and its top-2 needs no cross-thread
// traffic — and select by packed key (distance, index), which reproduces cv::BFMatcher's
// (distance asc, lowest index first) order exactly.  The second direction (knnMatch(d2, d1)) is the same
// kernel with the operands swapped (blockIdx.y): tensor work is cheap, cross-lane column reductions are not.
//
// Packed epilogue (PACK16).  One extra UMMA with small constant operands (row: -128, 1, 1; column c of a tile:
// -128, 127 - c, 1) adds 16384 + (128 - c) to every accumulator element, so the tensor core itself delivers the
// selection key   key16 = 128 * (256 - hamming) + (128 - c)  in [1, 32896]:  larger key == smaller distance, ties ->
// lower column, exactly cv::BFMatcher's order.  tcgen05.ld.pack::16b then delivers TWO columns per register and the
// top-2 selection runs directly on the loaded registers on the 16x2 SIMD min/max unit (VIMNMX.U16x2, 2.5 ALU
// instructions per pair of distances, no per-element key arithmetic at all).  Columns past the end of the set carry an
// all-zero operand row and an all-zero bias column: key 0, which loses to every real column.  The per-tile winners are
// folded into the global 32-bit keys (distance << 23 | index).
//
// Pipeline (warp-specialised, 288 threads, 2 CTAs per SM):
//   warps 0-3  epilogue: wait tfull[s] -> tcgen05.ld -> top-2 -> arrive tempty[s]
//   warps 4-7  producers: raw 32-byte descriptors (global) -> signed bytes (bit tricks, no table) ->
//              128-byte-swizzled K-major operand tile -> fence.proxy.async -> arrive bfull[s]
//   warp  8    one elected thread issues the UMMAs for tile j into TMEM stage j & 1 and commits to
//              bempty[s] (operand stage free) and tfull[s] (accumulator ready)
// Both the operand stage and the accumulator stage are double-buffered, so tile j+1 is expanded and multiplied
// while tile j is being selected from.
#include "common.cuh"
#include "knn_keys.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128;                    // rows per CTA  (UMMA M)
constexpr int TN = 128;                    // columns per tile (UMMA N)
constexpr int TC_THREADS = 288;
constexpr int KB_BYTES = 128 * 128;        // one K block: 128 rows x 128 bytes (one 128B-swizzle atom per 8 rows)
constexpr int OP_BYTES = 2 * KB_BYTES;     // 256 K-bytes per row -> two K blocks
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + OP_BYTES;
constexpr int OFF_BIAS = OFF_B + 2 * OP_BYTES;
constexpr int OFF_BAR = OFF_BIAS + 3 * 4096;     // bias operands: row side, column side (full tile), column side (last tile)
constexpr int TC_SMEM = OFF_BAR + 128;
constexpr int TMEM_COLS = 2 * TN;

constexpr uint32_t IDESC_I8 = umma::idesc(/*D s32*/ 2, /*A s8*/ 1, /*B s8*/ 1, TM, TN);

// raw 32-byte descriptor of tile row p -> 256 signed bytes in the K-major, 128-byte-swizzled operand layout:
// byte offset of (row p, 16-byte chunk c of K block kb) = kb * KB_BYTES + (p / 8) * 1024 + (p % 8) * 128 + ((c ^ (p % 8)) * 16).
// `rowp` = tile + (p / 8) * 1024 + (p % 8) * 128;  coff[c] = (c ^ (p % 8)) * 16 is computed once per thread.
// K order (the same for both operands, any fixed permutation of K leaves the dot product unchanged): raw word wi
// (32 bits) fills K bytes [32 wi, 32 wi + 32); 4-byte slot m of it holds bits m, m + 8, m + 16, m + 24.
//   COLS = false (row operand, once per CTA):  +1 / -1   = 0xFFFFFFFF - 254 * ((w >> m) & 0x01010101)
//   COLS = true  (column operand, per tile):   +64 / -64 = ((w << (7 - m)) & 0x80808080) ^ 0xC0C0C0C0
template <bool COLS>
__device__ __forceinline__ void expand_row(uint8_t* rowp, const uint32_t (&coff)[8], const uint4& r0, const uint4& r1,
                                           bool valid = true) {
    const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    if (COLS && !valid) {                                // column past the end of the set: all-zero operand row, dot = 0
#pragma unroll
        for (int c = 0; c < 16; c++)
            *reinterpret_cast<uint4*>(rowp + (c >> 3) * KB_BYTES + coff[c & 7]) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
#pragma unroll
    for (int c = 0; c < 16; c++) {                       // chunk c = slots 4 (c & 1) .. + 3 of raw word c / 2
        const uint32_t word = w[c >> 1];
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int m = 4 * (c & 1) + q;
            if (COLS) o[q] = ((word << (7 - m)) & 0x80808080u) ^ 0xC0C0C0C0u;
            else o[q] = ((word >> m) & 0x01010101u) * 0xFFFFFF02u + 0xFFFFFFFFu;
        }
        *reinterpret_cast<uint4*>(rowp + (c >> 3) * KB_BYTES + coff[c & 7]) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// top-2 LARGEST of packed u16x2 keys; two candidates at once: 5 instructions (the 3-input max is one VIMNMX3.U16x2)
__device__ __forceinline__ void top2max_insert2_u16x2(uint32_t& b0, uint32_t& b1, uint32_t x, uint32_t y) {
    const uint32_t hi = __vmaxu2(x, y), lo = __vminu2(x, y);
    const uint32_t t = __vminu2(b0, hi);
    b0 = __vmaxu2(b0, hi);
    b1 = __vimax3_u16x2(b1, t, lo);
}
__device__ __forceinline__ void top2max_merge_u16x2(uint32_t& a0, uint32_t& a1, uint32_t o0, uint32_t o1) {
    const uint32_t lo = __vminu2(a0, o0);
    const uint32_t hi2 = __vmaxu2(a1, o1);
    a0 = __vmaxu2(a0, o0);
    a1 = __vmaxu2(lo, hi2);
}

// One 128-column accumulator tile, packed epilogue: two tcgen05.ld.pack::16b (64 columns each) deliver the selection keys
// themselves (see the header); pairwise top-2 insertion on the 16x2 SIMD unit, then the tile's winners are folded into
// the global keys.  No per-element arithmetic and no masking: out-of-range columns arrive as key 0.
__device__ __forceinline__ void epilogue_tile_pack16(uint32_t taddr, uint32_t tempty_bar, int col0,
                                                     uint32_t& gb0, uint32_t& gb1) {
    uint32_t pb0[2] = {0u, 0u}, pb1[2] = {0u, 0u};
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t v[32];
        umma::tmem_ld32_pack16(taddr + half * 64, v);
        umma::tmem_wait_ld();
        if (half == 1) { umma::fence_before_sync(); umma::mbar_arrive(tempty_bar); }   // accumulator stage is free again
#pragma unroll
        for (int i = 0; i < 32; i += 2) top2max_insert2_u16x2(pb0[(i >> 1) & 1], pb1[(i >> 1) & 1], v[i], v[i + 1]);
    }
    top2max_merge_u16x2(pb0[0], pb1[0], pb0[1], pb1[1]);
    // a register holds columns (2i, 2i+1) of its 64-column half in its (low, high) 16 bits: the column inside the half
    // tile comes from the key itself (128 - c_tile), so halves need no bookkeeping
    const uint32_t k16[4] = {pb0[0] & 0xFFFFu, pb0[0] >> 16, pb1[0] & 0xFFFFu, pb1[0] >> 16};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t t = k16[q] - 1u;                         // 128 * (256 - hamming) + (127 - c)
        const uint32_t g = (k16[q] == 0u) ? KEY_INF
                                          : (((256u - (t >> 7)) << KEY_SHIFT) | (uint32_t)(col0 + 127 - (int)(t & 127u)));
        top2_insert(gb0, gb1, g);
    }
}

// 32-bit epilogue: four tcgen05.ld of 32 columns, key = hamming << 23 | column in one IMAD, 8 independent streams.
template <bool MASKED>
__device__ __forceinline__ void epilogue_tile_u32(uint32_t taddr, uint32_t tempty_bar, int col0, int n_cols,
                                                  uint32_t (&sb0)[8], uint32_t (&sb1)[8], int32_t* dump_row) {
#pragma unroll
    for (int chunk = 0; chunk < 4; chunk++) {
        uint32_t v[32];
        umma::tmem_ld32(taddr + chunk * 32, v);
        umma::tmem_wait_ld();
        if (chunk == 3) { umma::fence_before_sync(); umma::mbar_arrive(tempty_bar); }
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int col = col0 + chunk * 32 + i;
            uint32_t key = v[i] * 0xFFFF0000u + (0x40000000u + (uint32_t)col);   // (16384 - dot) << 16 = hamming << 23
            if (MASKED && col >= n_cols) key = KEY_INF;
            top2_insert(sb0[i & 7], sb1[i & 7], key);
            if (dump_row && col < n_cols) dump_row[col] = (int32_t)v[i];
        }
    }
}

template <bool PACK16>
__global__ void __launch_bounds__(TC_THREADS, 2)
knn2_hamming_tc_kernel(const uint8_t* __restrict__ d1, int n1_max, const int32_t* __restrict__ n1_arr,
                       const uint8_t* __restrict__ d2, int n2_max, const int32_t* __restrict__ n2_arr,
                       uint32_t* __restrict__ key12, uint32_t* __restrict__ key21, int32_t* __restrict__ dump,
                       int dump_ld) {
    const int prob = blockIdx.z, dir = blockIdx.y;
    const int n1 = n1_arr ? min(n1_arr[prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[prob], n2_max) : n2_max;
    const int n_rows = dir ? n2 : n1, n_cols = dir ? n1 : n2;
    const int row0 = blockIdx.x * TM;
    if (row0 >= n_rows) return;                                   // uniform per CTA, before any allocation
    const uint4* __restrict__ g_rows = reinterpret_cast<const uint4*>(dir ? d2 + (size_t)prob * n2_max * 32
                                                                          : d1 + (size_t)prob * n1_max * 32);
    const uint4* __restrict__ g_cols = reinterpret_cast<const uint4*>(dir ? d1 + (size_t)prob * n1_max * 32
                                                                          : d2 + (size_t)prob * n2_max * 32);
    uint32_t* keys_out = dir ? key21 + (size_t)prob * n2_max * 2 : key12 + (size_t)prob * n1_max * 2;
    const int T = (n_cols + TN - 1) / TN;

    extern __shared__ __align__(1024) uint8_t smem[];    // 128-byte swizzle atoms need a 1024-byte aligned base
    uint8_t* sA = smem + OFF_A;
    uint8_t* sB = smem + OFF_B;
    uint8_t* sBias = smem + OFF_BIAS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
    const uint32_t bar0 = umma::smem_u32(bars);
    // barrier ids: bfull[s] = s, bempty[s] = 2 + s, tfull[s] = 4 + s, tempty[s] = 6 + s
    auto BAR = [&](int id) { return bar0 + 8u * (uint32_t)id; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        umma::mbar_init(BAR(0), 128); umma::mbar_init(BAR(1), 128);     // bfull: the 128 producer threads
        umma::mbar_init(BAR(2), 1);   umma::mbar_init(BAR(3), 1);       // bempty: tcgen05.commit
        umma::mbar_init(BAR(4), 1);   umma::mbar_init(BAR(5), 1);       // tfull: tcgen05.commit
        umma::mbar_init(BAR(6), 128); umma::mbar_init(BAR(7), 128);     // tempty: the 128 epilogue threads
        umma::fence_mbar_init();
    }
    if (warp == 8) umma::tmem_alloc<TMEM_COLS>(umma::smem_u32(tmem_slot));
    // bias operands (no swizzle, K-major): 8-row core matrices of 16-byte rows, 128 bytes apart (SBO); the second
    // 16-byte K column 2048 bytes further (LBO).  Row side: K elements (-128, 1, 1, 0, ...); column side, column c of a
    // tile: (-128, 127 - c, 1, 0, ...), so the extra UMMA adds 16384 + (128 - c); columns past the end of the set are all
    // zero (only the last tile can have them).
    {
        const int last_valid = n_cols - (T - 1) * TN;
        for (int ci = tid; ci < 3 * 256; ci += TC_THREADS) {
            const int which = ci >> 8, r = ci & 255;
            uint32_t w0 = 0u;
            if (r < 128) {
                if (which == 0) w0 = 0x00010180u;
                else if (which == 1 || r < last_valid) w0 = 0x00010080u | ((uint32_t)(127 - r) << 8);
            }
            reinterpret_cast<uint4*>(sBias)[ci] = make_uint4(w0, 0u, 0u, 0u);
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if ((bar0 - OFF_BAR) & 1023u) __trap();                        // operand tiles must sit on a 1024-byte boundary

    if (warp < 4) {
        // ===================================== epilogue =====================================================
        const int row = row0 + warp * 32 + lane;
        uint32_t gb0 = KEY_INF, gb1 = KEY_INF;
        uint32_t sb0[8], sb1[8];                                       // 32-bit path: 8 independent streams
#pragma unroll
        for (int q = 0; q < 8; q++) { sb0[q] = KEY_INF; sb1[q] = KEY_INF; }
        for (int j = 0; j < T; j++) {
            const int s = j & 1, n = j >> 1;
            umma::mbar_wait(BAR(4 + s), n & 1);
            umma::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * TN);
            const int col0 = j * TN;
            const bool full = col0 + TN <= n_cols;
            if (PACK16) {
                epilogue_tile_pack16(taddr, BAR(6 + s), col0, gb0, gb1);
            } else {
                // the variant is chosen by a CTA-uniform condition: tcgen05.ld is warp-collective
                const bool dumping = dump != nullptr && dir == 0 && prob == 0;
                int32_t* dump_row = (dumping && row < n_rows) ? dump + (size_t)row * dump_ld : nullptr;
                if (full && !dumping) epilogue_tile_u32<false>(taddr, BAR(6 + s), col0, n_cols, sb0, sb1, nullptr);
                else epilogue_tile_u32<true>(taddr, BAR(6 + s), col0, n_cols, sb0, sb1, dump_row);
            }
        }
        if (!PACK16) {
#pragma unroll
            for (int q = 0; q < 8; q++) top2_merge(gb0, gb1, sb0[q], sb1[q]);
        }
        if (row < n_rows) *reinterpret_cast<uint2*>(keys_out + (size_t)row * 2) = make_uint2(gb0, gb1);
    } else if (warp < 8) {
        // ===================================== producers ====================================================
        const int p = tid - 128;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t row_off = (uint32_t)((p >> 3) * 1024 + (p & 7) * 128);
        uint32_t coff[8];
#pragma unroll
        for (int c = 0; c < 8; c++) coff[c] = (uint32_t)((c ^ (p & 7)) << 4);
        {
            const int r = row0 + p;
            const uint4 a0 = r < n_rows ? __ldg(g_rows + (size_t)r * 2) : zero;
            const uint4 a1 = r < n_rows ? __ldg(g_rows + (size_t)r * 2 + 1) : zero;
            expand_row<false>(sA + row_off, coff, a0, a1);
        }
        uint4 n0 = zero, n1v = zero;
        if (T > 0 && p < n_cols) { n0 = __ldg(g_cols + (size_t)p * 2); n1v = __ldg(g_cols + (size_t)p * 2 + 1); }
        for (int j = 0; j < T; j++) {
            const int s = j & 1, n = j >> 1;
            const uint4 c0 = n0, c1 = n1v;
            const bool cvalid = j * TN + p < n_cols;
            const int cn = (j + 1) * TN + p;
            if (j + 1 < T && cn < n_cols) { n0 = __ldg(g_cols + (size_t)cn * 2); n1v = __ldg(g_cols + (size_t)cn * 2 + 1); }
            else { n0 = zero; n1v = zero; }
            umma::mbar_wait(BAR(2 + s), (n & 1) ^ 1);                  // the UMMAs that read this stage have completed
            expand_row<true>(sB + s * OP_BYTES + row_off, coff, c0, c1, cvalid);
            umma::fence_proxy_async();
            umma::mbar_arrive(BAR(0 + s));
        }
    } else {
        // ===================================== UMMA issuer ==================================================
        if (lane == 0) {
            const uint32_t aA = umma::smem_u32(sA), aB = umma::smem_u32(sB), aBias = umma::smem_u32(sBias);
            const uint64_t abias_desc = umma::smem_desc(aBias, /*LBO*/ 2048, /*SBO*/ 128, umma::LAYOUT_NONE);
            const uint64_t bbias_full = umma::smem_desc(aBias + 4096, 2048, 128, umma::LAYOUT_NONE);
            const uint64_t bbias_last = umma::smem_desc(aBias + 8192, 2048, 128, umma::LAYOUT_NONE);
            for (int j = 0; j < T; j++) {
                const int s = j & 1, n = j >> 1;
                umma::mbar_wait(BAR(0 + s), n & 1);                    // operands of tile j are in shared memory
                umma::mbar_wait(BAR(6 + s), (n & 1) ^ 1);              // the epilogue has drained this accumulator stage
                umma::fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * TN);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t off = (uint32_t)((k >> 2) * KB_BYTES + (k & 3) * 32);
                    const uint64_t da = umma::smem_desc(aA + off, 16, 1024, umma::LAYOUT_SW128);
                    const uint64_t db = umma::smem_desc(aB + s * OP_BYTES + off, 16, 1024, umma::LAYOUT_SW128);
                    umma::mma_i8(d_tmem, da, db, IDESC_I8, k > 0 ? 1u : 0u);
                }
                if (PACK16) umma::mma_i8(d_tmem, abias_desc, j == T - 1 ? bbias_last : bbias_full, IDESC_I8, 1u);
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

}  // namespace

// Tensor-core implementation behind vsb_knn2_hamming_keys (csrc/knn_hamming.cu dispatches on ctx->knn_impl).
// `dump` (optional, debug): int32 [n1][dump_ld] raw dot products of problem 0, direction 1->2 (32-bit path only).
int vsb_knn2_hamming_tc(vsb_ctx* ctx, const uint8_t* d1, int n1_max, const int32_t* n1, const uint8_t* d2,
                        int n2_max, const int32_t* n2, int count, uint32_t* key12, uint32_t* key21, int pack16,
                        int32_t* dump, int dump_ld, cudaStream_t st) {
    if (!ctx || count < 0 || n1_max < 0 || n2_max < 0) return VSB_ERR_INVALID;
    if (n1_max > (int)KEY_IDX_MASK || n2_max > (int)KEY_IDX_MASK) return VSB_ERR_CAPACITY;
    if (count == 0 || (n1_max == 0 && n2_max == 0)) return VSB_OK;
    if (!ctx->attr_knn_tc_done) {
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_tc_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_tc_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->attr_knn_tc_done = 1;
    }
    const int row_tiles = vsb_div_up(n1_max > n2_max ? n1_max : n2_max, TM);
    for (int z0 = 0; z0 < count; z0 += 65535) {
        const int zc = count - z0 < 65535 ? count - z0 : 65535;
        dim3 grid(row_tiles, 2, zc);
        ProfScope ps(ctx, VSB_K_KNN_HAMMING, st);
        const uint8_t* a = d1 + (size_t)z0 * n1_max * 32;
        const uint8_t* b = d2 + (size_t)z0 * n2_max * 32;
        uint32_t* k12 = key12 + (size_t)z0 * n1_max * 2;
        uint32_t* k21 = key21 + (size_t)z0 * n2_max * 2;
        if (pack16)
            knn2_hamming_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM, st>>>(a, n1_max, n1 ? n1 + z0 : nullptr, b, n2_max,
                                                                            n2 ? n2 + z0 : nullptr, k12, k21, nullptr, 0);
        else
            knn2_hamming_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM, st>>>(a, n1_max, n1 ? n1 + z0 : nullptr, b, n2_max,
                                                                             n2 ? n2 + z0 : nullptr, k12, k21,
                                                                             z0 == 0 ? dump : nullptr, dump_ld);
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

// Debug / test entry (not part of include/vislam_b200.h): raw tensor-core dot products of one problem,
// dots[i][j] = sum_k a_ik b_jk = 64 * (256 - 2 * hamming(d1[i], d2[j])), through the 32-bit epilogue.
extern "C" int vsb_debug_knn_tc_dots(vsb_ctx_t* ctx, const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                                     int32_t* dots, int ld, void* stream) {
    if (!ctx || !d1 || !d2 || !dots || n1 <= 0 || n2 <= 0 || ld < n2) return VSB_ERR_INVALID;
    void* scratch = nullptr;
    int rc = vsb_scratch_reserve(ctx, ((size_t)n1 + n2) * 2 * sizeof(uint32_t) + 256, &scratch);
    if (rc) return rc;
    uint32_t* k12 = (uint32_t*)scratch;
    return vsb_knn2_hamming_tc(ctx, d1, n1, nullptr, d2, n2, nullptr, 1, k12, k12 + (size_t)n1 * 2, 0, dots, ld,
                               (cudaStream_t)stream);
}

// knn_mx.cu — Hamming kNN (k = 2) on the tensor cores with 4-bit operands: tcgen05.mma kind::mxf4.block_scale
// (knn_impl 3).  Same problem, pipeline and result as knn_tc.cu (Matcher::computeMatches, src/Matcher.cpp:83-94 /
// MatcherGPU::computeGPUMatches, src/MatcherGPU.cpp:44-66), with half the operand bytes and twice the tensor rate.
//
// Arithmetic.  Every descriptor bit becomes one e2m1 nibble, +1 (0x2) or -1 (0xA), in BOTH operands; a descriptor is then
// exactly one 128-byte row of the 128-byte-swizzled K-major operand tile, and four UMMAs of K = 64 cover its 256 bits.
// Block scaling is mandatory for this kind, but the scales are constants: the scale-factor columns in tensor memory are
// filled once per CTA with tcgen05.st (every byte the same, so their layout never matters) — 1 (0x7F, ue8m0) for the rows,
// 64 (0x85) for the columns — and the FP32 accumulator receives 64 * (256 - 2 * hamming), an exact integer.
// Two more UMMAs with constant operands add 2^23 + 2^14 (as 2^14 x 513: fourteen products 6 x 6 and one 3 x 3, row scale
// 2^14 = 0x8D) and 128 - c for column c of the tile (binary digits of 128 - c as products 1, 2, 4, 8, 16 and repeated 16s).
// The accumulator then holds 2^23 + key with  key = 128 * (256 - hamming) + (128 - c)  in [1, 32896] — the same selection key
// as knn_tc.cu's packed epilogue — and because 2^23 + key has the key as its low mantissa bits, the low 16 bits of the FP32
// bit pattern ARE the key: tcgen05.ld.pack::16b and the VIMNMX.U16x2 top-2 stay exactly as they were.  Columns past the end
// of the set carry all-zero operand and bias rows: accumulator 0.0, key 0, which loses to every real column.
//
// Tensor memory: two accumulator stages of 96 columns (tiles are 128 rows x 96 columns) + three 8-column regions of scale
// factors = 216 of the 256 columns allocated, so two CTAs still share an SM.
#include "common.cuh"
#include "knn_keys.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128;                    // rows per CTA  (UMMA M)
constexpr int TN = 96;                     // columns per tile (UMMA N)
constexpr int MX_THREADS = 288;
constexpr int OP_BYTES = 128 * 128;        // an operand tile: 128 rows x 128 bytes (256 nibbles), one swizzle atom per 8 rows
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + OP_BYTES;            // two stages
constexpr int OFF_BIAS = OFF_B + 2 * OP_BYTES;     // row side, column side (full tile), column side (last tile)
constexpr int OFF_BAR = OFF_BIAS + 3 * OP_BYTES;
constexpr int MX_SMEM = OFF_BAR + 128;
constexpr int TMEM_COLS = 256;
constexpr int SF_ONE = 192, SF_64 = 200, SF_2P14 = 208;      // TMEM columns of the constant scale factors

// block-scaled instruction descriptor: A/B format e2m1 (1) at [7,10) / [10,13), N >> 3 at [17,23), scale format ue8m0 at bit 23,
// M >> 4 at [24,29), K = 64 (bit 31 clear); both operands K-major
constexpr uint32_t IDESC_MX = (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | (1u << 23) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void mma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate, uint32_t sfa,
                                         uint32_t sfb) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC_MX), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// packed read of 32 columns: the low 16 bits of columns (2i, 2i+1) land in the (low, high) halves of v[i]
__device__ __forceinline__ void tmem_ld16_pack16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// 8 descriptor bits -> 8 e2m1 nibbles, element i of the byte in nibble i: two shift-or-mask steps spread the four 2-bit
// fields to nibble positions, one PRMT picks {0x22, 0x2A, 0xA2, 0xAA} for (bit 2i+1, bit 2i)
__device__ __forceinline__ uint32_t expand8(uint32_t v) {
    uint32_t x = (v | (v << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    return __byte_perm(0xAAA22A22u, 0u, x);
}
// raw 32-byte descriptor of tile row p -> 128 bytes of nibbles at rowp = tile + (p / 8) * 1024 + (p % 8) * 128, 16-byte chunk c
// at ((c ^ (p % 8)) * 16); chunk c holds descriptor bytes 4c .. 4c + 3
__device__ __forceinline__ void expand_row(uint8_t* rowp, int p7, const uint4& r0, const uint4& r1, bool valid) {
    const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int c = 0; c < 8; c++) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);                 // a row past the end of the set: all-zero operand, dot = 0
        if (valid) {
            const uint32_t word = w[c];
            o.x = expand8(word & 0xFFu);
            o.y = expand8((word >> 8) & 0xFFu);
            o.z = expand8((word >> 16) & 0xFFu);
            o.w = expand8(word >> 24);
        }
        *reinterpret_cast<uint4*>(rowp + ((c ^ p7) << 4)) = o;
    }
}

// top-2 LARGEST of packed u16x2 keys; two candidates at once (the 3-input max is one VIMNMX3.U16x2)
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

// 32-bit running pair, LARGEST first (the bulk kernel's G form, see its epilogue)
__device__ __forceinline__ void top2max_insert(uint32_t& b0, uint32_t& b1, uint32_t k) {
    const uint32_t lo = min(b0, k);
    b0 = max(b0, k);
    b1 = max(b1, lo);
}
__device__ __forceinline__ void top2max_merge(uint32_t& a0, uint32_t& a1, uint32_t o0, uint32_t o1) {
    const uint32_t lo = min(a0, o0);
    const uint32_t hi2 = max(a1, o1);
    a0 = max(a0, o0);
    a1 = max(lo, hi2);
}
// G = [257 - hamming : 9][7 bits, ignored][255 - tile : 8][x : 1][127 - c : 7]  ->  hamming << 23 | column;  no candidate (distance
// field 0) -> KEY_INF
__device__ __forceinline__ uint32_t g_to_key(uint32_t g) {
    const uint32_t hp = g >> 23;
    const uint32_t col = (255u - ((g >> 8) & 0xFFu)) * (uint32_t)TN + 127u - (g & 0x7Fu);
    return hp == 0u ? KEY_INF : (((257u - hp) << KEY_SHIFT) | col);
}

// one 96-column accumulator tile: 64 + 32 columns of packed keys, pairwise top-2 on the 16x2 unit, fold into the global keys
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, uint32_t tempty_bar, int col0, uint32_t& gb0, uint32_t& gb1) {
    uint32_t pb0[2] = {0u, 0u}, pb1[2] = {0u, 0u};
    uint32_t v[32], u[16];
    umma::tmem_ld32_pack16(taddr, v);
    tmem_ld16_pack16(taddr + 64, u);
    umma::tmem_wait_ld();
    umma::fence_before_sync();
    umma::mbar_arrive(tempty_bar);                                 // accumulator stage is free again
#pragma unroll
    for (int i = 0; i < 32; i += 2) top2max_insert2_u16x2(pb0[(i >> 1) & 1], pb1[(i >> 1) & 1], v[i], v[i + 1]);
#pragma unroll
    for (int i = 0; i < 16; i += 2) top2max_insert2_u16x2(pb0[(i >> 1) & 1], pb1[(i >> 1) & 1], u[i], u[i + 1]);
    top2max_merge_u16x2(pb0[0], pb1[0], pb0[1], pb1[1]);
    const uint32_t k16[4] = {pb0[0] & 0xFFFFu, pb0[0] >> 16, pb1[0] & 0xFFFFu, pb1[0] >> 16};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t t = k16[q] - 128u;                           // 128 * (256 - hamming) + (127 - c)
        const uint32_t g = (k16[q] == 0u) ? KEY_INF
                                          : (((256u - (t >> 7)) << KEY_SHIFT) | (uint32_t)(col0 + 127 - (int)(t & 127u)));
        top2_insert(gb0, gb1, g);
    }
}

// bias operand rows (K slice 0 = bytes 0..31: the 2^14 x 513 term; K slice 1 = bytes 32..63: the 255 - c term)
__device__ __forceinline__ uint32_t bias_byte(int which, int r, int b, int last_valid) {
    // which 0: row side (every row the same); 1: column side, full tile; 2: column side, last tile (columns >= last_valid are zero)
    if (r >= (which == 0 ? TM : TN)) return 0u;
    if (which == 2 && r >= last_valid) return 0u;
    if (b < 32) {                                                   // elements 0..13 = 6 (0x7), element 14 = 3 (0x5): sum of squares 513
        return b < 7 ? 0x77u : (b == 7 ? 0x05u : 0u);
    }
    const int e0 = 2 * (b - 32), e1 = e0 + 1;                       // the two elements of this byte
    auto row_elem = [](int e) -> uint32_t { return e == 0 ? 0x2u : e == 1 ? 0x4u : e <= 18 ? 0x6u : 0u; };     // 1, 2, 4, 4, 4, ...
    auto col_elem = [](int e, int v) -> uint32_t {                  // digits of v = 255 - c against the row constants
        if (e == 0) return (v & 1) ? 0x2u : 0u;                     // 1 x 1
        if (e == 1) return (v & 2) ? 0x2u : 0u;                     // 2 x 1
        if (e == 2) return (v & 4) ? 0x2u : 0u;                     // 4 x 1
        if (e == 3) return (v & 8) ? 0x4u : 0u;                     // 4 x 2
        if (e == 4) return (v & 16) ? 0x6u : 0u;                    // 4 x 4
        if (e <= 6) return (v & 32) ? 0x6u : 0u;                    // 2 x (4 x 4)
        if (e <= 10) return (v & 64) ? 0x6u : 0u;                   // 4 x (4 x 4)
        if (e <= 18) return (v & 128) ? 0x6u : 0u;                  // 8 x (4 x 4)
        return 0u;
    };
    if (which == 0) return row_elem(e0) | (row_elem(e1) << 4);
    const int v = 255 - r;      // 128 + (127 - c): the extra 128 makes every valid key's distance field >= 1 (0 = no candidate)
    return col_elem(e0, v) | (col_elem(e1, v) << 4);
}

// pre-expansion (PRE): every descriptor is turned into its 128 bytes of nibbles ONCE per call (both directions and all the
// CTAs that sweep it read the expanded form), so the producers of the matching kernel only copy: 8 x (LDG.128 + STS.128) per
// row instead of ~200 integer instructions
__global__ void __launch_bounds__(256) mx_expand_kernel(const uint8_t* __restrict__ d, long long n_chunks, uint4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;          // (row, 16-byte output chunk) = 4 descriptor bytes
    if (i >= n_chunks) return;
    const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(d) + i);
    out[i] = make_uint4(expand8(word & 0xFFu), expand8((word >> 8) & 0xFFu), expand8((word >> 16) & 0xFFu), expand8(word >> 24));
}

// BULK: the expanded rows are additionally stored in the SWIZZLED order of the operand tile (chunk c of row r at slot
// c ^ (r % 8); tiles start at multiples of 8 rows) and every problem is padded with zero rows to a multiple of 384 rows, so an
// operand tile is one contiguous block of global memory with exactly the shared-memory image the UMMA wants: ONE
// cp.async.bulk per tile, issued by one thread and completing on the tile's mbarrier, replaces the producer warps.
__global__ void __launch_bounds__(256)
mx_expand_swizzled_kernel(const uint8_t* __restrict__ d, int n_max, const int32_t* __restrict__ n_arr, int rows_pad, int count,
                          uint4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)count * rows_pad * 8) return;
    const int c = (int)(i & 7);
    const long long rr = i >> 3;
    const int prob = (int)(rr / rows_pad), r = (int)(rr - (long long)prob * rows_pad);
    const int n = n_arr ? min(n_arr[prob], n_max) : n_max;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (r < n) {
        const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(d + ((size_t)prob * n_max + r) * 32) + c);
        o = make_uint4(expand8(word & 0xFFu), expand8((word >> 8) & 0xFFu), expand8((word >> 16) & 0xFFu), expand8(word >> 24));
    }
    out[((size_t)prob * rows_pad + r) * 8 + (c ^ (r & 7))] = o;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// the bias operand rows do not depend on the problem: built once (mx_bias_init_kernel) and copied into every CTA's tiles
__device__ uint4 g_mx_bias[2 * 128 * 4];                            // [row side, column side][row][16-byte chunk 0..3]
__global__ void mx_bias_init_kernel() {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 2 * 128 * 4) return;
    const int which = i >> 9, r = (i >> 2) & 127, c = i & 3;
    uint32_t w[4];
    for (int q = 0; q < 4; q++) {
        uint32_t v = 0u;
        for (int j = 0; j < 4; j++) v |= bias_byte(which, r, 16 * c + 4 * q + j, TN) << (8 * j);
        w[q] = v;
    }
    g_mx_bias[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int MODE>        // 0: expand in the kernel; 1: copy pre-expanded rows; 2: one bulk copy per pre-swizzled tile
__global__ void __launch_bounds__(MX_THREADS, 2)
knn2_hamming_mx_kernel(const uint8_t* __restrict__ d1, int n1_max, const int32_t* __restrict__ n1_arr,
                       const uint8_t* __restrict__ d2, int n2_max, const int32_t* __restrict__ n2_arr,
                       uint32_t* __restrict__ key12, uint32_t* __restrict__ key21, int rows_pad) {
    // MODE 1, 2: d1 / d2 point at the EXPANDED descriptors (128 bytes per row; MODE 2: rows_pad rows per problem)
    constexpr bool PRE = MODE != 0;
    const int prob = blockIdx.z, dir = blockIdx.y;
    const int n1 = n1_arr ? min(n1_arr[prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[prob], n2_max) : n2_max;
    const int n_rows = dir ? n2 : n1, n_cols = dir ? n1 : n2;
    const int row0 = blockIdx.x * TM;
    if (row0 >= n_rows) return;                                   // uniform per CTA, before any allocation
    constexpr int ROWB = PRE ? 128 : 32;                          // bytes per descriptor row as stored
    const size_t rows1 = MODE == 2 ? (size_t)rows_pad : (size_t)n1_max, rows2 = MODE == 2 ? (size_t)rows_pad : (size_t)n2_max;
    const uint4* __restrict__ g_rows = reinterpret_cast<const uint4*>(dir ? d2 + (size_t)prob * rows2 * ROWB
                                                                          : d1 + (size_t)prob * rows1 * ROWB);
    const uint4* __restrict__ g_cols = reinterpret_cast<const uint4*>(dir ? d1 + (size_t)prob * rows1 * ROWB
                                                                          : d2 + (size_t)prob * rows2 * ROWB);
    uint32_t* keys_out = dir ? key21 + (size_t)prob * n2_max * 2 : key12 + (size_t)prob * n1_max * 2;
    const int T = (n_cols + TN - 1) / TN;

    extern __shared__ __align__(1024) uint8_t smem[];    // 128-byte swizzle atoms need a 1024-byte aligned base
    uint8_t* sA = smem + OFF_A;
    uint8_t* sB = smem + OFF_B;
    uint8_t* sBias = smem + OFF_BIAS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);     // after the 9 barriers
    const uint32_t bar0 = umma::smem_u32(bars);
    // barrier ids: bfull[s] = s, bempty[s] = 2 + s, tfull[s] = 4 + s, tempty[s] = 6 + s
    auto BAR = [&](int id) { return bar0 + 8u * (uint32_t)id; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        umma::mbar_init(BAR(0), MODE == 2 ? 1 : 128); umma::mbar_init(BAR(1), MODE == 2 ? 1 : 128);   // bfull: producers / the bulk copy
        umma::mbar_init(BAR(8), 1);                                     // afull (MODE 2): the row tile has landed
        umma::mbar_init(BAR(2), 1);   umma::mbar_init(BAR(3), 1);       // bempty: tcgen05.commit
        umma::mbar_init(BAR(4), 1);   umma::mbar_init(BAR(5), 1);       // tfull: tcgen05.commit
        umma::mbar_init(BAR(6), 128); umma::mbar_init(BAR(7), 128);     // tempty: the 128 epilogue threads
        umma::fence_mbar_init();
    }
    if (warp == 8) umma::tmem_alloc<TMEM_COLS>(umma::smem_u32(tmem_slot));
    // the three bias operand tiles (K slices 0 and 1 of each row; the rest of the row is never read)
    {
        const int last_valid = n_cols - (T - 1) * TN;
        for (int i = tid; i < 3 * 128 * 4; i += MX_THREADS) {       // (tile, row, 16-byte chunk 0..3)
            const int which = i >> 9, r = (i >> 2) & 127, c = i & 3;
            uint4 w = g_mx_bias[((which ? 1 : 0) << 9) | (i & 511)];
            if (which == 2 && r >= last_valid) w = make_uint4(0u, 0u, 0u, 0u);     // columns past the end of the set
            uint8_t* rowp = sBias + which * OP_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
            *reinterpret_cast<uint4*>(rowp + ((c ^ (r & 7)) << 4)) = w;
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if ((bar0 - OFF_BAR) & 1023u) __trap();                        // operand tiles must sit on a 1024-byte boundary
    if (warp < 4) {
        // constant block scales, every byte of the region the same: 1, 64, 2^14 (ue8m0: 127 + log2)
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 8; c += 4) {
            tmem_st4(lane_base + SF_ONE + c, 0x7F7F7F7Fu);
            tmem_st4(lane_base + SF_64 + c, 0x85858585u);
            tmem_st4(lane_base + SF_2P14 + c, 0x8D8D8D8Du);
        }
        tmem_wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    if (warp < 4) {
        // ===================================== epilogue =====================================================
        const int row = row0 + warp * 32 + lane;
        uint32_t gb0 = KEY_INF, gb1 = KEY_INF;
        for (int j = 0; j < T; j++) {
            const int s = j & 1, n = j >> 1;
            umma::mbar_wait(BAR(4 + s), n & 1);
            umma::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * TN);
            epilogue_tile(taddr, BAR(6 + s), j * TN, gb0, gb1);
        }
        if (row < n_rows) *reinterpret_cast<uint2*>(keys_out + (size_t)row * 2) = make_uint2(gb0, gb1);
    } else if (warp < 8) {
        // ===================================== producers ====================================================
        const int p = tid - 128;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t row_off = (uint32_t)((p >> 3) * 1024 + (p & 7) * 128);
        if (MODE == 2) {
            if (p == 0) {                                              // the loader: one thread, one bulk copy per tile
                mbar_expect_tx(BAR(8), TM * 128);
                bulk_g2s(umma::smem_u32(sA), g_rows + (size_t)row0 * 8, TM * 128, BAR(8));
                for (int j = 0; j < T; j++) {
                    const int s = j & 1, n = j >> 1;
                    umma::mbar_wait(BAR(2 + s), (n & 1) ^ 1);          // the UMMAs that read this stage have completed
                    mbar_expect_tx(BAR(0 + s), TN * 128);
                    bulk_g2s(umma::smem_u32(sB + s * OP_BYTES), g_cols + (size_t)j * TN * 8, TN * 128, BAR(0 + s));
                }
            }
        } else if (PRE) {
            // copy of pre-expanded rows: chunk c of the row goes to the swizzled slot (c ^ (p % 8))
            {
                const int r = row0 + p;
                uint8_t* rowp = sA + row_off;
#pragma unroll
                for (int c = 0; c < 8; c++)
                    *reinterpret_cast<uint4*>(rowp + ((c ^ (p & 7)) << 4)) = r < n_rows ? __ldg(g_rows + (size_t)r * 8 + c) : zero;
            }
            uint4 nx[8];
#pragma unroll
            for (int c = 0; c < 8; c++) nx[c] = (T > 0 && p < TN && p < n_cols) ? __ldg(g_cols + (size_t)p * 8 + c) : zero;
            for (int j = 0; j < T; j++) {
                const int s = j & 1, n = j >> 1;
                uint4 cur[8];
#pragma unroll
                for (int c = 0; c < 8; c++) cur[c] = nx[c];
                const int cn = (j + 1) * TN + p;
                const bool more = j + 1 < T && p < TN && cn < n_cols;
#pragma unroll
                for (int c = 0; c < 8; c++) nx[c] = more ? __ldg(g_cols + (size_t)cn * 8 + c) : zero;
                umma::mbar_wait(BAR(2 + s), (n & 1) ^ 1);              // the UMMAs that read this stage have completed
                if (p < TN) {
                    uint8_t* rowp = sB + s * OP_BYTES + row_off;
#pragma unroll
                    for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(rowp + ((c ^ (p & 7)) << 4)) = cur[c];
                }
                umma::fence_proxy_async();
                umma::mbar_arrive(BAR(0 + s));
            }
        } else {
        {
            const int r = row0 + p;
            const bool valid = r < n_rows;
            const uint4 a0 = valid ? __ldg(g_rows + (size_t)r * 2) : zero;
            const uint4 a1 = valid ? __ldg(g_rows + (size_t)r * 2 + 1) : zero;
            expand_row(sA + row_off, p & 7, a0, a1, valid);
        }
        uint4 n0 = zero, n1v = zero;
        if (T > 0 && p < TN && p < n_cols) { n0 = __ldg(g_cols + (size_t)p * 2); n1v = __ldg(g_cols + (size_t)p * 2 + 1); }
        for (int j = 0; j < T; j++) {
            const int s = j & 1, n = j >> 1;
            const uint4 c0 = n0, c1 = n1v;
            const bool cvalid = p < TN && j * TN + p < n_cols;
            const int cn = (j + 1) * TN + p;
            if (j + 1 < T && p < TN && cn < n_cols) { n0 = __ldg(g_cols + (size_t)cn * 2); n1v = __ldg(g_cols + (size_t)cn * 2 + 1); }
            else { n0 = zero; n1v = zero; }
            umma::mbar_wait(BAR(2 + s), (n & 1) ^ 1);                  // the UMMAs that read this stage have completed
            if (p < TN) expand_row(sB + s * OP_BYTES + row_off, p & 7, c0, c1, cvalid);
            umma::fence_proxy_async();
            umma::mbar_arrive(BAR(0 + s));
        }
        }
    } else {
        // ===================================== UMMA issuer ==================================================
        if (lane == 0) {
            const uint32_t aA = umma::smem_u32(sA), aB = umma::smem_u32(sB), aBias = umma::smem_u32(sBias);
            const uint32_t sf1 = tmem_base + SF_ONE, sf64 = tmem_base + SF_64, sf2p14 = tmem_base + SF_2P14;
            if (MODE == 2) umma::mbar_wait(BAR(8), 0);                 // the row tile
            for (int j = 0; j < T; j++) {
                const int s = j & 1, n = j >> 1;
                umma::mbar_wait(BAR(0 + s), n & 1);                    // operands of tile j are in shared memory
                umma::mbar_wait(BAR(6 + s), (n & 1) ^ 1);              // the epilogue has drained this accumulator stage
                umma::fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * TN);
#pragma unroll
                for (int k = 0; k < 4; k++) {                          // K = 64 nibbles = 32 bytes per instruction
                    const uint64_t da = umma::smem_desc(aA + k * 32, 16, 1024, umma::LAYOUT_SW128);
                    const uint64_t db = umma::smem_desc(aB + s * OP_BYTES + k * 32, 16, 1024, umma::LAYOUT_SW128);
                    mma_mxf4(d_tmem, da, db, k > 0 ? 1u : 0u, sf1, sf64);
                }
                const uint32_t bcol = aBias + (j == T - 1 ? 2 : 1) * OP_BYTES;
                mma_mxf4(d_tmem, umma::smem_desc(aBias, 16, 1024, umma::LAYOUT_SW128),
                         umma::smem_desc(bcol, 16, 1024, umma::LAYOUT_SW128), 1u, sf2p14, sf1);          // + 2^23 + 2^14
                mma_mxf4(d_tmem, umma::smem_desc(aBias + 32, 16, 1024, umma::LAYOUT_SW128),
                         umma::smem_desc(bcol + 32, 16, 1024, umma::LAYOUT_SW128), 1u, sf1, sf1);        // + 128 - c
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


// ===== knn_impl 5: persistent CTAs, bulk-copy pipeline ==========================================================================
// Measured on the per-row-tile kernels above: a CTA costs about 5 us before and after its tiles (launch, barrier and TMEM set-up,
// first operand latencies, tear-down) and only ~0.5 us per tile, so at 1000 x 1000 (11 tiles per CTA) the fixed part is half of
// the run time.  Here two CTAs per SM stay resident and walk over the work items (problem, direction, row tile): barriers,
// tensor memory, scale factors and the bias tile are set up once; the row tile is double-buffered, so the next item's rows are
// in flight while the current item is multiplied; the column tiles run through a ring of BS stages.  Operand tiles are fetched
// by ONE thread with one cp.async.bulk each from descriptors that were expanded, swizzled and zero-padded once per call
// (mx_expand_swizzled_kernel).  One constant tile (g_mx_bias_tile, also a bulk copy) carries the four bias operand slices:
// K slices 0, 1 = row side, 2, 3 = column side.  Columns past the end of the set are masked in the epilogue of an item's
// last tile (their operand rows are zero, so the accumulator holds the bias alone).  Two groups of four epilogue warps
// alternate over the tiles (accumulator stage = tile parity); a row's two partial results meet in shared memory per item.
constexpr int BS = 5;                                   // column-tile stages
constexpr int B_STAGE = TN * 128;                       // 12 KB
constexpr int BK_OFF_A = 0;                             // two row-tile stages
constexpr int BK_OFF_BIAS = 2 * OP_BYTES;
constexpr int BK_OFF_B = 3 * OP_BYTES;
constexpr int BK_OFF_BAR = BK_OFF_B + BS * B_STAGE;
constexpr int BK_OFF_PART = BK_OFF_BAR + 256;           // [2][128] partial top-2 of the odd-tile group
constexpr int BK_SMEM = BK_OFF_PART + 2 * 1024;
constexpr int BK_THREADS = 352;                         // 8 epilogue warps, the loader warp, two UMMA issuer warps (even / odd tiles)
__device__ uint4 g_mx_bias_tile[128 * 8];               // the swizzled shared-memory image of the bias tile
__global__ void mx_bias_tile_init_kernel() {
    const int i = blockIdx.x * 256 + threadIdx.x;       // (row, 16-byte chunk of the 128-byte row)
    if (i >= 128 * 8) return;
    const int r = i >> 3, c = i & 7;
    uint32_t w[4];
    for (int q = 0; q < 4; q++) {
        uint32_t v = 0u;
        for (int j = 0; j < 4; j++) {
            const int b = 16 * c + 4 * q + j;           // byte of the row: 0..63 row-side slices, 64..127 column-side slices
            v |= (b < 64 ? bias_byte(0, r, b, TN) : bias_byte(1, r, b - 64, TN)) << (8 * j);
        }
        w[q] = v;
    }
    g_mx_bias_tile[r * 8 + (c ^ (r & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
}

struct MxItem { int prob, dir, row0, n_rows, n_cols, T; };
__device__ __forceinline__ bool mx_item(long long it, int row_tiles, int n1_max, const int32_t* n1_arr, int n2_max,
                                        const int32_t* n2_arr, MxItem& w) {
    w.prob = (int)(it / (2 * row_tiles));
    const int rem = (int)(it - (long long)w.prob * 2 * row_tiles);
    w.dir = rem / row_tiles;
    w.row0 = (rem - w.dir * row_tiles) * TM;
    const int n1 = n1_arr ? min(n1_arr[w.prob], n1_max) : n1_max;
    const int n2 = n2_arr ? min(n2_arr[w.prob], n2_max) : n2_max;
    w.n_rows = w.dir ? n2 : n1;
    w.n_cols = w.dir ? n1 : n2;
    w.T = (w.n_cols + TN - 1) / TN;
    return w.row0 < w.n_rows;                                     // false: nothing to do for this item (every role agrees)
}

__global__ void __launch_bounds__(BK_THREADS, 2)
knn2_hamming_mx_bulk_kernel(const uint8_t* __restrict__ e1, int n1_max, const int32_t* __restrict__ n1_arr,
                            const uint8_t* __restrict__ e2, int n2_max, const int32_t* __restrict__ n2_arr,
                            uint32_t* __restrict__ key12, uint32_t* __restrict__ key21, int rows_pad, int row_tiles,
                            long long n_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem + BK_OFF_A;
    uint8_t* sBias = smem + BK_OFF_BIAS;
    uint8_t* sB = smem + BK_OFF_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BK_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BK_OFF_BAR + 224);
    uint2* s_part = reinterpret_cast<uint2*>(smem + BK_OFF_PART);
    const uint32_t bar0 = umma::smem_u32(bars);
    // barrier ids: bfull[s] = s, bempty[s] = BS + s, tfull[t] = 2 BS + t, tempty[t] = 2 BS + 2 + t, afull[a] = 2 BS + 4 + a,
    // aempty[a] = 2 BS + 6 + a, biasfull = 2 BS + 8
    auto BAR = [&](int id) { return bar0 + 8u * (uint32_t)id; };
    constexpr int TFULL = 2 * BS, TEMPTY = 2 * BS + 2, AFULL = 2 * BS + 4, AEMPTY = 2 * BS + 6, BIASFULL = 2 * BS + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < BS; s++) { umma::mbar_init(BAR(s), 1); umma::mbar_init(BAR(BS + s), 1); }
        umma::mbar_init(BAR(TFULL), 1); umma::mbar_init(BAR(TFULL + 1), 1);                 // tcgen05.commit
        umma::mbar_init(BAR(TEMPTY), 128); umma::mbar_init(BAR(TEMPTY + 1), 128);           // one epilogue group each
        umma::mbar_init(BAR(AFULL), 1); umma::mbar_init(BAR(AFULL + 1), 1);
        umma::mbar_init(BAR(AEMPTY), 2); umma::mbar_init(BAR(AEMPTY + 1), 2);               // both issuers
        umma::mbar_init(BAR(BIASFULL), 1);
        umma::fence_mbar_init();
    }
    if (warp == 9) umma::tmem_alloc<TMEM_COLS>(umma::smem_u32(tmem_slot));
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if ((bar0 - BK_OFF_BAR) & 1023u) __trap();

    if (warp < 8) {
        // ===================================== epilogue =====================================================
        const int grp = warp >> 2, wq = warp & 3;
        if (grp == 0) {   // constant block scales (see knn2_hamming_mx_kernel)
            const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
#pragma unroll
            for (int c = 0; c < 8; c += 4) {
                tmem_st4(lane_base + SF_ONE + c, 0x7F7F7F7Fu);
                tmem_st4(lane_base + SF_64 + c, 0x85858585u);
                tmem_st4(lane_base + SF_2P14 + c, 0x8D8D8D8Du);
            }
            tmem_wait_st();
            umma::fence_before_sync();
            asm volatile("bar.sync 1, 192;" ::: "memory");                // group 0 + the two UMMA issuer warps
        }
        uint32_t tc = 0;                                                  // tiles issued so far by this CTA (all items)
        int ic = 0;
        for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
            MxItem w;
            if (!mx_item(it, row_tiles, n1_max, n1_arr, n2_max, n2_arr, w)) continue;
            const int last_valid = w.n_cols - (w.T - 1) * TN;
            const int row = w.row0 + wq * 32 + lane;
            uint32_t gb0 = 0u, gb1 = 0u;                                  // running pair in G form, largest first (0 = none)
            for (int j = (int)((tc ^ (uint32_t)grp) & 1u); j < w.T; j += 2) {      // this group's tiles of the item
                const uint32_t g = tc + (uint32_t)j;
                const int t = grp;
                umma::mbar_wait(BAR(TFULL + t), (g >> 1) & 1u);
                umma::fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(t * TN);
                uint32_t pb0[2] = {0u, 0u}, pb1[2] = {0u, 0u};
                uint32_t v[32], u[16];
                umma::tmem_ld32_pack16(taddr, v);
                tmem_ld16_pack16(taddr + 64, u);
                umma::tmem_wait_ld();
                umma::fence_before_sync();
                umma::mbar_arrive(BAR(TEMPTY + t));                       // accumulator stage is free again
                if (j == w.T - 1 && last_valid < TN) {                    // columns past the end of the set: key 0
                    // register i holds columns 2 i and 2 i + 1: mask 0xFFFFFFFF >> 16 x (its columns past the end); shr clamps
                    // amounts above 31, so a register entirely past the end is cleared
                    int lv = last_valid;
                    asm volatile("" : "+r"(lv));              // (opaque: keeps the 48 shift amounts from being hoisted out of the tile loop and spilled)
                    const int sh0 = 32 - 16 * lv;
#pragma unroll
                    for (int i = 0; i < 48; i++) {
                        uint32_t m;
                        asm("shr.b32 %0, %1, %2;" : "=r"(m) : "r"(0xFFFFFFFFu), "r"(max(sh0 + 32 * i, 0)));
                        if (i < 32) v[i] &= m; else u[i - 32] &= m;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) top2max_insert2_u16x2(pb0[(i >> 1) & 1], pb1[(i >> 1) & 1], v[i], v[i + 1]);
#pragma unroll
                for (int i = 0; i < 16; i += 2) top2max_insert2_u16x2(pb0[(i >> 1) & 1], pb1[(i >> 1) & 1], u[i], u[i + 1]);
                top2max_merge_u16x2(pb0[0], pb1[0], pb0[1], pb1[1]);
                // fold into the item's running pair without decoding: a 16-bit key k = [257 - hamming : 9][127 - c : 7] of
                // tile j becomes G = [k : 16][255 - j : 8][low byte of k : 8] (one PRMT).  Inside a tile unsigned order of G is
                // the order of k; the running pair gets its 7 column bits of the HIGH half forced to ones first, so a candidate
                // of a later tile can only win on the distance field, an equal distance keeps the earlier tile (its tile byte
                // is larger), and the true column bits survive in the low byte.
                const uint32_t tsel = (uint32_t)(255 - j);
                gb0 |= 0x007F0000u; gb1 |= 0x007F0000u;
                top2max_insert(gb0, gb1, __byte_perm(pb0[0], tsel, 0x1040));
                top2max_insert(gb0, gb1, __byte_perm(pb0[0], tsel, 0x3242));
                top2max_insert(gb0, gb1, __byte_perm(pb1[0], tsel, 0x1040));
                top2max_insert(gb0, gb1, __byte_perm(pb1[0], tsel, 0x3242));
            }
            tc += (uint32_t)w.T;
            // the two groups' partial results of this item (double-buffered by item parity: the barrier of item ic + 1 orders the
            // next write of a buffer after this read)
            uint2* part = s_part + (ic & 1) * 128;
            gb0 |= 0x007F0000u; gb1 |= 0x007F0000u;                       // different tiles from here on: distance, tile, column
            if (grp == 1) part[wq * 32 + lane] = make_uint2(gb0, gb1);
            asm volatile("bar.sync 2, 256;" ::: "memory");                // the eight epilogue warps
            if (grp == 0) {
                const uint2 o = part[wq * 32 + lane];
                top2max_merge(gb0, gb1, o.x, o.y);
                gb0 = g_to_key(gb0); gb1 = g_to_key(gb1);
                uint32_t* keys_out = w.dir ? key21 + (size_t)w.prob * n2_max * 2 : key12 + (size_t)w.prob * n1_max * 2;
                if (row < w.n_rows) *reinterpret_cast<uint2*>(keys_out + (size_t)row * 2) = make_uint2(gb0, gb1);
            }
            ic++;
        }
    } else if (warp == 8) {
        // ===================================== loader: one thread ============================================
        if (lane == 0) {
            mbar_expect_tx(BAR(BIASFULL), OP_BYTES);
            bulk_g2s(umma::smem_u32(sBias), g_mx_bias_tile, OP_BYTES, BAR(BIASFULL));
            int s = 0;
            uint32_t sphase = 0;                                          // completed passes over the ring, mod 2
            int ic = 0;
            for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
                MxItem w;
                if (!mx_item(it, row_tiles, n1_max, n1_arr, n2_max, n2_arr, w)) continue;
                const uint8_t* g_rows = (w.dir ? e2 : e1) + (size_t)w.prob * rows_pad * 128;
                const uint8_t* g_cols = (w.dir ? e1 : e2) + (size_t)w.prob * rows_pad * 128;
                const int a = ic & 1;
                umma::mbar_wait(BAR(AEMPTY + a), (uint32_t)(((ic >> 1) & 1) ^ 1));       // the item two back has been multiplied
                mbar_expect_tx(BAR(AFULL + a), OP_BYTES);
                bulk_g2s(umma::smem_u32(sA + a * OP_BYTES), g_rows + (size_t)w.row0 * 128, OP_BYTES, BAR(AFULL + a));
                const uint32_t sB0 = umma::smem_u32(sB);
                for (int j = 0; j < w.T; j++) {
                    umma::mbar_wait(BAR(BS + s), sphase ^ 1u);                           // the UMMAs that read this stage have completed
                    mbar_expect_tx(BAR(s), B_STAGE);
                    bulk_g2s(sB0 + s * B_STAGE, g_cols + (size_t)j * B_STAGE, B_STAGE, BAR(s));
                    if (++s == BS) { s = 0; sphase ^= 1u; }
                }
                ic++;
            }
        }
        __syncwarp();
    } else {
        // ===================================== UMMA issuers ================================================
        // One thread issuing every tile's six UMMAs, two waits and two commits was what paced a tile (measured: trimming its
        // address arithmetic alone gained 9 %), so two threads share the tiles: issuer i takes the tiles of parity i, i.e. the
        // accumulator stage i; the tensor pipe serialises their UMMAs anyway and the accumulators are independent.
        const int iss = warp - 9;
        asm volatile("bar.sync 1, 192;" ::: "memory");                    // the scale factors are in tensor memory
        umma::fence_after_sync();
        if (lane == 0) {
            const uint32_t aA = umma::smem_u32(sA), aB = umma::smem_u32(sB), aBias = umma::smem_u32(sBias);
            const uint32_t sf1 = tmem_base + SF_ONE, sf64 = tmem_base + SF_64, sf2p14 = tmem_base + SF_2P14;
            const uint64_t da_b1 = umma::smem_desc(aBias, 16, 1024, umma::LAYOUT_SW128);
            const uint64_t da_b2 = umma::smem_desc(aBias + 32, 16, 1024, umma::LAYOUT_SW128);
            const uint64_t db_b1 = umma::smem_desc(aBias + 64, 16, 1024, umma::LAYOUT_SW128);
            const uint64_t db_b2 = umma::smem_desc(aBias + 96, 16, 1024, umma::LAYOUT_SW128);
            // descriptors differ only in their 14-bit start-address field (bytes >> 4): add the offset to a base descriptor
            const uint64_t dA0 = umma::smem_desc(aA, 16, 1024, umma::LAYOUT_SW128);
            const uint64_t dB0 = umma::smem_desc(aB, 16, 1024, umma::LAYOUT_SW128);
            umma::mbar_wait(BAR(BIASFULL), 0);
            uint32_t tc = 0;
            int ic = 0;
            for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
                MxItem w;
                if (!mx_item(it, row_tiles, n1_max, n1_arr, n2_max, n2_arr, w)) continue;
                const int a = ic & 1;
                umma::mbar_wait(BAR(AFULL + a), (uint32_t)((ic >> 1) & 1));              // this item's row tile
                const uint64_t dA = dA0 + (uint64_t)((a * OP_BYTES) >> 4);
                for (int j = (int)((tc ^ (uint32_t)iss) & 1u); j < w.T; j += 2) {          // this issuer's tiles of the item
                    const uint32_t g = tc + (uint32_t)j;
                    const int s = (int)(g % BS), t = iss;
                    umma::mbar_wait(BAR(s), (g / BS) & 1u);                               // operands of the tile are in shared memory
                    umma::mbar_wait(BAR(TEMPTY + t), ((g >> 1) & 1u) ^ 1u);               // the epilogue has drained this stage
                    umma::fence_after_sync();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(t * TN);
                    const uint64_t dB = dB0 + (uint64_t)((s * B_STAGE) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; k++) mma_mxf4(d_tmem, dA + 2 * k, dB + 2 * k, k > 0 ? 1u : 0u, sf1, sf64);   // + 32 bytes per K step
                    mma_mxf4(d_tmem, da_b1, db_b1, 1u, sf2p14, sf1);      // + 2^23 + 2^14
                    mma_mxf4(d_tmem, da_b2, db_b2, 1u, sf1, sf1);         // + 128 - c
                    umma::commit(BAR(BS + s));
                    umma::commit(BAR(TFULL + t));
                }
                umma::commit(BAR(AEMPTY + a));                             // this issuer is done with the row tile
                tc += (uint32_t)w.T;
                ic++;
            }
        }
        __syncwarp();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        umma::fence_after_sync();
        umma::tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

}  // namespace

// 4-bit tensor-core implementation behind vsb_knn2_hamming_keys (csrc/knn_hamming.cu dispatches on ctx->knn_impl == 3)
int vsb_knn2_hamming_mx(vsb_ctx* ctx, const uint8_t* d1, int n1_max, const int32_t* n1, const uint8_t* d2, int n2_max,
                        const int32_t* n2, int count, uint32_t* key12, uint32_t* key21, int pre, cudaStream_t st) {
    if (!ctx || count < 0 || n1_max < 0 || n2_max < 0) return VSB_ERR_INVALID;
    if (n1_max > (int)KEY_IDX_MASK || n2_max > (int)KEY_IDX_MASK) return VSB_ERR_CAPACITY;
    if (count == 0 || (n1_max == 0 && n2_max == 0)) return VSB_OK;
    if (!ctx->attr_knn_mx_done) {
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_mx_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_mx_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_mx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MX_SMEM));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_mx_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_mx_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
        VSB_CUDA(ctx, cudaFuncSetAttribute(knn2_hamming_mx_bulk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        mx_bias_tile_init_kernel<<<4, 256, 0, st>>>();
        VSB_LAUNCHED(ctx);
        mx_bias_init_kernel<<<4, 256, 0, st>>>();
        VSB_LAUNCHED(ctx);
        VSB_CUDA(ctx, cudaStreamSynchronize(st));      // first use only: the constant tables must exist before ANY stream reads them
        ctx->attr_knn_mx_done = 1;
    }
    const int row_tiles = vsb_div_up(n1_max > n2_max ? n1_max : n2_max, TM);
    for (int z0 = 0; z0 < count; z0 += 65535) {
        const int zc = count - z0 < 65535 ? count - z0 : 65535;
        dim3 grid(row_tiles, 2, zc);
        const uint8_t* a = d1 + (size_t)z0 * n1_max * 32;
        const uint8_t* b = d2 + (size_t)z0 * n2_max * 32;
        uint32_t* k12 = key12 + (size_t)z0 * n1_max * 2;
        uint32_t* k21 = key21 + (size_t)z0 * n2_max * 2;
        ProfScope ps(ctx, VSB_K_KNN_HAMMING, st);
        if (pre == 2) {
            const int rows_pad = ((n1_max > n2_max ? n1_max : n2_max) + 383) / 384 * 384;
            const size_t bb = (size_t)zc * rows_pad * 128;
            // a SEQUENCE (set 2 of pair k is set 1 of pair k + 1: d2 = d1 + one set, the tracker's sequence entries) has every
            // frame's descriptors expanded once instead of once per pair side
            const bool chained = n1_max == n2_max && d2 == d1 + (size_t)n1_max * 32 && ((!n1 && !n2) || (n1 && n2 == n1 + 1));
            void* scratch = nullptr;
            int rc = vsb_stream_ws_reserve(ctx, st, (chained ? bb + (size_t)rows_pad * 128 : 2 * bb) + 256, &scratch);
            if (rc) return rc;
            uint8_t* e1 = static_cast<uint8_t*>(scratch);
            uint8_t* e2 = chained ? e1 + (size_t)rows_pad * 128 : e1 + bb;
            if (chained) {
                const long long chunks = (long long)(zc + 1) * rows_pad * 8;
                mx_expand_swizzled_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(a, n1_max, n1 ? n1 + z0 : nullptr, rows_pad, zc + 1,
                                                                                             reinterpret_cast<uint4*>(e1));
                VSB_LAUNCHED(ctx);
            } else {
                const long long chunks = (long long)zc * rows_pad * 8;
                mx_expand_swizzled_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(a, n1_max, n1 ? n1 + z0 : nullptr, rows_pad, zc,
                                                                                             reinterpret_cast<uint4*>(e1));
                VSB_LAUNCHED(ctx);
                mx_expand_swizzled_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(b, n2_max, n2 ? n2 + z0 : nullptr, rows_pad, zc,
                                                                                             reinterpret_cast<uint4*>(e2));
                VSB_LAUNCHED(ctx);
            }
            const long long n_items = (long long)zc * 2 * row_tiles;
            const int resident = 2 * (ctx->sm_count > 0 ? ctx->sm_count : 148);
            const int ctas = (int)(n_items < resident ? n_items : resident);
            knn2_hamming_mx_bulk_kernel<<<ctas, BK_THREADS, BK_SMEM, st>>>(e1, n1_max, n1 ? n1 + z0 : nullptr, e2, n2_max,
                                                                           n2 ? n2 + z0 : nullptr, k12, k21, rows_pad, row_tiles, n_items);
        } else if (pre) {
            const size_t b1 = ((size_t)zc * n1_max * 128 + 255) & ~(size_t)255, b2 = ((size_t)zc * n2_max * 128 + 255) & ~(size_t)255;
            void* scratch = nullptr;
            int rc = vsb_stream_ws_reserve(ctx, st, b1 + b2 + 256, &scratch);
            if (rc) return rc;
            uint8_t* e1 = static_cast<uint8_t*>(scratch);
            uint8_t* e2 = e1 + b1;
            const long long c1 = (long long)zc * n1_max * 8, c2 = (long long)zc * n2_max * 8;
            if (c1) { mx_expand_kernel<<<(unsigned)((c1 + 255) / 256), 256, 0, st>>>(a, c1, reinterpret_cast<uint4*>(e1)); VSB_LAUNCHED(ctx); }
            if (c2) { mx_expand_kernel<<<(unsigned)((c2 + 255) / 256), 256, 0, st>>>(b, c2, reinterpret_cast<uint4*>(e2)); VSB_LAUNCHED(ctx); }
            knn2_hamming_mx_kernel<1><<<grid, MX_THREADS, MX_SMEM, st>>>(e1, n1_max, n1 ? n1 + z0 : nullptr, e2, n2_max,
                                                                         n2 ? n2 + z0 : nullptr, k12, k21, 0);
        } else {
            knn2_hamming_mx_kernel<0><<<grid, MX_THREADS, MX_SMEM, st>>>(a, n1_max, n1 ? n1 + z0 : nullptr, b, n2_max,
                                                                         n2 ? n2 + z0 : nullptr, k12, k21, 0);
        }
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

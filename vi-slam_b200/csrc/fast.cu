// fast.cu — FAST-9/16 corner detection with non-maximum suppression on the device, batched over frames: the key-point
// stage of the feature detectors the reference calls before the tracking path (cv::ORB, src/Camera.cpp:124-129, and
// cv::cuda::ORB, src/CameraGPU.cpp:99-104, both detect with FAST-9/16, threshold 20) — SURVEY.md 8f row N-4, first stage.
// Semantics are cv::FAST(TYPE_9_16)'s, restated in oracle/fast.c and pinned there against cv2: corner test, score
// (largest threshold that keeps the pixel a corner, minus one), strict 3x3 suppression, row-major output order.
//
//   fast_score_kernel    one 64x32 tile (+halo) per CTA in shared memory (one 16-byte vector per thread where the frame allows
//                        it), 4 adjacent pixels x 2 rows per thread.  The flag pass is the NECESSARY condition only — two
//                        neighbouring compass points of the ring both brighter or both darker, 8 comparisons on packed bytes
//                        (SWAR) — and the pixels that pass (a few percent up to a fifth, by texture) are queued with one
//                        shared-memory atomic per warp and get the full arc test from their score, one thread per candidate,
//                        both polarities in the two 16-bit lanes of a register.  Writes score + 1 per pixel (0 = no corner).
//   fast_count16_kernel  one warp per image row, 16 pixels per lane: the 3x3 suppression runs on the (few) corners only and leaves
//                        one bit per pixel next to the row's corner count (score rows are padded to 16 bytes).
//   fast_scan_kernel     exclusive scan of the row counts of each frame (row-major order needs the offsets).
//   fast_write16_kernel  reads the bits, warp scan, ordered write of (x, y, score) while the slot is below the caller's capacity.
//   fast_count4_kernel / fast_write4_kernel: the previous compaction (4 pixels per lane, suppression in both passes; fast_impl 1).
#include "common.cuh"

namespace {


// ring offsets (x, y) of the radius-3 Bresenham circle, clockwise from (0, 3) — cv::FAST's order.  Kept as local constant
// arrays inside the kernel: after unrolling they fold into immediate shared-memory offsets.
#define FAST_RING_X {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}
#define FAST_RING_Y {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3}

__device__ __forceinline__ bool has_arc9(uint32_t m16) {      // 9 contiguous set bits in the circular 16-bit mask
    const uint32_t x = m16 | (m16 << 16);
    uint32_t t = x & (x >> 1);
    t &= t >> 2;
    t &= t >> 4;                                               // bit i: bits i..i+7 set
    t &= x >> 8;                                               // ... and bit i+8
    return (t & 0xFFFFu) != 0u;
}

// max over the 16 arcs of the minimum of a[k..k+8] (indices mod 16)
__device__ __forceinline__ int best_arc_min(const int (&a)[16]) {
    int m2[16], m4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m2[k] = min(a[k], a[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m4[k] = min(m2[k], m2[(k + 2) & 15]);
    int best = -(1 << 30);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int m8 = min(m4[k], m4[(k + 4) & 15]);
        best = max(best, min(m8, a[(k + 8) & 15]));
    }
    return best;
}

// Byte-wise unsigned compare of four packed pixels, result in bit 7 of every byte (the other bits are garbage):
//   gt7(a, b):  a_j > b_j.   The carry out of bit 6 of (a & 0x7f) + (~b & 0x7f) tells a7 > b7 for the low 7 bits, and the
//   majority-like LOP3 0xb2 folds in the top bits — the first four instructions of the compiler's own __vcmpgtu4 (which
//   then spends one more to smear bit 7 over the byte; the arc test below only needs the bit).
__device__ __forceinline__ uint32_t gt7(uint32_t a, uint32_t b) {
    uint32_t r;
    const uint32_t t = (a & 0x7f7f7f7fu) + (~b & 0x7f7f7f7fu);
    asm("lop3.b32 %0, %1, %2, %3, 0xb2;" : "=r"(r) : "r"(a), "r"(b), "r"(t));
    return r;
}

// One 64x32 tile per CTA, FOUR horizontally adjacent pixels x TWO rows (ty, ty + 16) per thread, all ring comparisons on
// packed bytes.  Shared tile: 38 rows x 72 bytes (x0 - 4 .. x0 + 67), so a thread's centre word is aligned and the ring
// positions of its four pixels are funnel shifts of aligned words; rows are 48 words apart so that the two tile rows a warp
// touches fall into disjoint banks.  The loader requests a thread's (up to three) words before it stores any of them, so
// their latencies overlap; the tile's height halves the halo share (38 rows read for 32) and the per-thread set-up is spread
// over eight pixels instead of four.
// Pixels that pass the flag pass (a percent or so up to a fifth, by texture) are queued in shared memory and scored densely,
// one thread per candidate, after the tile's flag pass — inside the flag pass the scalar score would run for whole warps
// whenever one lane has a candidate.
constexpr int FT_W = 64, FT_H = 32, FT_HALO = 3, FT_SH = FT_H + 2 * FT_HALO;
constexpr int FT_WORDS = 18, FT_PITCH = 48;
constexpr int FT_LOADS = (FT_SH * FT_WORDS + 255) / 256;              // words per thread of the loader (3)

// VEC: frames whose base, stride and pitch are multiples of 16 (every level image of the scale pyramid, 752- and 640-wide
// frames): the tile is loaded as 16-byte vectors from x0 - 16 on, 6 per row, ONE per thread, instead of three words per thread
// with their index arithmetic and byte-wise border paths (a third of the kernel's instructions); bytes between w and the pitch
// are read as they are — they only ever reach the 3-pixel border, which is never a corner.
template <bool VEC>
__global__ void __launch_bounds__(256, 4)
fast_score_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, int w, int h, int threshold,
                  uint8_t* __restrict__ score1, int sw) {
    constexpr int WOFF = VEC ? 4 : 1;                                   // word column of the tile's first pixel
    __shared__ __align__(16) uint32_t s[FT_SH][FT_PITCH];
    __shared__ uint16_t s_list[FT_W * FT_H];                            // tile-local (y << 6 | x) of the candidates found
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    const int frame = blockIdx.z;
    const uint8_t* in = img + (size_t)frame * img_stride;
    uint8_t* out = score1 + (size_t)frame * sw * h;      // score rows are sw = round_up(w, 16) bytes apart; the pad is 0
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
    if (VEC) {
        if (threadIdx.x < FT_SH * 6) {
            const int py = threadIdx.x / 6, pc = threadIdx.x - 6 * py;
            const int gx = x0 - 16 + 16 * pc, gy = y0 + py - FT_HALO;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (gy >= 0 && gy < h && gx >= 0 && gx + 15 < pitch) v = __ldg(reinterpret_cast<const uint4*>(in + (size_t)gy * pitch + gx));
            *reinterpret_cast<uint4*>(&s[py][4 * pc]) = v;
        }
    } else {
    const bool aligned_in = ((reinterpret_cast<uintptr_t>(in) | (uintptr_t)pitch) & 3u) == 0;
        uint32_t v[FT_LOADS];
        int slot[FT_LOADS];
#pragma unroll
        for (int q = 0; q < FT_LOADS; q++) {
            const int p = threadIdx.x + 256 * q;
            const int py = (p * 3641) >> 16, pw = p - py * FT_WORDS;   // p / 18 for p < 1024
            const int gx = x0 - 4 + 4 * pw, gy = y0 + py - FT_HALO;
            slot[q] = p < FT_SH * FT_WORDS ? py * FT_PITCH + pw : -1;
            v[q] = 0u;
            if (slot[q] >= 0 && gy >= 0 && gy < h) {
                const uint8_t* row = in + (size_t)gy * pitch;
                if (aligned_in && gx >= 0 && gx + 3 < w) {
                    v[q] = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (gx + j >= 0 && gx + j < w) v[q] |= (uint32_t)__ldg(row + gx + j) << (8 * j);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < FT_LOADS; q++)
            if (slot[q] >= 0) (&s[0][0])[slot[q]] = v[q];
    }
    __syncthreads();
    const int ty0 = threadIdx.x >> 4, wc = (threadIdx.x & 15) + WOFF;  // first row in the tile, word column of the centre word
    const int x = x0 + 4 * (wc - WOFF);
    const uint32_t T4 = (uint32_t)threshold * 0x01010101u;
    // Flag pass = the necessary condition only.  Any 9 contiguous ring positions contain two NEIGHBOURING compass points
    // (ring positions 0, 4, 8, 12), so a corner needs two neighbouring compass pixels that are both brighter than centre + t or
    // both darker than centre - t.  That test needs 4 of the 16 ring positions (rows y-3, y, y+3 only); the pixels that pass are
    // queued and get the full arc test from their score.
    uint32_t inside = 0u;                                              // the 3-pixel border is never a corner
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (x + j >= 3 && x + j < w - 3) inside |= 0x80u << (8 * j);
    uint32_t cand = 0u;                                                // bit 8 j + half: pixel j of row ty0 + 16 half passed
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int ty = ty0 + 16 * half, y = y0 + ty;
        const bool in_image = x < w && y < h;
        if (in_image && y >= 3 && y < h - 3) {
            const uint32_t C = s[ty + 3][wc];
            const uint32_t hi = __vaddus4(C, T4), lo = __vsubus4(C, T4);   // saturating: a ring byte can never beat 255 / 0
            const uint32_t up = s[ty + 6][wc], dn = s[ty][wc];            // ring positions 0 (0, +3) and 8 (0, -3)
            const uint32_t rt = __funnelshift_r(C, s[ty + 3][wc + 1], 24);                  // position 4 (+3, 0)
            const uint32_t lf = __funnelshift_r(s[ty + 3][wc - 1], C, 8);                   // position 12 (-3, 0)
            // gt7 with the parts that depend on the centre only formed once: five instructions per ring position for both tests
            const uint32_t nh = ~hi & 0x7f7f7f7fu, lm = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
            uint32_t bq[4], dq[4];
            const uint32_t ring4[4] = {up, rt, dn, lf};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t m = ring4[q] & 0x7f7f7f7fu;
                const uint32_t tb = m + nh, td = lm - m;               // (ring & 0x7f) + (~hi & 0x7f);  (lo & 0x7f) + (~ring & 0x7f)
                asm("lop3.b32 %0, %1, %2, %3, 0xb2;" : "=r"(bq[q]) : "r"(ring4[q]), "r"(hi), "r"(tb));
                asm("lop3.b32 %0, %1, %2, %3, 0xb2;" : "=r"(dq[q]) : "r"(lo), "r"(ring4[q]), "r"(td));
            }
            const uint32_t b0 = bq[0], b4 = bq[1], b8 = bq[2], b12 = bq[3];
            const uint32_t d0 = dq[0], d4 = dq[1], d8 = dq[2], d12 = dq[3];
            const uint32_t bb = ((b0 | b8) & (b4 | b12));                  // (b0&b4)|(b4&b8)|(b8&b12)|(b12&b0)
            const uint32_t dd = ((d0 | d8) & (d4 | d12));
            cand |= ((bb | dd) & inside) >> (7 - half);
        }
        // every pixel gets its byte now (0 = no corner); the candidates are queued for the arc test / score pass below
        if (x < sw && y < h) *reinterpret_cast<uint32_t*>(out + (size_t)y * sw + x) = 0u;   // x is a multiple of 4 and x + 3 < sw; the row's padding is zeroed too
    }
    // Queue the candidates: ONE shared-memory atomic per warp (inclusive scan of the per-thread counts), then every thread
    // writes its own.  (One atomic per pixel position cost eight warp-aggregated atomics per thread as soon as one lane of the
    // warp had a candidate there — with a tenth of the pixels passing, always.)  The order of the queue does not matter: every
    // entry is scored on its own and written to its own pixel.
    {
        const int lane = threadIdx.x & 31;
        const int n = __popc(cand);
        int incl = n;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            int base = 0;
            if (lane == 31) base = atomicAdd(&s_count, total);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - n;
            while (cand) {
                const int b = __ffs(cand) - 1;
                cand &= cand - 1u;
                s_list[base++] = (uint16_t)(((ty0 + 16 * (b & 1)) << 6) | (4 * (wc - WOFF) + (b >> 3)));
            }
        }
    }
    __syncthreads();
    // ---- score pass: one thread per queued candidate (dense), ring bytes from the shared tile ------------------------------
    // Both polarities at once in the two 16-bit lanes of a register, biased by 256 so that the lanes stay unsigned:
    // low = 256 + centre - ring, high = 256 + ring - centre = kv + ring * 0xFFFF (the low lane stays >= 1, so no borrow crosses
    // the lanes).  min over the 9 positions of an arc as min3(min3, min3, min3) of shared triples, max over the 16 arcs.
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(&s[0][0]);
    const int ring_x[16] = FAST_RING_X, ring_y[16] = FAST_RING_Y;
    for (int e = threadIdx.x; e < s_count; e += 256) {
        const int ly = s_list[e] >> 6, lx = s_list[e] & 63;
        const uint8_t* c = sb + (size_t)(ly + FT_HALO) * (FT_PITCH * 4) + lx + 4 * WOFF;
        const uint32_t v = c[0];
        const uint32_t kv = (v + 256u) + ((256u - v) << 16);
        uint32_t a[16], t3[16];
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = kv + (uint32_t)c[ring_y[k] * (FT_PITCH * 4) + ring_x[k]] * 0xFFFFu;
#pragma unroll
        for (int k = 0; k < 16; k++) t3[k] = __vminu2(__vminu2(a[k], a[(k + 1) & 15]), a[(k + 2) & 15]);
        uint32_t best2 = 0u;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const uint32_t arc0 = __vminu2(__vminu2(t3[k], t3[(k + 3) & 15]), t3[(k + 6) & 15]);
            const uint32_t arc1 = __vminu2(__vminu2(t3[k + 1], t3[(k + 4) & 15]), t3[(k + 7) & 15]);
            best2 = __vmaxu2(__vmaxu2(arc0, arc1), best2);
        }
        // largest threshold for which the pixel still has a 9-arc, plus one: the pixel is a corner iff that exceeds the
        // threshold (cv::FAST's corner test and cornerScore agree by construction); stored as score + 1, in 1..255
        const int best = max((int)(best2 & 0xFFFFu), (int)(best2 >> 16)) - 256;
        if (best > threshold) out[(size_t)(y0 + ly) * sw + x0 + lx] = (uint8_t)best;
    }
}

__global__ void __launch_bounds__(256)
fast_scan_kernel(const int32_t* __restrict__ row_count, int h, int32_t* __restrict__ row_offset, int32_t* __restrict__ n_kp) {
    const int frame = blockIdx.x;
    const int32_t* c = row_count + (size_t)frame * h;
    int32_t* o = row_offset + (size_t)frame * h;
    __shared__ int s_warp[8];
    const int per = (h + 255) / 256;
    const int b = threadIdx.x * per, e = min(b + per, h);
    int t = 0;
    for (int i = b; i < e; i++) t += c[i];
    // exclusive scan of the 256 partial sums (warp scans; a walk by one thread was 256 dependent shared-memory accesses)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w8 = lane < 8 ? s_warp[lane] : 0;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w8, off); if (lane >= off) w8 += v; }
        if (lane < 8) s_warp[lane] = w8;
    }
    __syncthreads();
    if (threadIdx.x == 0) n_kp[frame] = s_warp[7];
    int acc = (warp ? s_warp[warp - 1] : 0) + inc - t;
    for (int i = b; i < e; i++) { o[i] = acc; acc += c[i]; }
}

// ---- compaction, 4 pixels per lane (score rows are 4-byte aligned: their stride is round_up(w, 4)) -----------------------------------------------
// One warp per image row.  A lane looks at 4 consecutive pixels: their score bytes and the words left / right of them for
// the rows y-1, y, y+1; the 3x3 neighbour maximum and the "strictly greater" test run on all four bytes at once
// (__vmaxu4 / __vcmpgtu4).  score1 bytes are score + 1 with 0 = no corner, so "keep" is  cur > max(neighbours, 1).
// Words that straddle the row ends only ever pick up pixels of the 3-pixel border, which are never corners (0).
__device__ __forceinline__ uint32_t row_max3(uint32_t prev, uint32_t cur, uint32_t next, bool centre) {
    const uint32_t L = (cur << 8) | (prev >> 24);          // byte j = pixel j - 1
    const uint32_t R = (cur >> 8) | (next << 24);          // byte j = pixel j + 1
    const uint32_t m = __vmaxu4(L, R);
    return centre ? __vmaxu4(m, cur) : m;
}
// keep mask of the four pixels of word k whose score bytes are `cur` (already loaded, non-zero)
__device__ __forceinline__ uint32_t keep_of4(const uint32_t* __restrict__ r0, const uint32_t* __restrict__ r1,
                                             const uint32_t* __restrict__ r2, int k, int nonmax, uint32_t cur) {
    if (!nonmax) return __vcmpgtu4(cur, 0u);
    uint32_t n = row_max3(__ldg(r1 + k - 1), cur, __ldg(r1 + k + 1), false);
    n = __vmaxu4(n, row_max3(__ldg(r0 + k - 1), __ldg(r0 + k), __ldg(r0 + k + 1), true));
    n = __vmaxu4(n, row_max3(__ldg(r2 + k - 1), __ldg(r2 + k), __ldg(r2 + k + 1), true));
    return __vcmpgtu4(cur, __vmaxu4(n, 0x01010101u));
}

__global__ void __launch_bounds__(256)
fast_count4_kernel(const uint8_t* __restrict__ score1, int w, int h, int nonmax, int32_t* __restrict__ row_count) {
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5), frame = blockIdx.y, lane = threadIdx.x & 31;
    if (y >= h) return;
    int n = 0;
    if (y >= 3 && y < h - 3) {
        const uint8_t* sc = score1 + (size_t)frame * w * h;
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(sc + (size_t)(y - 1) * w);
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(sc + (size_t)y * w);
        const uint32_t* r2 = reinterpret_cast<const uint32_t*>(sc + (size_t)(y + 1) * w);
        // four words per lane and round, requested together (a row is a handful of dependent-free loads per lane); words
        // without a corner — almost all — cost nothing more
        const int words = w >> 2;
        for (int k0 = 0; k0 < words; k0 += 128) {
            uint32_t cur[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = k0 + 32 * j + lane;
                cur[j] = k < words ? __ldg(r1 + k) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (cur[j] != 0u) n += __popc(keep_of4(r0, r1, r2, k0 + 32 * j + lane, nonmax, cur[j]) & 0x01010101u);
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n += __shfl_down_sync(0xffffffffu, n, off);
    if (lane == 0) row_count[(size_t)frame * h + y] = n;
}

__global__ void __launch_bounds__(256)
fast_write4_kernel(const uint8_t* __restrict__ score1, int w, int h, int nonmax, const int32_t* __restrict__ row_offset,
                   const int32_t* __restrict__ row_count, int cap, int32_t* __restrict__ kp_xy, int32_t* __restrict__ kp_score) {
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5), frame = blockIdx.y, lane = threadIdx.x & 31;
    if (y < 3 || y >= h - 3) return;
    if (row_count[(size_t)frame * h + y] == 0) return;     // about half of the rows of a natural image hold no corner
    int base = row_offset[(size_t)frame * h + y];
    if (base >= cap) return;
    const uint8_t* sc = score1 + (size_t)frame * w * h;
    const uint32_t* r0 = reinterpret_cast<const uint32_t*>(sc + (size_t)(y - 1) * w);
    const uint32_t* r1 = reinterpret_cast<const uint32_t*>(sc + (size_t)y * w);
    const uint32_t* r2 = reinterpret_cast<const uint32_t*>(sc + (size_t)(y + 1) * w);
    int32_t* oxy = kp_xy + (size_t)frame * cap * 2;
    int32_t* osc = kp_score + (size_t)frame * cap;
    const int words = w >> 2;
    for (int k0 = 0; k0 < words; k0 += 32) {
        const int k = k0 + lane;
        uint32_t cur = 0u, keep = 0u;
        if (k < words) cur = __ldg(r1 + k);
        if (cur != 0u) keep = keep_of4(r0, r1, r2, k, nonmax, cur);
        if (!__any_sync(0xffffffffu, keep != 0u)) continue;
        const int cnt = __popc(keep & 0x01010101u);
        int incl = cnt;                                    // inclusive warp scan of the per-lane counts
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        int slot = base + incl - cnt;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((keep >> (8 * j)) & 1u) {
                if (slot < cap) {
                    oxy[2 * slot] = 4 * k + j; oxy[2 * slot + 1] = y;
                    osc[slot] = nonmax ? (int)((cur >> (8 * j)) & 0xFFu) - 1 : 0;
                }
                slot++;
            }
        base += __shfl_sync(0xffffffffu, incl, 31);
        if (base >= cap) return;
    }
}

// ---- compaction, 16 pixels per lane (score rows 16 bytes aligned) -------------------------------------------------------
// The suppression runs ONCE: the count pass loads 16 score bytes per lane (almost always zero: nothing more to do), runs
// keep_of4 on the non-zero words and leaves one bit per pixel (a uint16 per 16 pixels) next to the row's count; the write pass
// reads the bits only, plus the score byte of every corner it writes.
__global__ void __launch_bounds__(256)
fast_count16_kernel(const uint8_t* __restrict__ score1, int sw, int h, int nonmax, int32_t* __restrict__ row_count,
                    uint16_t* __restrict__ bits) {
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5), frame = blockIdx.y, lane = threadIdx.x & 31;
    if (y >= h) return;
    int n = 0;
    if (y >= 3 && y < h - 3) {
        const uint8_t* sc = score1 + (size_t)frame * sw * h;
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(sc + (size_t)(y - 1) * sw);
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(sc + (size_t)y * sw);
        const uint32_t* r2 = reinterpret_cast<const uint32_t*>(sc + (size_t)(y + 1) * sw);
        const uint4* v1 = reinterpret_cast<const uint4*>(r1);
        const int vecs = sw >> 4;
        uint16_t* brow = bits + ((size_t)frame * h + y) * vecs;
        for (int q0 = 0; q0 < vecs; q0 += 64) {
            uint4 c[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int q = q0 + 32 * j + lane;
                c[j] = q < vecs ? __ldg(v1 + q) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int q = q0 + 32 * j + lane;
                if (q >= vecs) continue;
                const uint32_t cw[4] = {c[j].x, c[j].y, c[j].z, c[j].w};
                uint32_t nz = 0u;                                  // one bit per non-zero score byte
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t b7 = (cw[t] | ((cw[t] & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
                    nz |= ((((b7 >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xFu) << (4 * t);
                }
                // corners are a few percent of the pixels: each one is tested on its own bytes (nine byte loads that hit L1)
                // instead of running the packed 3x3 maximum on every word that holds one
                uint32_t m = 0u;
                const uint8_t* b0 = reinterpret_cast<const uint8_t*>(r0);
                const uint8_t* b1 = reinterpret_cast<const uint8_t*>(r1);
                const uint8_t* b2 = reinterpret_cast<const uint8_t*>(r2);
                while (nz) {
                    const int b = __ffs(nz) - 1;
                    nz &= nz - 1u;
                    const int x = 16 * q + b;
                    bool keep = true;
                    if (nonmax) {
                        const int cur = __ldg(b1 + x);
                        const int n0 = max(max((int)__ldg(b0 + x - 1), (int)__ldg(b0 + x)), (int)__ldg(b0 + x + 1));
                        const int n1 = max((int)__ldg(b1 + x - 1), (int)__ldg(b1 + x + 1));
                        const int n2 = max(max((int)__ldg(b2 + x - 1), (int)__ldg(b2 + x)), (int)__ldg(b2 + x + 1));
                        keep = cur > max(max(n0, n1), max(n2, 1));
                    }
                    if (keep) m |= 1u << b;
                }
                brow[q] = (uint16_t)m;
                n += __popc(m);
            }
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n += __shfl_down_sync(0xffffffffu, n, off);
    if (lane == 0) row_count[(size_t)frame * h + y] = n;
}

__global__ void __launch_bounds__(256)
fast_write16_kernel(const uint8_t* __restrict__ score1, int sw, int h, int nonmax, const int32_t* __restrict__ row_offset,
                    const int32_t* __restrict__ row_count, const uint16_t* __restrict__ bits, int cap,
                    int32_t* __restrict__ kp_xy, int32_t* __restrict__ kp_score) {
    const int y = blockIdx.x * 8 + (threadIdx.x >> 5), frame = blockIdx.y, lane = threadIdx.x & 31;
    if (y < 3 || y >= h - 3) return;
    if (row_count[(size_t)frame * h + y] == 0) return;     // about half of the rows of a natural image hold no corner
    int base = row_offset[(size_t)frame * h + y];
    if (base >= cap) return;
    const uint8_t* srow = score1 + ((size_t)frame * h + y) * sw;
    const int vecs = sw >> 4;
    const uint16_t* brow = bits + ((size_t)frame * h + y) * vecs;
    int32_t* oxy = kp_xy + (size_t)frame * cap * 2;
    int32_t* osc = kp_score + (size_t)frame * cap;
    for (int q0 = 0; q0 < vecs; q0 += 32) {
        const int q = q0 + lane;
        uint32_t m = q < vecs ? (uint32_t)brow[q] : 0u;
        if (!__any_sync(0xffffffffu, m != 0u)) continue;
        const int cnt = __popc(m);
        int incl = cnt;                                    // inclusive warp scan of the per-lane counts
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        int slot = base + incl - cnt;
        while (m) {
            const int x = 16 * q + __ffs(m) - 1;
            m &= m - 1u;
            if (slot < cap) {
                oxy[2 * slot] = x; oxy[2 * slot + 1] = y;
                osc[slot] = nonmax ? (int)__ldg(srow + x) - 1 : 0;
            }
            slot++;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
        if (base >= cap) return;
    }
}

}  // namespace

// Scratch of one vsb_fast_detect_ws call: the score image (rows padded to 16 bytes), two int32 per image row, one bit per pixel.
size_t vsb_fast_scratch_bytes(int w, int h, int count) {
    const int sw = (w + 15) & ~15;
    const size_t img_bytes = ((size_t)count * sw * h + 255) & ~(size_t)255;
    const size_t rows_bytes = ((size_t)count * h * sizeof(int32_t) + 255) & ~(size_t)255;
    const size_t bits_bytes = ((size_t)count * h * (sw >> 4) * sizeof(uint16_t) + 255) & ~(size_t)255;
    return img_bytes + 2 * rows_bytes + bits_bytes + 256;
}

// The detector on a caller-provided scratch block (at least vsb_fast_scratch_bytes, 256-byte aligned): what the tracker's
// per-slot workspaces use, so that two chunks of a sequence can be in flight on two streams.
int vsb_fast_detect_ws(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count, int threshold,
                       int nonmax, int cap, int32_t* kp_xy, int32_t* kp_score, int32_t* n_kp, void* scratch, void* stream) {
    if (!ctx || !img || !kp_xy || !kp_score || !n_kp || !scratch) return VSB_ERR_INVALID;
    if (w <= 0 || h <= 0 || pitch < w || count < 0 || cap < 0 || threshold < 0 || threshold > 255) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // score rows are padded to a multiple of 16 bytes (zero pad = "no corner"), so the 16-pixels-per-lane compaction kernels
    // serve every width; words that straddle a row end only ever see border / pad zeros
    const int sw = (w + 15) & ~15;
    const size_t img_bytes = ((size_t)count * sw * h + 255) & ~(size_t)255;
    const size_t rows_bytes = ((size_t)count * h * sizeof(int32_t) + 255) & ~(size_t)255;
    uint8_t* score1 = static_cast<uint8_t*>(scratch);
    int32_t* row_count = reinterpret_cast<int32_t*>(score1 + img_bytes);
    int32_t* row_offset = reinterpret_cast<int32_t*>(score1 + img_bytes + rows_bytes);
    uint16_t* bits = reinterpret_cast<uint16_t*>(score1 + img_bytes + 2 * rows_bytes);
    const bool per_word = ctx->fast_impl == 1;               // the previous compaction (suppression in both passes), kept for comparison
    for (int z0 = 0; z0 < count; z0 += 65535) {
        const int zc = count - z0 < 65535 ? count - z0 : 65535;
        const uint8_t* in = img + (size_t)z0 * img_stride;
        uint8_t* sc = score1 + (size_t)z0 * sw * h;
        uint16_t* bz = bits + (size_t)z0 * h * (sw >> 4);
        {
            ProfScope ps(ctx, VSB_K_FAST_SCORE, st);
            const bool vec = ((reinterpret_cast<uintptr_t>(in) | (uintptr_t)img_stride | (uintptr_t)pitch) & 15u) == 0 && ctx->fast_impl != 1;
            const dim3 grid(vsb_div_up(w, FT_W), vsb_div_up(h, FT_H), zc);
            if (vec) fast_score_kernel<true><<<grid, 256, 0, st>>>(in, img_stride, pitch, w, h, threshold, sc, sw);
            else fast_score_kernel<false><<<grid, 256, 0, st>>>(in, img_stride, pitch, w, h, threshold, sc, sw);
            VSB_LAUNCHED(ctx);
        }
        ProfScope ps(ctx, VSB_K_FAST_COMPACT, st);
        if (per_word) fast_count4_kernel<<<dim3(vsb_div_up(h, 8), zc), 256, 0, st>>>(sc, sw, h, nonmax, row_count + (size_t)z0 * h);
        else fast_count16_kernel<<<dim3(vsb_div_up(h, 8), zc), 256, 0, st>>>(sc, sw, h, nonmax, row_count + (size_t)z0 * h, bz);
        VSB_LAUNCHED(ctx);
        fast_scan_kernel<<<zc, 256, 0, st>>>(row_count + (size_t)z0 * h, h, row_offset + (size_t)z0 * h, n_kp + z0);
        VSB_LAUNCHED(ctx);
        if (cap > 0) {
            if (per_word)
                fast_write4_kernel<<<dim3(vsb_div_up(h, 8), zc), 256, 0, st>>>(sc, sw, h, nonmax, row_offset + (size_t)z0 * h,
                                                                               row_count + (size_t)z0 * h, cap,
                                                                               kp_xy + (size_t)z0 * cap * 2, kp_score + (size_t)z0 * cap);
            else
                fast_write16_kernel<<<dim3(vsb_div_up(h, 8), zc), 256, 0, st>>>(sc, sw, h, nonmax, row_offset + (size_t)z0 * h,
                                                                                row_count + (size_t)z0 * h, bz, cap,
                                                                                kp_xy + (size_t)z0 * cap * 2, kp_score + (size_t)z0 * cap);
            VSB_LAUNCHED(ctx);
        }
    }
    return VSB_OK;
}

// FAST-9/16 corners of `count` frames (cv::FAST, TYPE_9_16).  img: frames of h rows x pitch bytes, img_stride bytes apart.
// kp_xy [count][cap][2] int32 (x, y) and kp_score [count][cap] int32 in row-major order; n_kp [count] = corners FOUND
// (may exceed cap; only the first cap are stored).  Scratch comes from the context.
extern "C" int vsb_fast_detect(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count,
                               int threshold, int nonmax, int cap, int32_t* kp_xy, int32_t* kp_score, int32_t* n_kp,
                               void* stream) {
    if (!ctx || !img || !kp_xy || !kp_score || !n_kp) return VSB_ERR_INVALID;
    if (w <= 0 || h <= 0 || pitch < w || count < 0 || cap < 0 || threshold < 0 || threshold > 255) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    void* scratch = nullptr;
    int rc = vsb_scratch_reserve(ctx, vsb_fast_scratch_bytes(w, h, count), &scratch);
    if (rc) return rc;
    return vsb_fast_detect_ws(ctx, img, img_stride, pitch, w, h, count, threshold, nonmax, cap, kp_xy, kp_score, n_kp, scratch, stream);
}

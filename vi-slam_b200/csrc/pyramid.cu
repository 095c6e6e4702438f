// pyramid.cu — per-frame front end on the device: image pyramid, Scharr gradients, candidate patch points.
// Replaces Camera::Update (reference src/Camera.cpp:63-72), Camera::computeGradient (:167-184) and
// Camera::ObtainPatchesPointsPreviousFrame (:358-409).  All integer-exact, so results are bit-identical
// to cv::resize(0.5) / cv::Scharr(scale 3) (checked against cv2 through the oracle).
#include "common.cuh"
#include "se3.cuh"

int vsb_pyramid_build_levels(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int count,
                             const vsb_pyr_layout_t* layout, uint8_t* pyr, int copy_l0, void* stream);

namespace {

// ---------------------------------------------------------------------------------------------- pyramid
// One CTA cascades a 64x64 level-0 tile through all levels in shared memory: level l+1 pixel (x,y) is
// the rounded mean of the level-l block (2x..2x+1, 2y..2y+1) clipped to the image, which is what
// cv::resize(src, dst, Size(), 0.5, 0.5) computes (INTER_LINEAR at scale exactly 2 takes OpenCV's
// area fast path; clipped blocks are averaged in float and rounded half-to-even).
constexpr int PT = 64;

struct PyrParams {
    vsb_pyr_layout_t lay;
    int copy_l0;          // 1 = Camera::Update's copy of the frame as level 0 is written; 0 = the readers take level 0 from the frames
};

__device__ __forceinline__ uint8_t mean_clipped(int sum, int count) {
    if (count == 4) return (uint8_t)((sum + 2) >> 2);
    if (count == 0) return 0;
    float v = __fdiv_rn((float)sum, (float)count);
    return (uint8_t)min(__float2int_rn(v), 255);
}

__global__ void __launch_bounds__(256)
pyramid_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, PyrParams P,
               uint8_t* __restrict__ pyr) {
    __shared__ uint8_t s[2][PT * PT];
    const vsb_pyr_layout_t& L = P.lay;
    const int frame = blockIdx.z;
    const int x0 = blockIdx.x * PT, y0 = blockIdx.y * PT;
    uint8_t* out = pyr + (size_t)frame * L.frame_stride;
    const uint8_t* in = img ? img + (size_t)frame * img_stride : out;  // img == NULL: level 0 already in place
    const int in_pitch = img ? pitch : L.w[0];
    const int tid = threadIdx.x;
    const int w0 = L.w[0], h0 = L.h[0];

    // level 0: load the tile (16 bytes per thread when rows are 16-byte aligned; otherwise byte by byte with consecutive
    // threads on consecutive bytes, so a warp's access is one contiguous run), copy it out unless it is in place
    {
        const bool vec_ok = ((in_pitch & 15) == 0) && ((w0 & 15) == 0) && ((((size_t)in) & 15) == 0);
        if (vec_ok) {
            const int ty = tid >> 2, tx = (tid & 3) * 16;
            const int gy = y0 + ty, gx = x0 + tx;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (gy < h0 && gx + 15 < w0) {
                v = __ldg(reinterpret_cast<const uint4*>(in + (size_t)gy * in_pitch + gx));
                if (img && P.copy_l0) *reinterpret_cast<uint4*>(out + (size_t)gy * w0 + gx) = v;
            }
            *reinterpret_cast<uint4*>(&s[0][ty * PT + tx]) = v;
        } else {
#pragma unroll 4
            for (int i = tid; i < PT * PT; i += 256) {
                const int ty = i >> 6, tx = i & 63;
                const int gy = y0 + ty, gx = x0 + tx;
                uint8_t v = 0;
                if (gy < h0 && gx < w0) {
                    v = __ldg(in + (size_t)gy * in_pitch + gx);
                    if (img && P.copy_l0) out[(size_t)gy * w0 + gx] = v;
                }
                s[0][i] = v;
            }
        }
    }
    __syncthreads();
    int cur = 0;
    int tw = PT;                         // tile extent at the current level
    int lx0 = x0, ly0 = y0;              // tile origin at the current level
    for (int l = 1; l < L.levels; l++) {
        const int sw = L.w[l - 1], sh = L.h[l - 1];
        const int dw = L.w[l], dh = L.h[l];
        const int ntw = tw >> 1;
        const int nx0 = lx0 >> 1, ny0 = ly0 >> 1;
        uint8_t* dst = out + L.offset[l];
        for (int p = tid; p < ntw * ntw; p += 256) {
            const int py = p / ntw, px = p - py * ntw;
            const int dx = nx0 + px, dy = ny0 + py;
            uint8_t v = 0;
            if (dx < dw && dy < dh) {
                int sum = 0, cnt = 0;
#pragma unroll
                for (int sy = 0; sy < 2; sy++) {
                    if (2 * dy + sy >= sh) break;
#pragma unroll
                    for (int sx = 0; sx < 2; sx++) {
                        if (2 * dx + sx >= sw) break;
                        sum += s[cur][(2 * py + sy) * tw + 2 * px + sx];
                        cnt++;
                    }
                }
                v = mean_clipped(sum, cnt);
                dst[(size_t)dy * dw + dx] = v;
            }
            s[cur ^ 1][py * ntw + px] = v;
        }
        __syncthreads();
        cur ^= 1; tw = ntw; lx0 = nx0; ly0 = ny0;
    }
}

// Fast path for frames whose width and height are multiples of 16 (752x480, 640x480): every level is an exact
// halving, so a 16x16 level-0 block maps to 8x8 / 4x4 / 2x2 / 1 pixels of levels 1..4 with no clipping.  One
// THREAD owns one block and keeps it in registers (16 rows x uint4); the 2x2 means are formed four bytes at a time in
// 16-bit lanes ((a+b+c+d+2)>>2, exact), so there is no shared memory, no barrier and no idle thread at the small
// levels.  Lanes of a warp own horizontally consecutive blocks: every load/store instruction of a warp touches one
// contiguous run per image row (512 B at level 0, 256/128/64/32 B at levels 1..4).
constexpr int P16_THREADS = 128;

__device__ __forceinline__ uint32_t hsum2(uint32_t w) {          // bytes p0..p3 -> (p0+p1) | (p2+p3) << 16
    return (w & 0x00FF00FFu) + ((w >> 8) & 0x00FF00FFu);
}
__device__ __forceinline__ uint32_t mean4(uint32_t s) {          // two 16-bit sums of four pixels -> two rounded means
    return ((s + 0x00020002u) >> 2) & 0x00FF00FFu;
}
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b) {   // bytes a0, a2, b0, b2
    return __byte_perm(a, b, 0x6420);
}

__global__ void __launch_bounds__(P16_THREADS)
pyramid16_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, PyrParams P,
                 uint8_t* __restrict__ pyr) {
    const vsb_pyr_layout_t& L = P.lay;
    const int w0 = L.w[0];
    const int bw = w0 >> 4, nb = bw * (L.h[0] >> 4);
    const int b = blockIdx.x * P16_THREADS + threadIdx.x;
    if (b >= nb) return;
    const int by = b / bw, bx = b - by * bw;
    const int frame = blockIdx.y;
    uint8_t* out = pyr + (size_t)frame * L.frame_stride;
    const uint8_t* in = img ? img + (size_t)frame * img_stride : out;
    const int in_pitch = img ? pitch : w0;

    uint4 r[16];
    const uint8_t* src = in + (size_t)(by * 16) * in_pitch + bx * 16;
#pragma unroll
    for (int k = 0; k < 16; k++) r[k] = __ldcs(reinterpret_cast<const uint4*>(src + (size_t)k * in_pitch));
    if (img && P.copy_l0) {                                       // Camera::Update keeps a copy of the frame as level 0
        uint8_t* dst = out + (size_t)(by * 16) * w0 + bx * 16;
#pragma unroll
        for (int k = 0; k < 16; k++) *reinterpret_cast<uint4*>(dst + (size_t)k * w0) = r[k];
    }
    if (L.levels < 2) return;
    uint2 l1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint4 a = r[2 * k], c = r[2 * k + 1];
        const uint32_t s0 = mean4(hsum2(a.x) + hsum2(c.x)), s1 = mean4(hsum2(a.y) + hsum2(c.y));
        const uint32_t s2 = mean4(hsum2(a.z) + hsum2(c.z)), s3 = mean4(hsum2(a.w) + hsum2(c.w));
        l1[k] = make_uint2(pack4(s0, s1), pack4(s2, s3));
    }
    {
        const int w1 = L.w[1];
        uint8_t* dst = out + L.offset[1] + (size_t)(by * 8) * w1 + bx * 8;
#pragma unroll
        for (int k = 0; k < 8; k++) *reinterpret_cast<uint2*>(dst + (size_t)k * w1) = l1[k];
    }
    if (L.levels < 3) return;
    uint32_t l2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint2 a = l1[2 * k], c = l1[2 * k + 1];
        l2[k] = pack4(mean4(hsum2(a.x) + hsum2(c.x)), mean4(hsum2(a.y) + hsum2(c.y)));
    }
    {
        const int w2 = L.w[2];
        uint8_t* dst = out + L.offset[2] + (size_t)(by * 4) * w2 + bx * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) *reinterpret_cast<uint32_t*>(dst + (size_t)k * w2) = l2[k];
    }
    if (L.levels < 4) return;
    uint32_t l3[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t s = mean4(hsum2(l2[2 * k]) + hsum2(l2[2 * k + 1]));
        l3[k] = (s & 0xFFu) | ((s >> 8) & 0xFF00u);
    }
    {
        const int w3 = L.w[3];
        uint8_t* dst = out + L.offset[3] + (size_t)(by * 2) * w3 + bx * 2;
        *reinterpret_cast<uint16_t*>(dst) = (uint16_t)l3[0];
        *reinterpret_cast<uint16_t*>(dst + w3) = (uint16_t)l3[1];
    }
    if (L.levels < 5) return;
    const uint32_t s4 = (l3[0] & 0xFFu) + (l3[0] >> 8) + (l3[1] & 0xFFu) + (l3[1] >> 8);
    out[L.offset[4] + (size_t)by * L.w[4] + bx] = (uint8_t)((s4 + 2u) >> 2);
}

// The same register-blocked form for ANY frame size and alignment (KITTI's 1241x376: rows start on every byte alignment, the
// levels are 620x188, 310x94, 155x47, 77x23).  Level l is (w >> l) x (h >> l) (Camera.cpp:44-47), so pixel (x, y) of level l
// is still the rounded mean of the level-0 block [x 2^l, (x + 1) 2^l) x [y 2^l, (y + 1) 2^l), which lies inside the frame
// whenever the pixel exists: a thread owns a 16x16 level-0 block as before, the blocks of the last column / row are partial.
// A block row is read as five aligned words and funnel-shifted into place (bytes past the frame's width only reach pixels
// that do not exist); stores take the widest unit the destination's alignment allows (level 1 of a 1241-wide frame is on
// 4-byte boundaries, level 2 on 2-byte ones) and stop at the level's width.
template <int NW>
__device__ __forceinline__ void store_run(uint8_t* dst, const uint32_t (&wd)[NW], int nbytes) {
    if (nbytes <= 0) return;
    const unsigned a = (unsigned)(reinterpret_cast<uintptr_t>(dst) & 3u);
    if (a == 0u) {
#pragma unroll
        for (int j = 0; j < NW; j++) {
            if (4 * j + 3 < nbytes) *reinterpret_cast<uint32_t*>(dst + 4 * j) = wd[j];
            else {
#pragma unroll
                for (int t = 0; t < 4; t++) if (4 * j + t < nbytes) dst[4 * j + t] = (uint8_t)(wd[j] >> (8 * t));
            }
        }
    } else if (a == 2u) {
#pragma unroll
        for (int j = 0; j < 2 * NW; j++) {
            const uint32_t hw = (wd[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
            if (2 * j + 1 < nbytes) *reinterpret_cast<uint16_t*>(dst + 2 * j) = (uint16_t)hw;
            else if (2 * j < nbytes) dst[2 * j] = (uint8_t)hw;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4 * NW; i++) if (i < nbytes) dst[i] = (uint8_t)(wd[i >> 2] >> (8 * (i & 3)));
    }
}

__global__ void __launch_bounds__(P16_THREADS)
pyramid16u_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, PyrParams P, uint8_t* __restrict__ pyr) {
    const vsb_pyr_layout_t& L = P.lay;
    const int w0 = L.w[0], h0 = L.h[0];
    const int bw = (w0 + 15) >> 4, nb = bw * ((h0 + 15) >> 4);
    const int b = blockIdx.x * P16_THREADS + threadIdx.x;
    if (b >= nb) return;
    const int by = b / bw, bx = b - by * bw;
    const int frame = blockIdx.y;
    uint8_t* out = pyr + (size_t)frame * L.frame_stride;
    const uint8_t* in = img ? img + (size_t)frame * img_stride : out;
    const int in_pitch = img ? pitch : w0;
    const int nx0 = min(16, w0 - 16 * bx), ny0 = min(16, h0 - 16 * by);      // valid level-0 bytes / rows of this block

    uint4 r[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        uint32_t wv[5] = {0u, 0u, 0u, 0u, 0u};
        const uint8_t* a = in + (size_t)(by * 16 + k) * in_pitch + bx * 16;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(a) & 3u);
        if (k < ny0) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(a - mis);
            const int need = (int)mis + nx0;                              // bytes from the aligned start to the last valid one
#pragma unroll
            for (int j = 0; j < 5; j++) if (4 * j < need) wv[j] = __ldg(wp + j);
        }
        const unsigned sh = 8u * mis;
        r[k] = make_uint4(__funnelshift_r(wv[0], wv[1], sh), __funnelshift_r(wv[1], wv[2], sh),
                          __funnelshift_r(wv[2], wv[3], sh), __funnelshift_r(wv[3], wv[4], sh));
    }
    if (img && P.copy_l0) {                                       // Camera::Update keeps a copy of the frame as level 0
        uint8_t* dst = out + (size_t)(by * 16) * w0 + bx * 16;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (k < ny0) { const uint32_t wd[4] = {r[k].x, r[k].y, r[k].z, r[k].w}; store_run<4>(dst + (size_t)k * w0, wd, nx0); }
    }
    if (L.levels < 2) return;
    uint2 l1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint4 a = r[2 * k], c = r[2 * k + 1];
        const uint32_t s0 = mean4(hsum2(a.x) + hsum2(c.x)), s1 = mean4(hsum2(a.y) + hsum2(c.y));
        const uint32_t s2 = mean4(hsum2(a.z) + hsum2(c.z)), s3 = mean4(hsum2(a.w) + hsum2(c.w));
        l1[k] = make_uint2(pack4(s0, s1), pack4(s2, s3));
    }
    {
        const int w1 = L.w[1], nx = min(8, w1 - 8 * bx), ny = min(8, L.h[1] - 8 * by);
        uint8_t* dst = out + L.offset[1] + (size_t)(by * 8) * w1 + bx * 8;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < ny) { const uint32_t wd[2] = {l1[k].x, l1[k].y}; store_run<2>(dst + (size_t)k * w1, wd, nx); }
    }
    if (L.levels < 3) return;
    uint32_t l2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint2 a = l1[2 * k], c = l1[2 * k + 1];
        l2[k] = pack4(mean4(hsum2(a.x) + hsum2(c.x)), mean4(hsum2(a.y) + hsum2(c.y)));
    }
    {
        const int w2 = L.w[2], nx = min(4, w2 - 4 * bx), ny = min(4, L.h[2] - 4 * by);
        uint8_t* dst = out + L.offset[2] + (size_t)(by * 4) * w2 + bx * 4;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < ny) { const uint32_t wd[1] = {l2[k]}; store_run<1>(dst + (size_t)k * w2, wd, nx); }
    }
    if (L.levels < 4) return;
    uint32_t l3[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t s = mean4(hsum2(l2[2 * k]) + hsum2(l2[2 * k + 1]));
        l3[k] = (s & 0xFFu) | ((s >> 8) & 0xFF00u);
    }
    {
        const int w3 = L.w[3], nx = min(2, w3 - 2 * bx), ny = min(2, L.h[3] - 2 * by);
        uint8_t* dst = out + L.offset[3] + (size_t)(by * 2) * w3 + bx * 2;
#pragma unroll
        for (int k = 0; k < 2; k++)
            if (k < ny) { const uint32_t wd[1] = {l3[k]}; store_run<1>(dst + (size_t)k * w3, wd, nx); }
    }
    if (L.levels < 5) return;
    if (bx < L.w[4] && by < L.h[4]) {
        const uint32_t s4 = (l3[0] & 0xFFu) + (l3[0] >> 8) + (l3[1] & 0xFFu) + (l3[1] >> 8);
        out[L.offset[4] + (size_t)by * L.w[4] + bx] = (uint8_t)((s4 + 2u) >> 2);
    }
}

// Level sizes are cvRound(size / 2) (vsb_pyr_layout), which rounds an odd half UP when it is odd (155 -> 78, 47 -> 24): the
// last column / row of such a level has only one source column / row (the clipped mean of pyramid_kernel: sum / count, rounded
// to nearest even), and everything computed from it inherits that.  Whatever the block kernel does wrong therefore stays inside
// the last two columns and rows of every level (pixel x of level l reads 2x and 2x + 1 of level l - 1), and this kernel — one
// block per frame, level after level — recomputes exactly those from the level below with the generic kernel's arithmetic.
__global__ void __launch_bounds__(256)
pyramid_edge_fix_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, PyrParams P, uint8_t* __restrict__ pyr) {
    const vsb_pyr_layout_t& L = P.lay;
    const int frame = blockIdx.x;
    uint8_t* out = pyr + (size_t)frame * L.frame_stride;
    for (int l = 1; l < L.levels; l++) {
        const int sw = L.w[l - 1], sh = L.h[l - 1], dw = L.w[l], dh = L.h[l];
        const uint8_t* src = (l == 1 && img) ? img + (size_t)frame * img_stride : out + L.offset[l - 1];
        const int sp = (l == 1 && img) ? pitch : sw;
        uint8_t* dst = out + L.offset[l];
        const int ncol = 2 * dh, nrow = 2 * dw;                    // the last two columns, then the last two rows
        for (int i = threadIdx.x; i < ncol + nrow; i += 256) {
            int dx, dy;
            if (i < ncol) { dy = i >> 1; dx = dw - 1 - (i & 1); }
            else { const int j = i - ncol; dx = j >> 1; dy = dh - 1 - (j & 1); }
            if (dx < 0 || dy < 0) continue;
            int sum = 0, cnt = 0;
#pragma unroll
            for (int sy = 0; sy < 2; sy++)
#pragma unroll
                for (int sx = 0; sx < 2; sx++)
                    if (2 * dy + sy < sh && 2 * dx + sx < sw) { sum += src[(size_t)(2 * dy + sy) * sp + 2 * dx + sx]; cnt++; }
            dst[(size_t)dy * dw + dx] = mean_clipped(sum, cnt);
        }
        __syncthreads();                                            // the next level reads what this one wrote
    }
}

// ---------------------------------------------------------------------------------------------- Scharr
// cv::Scharr(src, dst, CV_16S, dx, dy, scale = 3, 0, BORDER_REFLECT_101): derivative [-1 0 1], smoothing
// [3 10 3], times 3 (Camera.cpp:171-172; the literal 3 is `scale`).  One launch covers every
// level of every frame; a CTA stages a (GW+2) x (GH+2) tile (reflected at the borders) in shared memory.
constexpr int GW = 64, GH = 16;

struct GradParams {
    vsb_pyr_layout_t lay;
    int tile_begin[VSB_MAX_LEVELS + 1];   // prefix of tiles per level
    int tiles_x[VSB_MAX_LEVELS];
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    if (p < 0) return -p;
    if (p >= len) return 2 * len - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256)
gradient_kernel(const uint8_t* __restrict__ pyr, GradParams P, int16_t* __restrict__ gx, int16_t* __restrict__ gy,
                uint8_t* __restrict__ gmag) {
    __shared__ uint8_t s[(GH + 2)][GW + 4];
    const vsb_pyr_layout_t& L = P.lay;
    int lvl = 0;
    while (lvl + 1 < L.levels && (int)blockIdx.x >= P.tile_begin[lvl + 1]) lvl++;
    const int t = blockIdx.x - P.tile_begin[lvl];
    const int tyi = t / P.tiles_x[lvl], txi = t - tyi * P.tiles_x[lvl];
    const int w = L.w[lvl], h = L.h[lvl];
    const int x0 = txi * GW, y0 = tyi * GH;
    const size_t base = (size_t)blockIdx.y * L.frame_stride + L.offset[lvl];
    const uint8_t* src = pyr + base;
    for (int p = threadIdx.x; p < (GH + 2) * (GW + 2); p += 256) {
        const int py = p / (GW + 2), px = p - py * (GW + 2);
        const int yy = reflect101(min(y0 + py - 1, h), h);
        const int xx = reflect101(min(x0 + px - 1, w), w);
        s[py][px] = __ldg(src + (size_t)yy * w + xx);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < GH * GW; p += 256) {
        const int py = p / GW, px = p - py * GW;
        const int x = x0 + px, y = y0 + py;
        if (x >= w || y >= h) continue;
        const int a00 = s[py][px], a01 = s[py][px + 1], a02 = s[py][px + 2];
        const int a10 = s[py + 1][px], a12 = s[py + 1][px + 2];
        const int a20 = s[py + 2][px], a21 = s[py + 2][px + 1], a22 = s[py + 2][px + 2];
        const int dx = 3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20);
        const int dy = 3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02);
        const size_t o = base + (size_t)y * w + x;
        const int vx = 3 * dx, vy = 3 * dy;
        gx[o] = (int16_t)vx;
        gy[o] = (int16_t)vy;
        if (gmag) {   // convertScaleAbs + addWeighted(0.5, 0.5), Camera.cpp:174-180
            float m = __fadd_rn(__fmul_rn((float)min(abs(vx), 255), 0.5f), __fmul_rn((float)min(abs(vy), 255), 0.5f));
            gmag[o] = (uint8_t)min(__float2int_rn(m), 255);
        }
    }
}

// ---------------------------------------------------------------------------------------------- candidates
struct CandParams {
    int levels;
    int lw[VSB_MAX_LEVELS], lh[VSB_MAX_LEVELS];
    // fused attribute pass of the solver (optional): for levels last_lvl..first_lvl also write, per point, the record
    // {Scharr gx | gy << 16, I_prev} that gn_prepare_kernel would gather (grad_mode 1 arithmetic, bit-identical)
    const uint8_t* prev_pyr;
    long long pair_stride;
    const uint8_t* prev_l0;       // level 0 of the previous frames outside the packed pyramid (frame p at prev_l0 + p * l0_stride), or NULL
    long long l0_stride;
    vsb_pyr_layout_t lay;
    int first_lvl, last_lvl;
    uint2* patt;
    // ... and, instead of the float4 rows, the back-projected (X, Y) = (x * invfx + bx, y * invfy + by) of the unit-depth
    // points as doubles (VISystem.cpp:1519-1524 with z = 1; bx = backproj_offset(cx, invfx), see se3.cuh), which is all the
    // solver needs of a point
    double2* xy;
    // ... or (gn_track.cu's form, rec_abs) nothing per point but the record, which then also carries the point's column
    // and row — the solver's back-projection tables are indexed by them: {gx | gy << 16, I_prev | column << 8 | row << 20}.
    // Levels in dedup_mask (small images, where the patches of different features overlap heavily: level 3 of a 752x480
    // frame is 94x60 pixels for up to 200 x 121 candidate points) get ONE record per distinct pixel with its multiplicity:
    // {gx | gy << 16, I_prev | column << 8 | row << 16 | multiplicity << 24}.  n_pts = candidate points before merging.
    int rec_abs;
    uint32_t dedup_mask;
    int32_t* n_pts;
    float bx[VSB_MAX_LEVELS], by[VSB_MAX_LEVELS], invfx[VSB_MAX_LEVELS], invfy[VSB_MAX_LEVELS];
};

__global__ void __launch_bounds__(256, 5)
candidates_kernel(const float* __restrict__ good_xy, int good_cap, const int32_t* __restrict__ n_good, CandParams P,
                  float4* __restrict__ cand, int cand_cap, int32_t* __restrict__ n_cand) {
    const int prob = blockIdx.x;
    const int lvl = blockIdx.y;
    const int tid = threadIdx.x;
    if (P.patt != nullptr && lvl > P.first_lvl) {           // fused (tracker) form: levels the solver never visits stay empty
        if (tid == 0) n_cand[(size_t)prob * P.levels + lvl] = 0;
        return;
    }
    __shared__ int s_cnt[256], s_off[257];
    __shared__ int s_ia[256], s_ja[256], s_nj[256];
    const int patch_size[VSB_MAX_LEVELS] = {5, 3, 2, 5, 5};              // Camera.cpp:369-373
    const int nf = min(min(n_good[prob], good_cap), VSB_MAX_GN_FEATURES);  // Camera.cpp:382
    const float factor_lvl = (float)(1.0 / (double)(1 << lvl));           // Camera.cpp:379
    const int sp = patch_size[lvl] - 1 / 2;                               // Camera.cpp:381 (int division)
    const int lw = P.lw[lvl], lh = P.lh[lvl];
    int cnt = 0;
    if (tid < nf) {
        const float px = good_xy[((size_t)prob * good_cap + tid) * 2];
        const float py = good_xy[((size_t)prob * good_cap + tid) * 2 + 1];
        const float x = (float)(((double)px + 0.5) * (double)factor_lvl - 0.5);   // Camera.cpp:384
        const float y = (float)(((double)py + 0.5) * (double)factor_lvl - 0.5);   // Camera.cpp:385
        // for (int i = x - sp; i <= x + sp; i++): float -> int truncation, float compare (Camera.cpp:391-392)
        const int i0 = (int)__fsub_rn(x, (float)sp), i1 = (int)floorf(__fadd_rn(x, (float)sp));
        const int j0 = (int)__fsub_rn(y, (float)sp), j1 = (int)floorf(__fadd_rn(y, (float)sp));
        const int ia = max(i0, 1), ib = min(i1, lw - 1);                  // 0 < i < w (Camera.cpp:393)
        const int ja = max(j0, 1), jb = min(j1, lh - 1);
        const int ni = max(ib - ia + 1, 0), nj = max(jb - ja + 1, 0);
        cnt = ni * nj;
        s_ia[tid] = ia; s_ja[tid] = ja; s_nj[tid] = nj;
    }
    s_cnt[tid] = cnt;
    // exclusive prefix of the per-feature counts (cnt is 0 from feature nf on): warp scans, then the eight warp totals
    __shared__ int s_wsum[8];
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((tid & 31) >= d) incl += v;
    }
    if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) base += k < (tid >> 5) ? s_wsum[k] : 0;
    s_off[tid] = base + incl - cnt;
    if (tid == 255) {
        s_off[256] = base + incl;
        n_cand[(size_t)prob * P.levels + lvl] = min(base + incl, cand_cap);
    }
    __syncthreads();
    float4* out = cand + ((size_t)prob * P.levels + lvl) * cand_cap;
    const int total = min(s_off[nf], cand_cap);
    const bool attrs = P.patt != nullptr && lvl <= P.first_lvl && lvl >= P.last_lvl;
    uint2* pout = attrs ? P.patt + ((size_t)prob * P.levels + lvl) * cand_cap : nullptr;
    const uint8_t* image1 = !attrs ? nullptr : (lvl == 0 && P.prev_l0) ? P.prev_l0 + (size_t)prob * P.l0_stride
                                                                       : P.prev_pyr + (size_t)prob * P.pair_stride + P.lay.offset[lvl];
    const int cols = P.lay.w[lvl], rows = P.lay.h[lvl];
    // one warp per feature: lanes stride over the feature's points (i outer, j inner — reference row order)
    const int warp = tid >> 5, lane = tid & 31;
    if (P.rec_abs && P.n_pts != nullptr && tid == 0) P.n_pts[(size_t)prob * P.levels + lvl] = total;
    if (attrs && P.rec_abs && ((P.dedup_mask >> lvl) & 1u)) {
        // ---- one record per distinct pixel, with its multiplicity -------------------------------------------------
        __shared__ uint32_t s_mult[VSB_DEDUP_PIX / 4];                  // one byte per pixel (at most 200 features cover it)
        const int npix = cols * rows, words = (npix + 3) >> 2;
        for (int k = tid; k < words; k += 256) s_mult[k] = 0u;
        __syncthreads();
        for (int f = warp; f < nf; f += 8) {
            const int c = s_cnt[f], njs = max(s_nj[f], 1), ia = s_ia[f], ja = s_ja[f];
            for (int p = lane; p < c; p += 32) {
                const int ii = p / njs, jj = p - ii * njs;
                const int pix = (ja + jj) * cols + ia + ii;
                atomicAdd(&s_mult[pix >> 2], 1u << ((pix & 3) * 8));
            }
        }
        __syncthreads();
        // every thread owns a contiguous span of words; records come out in pixel order
        const int wspan = (words + 255) >> 8;
        const int w0 = min(tid * wspan, words), w1 = min(w0 + wspan, words);
        int mine = 0;
        for (int w = w0; w < w1; w++) {
            const uint32_t v = s_mult[w];
            mine += ((v & 0xFFu) != 0u) + ((v & 0xFF00u) != 0u) + ((v & 0xFF0000u) != 0u) + ((v >> 24) != 0u);
        }
        int inc2 = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
            if (lane >= d) inc2 += v;
        }
        __syncthreads();                                                  // s_wsum is reused
        if (lane == 31) s_wsum[warp] = inc2;
        __syncthreads();
        int o = inc2 - mine;
#pragma unroll
        for (int k = 0; k < 8; k++) o += k < warp ? s_wsum[k] : 0;
        if (tid == 255) n_cand[(size_t)prob * P.levels + lvl] = o + mine;
        for (int w = w0; w < w1; w++) {
            const uint32_t v = s_mult[w];
            if (v == 0u) continue;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t m = (v >> (8 * b)) & 0xFFu;
                if (m == 0u) continue;
                const int pix = 4 * w + b;
                const int j = pix / cols, i = pix - j * cols;
                // 0 < i < lw <= cols and 0 < j < lh <= rows: the centre is inside, the neighbours reflect (101) at the border
                const int xm = i - 1, xp = reflect101(i + 1, cols), ym = j - 1, yp = reflect101(j + 1, rows);
                const uint8_t* q0 = image1 + (size_t)ym * cols;
                const uint8_t* q1 = image1 + (size_t)j * cols;
                const uint8_t* q2 = image1 + (size_t)yp * cols;
                const int a00 = __ldg(q0 + xm), a01 = __ldg(q0 + i), a02 = __ldg(q0 + xp);
                const int a10 = __ldg(q1 + xm), a11 = __ldg(q1 + i), a12 = __ldg(q1 + xp);
                const int a20 = __ldg(q2 + xm), a21 = __ldg(q2 + i), a22 = __ldg(q2 + xp);
                const int gx = 3 * (3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20));
                const int gy = 3 * (3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02));
                pout[o++] = make_uint2(((uint32_t)gx & 0xFFFFu) | ((uint32_t)gy << 16),
                                       (uint32_t)a11 | (uint32_t)i << 8 | (uint32_t)j << 16 | m << 24);
            }
        }
        return;
    }
    // the feature's patch plus its one-pixel Scharr apron (<= 13 x 13 bytes) is staged row by row (contiguous bytes per image
    // row) — the points themselves run down the columns, so reading the nine neighbours straight from the image would touch
    // a different row per lane
    __shared__ uint8_t s_tile[8][13][16];
    // ... and fetched one feature ahead into registers (two tile rows per pass, sixteen lanes a row, at most seven passes), so
    // the image latency of feature f + 8 hides behind the arithmetic of feature f.  reflect101 is the identity inside the
    // image, so one formula serves the apron and the interior
    const int tu = lane & 15, tv = lane >> 4;
    uint32_t pre[7];
    auto fetch = [&](int f) {
        if (!attrs || f >= nf) return;
        const int c = s_cnt[f], nj = s_nj[f];
        if (c <= 0 || tu >= c / nj + 2) return;
        const uint8_t* col = image1 + reflect101(s_ia[f] - 1 + tu, cols);
        const int ja = s_ja[f];
#pragma unroll
        for (int k = 0; k < 7; k++)
            if (tv + 2 * k < nj + 2) pre[k] = __ldg(col + (size_t)reflect101(ja - 1 + tv + 2 * k, rows) * cols);
    };
    fetch(warp);
    for (int f = warp; f < nf; f += 8) {
        const int c = s_cnt[f], off = s_off[f], nj = s_nj[f], ia = s_ia[f], ja = s_ja[f];
        if (attrs) {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 7; k++)
                if (tv + 2 * k < 13) s_tile[warp][tv + 2 * k][tu] = (uint8_t)pre[k];   // rows past the patch hold stale bytes nobody reads
            __syncwarp();
            fetch(f + 8);
        }
        // the lane's first point and the step of 32 points as (column, row), the row running fastest: one division per feature
        const int njs = max(nj, 1);
        int ii = lane / njs, jj = lane - ii * njs;
        const int di = 32 / njs, dj = 32 - di * njs;
        for (int p = lane; p < c; p += 32) {
            if (off + p >= total) break;
            // (float) of a small non-negative integer without the conversion pipe: 2^23 + n is exact, so is the subtraction
            const float xf = __fsub_rn(__int_as_float(0x4B000000 | (ia + ii)), 8388608.0f);
            const float yf = __fsub_rn(__int_as_float(0x4B000000 | (ja + jj)), 8388608.0f);
            if (attrs && P.rec_abs) {
                // no per-point coordinates beside the record: it carries the column and the row
            } else if (attrs && P.xy) {
                const float X = __fadd_rn(__fmul_rn(xf, P.invfx[lvl]), P.bx[lvl]);   // * z (= 1) is the identity
                const float Y = __fadd_rn(__fmul_rn(yf, P.invfy[lvl]), P.by[lvl]);
                P.xy[((size_t)prob * P.levels + lvl) * cand_cap + off + p] = make_double2((double)X, (double)Y);
            } else if (cand != nullptr) {
                out[off + p] = make_float4(xf, yf, 1.0f, 1.0f);
            }
            if (attrs) {
                // point (ia + ii, ja + jj) sits at tile[jj + 1][ii + 1]: 0 < x < lw <= cols and 0 < y < lh <= rows, so the
                // centre needs no clamping and its neighbours are the reflect101 values staged above
                const uint8_t* q0 = &s_tile[warp][jj][ii];
                const uint8_t* q1 = q0 + 16;
                const uint8_t* q2 = q0 + 32;
                const int a00 = q0[0], a01 = q0[1], a02 = q0[2];
                const int a10 = q1[0], a11 = q1[1], a12 = q1[2];
                const int a20 = q2[0], a21 = q2[1], a22 = q2[2];
                const int gx = 3 * (3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20));
                const int gy = 3 * (3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02));
                uint32_t tag = (uint32_t)a11;
                if (P.rec_abs) tag |= (uint32_t)(ia + ii) << 8 | (uint32_t)(ja + jj) << 20;
                pout[off + p] = make_uint2(((uint32_t)gx & 0xFFFFu) | ((uint32_t)gy << 16), tag);
            }
            ii += di;
            jj += dj;
            if (jj >= njs) { jj -= njs; ii++; }
        }
    }
}

}  // namespace

static int round_half_even_half(int v) {  // cvRound(v * 0.5)
    int k = v >> 1;
    return (v & 1) ? k + (k & 1) : k;
}

extern "C" int vsb_pyr_layout(int w, int h, int levels, vsb_pyr_layout_t* out) {
    if (!out || w <= 0 || h <= 0 || levels < 1 || levels > VSB_MAX_LEVELS) return VSB_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->levels = levels;
    int64_t off = 0;
    for (int l = 0; l < levels; l++) {
        out->w[l] = l ? round_half_even_half(out->w[l - 1]) : w;
        out->h[l] = l ? round_half_even_half(out->h[l - 1]) : h;
        if (out->w[l] <= 0 || out->h[l] <= 0) return VSB_ERR_INVALID;
        out->offset[l] = off;
        off += ((int64_t)out->w[l] * out->h[l] + 255) & ~(int64_t)255;
    }
    out->frame_stride = off;
    return VSB_OK;
}

extern "C" int vsb_pyramid_build(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int count,
                                 const vsb_pyr_layout_t* layout, uint8_t* pyr, void* stream) {
    return vsb_pyramid_build_levels(ctx, img, img_stride, pitch, count, layout, pyr, 1, stream);
}

// Internal form (tracker): copy_l0 = 0 leaves level 0 of the packed pyramid unwritten — the candidate pass and the solver then
// read level 0 from the caller's frames (prev_l0 / cur_l0), which saves a write and later reads of w h bytes per frame.
int vsb_pyramid_build_levels(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int count,
                             const vsb_pyr_layout_t* layout, uint8_t* pyr, int copy_l0, void* stream) {
    if (!ctx || !layout || !pyr || count < 0) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    PyrParams P;
    P.lay = *layout;
    P.copy_l0 = copy_l0;
    // register-blocked fast path: exact halving at every level and 16-byte aligned rows
    const int in_pitch = img ? pitch : layout->w[0];
    const bool fast = ctx->pyr_impl != 0 && (layout->w[0] % 16 == 0) && (layout->h[0] % 16 == 0) &&
                      (in_pitch % 16 == 0) && (!img || (((uintptr_t)img | (uintptr_t)img_stride) % 16 == 0)) &&
                      ((uintptr_t)pyr % 16 == 0);
    if (fast) {
        const int nb = (layout->w[0] / 16) * (layout->h[0] / 16);
        for (int z0 = 0; z0 < count; z0 += 65535) {
            int zc = count - z0 < 65535 ? count - z0 : 65535;
            dim3 grid(vsb_div_up(nb, P16_THREADS), zc);
            ProfScope ps(ctx, VSB_K_PYRAMID, (cudaStream_t)stream);
            pyramid16_kernel<<<grid, P16_THREADS, 0, (cudaStream_t)stream>>>(
                img ? img + (size_t)z0 * img_stride : nullptr, img_stride, pitch, P, pyr + (size_t)z0 * layout->frame_stride);
            VSB_LAUNCHED(ctx);
        }
        return VSB_OK;
    }
    // any other size / alignment: the same blocks read as aligned words and funnel-shifted (levels are floor halvings up to
    // level 4, which is all a 16x16 block determines)
    if (ctx->pyr_impl != 0 && layout->levels <= 5) {
        // level sizes a 16x16 block hierarchy covers: floor or ceil of half the level below (cvRound(size / 2) is one of them)
        bool halvings = true, exact = true;
        for (int l = 1; l < layout->levels; l++) {
            const int pw = layout->w[l - 1], ph = layout->h[l - 1];
            if ((layout->w[l] != pw / 2 && layout->w[l] != (pw + 1) / 2) || (layout->h[l] != ph / 2 && layout->h[l] != (ph + 1) / 2)) halvings = false;
            if (layout->w[l] != pw / 2 || layout->h[l] != ph / 2) exact = false;
        }
        if (halvings) {
            const int nb = ((layout->w[0] + 15) / 16) * ((layout->h[0] + 15) / 16);
            // Camera::Update's copy of the frame as level 0: rows of any alignment are best left to the copy engine's
            // arithmetic (one strided device-to-device copy for the batch) rather than byte stores in the kernel
            if (img && copy_l0 && pitch == layout->w[0]) {
                ProfScope ps(ctx, VSB_K_PYRAMID, (cudaStream_t)stream);
                VSB_CUDA(ctx, cudaMemcpy2DAsync(pyr, (size_t)layout->frame_stride, img, (size_t)img_stride, (size_t)layout->w[0] * layout->h[0],
                                                (size_t)count, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
                P.copy_l0 = 0;
            }
            for (int z0 = 0; z0 < count; z0 += 65535) {
                int zc = count - z0 < 65535 ? count - z0 : 65535;
                dim3 grid(vsb_div_up(nb, P16_THREADS), zc);
                ProfScope ps(ctx, VSB_K_PYRAMID, (cudaStream_t)stream);
                pyramid16u_kernel<<<grid, P16_THREADS, 0, (cudaStream_t)stream>>>(
                    img ? img + (size_t)z0 * img_stride : nullptr, img_stride, pitch, P, pyr + (size_t)z0 * layout->frame_stride);
                VSB_LAUNCHED(ctx);
                if (!exact) {
                    pyramid_edge_fix_kernel<<<zc, 256, 0, (cudaStream_t)stream>>>(
                        img ? img + (size_t)z0 * img_stride : nullptr, img_stride, pitch, P, pyr + (size_t)z0 * layout->frame_stride);
                    VSB_LAUNCHED(ctx);
                }
            }
            return VSB_OK;
        }
    }
    for (int z0 = 0; z0 < count; z0 += 65535) {
        int zc = count - z0 < 65535 ? count - z0 : 65535;
        dim3 grid(vsb_div_up(layout->w[0], PT), vsb_div_up(layout->h[0], PT), zc);
        ProfScope ps(ctx, VSB_K_PYRAMID, (cudaStream_t)stream);
        pyramid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img ? img + (size_t)z0 * img_stride : nullptr, img_stride,
                                                                pitch, P, pyr + (size_t)z0 * layout->frame_stride);
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

extern "C" int vsb_gradient_build(vsb_ctx_t* ctx, const uint8_t* pyr, int count, const vsb_pyr_layout_t* layout,
                                  int16_t* gx, int16_t* gy, uint8_t* gmag, void* stream) {
    if (!ctx || !layout || !pyr || !gx || !gy || count < 0) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    GradParams P;
    P.lay = *layout;
    int acc = 0;
    for (int l = 0; l < layout->levels; l++) {
        P.tile_begin[l] = acc;
        P.tiles_x[l] = vsb_div_up(layout->w[l], GW);
        acc += P.tiles_x[l] * vsb_div_up(layout->h[l], GH);
    }
    for (int l = layout->levels; l <= VSB_MAX_LEVELS; l++) P.tile_begin[l] = acc;
    for (int z0 = 0; z0 < count; z0 += 65535) {
        int zc = count - z0 < 65535 ? count - z0 : 65535;
        dim3 grid(acc, zc);
        ProfScope ps(ctx, VSB_K_GRADIENT, (cudaStream_t)stream);
        size_t o = (size_t)z0 * layout->frame_stride;
        gradient_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pyr + o, P, gx + o, gy + o, gmag ? gmag + o : nullptr);
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

// Internal entry (tracker): candidate points AND, for the solver's levels, the per-point attribute records in one pass
// (prev_pyr == NULL: candidates only).  The candidate level sizes lw/lh must equal the pyramid layout's (exact halving).
int vsb_candidates_prepare(vsb_ctx_t* ctx, const float* good_xy, int good_cap, const int32_t* n_good, int count, int levels,
                           const int* lw, const int* lh, float* cand, int cand_cap, int32_t* n_cand,
                           const uint8_t* prev_pyr, int64_t pair_stride, const vsb_pyr_layout_t* layout, int first_lvl,
                           int last_lvl, void* patt, void* xy, const vsb_intr_t* K, int rec_abs, uint32_t dedup_mask,
                           int32_t* n_pts, const uint8_t* prev_l0, int64_t l0_stride, void* stream) {
    if (!ctx || !good_xy || !n_good || (!cand && !rec_abs) || !n_cand || !lw || !lh) return VSB_ERR_INVALID;
    if (rec_abs && (!prev_pyr || !layout || !patt)) return VSB_ERR_INVALID;
    if (levels < 1 || levels > VSB_MAX_LEVELS || count < 0 || cand_cap < 0) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    CandParams P;
    P.levels = levels;
    for (int l = 0; l < levels; l++) { P.lw[l] = lw[l]; P.lh[l] = lh[l]; }
    P.prev_pyr = nullptr; P.pair_stride = 0; P.first_lvl = -1; P.last_lvl = 0; P.patt = nullptr; P.xy = nullptr;
    P.rec_abs = 0; P.dedup_mask = 0u; P.n_pts = nullptr; P.prev_l0 = nullptr; P.l0_stride = 0;
    for (int l = 0; l < VSB_MAX_LEVELS; l++) { P.bx[l] = P.by[l] = 0.f; P.invfx[l] = P.invfy[l] = 0.f; }
    memset(&P.lay, 0, sizeof(P.lay));
    if (prev_pyr && layout && patt) {
        P.prev_pyr = prev_pyr; P.pair_stride = pair_stride; P.lay = *layout; P.prev_l0 = prev_l0; P.l0_stride = l0_stride;
        P.first_lvl = first_lvl; P.last_lvl = last_lvl; P.patt = reinterpret_cast<uint2*>(patt);
        if (rec_abs) {
            P.rec_abs = 1; P.n_pts = n_pts;
            for (int l = 0; l < levels; l++)      // records hold 12-bit columns / rows; merged levels 8-bit ones and a byte map
                if (layout->w[l] > 4095 || layout->h[l] > 4095) return VSB_ERR_UNSUPPORTED;
            for (int l = 0; l < levels; l++)
                if (((dedup_mask >> l) & 1u) && layout->w[l] <= 255 && layout->h[l] <= 255 && layout->w[l] * layout->h[l] <= VSB_DEDUP_PIX)
                    P.dedup_mask |= 1u << l;
        } else if (xy && K) {
            P.xy = reinterpret_cast<double2*>(xy);
            for (int l = 0; l < VSB_MAX_LEVELS; l++) {
                P.bx[l] = vsb::backproj_offset(K[l].cx, K[l].invfx); P.by[l] = vsb::backproj_offset(K[l].cy, K[l].invfy);
                P.invfx[l] = K[l].invfx; P.invfy[l] = K[l].invfy;
            }
        }
    }
    dim3 grid(count, levels);
    ProfScope ps(ctx, VSB_K_CANDIDATES, (cudaStream_t)stream);
    candidates_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(good_xy, good_cap, n_good, P,
                                                              reinterpret_cast<float4*>(cand), cand_cap, n_cand);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_candidates_build(vsb_ctx_t* ctx, const float* good_xy, int good_cap, const int32_t* n_good,
                                    int count, int levels, const int* lw, const int* lh, float* cand, int cand_cap,
                                    int32_t* n_cand, void* stream) {
    return vsb_candidates_prepare(ctx, good_xy, good_cap, n_good, count, levels, lw, lh, cand, cand_cap, n_cand, nullptr, 0,
                                  nullptr, 0, 0, nullptr, nullptr, nullptr, 0, 0u, nullptr, nullptr, 0, stream);
}


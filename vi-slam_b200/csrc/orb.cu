// orb.cu — ORB key points and descriptors on the device: what cv::ORB::detectAndCompute does for ONE
// pyramid level (cv::ORB::create(n, 1.2f, 1)), the detector the reference runs in Camera::detectAndComputeFeatures
// (src/Camera.cpp:79-86, 124-129) / cv::cuda::ORB in CameraGPU (src/CameraGPU.cpp:99-104) before the tracked path:
//   FAST-9/16 (fast.cu) -> border filter (edgeThreshold 31) -> retainBest(2n) on the FAST score -> Harris response (7x7,
//   k = 0.04) -> retainBest(n) -> intensity-centroid angle (31-pixel circular patch, cv::fastAtan2) -> 7x7 sigma-2 blur ->
//   steered rBRIEF (256 tests).
// Results are bit-identical to oracle/orb.c, which is pinned bit for bit against cv2 4.13: same key-point set (ties at the
// two selection thresholds are kept, as cv::KeyPointsFilter::retainBest keeps them), same responses, angles, descriptors;
// key points come out in row-major (y, x) order (OpenCV's own order is whatever std::nth_element leaves).
// All float arithmetic that OpenCV evaluates without fused multiply-add uses explicitly rounded intrinsics; the blur is the
// float separable filter with the fused steps the oracle documents.
// vsb_orb_detect_compute_pyr adds the scale pyramid (cv::ORB::create(n): 8 levels, factor 1.2; INTER_LINEAR_EXACT resize in 8.8
// fixed point); when one block of scratch per level fits the workspace the levels run on separate streams.
#include "common.cuh"

size_t vsb_fast_scratch_bytes(int w, int h, int count);
int vsb_orb_lp_streams(vsb_ctx_t* ctx);            // capi.cu: creates the context's per-level streams and events on first use
int vsb_fast_detect_ws(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count, int threshold,
                       int nonmax, int cap, int32_t* kp_xy, int32_t* kp_score, int32_t* n_kp, void* scratch, void* stream);

namespace {

constexpr int ORB_EDGE = 31;
constexpr int ORB_HALF = 15;
constexpr int ORB_KP_PER_WARP = 2;     // key points a warp of the angle / descriptor kernels is sized for (fewer, longer-lived blocks)

__device__ __align__(16) const signed char c_pattern[256 * 4] = {
#include "orb_pattern.inc"
};
// end of each row of the circular patch of radius 15 (orb.cpp: umax)
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
// cv::getGaussianKernel(7, 2, CV_32F), bit patterns
__constant__ uint32_t c_gauss[4] = {1032826801u, 1040595070u, 1044597305u, 1046301408u};

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    // 256 threads; returns the exclusive prefix of v over the block, total = block sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < 8 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        if (lane < 8) s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[7];
    const int base = warp ? s_warp[warp - 1] : 0;
    __syncthreads();
    return base + inc - v;
}

// The bin a descending walk over a 256-bin histogram stops in: the largest b >= floor_bin with sum(hist[b..255]) >= want, and
// the sum over the bins above it; floor_bin (with the sum over the bins above IT) when no bin reaches `want`.  All 256 threads
// call it; the result lands in s_out[0..1].  (The walk itself, on one thread, was 256 dependent shared-memory loads per radix
// pass: ~10 k cycles, most of these kernels' time and of the detector's fixed cost per call.)
__device__ __forceinline__ void hist_select_desc(const int* s_hist, int want, int floor_bin, int* s_warp, int* s_out) {
    const int b = 255 - (int)threadIdx.x;
    const int v = b >= floor_bin ? s_hist[b] : 0;
    int total;
    const int excl = block_exclusive_scan(v, s_warp, total);       // over the bins above b
    if (b > floor_bin ? (excl + v >= want && excl < want) : (b == floor_bin && excl < want)) { s_out[0] = b; s_out[1] = excl; }
    __syncthreads();
}

// ---- stage 1: border filter + retainBest(2n) on the integer FAST score, order preserved --------------------------------
// One block per frame.  Scores are integers 1..255, so the n-th best is found on a 256-bin histogram.
__global__ void __launch_bounds__(256)
orb_select_fast_kernel(const int32_t* __restrict__ fxy, const int32_t* __restrict__ fscore, const int32_t* __restrict__ nfast,
                       int fcap, int w, int h, int keep_n, int32_t* __restrict__ sel_xy, int32_t* __restrict__ n_sel) {
    __shared__ int s_hist[256];
    __shared__ int s_warp[8];
    __shared__ int s_sel[2];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(nfast[f], fcap);
    const int32_t* xy = fxy + (size_t)f * fcap * 2;
    const int32_t* sc = fscore + (size_t)f * fcap;
    int32_t* out = sel_xy + (size_t)f * fcap * 2;
    s_hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int x = xy[2 * i], y = xy[2 * i + 1];
        if (x >= ORB_EDGE && x < w - ORB_EDGE && y >= ORB_EDGE && y < h - ORB_EDGE) atomicAdd(&s_hist[min(max(sc[i], 0), 255)], 1);
    }
    __syncthreads();
    // threshold = the largest score s with at least keep_n corners of score >= s (everything when there are no more than keep_n)
    hist_select_desc(s_hist, keep_n > 0 ? keep_n : 1, 0, s_warp, s_sel);
    // (a bin >= 1 that reaches keep_n is the threshold whether or not there are more than keep_n corners in all — when there are
    // not, nothing lies below it; bin 0 keeps everything)
    const int thr = keep_n < 0 ? 0 : (keep_n == 0 ? 256 : s_sel[0]);
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        int x = 0, y = 0, flag = 0;
        if (i < n) {
            x = xy[2 * i]; y = xy[2 * i + 1];
            flag = (x >= ORB_EDGE && x < w - ORB_EDGE && y >= ORB_EDGE && y < h - ORB_EDGE && sc[i] >= thr) ? 1 : 0;
        }
        int total;
        const int pos = block_exclusive_scan(flag, s_warp, total);
        if (flag) { out[2 * (base + pos)] = x; out[2 * (base + pos) + 1] = y; }
        base += total;
    }
    if (tid == 0) n_sel[f] = base;
}

// ---- stage 2: Harris response of every selected corner (HarrisResponses, blockSize 7, k 0.04) --------------------------
__global__ void __launch_bounds__(128, 8)
orb_harris_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, const int32_t* __restrict__ sel_xy,
                  const int32_t* __restrict__ n_sel, int fcap, float* __restrict__ resp) {
    const int f = blockIdx.y;
    const int n = n_sel[f];
    for (int i = blockIdx.x * 128 + threadIdx.x; i < n; i += gridDim.x * 128) {
    const int x0 = sel_xy[((size_t)f * fcap + i) * 2], y0 = sel_xy[((size_t)f * fcap + i) * 2 + 1];
    const uint8_t* base = img + (size_t)f * img_stride + (size_t)(y0 - 4) * pitch + (x0 - 4);
    // 9x9 neighbourhood, three rows at a time
    int a = 0, b = 0, c = 0;
    int r0[9], r1[9], r2[9];
#pragma unroll
    for (int j = 0; j < 9; j++) { r0[j] = __ldg(base + j); r1[j] = __ldg(base + pitch + j); }
#pragma unroll
    for (int row = 0; row < 7; row++) {
        const uint8_t* p = base + (size_t)(row + 2) * pitch;
#pragma unroll
        for (int j = 0; j < 9; j++) r2[j] = __ldg(p + j);
#pragma unroll
        for (int j = 1; j < 8; j++) {
            const int Ix = (r1[j + 1] - r1[j - 1]) * 2 + (r0[j + 1] - r0[j - 1]) + (r2[j + 1] - r2[j - 1]);
            const int Iy = (r2[j] - r0[j]) * 2 + (r2[j - 1] - r0[j - 1]) + (r2[j + 1] - r0[j + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
#pragma unroll
        for (int j = 0; j < 9; j++) { r0[j] = r1[j]; r1[j] = r2[j]; }
    }
    const float scale = __fdiv_rn(1.f, __fmul_rn((float)((1 << 2) * 7), 255.f));
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float s = __fadd_rn(fa, fb);
    const float t3 = __fmul_rn(__fmul_rn(0.04f, s), s);
    const float d = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), t3);
    resp[(size_t)f * fcap + i] = __fmul_rn(d, s4);
    }
}

// ---- stage 3: retainBest(n) on the float response (radix select of the n-th largest), order preserved -------------------
__device__ __forceinline__ uint32_t float_key(float v) {      // unsigned order == float order
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
orb_select_harris_kernel(const int32_t* __restrict__ sel_xy, const float* __restrict__ resp, const int32_t* __restrict__ n_sel,
                         int fcap, int keep_n, int cap, int32_t* __restrict__ kp_xy, float* __restrict__ kp_resp,
                         int32_t* __restrict__ n_kp) {
    __shared__ int s_hist[256];
    __shared__ int s_warp[8];
    __shared__ uint32_t s_prefix;
    __shared__ int s_want;
    __shared__ int s_sel[2];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = n_sel[f];
    const int32_t* xy = sel_xy + (size_t)f * fcap * 2;
    const float* r = resp + (size_t)f * fcap;
    uint32_t thr_key = 0;                                 // keep everything
    if (keep_n >= 0 && n > keep_n) {
        if (keep_n == 0) thr_key = 0xffffffffu;
        else {
            // the keep_n-th largest key, most significant byte first
            if (tid == 0) { s_prefix = 0; s_want = keep_n; }
            __syncthreads();
            for (int shift = 24; shift >= 0; shift -= 8) {
                s_hist[tid] = 0;
                __syncthreads();
                const uint32_t prefix = s_prefix;
                const uint32_t mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
                for (int i = tid; i < n; i += 256) {
                    const uint32_t k = float_key(r[i]);
                    if ((k & mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255], 1);
                }
                __syncthreads();
                hist_select_desc(s_hist, s_want, 0, s_warp, s_sel);
                if (tid == 0) {
                    s_prefix = prefix | ((uint32_t)s_sel[0] << shift);
                    s_want -= s_sel[1];
                }
                __syncthreads();
            }
            thr_key = s_prefix;
        }
    }
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        int flag = 0;
        float v = 0.f;
        if (i < n) { v = r[i]; flag = float_key(v) >= thr_key ? 1 : 0; }
        if (keep_n == 0) flag = 0;
        int total;
        const int pos = block_exclusive_scan(flag, s_warp, total);
        if (flag && base + pos < cap) {
            kp_xy[((size_t)f * cap + base + pos) * 2] = xy[2 * i];
            kp_xy[((size_t)f * cap + base + pos) * 2 + 1] = xy[2 * i + 1];
            kp_resp[(size_t)f * cap + base + pos] = v;
        }
        base += total;
    }
    if (tid == 0) n_kp[f] = base;
}

// ---- stage 4: intensity-centroid orientation (ICAngles + cv::fastAtan2), one warp per key point -------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float c = (float)(180 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, c), p3 = __fmul_rn(-0.3258083974640975f, c),
                p5 = __fmul_rn(0.1555786518463281f, c), p7 = __fmul_rn(-0.04432655554792128f, c);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    const bool xs = ax >= ay;
    const float q = xs ? __fdiv_rn(ay, __fadd_rn(ax, eps)) : __fdiv_rn(ax, __fadd_rn(ay, eps));
    const float q2 = __fmul_rn(q, q);
    float a = __fmul_rn(p7, q2);
    a = __fadd_rn(a, p5); a = __fmul_rn(a, q2); a = __fadd_rn(a, p3); a = __fmul_rn(a, q2); a = __fadd_rn(a, p1);
    a = __fmul_rn(a, q);
    if (!xs) a = __fsub_rn(90.f, a);
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}
__global__ void __launch_bounds__(256, 8)
orb_angle_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, const int32_t* __restrict__ kp_xy,
                 const int32_t* __restrict__ n_kp, int cap, float* __restrict__ kp_angle) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int nk = min(n_kp[f], cap);
    for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nk; k += gridDim.x * 8) {
    const int x0 = kp_xy[((size_t)f * cap + k) * 2], y0 = kp_xy[((size_t)f * cap + k) * 2 + 1];
    const uint8_t* center = img + (size_t)f * img_stride + (size_t)y0 * pitch + x0;
    // lane u-index: u = lane - 15 for lanes 0..30 (31 columns); rows v = 0..15 handled by every lane for its column
    int m01 = 0, m10 = 0;
    const int u = lane - ORB_HALF;
    if (lane < 31) {
        m10 += u * (int)__ldg(center + u);
        const int au = u < 0 ? -u : u;
#pragma unroll
        for (int v = 1; v <= ORB_HALF; v++) {
            if (au <= c_umax[v]) {
                const int vp = __ldg(center + u + v * pitch), vm = __ldg(center + u - v * pitch);
                m01 += v * (vp - vm);
                m10 += u * (vp + vm);
            }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { m01 += __shfl_down_sync(0xffffffffu, m01, o); m10 += __shfl_down_sync(0xffffffffu, m10, o); }
    if (lane == 0) kp_angle[(size_t)f * cap + k] = fast_atan2_deg((float)m01, (float)m10);
    }
}

// ---- stage 5: the blur ORB applies before describing: 7x7 sigma 2, generic float separable filter, reflect-101 ---------
// One 64x32 output tile per block.  The input tile (+3 halo, 72 bytes per row so rows start on the word before the tile) is
// loaded as aligned 32-bit words where the image allows it; the row pass produces four adjacent outputs per thread, the column
// pass four adjacent outputs packed into one 32-bit store.  Every output is the same sequence of rounded operations as the
// oracle's: s = x0 k0, s = fma(x_i, k_i, s) along the row; s = r3 k3, s = fma(r[3+i] + r[3-i], k[3+i], s) down the column.
constexpr int BTW = 64, BTH = 32;
// BORDER_REFLECT_101 index for positions at most n - 1 outside the range (the blur reaches 3 outside an image of more than 62
// pixels), branch-free; positions further out — tile padding nobody reads — are clamped into the range.
__device__ __forceinline__ int refl101(int p, int n) {
    const int a = abs(p);
    return max(min(a, 2 * n - 2 - a), 0);
}
// VEC: frames whose base, stride and pitch are multiples of 16 (every level image; 752- and 640-wide frames): the tile is loaded
// as 16-byte vectors from x0 - 16 on (6 per row, one per thread, rows outside the image from their reflected row), and the
// up to six columns outside the image are then copied from their reflections inside the tile.
template <bool VEC>
__global__ void __launch_bounds__(256)
orb_blur_kernel(const uint8_t* __restrict__ img, long long img_stride, int pitch, int w, int h, uint8_t* __restrict__ out, int opitch) {
    constexpr int XOFF = VEC ? 16 : 4;                            // tile column of x0
    constexpr int BTP = VEC ? 192 : 72;                           // row pitch of the tile in bytes
    __shared__ __align__(16) uint8_t s_in[BTH + 6][BTP];          // columns x0 - XOFF ..
    __shared__ __align__(16) float s_row[BTH + 6][BTW];
    const int f = blockIdx.z, x0 = blockIdx.x * BTW, y0 = blockIdx.y * BTH, tid = threadIdx.x;
    const uint8_t* src = img + (size_t)f * img_stride;
    if (VEC) {
        if (tid < (BTH + 6) * 6) {
            const int ry = tid / 6, pc = tid - 6 * ry;
            const int gx = x0 - 16 + 16 * pc;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (gx >= 0 && gx + 15 < pitch) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)refl101(y0 + ry - 3, h) * pitch + gx));
            *reinterpret_cast<uint4*>(&s_in[ry][16 * pc]) = v;
        }
        if (x0 == 0 || x0 + BTW + 3 > w) {                       // (uniform) a tile that reaches past the left or right border
            __syncthreads();
            if (tid < (BTH + 6) * 6) {
                const int ry = tid / 6, k = tid - 6 * ry;
                const int x = k < 3 ? k - 3 : w + k - 3;          // -3 .. -1, w .. w + 2
                if (x >= x0 - 3 && x < x0 + BTW + 3) s_in[ry][x - x0 + 16] = s_in[ry][refl101(x, w) - x0 + 16];
            }
        }
    } else {
    const bool words_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)pitch) & 3u) == 0;
    for (int i = tid; i < (BTH + 6) * (BTP / 4); i += 256) {
        const int ry = i / (BTP / 4), wx = i - ry * (BTP / 4);
        const int gy = y0 + ry - 3, gx = x0 - 4 + 4 * wx;
        uint32_t v;
        const uint8_t* row = src + (size_t)refl101(gy, h) * pitch;
        if (words_ok && gx >= 0 && gx + 3 < w) {
            v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
        } else {
            v = 0u;
#pragma unroll
            for (int j = 0; j < 4; j++) v |= (uint32_t)__ldg(row + refl101(gx + j, w)) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(&s_in[ry][4 * wx]) = v;
    }
    }
    __syncthreads();
    float k[7];
#pragma unroll
    for (int i = 0; i < 4; i++) { k[i] = __uint_as_float(c_gauss[i]); k[6 - i] = k[i]; }
    // row pass: (BTH + 6) rows x 16 quads
    for (int i = tid; i < (BTH + 6) * (BTW / 4); i += 256) {
        const int ry = i / (BTW / 4), q = i - ry * (BTW / 4);
        // outputs x0 + 4q .. +3 need inputs x0 + 4q - 3 .. x0 + 4q + 6 = tile bytes 4q + 1 .. 4q + 10: three aligned words
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(&s_in[ry][4 * q + XOFF - 4]);
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        float px[12];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            px[j] = (float)((w0 >> (8 * j)) & 0xFFu);
            px[4 + j] = (float)((w1 >> (8 * j)) & 0xFFu);
            px[8 + j] = (float)((w2 >> (8 * j)) & 0xFFu);
        }
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float acc = __fmul_rn(px[1 + j], k[0]);
#pragma unroll
            for (int t = 1; t < 7; t++) acc = __fmaf_rn(px[1 + j + t], k[t], acc);
            op[j] = acc;
        }
        *reinterpret_cast<float4*>(&s_row[ry][4 * q]) = o;
    }
    __syncthreads();
    // column pass: BTH rows x 16 quads, one packed store per quad
    uint8_t* dst = out + (size_t)f * opitch * h;          // rows opitch bytes apart (a multiple of 4 wherever the caller can pad)
    const bool store_words = (opitch & 3) == 0;
    for (int i = tid; i < BTH * (BTW / 4); i += 256) {
        const int cy = i / (BTW / 4), q = i - cy * (BTW / 4);
        const int x = x0 + 4 * q, y = y0 + cy;
        if (x >= w || y >= h) continue;
        float4 r[7];
#pragma unroll
        for (int t = 0; t < 7; t++) r[t] = *reinterpret_cast<const float4*>(&s_row[cy + t][4 * q]);
        uint32_t packed = 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float* c0 = reinterpret_cast<const float*>(&r[0]) + j;      // stride 4 floats between rows
            float acc = __fmul_rn(c0[12], k[3]);
#pragma unroll
            for (int t = 1; t <= 3; t++) acc = __fmaf_rn(__fadd_rn(c0[4 * (3 + t)], c0[4 * (3 - t)]), k[3 + t], acc);
            uint32_t v;                                           // cv::saturate_cast<uchar>(cvRound(acc)): the conversion saturates by itself
            asm("cvt.rni.u8.f32 %0, %1;" : "=r"(v) : "f"(acc));
            packed |= v << (8 * j);
        }
        uint8_t* o = dst + (size_t)y * opitch + x;
        if (store_words && x + 3 < opitch) {                 // (bytes past w land in the row's padding)
            *reinterpret_cast<uint32_t*>(o) = packed;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (x + j < w) o[j] = (uint8_t)(packed >> (8 * j));
        }
    }
}

// ---- stage 6: steered rBRIEF (computeOrbDescriptors, WTA_K = 2): one warp per key point, one descriptor byte per lane -----
// cos / sin of every key point's angle, one THREAD per key point: inside the descriptor kernel the double-precision pair ran on
// one lane of a warp while the other 31 waited (40 % of that kernel's issue slots, all of them latency-bound).
__global__ void __launch_bounds__(256)
orb_trig_kernel(const float* __restrict__ kp_angle, const int32_t* __restrict__ n_kp, int cap, int cs_stride, float2* __restrict__ cs) {
    const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= min(n_kp[f], cap)) return;
    const float angle = __fmul_rn(kp_angle[(size_t)f * cap + i], (float)(3.14159265358979323846 / 180.f));
    cs[(size_t)f * cs_stride + i] = make_float2((float)cos((double)angle), (float)sin((double)angle));
}
// A lane's eight tests (its descriptor byte) stay in eight registers for every key point the warp describes: four signed bytes
// (x0, y0, x1, y1) per test, converted straight out of the register bytes.  (Kept in shared memory the pattern was read one
// byte at a time at a 32-byte lane stride: eight-way bank conflicts, 55 % of the shared-memory pipe.)
__global__ void __launch_bounds__(256, 4)
orb_describe_kernel(const uint8_t* __restrict__ blurred, int w, int h, int bpitch, const int32_t* __restrict__ kp_xy,
                    const float* __restrict__ kp_angle, const float2* __restrict__ cs, int cs_stride,
                    const int32_t* __restrict__ n_kp, int cap, uint8_t* __restrict__ desc) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    uint32_t pat[8];
    {
        const uint4* p4 = reinterpret_cast<const uint4*>(c_pattern) + 2 * lane;
        const uint4 a = __ldg(p4), b = __ldg(p4 + 1);
        pat[0] = a.x; pat[1] = a.y; pat[2] = a.z; pat[3] = a.w; pat[4] = b.x; pat[5] = b.y; pat[6] = b.z; pat[7] = b.w;
    }
    const int nk = min(n_kp[f], cap);
    for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nk; k += gridDim.x * 8) {
    const int x0 = kp_xy[((size_t)f * cap + k) * 2], y0 = kp_xy[((size_t)f * cap + k) * 2 + 1];
    float a, b;
    if (cs) {
        const float2 t = __ldg(cs + (size_t)f * cs_stride + k);
        a = t.x; b = t.y;
    } else {
        float angle = kp_angle[(size_t)f * cap + k];
        angle = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
        a = 0.f; b = 0.f;
        if (lane == 0) { a = (float)cos((double)angle); b = (float)sin((double)angle); }      // one double-precision pair per key point
        a = __shfl_sync(0xffffffffu, a, 0);
        b = __shfl_sync(0xffffffffu, b, 0);
    }
    const uint8_t* center = blurred + (size_t)f * bpitch * h + (size_t)y0 * bpitch + x0;
    int val = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int v[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float px = (float)(signed char)(pat[i] >> (16 * q)), py = (float)(signed char)(pat[i] >> (16 * q + 8));
            const float x = __fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b));
            const float y = __fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a));
            v[q] = __ldg(center + __float2int_rn(y) * bpitch + __float2int_rn(x));
        }
        val |= (v[0] < v[1]) << i;
    }
    desc[((size_t)f * cap + k) * 32 + lane] = (uint8_t)val;
    }
}

// ---- the scale pyramid: cv::resize(.., INTER_LINEAR_EXACT) for 8-bit images (resize_bitExact, 8.8 fixed point) -----------
__device__ __forceinline__ void lin_exact_coeff(int v, int ssize, int dsize, int& ofs, int& c1, int& edge) {
    const double inv = __ddiv_rn((double)dsize, (double)ssize);
    const double scale = __ddiv_rn(1.0, inv);
    const double fval = __dsub_rn(__dmul_rn(scale, __dadd_rn((double)v, 0.5)), 0.5);
    const int ival = (int)floor(fval);
    edge = 0; ofs = 0; c1 = 0;
    if (ival >= 0 && ssize > 1) {
        if (ival < ssize - 1) { ofs = ival; c1 = __double2int_rn(__dmul_rn(__dsub_rn(fval, (double)ival), 256.0)); }
        else { ofs = ssize - 1; edge = 1; }
    } else {
        edge = -1;
    }
}
// coefficient table of one axis: (offset, 8.8 weight of the second sample, edge flag) per destination index
__global__ void orb_resize_coeff_kernel(int sw, int sh, int dw, int dh, int4* __restrict__ tab) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= dw + dh) return;
    int ofs, c1, edge;
    if (i < dw) lin_exact_coeff(i, sw, dw, ofs, c1, edge); else lin_exact_coeff(i - dw, sh, dh, ofs, c1, edge);
    tab[i] = make_int4(ofs, c1, edge, 0);
}
__global__ void __launch_bounds__(256)
orb_resize_kernel(const uint8_t* __restrict__ src, long long src_stride, int spitch, int sw, int sh, uint8_t* __restrict__ dst,
                  int dw, int dh, int dpitch, const int4* __restrict__ tab) {
    // four adjacent output pixels per thread, one packed store
    const int xq = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, y = blockIdx.y * 8 + (threadIdx.x >> 5), f = blockIdx.z;
    if (xq >= dw || y >= dh) return;
    const int4 ty = __ldg(tab + dw + y);
    const int oy = ty.x, ay = ty.y, ey = ty.z;
    const uint8_t* img = src + (size_t)f * src_stride;
    const int y0 = ey < 0 ? 0 : oy, y1 = ey != 0 ? y0 : oy + 1;
    const uint8_t* r0 = img + (size_t)y0 * spitch;
    const uint8_t* r1 = img + (size_t)y1 * spitch;
    uint32_t packed = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int x = xq + j;
        if (x >= dw) break;
        const int4 tx = __ldg(tab + x);
        const int ox = tx.x, ax = tx.y, ex = tx.z;
        uint32_t h0, h1;
        if (ex < 0) { h0 = (uint32_t)__ldg(r0) << 8; h1 = (uint32_t)__ldg(r1) << 8; }
        else if (ex > 0) { h0 = (uint32_t)__ldg(r0 + sw - 1) << 8; h1 = (uint32_t)__ldg(r1 + sw - 1) << 8; }
        else {
            h0 = (uint32_t)(256 - ax) * __ldg(r0 + ox) + (uint32_t)ax * __ldg(r0 + ox + 1);
            h1 = (uint32_t)(256 - ax) * __ldg(r1 + ox) + (uint32_t)ax * __ldg(r1 + ox + 1);
        }
        uint32_t v = ey != 0 ? h0 << 8 : (uint32_t)(256 - ay) * h0 + (uint32_t)ay * h1;
        v = (v + (1u << 15)) >> 16;
        packed |= min(v, 255u) << (8 * j);
    }
    uint8_t* o = dst + ((size_t)f * dh + y) * dpitch + xq;
    if ((dpitch & 3) == 0) {
        *reinterpret_cast<uint32_t*>(o) = packed;                  // rows padded to a multiple of 4: the address is aligned, bytes past dw are padding
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (xq + j < dw) o[j] = (uint8_t)(packed >> (8 * j));
    }
}

// The same resize for the case every level of the scale pyramid is in: source rows on 4-byte boundaries and a scale of at most
// 2.  One thread owns FOUR adjacent destination columns for RZ_ROWS destination rows.  Per source row it loads the (up to three)
// aligned words that cover its taps, funnel-shifts them into an 8-byte window that starts at its first tap, picks every
// column's two neighbouring bytes with one PRMT (selectors formed once per thread) and forms (256 - a) p0 + a p1 with one
// two-way 16 x 8-bit dot product; the border cases are ordinary taps in the folded table (before the first sample: offset 0,
// weight 0; past the last: offset size - 2, weight 256), so nothing branches per pixel.  A source row's horizontal pass is
// kept for the next destination row when that one starts on it (scale 1.2: most of the time).  Rounded sums are below
// 2^24 with the result in byte 2, so the four results are packed with three PRMTs.  ~12 instructions per output (the
// per-pixel kernel above: ~55).
constexpr int RZ_ROWS = 8;
__global__ void orb_resize_coeff_folded_kernel(int sw, int sh, int dw, int dh, int2* __restrict__ tab) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= dw + dh) return;
    int ofs, c1, edge;
    const int ssize = i < dw ? sw : sh;
    if (i < dw) lin_exact_coeff(i, sw, dw, ofs, c1, edge); else lin_exact_coeff(i - dw, sh, dh, ofs, c1, edge);
    if (edge < 0) { ofs = 0; c1 = 0; }
    else if (edge > 0) { ofs = ssize - 2; c1 = 256; }
    tab[i] = make_int2(ofs, c1);
}
__device__ __forceinline__ void resize_row_taps(const uint32_t* __restrict__ rp, int wi0, int wi1, int wi2, uint32_t shift,
                                                const uint32_t (&sel)[4], const uint32_t (&cf)[4], uint32_t (&hout)[4]) {
    const uint32_t w0 = __ldg(rp + wi0), w1 = __ldg(rp + wi1), w2 = __ldg(rp + wi2);
    const uint32_t a = __funnelshift_r(w0, w1, shift), b = __funnelshift_r(w1, w2, shift);
#pragma unroll
    for (int j = 0; j < 4; j++) hout[j] = __dp2a_lo(cf[j], __byte_perm(a, b, sel[j]), 0u);
}
__global__ void __launch_bounds__(256)
orb_resize4_kernel(const uint8_t* __restrict__ src, long long src_stride, int spitch, uint8_t* __restrict__ dst, int dw, int dh,
                   int dpitch, const int2* __restrict__ tab) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int xq = (blockIdx.x * 32 + lane) * 4, yb = (blockIdx.y * 8 + wrp) * RZ_ROWS, f = blockIdx.z;
    if (xq >= dw || yb >= dh) return;
    // column constants: window start, byte selectors inside the window, packed weights, mask of the columns that exist
    int ox[4];
    uint32_t cf[4], sel[4], mask = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int2 t = __ldg(tab + min(xq + j, dw - 1));
        ox[j] = t.x;
        cf[j] = (uint32_t)(256 - t.y) | ((uint32_t)t.y << 16);
        if (xq + j < dw) mask |= 0xFFu << (8 * j);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t d = (uint32_t)(ox[j] - ox[0]);                  // 0 .. 6 for a scale of at most 2
        sel[j] = d | ((d + 1u) << 4);
    }
    const int last_word = (spitch >> 2) - 1;
    const int wi0 = ox[0] >> 2, wi1 = min(wi0 + 1, last_word), wi2 = min(wi0 + 2, last_word);   // (a clamped word holds no tap)
    const uint32_t shift = (uint32_t)(ox[0] & 3) * 8u;
    const uint32_t* img = reinterpret_cast<const uint32_t*>(src + (size_t)f * src_stride);
    const int wpitch = spitch >> 2;
    uint8_t* out = dst + ((size_t)f * dh + yb) * dpitch + xq;
    int have = -1;
    uint32_t hp[4] = {0u, 0u, 0u, 0u};
#pragma unroll 2
    for (int k = 0; k < RZ_ROWS; k++) {
        if (yb + k >= dh) break;
        const int2 ty = __ldg(tab + dw + yb + k);
        const int oy = ty.x;
        const uint32_t ay = (uint32_t)ty.y;
        uint32_t h0[4], h1[4];
        if (oy == have) {                                            // (uniform over the warp: one destination row per warp)
#pragma unroll
            for (int j = 0; j < 4; j++) h0[j] = hp[j];
        } else {
            resize_row_taps(img + (size_t)oy * wpitch, wi0, wi1, wi2, shift, sel, cf, h0);
        }
        resize_row_taps(img + (size_t)(oy + 1) * wpitch, wi0, wi1, wi2, shift, sel, cf, h1);
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j] = (256u - ay) * h0[j] + (ay * h1[j] + (1u << 15));   // < 2^24: the rounded result is byte 2
            hp[j] = h1[j];
        }
        have = oy + 1;
        const uint32_t lo = __byte_perm(v[0], v[1], 0x0062), hi = __byte_perm(v[2], v[3], 0x0062);
        *reinterpret_cast<uint32_t*>(out + (size_t)k * dpitch) = __byte_perm(lo, hi, 0x5410) & mask;   // dpitch % 4 == 0: aligned, bytes past dw are padding
    }
}

// The same arithmetic with the source rows staged in shared memory: orb_resize4_kernel is bound by the latency of its global loads
// (ncu: 12 of 17 cycles per issued instruction on the long scoreboard, DRAM 20 % busy, 3-6 four-byte loads in flight per thread).
// Here a block (128 destination columns x 64 destination rows) first requests the source region it needs as 16-byte vectors, all
// of a thread's loads in flight at once, and the taps then come from shared memory.  Needs source base, stride and pitch to be
// multiples of 16 (every level image, 752- / 640-wide frames); nvec = 16-byte vectors per tile row, nrows_cap = tile rows.
__global__ void __launch_bounds__(256, 4)
orb_resize_tile_kernel(const uint8_t* __restrict__ src, long long src_stride, int spitch, uint8_t* __restrict__ dst, int dw, int dh,
                       int dpitch, const int2* __restrict__ tab, int nvec, int nrows_cap) {
    extern __shared__ uint4 s_tile[];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, f = blockIdx.z;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * (8 * RZ_ROWS);
    const int xs = __ldg(tab + x0).x & ~15;                            // first source byte column of the tile
    const int oy_first = __ldg(tab + dw + y0).x;
    const int nrows = min(__ldg(tab + dw + min(y0 + 8 * RZ_ROWS - 1, dh - 1)).x + 2 - oy_first, nrows_cap);
    const uint8_t* img = src + (size_t)f * src_stride;
    // 16 x 16 threads over (vector column, row); pointers advance by additions only
    for (int c = threadIdx.x & 15; c < nvec; c += 16) {
        const int gx = xs + 16 * c;
        const int r0 = threadIdx.x >> 4;
        uint4* sp = s_tile + r0 * nvec + c;
        if (gx < spitch) {
            const uint8_t* gp = img + (size_t)(oy_first + r0) * spitch + gx;
            const size_t gstep = (size_t)16 * spitch;
#pragma unroll 4
            for (int r = r0; r < nrows; r += 16) {
                *sp = __ldg(reinterpret_cast<const uint4*>(gp));
                gp += gstep; sp += 16 * nvec;
            }
        } else {
            for (int r = r0; r < nrows; r += 16) { *sp = make_uint4(0u, 0u, 0u, 0u); sp += 16 * nvec; }
        }
    }
    __syncthreads();
    const int xq = x0 + lane * 4, yb = y0 + wrp * RZ_ROWS;
    if (xq >= dw || yb >= dh) return;
    int ox[4];
    uint32_t cf[4], sel[4], mask = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int2 t = __ldg(tab + min(xq + j, dw - 1));
        ox[j] = t.x;
        cf[j] = (uint32_t)(256 - t.y) | ((uint32_t)t.y << 16);
        if (xq + j < dw) mask |= 0xFFu << (8 * j);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t d = (uint32_t)(ox[j] - ox[0]);
        sel[j] = d | ((d + 1u) << 4);
    }
    const int wrow = nvec * 4;                                         // words per tile row
    const int wi0 = min((ox[0] - xs) >> 2, wrow - 3);                 // (the tile row is sized for the 12-byte window of its last thread)
    const uint32_t shift = (uint32_t)(ox[0] & 3) * 8u;                 // (xs is a multiple of 4)
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(s_tile);
    uint8_t* out = dst + ((size_t)f * dh + yb) * dpitch + xq;
    int have = -1;
    uint32_t hp[4] = {0u, 0u, 0u, 0u};
#pragma unroll 2
    for (int k = 0; k < RZ_ROWS; k++) {
        if (yb + k >= dh) break;
        const int2 ty = __ldg(tab + dw + yb + k);
        const int oy = ty.x;
        const uint32_t ay = (uint32_t)ty.y;
        uint32_t h0[4], h1[4];
        if (oy == have) {
#pragma unroll
            for (int j = 0; j < 4; j++) h0[j] = hp[j];
        } else {
            const uint32_t* rp = tile + (oy - oy_first) * wrow + wi0;
            const uint32_t a = __funnelshift_r(rp[0], rp[1], shift), b = __funnelshift_r(rp[1], rp[2], shift);
#pragma unroll
            for (int j = 0; j < 4; j++) h0[j] = __dp2a_lo(cf[j], __byte_perm(a, b, sel[j]), 0u);
        }
        {
            const uint32_t* rp = tile + (oy + 1 - oy_first) * wrow + wi0;
            const uint32_t a = __funnelshift_r(rp[0], rp[1], shift), b = __funnelshift_r(rp[1], rp[2], shift);
#pragma unroll
            for (int j = 0; j < 4; j++) h1[j] = __dp2a_lo(cf[j], __byte_perm(a, b, sel[j]), 0u);
        }
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j] = (256u - ay) * h0[j] + (ay * h1[j] + (1u << 15));
            hp[j] = h1[j];
        }
        have = oy + 1;
        const uint32_t lo = __byte_perm(v[0], v[1], 0x0062), hi = __byte_perm(v[2], v[3], 0x0062);
        *reinterpret_cast<uint32_t*>(out + (size_t)k * dpitch) = __byte_perm(lo, hi, 0x5410) & mask;
    }
}

// appends one level's key points to the frame's list: pt = level coordinates * scale (float), octave = level
__global__ void __launch_bounds__(256)
orb_append_kernel(const int32_t* __restrict__ lxy, const float* __restrict__ lresp, const float* __restrict__ langle,
                  const uint8_t* __restrict__ ldesc, const int32_t* __restrict__ ln, int cap, float scale, int level,
                  float* __restrict__ kp_xy, int32_t* __restrict__ kp_octave, float* __restrict__ kp_resp,
                  float* __restrict__ kp_angle, uint8_t* __restrict__ desc, const int32_t* __restrict__ n_total) {
    const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    const int n = min(ln[f], cap);
    if (i >= n) return;
    const int o = n_total[f] + i;
    if (o >= cap) return;
    const size_t src = (size_t)f * cap + i, dst = (size_t)f * cap + o;
    kp_xy[2 * dst] = __fmul_rn((float)lxy[2 * src], scale);
    kp_xy[2 * dst + 1] = __fmul_rn((float)lxy[2 * src + 1], scale);
    if (kp_octave) kp_octave[dst] = level;
    kp_resp[dst] = lresp[src];
    kp_angle[dst] = langle[src];
    if (desc) {
        const uint4* s4 = reinterpret_cast<const uint4*>(ldesc + src * 32);
        uint4* d4 = reinterpret_cast<uint4*>(desc + dst * 32);
        d4[0] = s4[0]; d4[1] = s4[1];
    }
}
__global__ void orb_advance_kernel(int32_t* __restrict__ n_total, const int32_t* __restrict__ ln, int count) {
    const int f = blockIdx.x * 256 + threadIdx.x;
    if (f < count) n_total[f] += ln[f];
}

struct OrbScratch {
    int32_t *fxy, *fsc, *sel, *nfast, *nsel;
    float* resp;
    uint8_t* blurred;
    int fcap;
};
inline size_t orb_al(size_t b) { return (b + 255) & ~(size_t)255; }
inline size_t orb_scratch_bytes(int chunk, int w, int h, bool describe) {
    const size_t fcap = (size_t)w * h / 4 + 16;
    return 2 * orb_al(chunk * fcap * 8) + orb_al(chunk * fcap * 4) + orb_al(chunk * fcap * 4) + 2 * orb_al((size_t)chunk * 4) +
           (describe ? orb_al((size_t)chunk * ((w + 15) & ~15) * h) : 0);       // (rows of the level images are padded to 16 bytes)
}
inline OrbScratch orb_scratch_carve(uint8_t*& p, int chunk, int w, int h, bool describe) {
    OrbScratch s;
    s.fcap = w * h / 4 + 16;
    const size_t fcap = (size_t)s.fcap;
    s.fxy = reinterpret_cast<int32_t*>(p); p += orb_al(chunk * fcap * 8);
    s.sel = reinterpret_cast<int32_t*>(p); p += orb_al(chunk * fcap * 8);
    s.fsc = reinterpret_cast<int32_t*>(p); p += orb_al(chunk * fcap * 4);
    s.resp = reinterpret_cast<float*>(p); p += orb_al(chunk * fcap * 4);
    s.nfast = reinterpret_cast<int32_t*>(p); p += orb_al((size_t)chunk * 4);
    s.nsel = reinterpret_cast<int32_t*>(p); p += orb_al((size_t)chunk * 4);
    s.blurred = describe ? p : nullptr;
    if (describe) p += orb_al((size_t)chunk * ((w + 15) & ~15) * h);
    return s;
}

// the one-level pipeline on zc frames of w x h (fcap is taken from the scratch, sized for the largest level)
int orb_one_level(vsb_ctx_t* ctx, const OrbScratch& S, void* fast_scratch, const uint8_t* in, int64_t img_stride, int pitch, int w, int h, int zc,
                  int nfeatures, int fast_threshold, int cap, int32_t* kp_xy, float* kp_resp, float* kp_angle, uint8_t* desc,
                  int32_t* n_kp, cudaStream_t st, int bpitch = 0) {
    if (bpitch <= 0) bpitch = w;                 // row pitch of the blurred copy (the scratch holds w0 x h0 bytes per frame)
    const int fcap = S.fcap;
    int rc = vsb_fast_detect_ws(ctx, in, img_stride, pitch, w, h, zc, fast_threshold, 1, fcap, S.fxy, S.fsc, S.nfast, fast_scratch, (void*)st);
    if (rc) return rc;
    ProfScope ps(ctx, VSB_K_ORB, st);
    orb_select_fast_kernel<<<zc, 256, 0, st>>>(S.fxy, S.fsc, S.nfast, fcap, w, h, 2 * nfeatures, S.sel, S.nsel);
    VSB_LAUNCHED(ctx);
    orb_harris_kernel<<<dim3(min(vsb_div_up(w * h / 4 + 16, 128), max(8, vsb_div_up(4 * nfeatures, 128))), zc), 128, 0, st>>>(in, img_stride, pitch, S.sel, S.nsel, fcap, S.resp);
    VSB_LAUNCHED(ctx);
    orb_select_harris_kernel<<<zc, 256, 0, st>>>(S.sel, S.resp, S.nsel, fcap, nfeatures, cap, kp_xy, kp_resp, n_kp);
    VSB_LAUNCHED(ctx);
    orb_angle_kernel<<<dim3(min(vsb_div_up(cap, 8), max(4, vsb_div_up(nfeatures + 64, 8 * ORB_KP_PER_WARP))), zc), 256, 0, st>>>(in, img_stride, pitch, kp_xy, n_kp, cap, kp_angle);
    VSB_LAUNCHED(ctx);
    if (desc) {
        const dim3 bgrid(vsb_div_up(w, BTW), vsb_div_up(h, BTH), zc);
        if ((((uintptr_t)in | (uintptr_t)img_stride | (uintptr_t)pitch) & 15u) == 0 && !(ctx->orb_impl & 4))
            orb_blur_kernel<true><<<bgrid, 256, 0, st>>>(in, img_stride, pitch, w, h, S.blurred, bpitch);
        else
            orb_blur_kernel<false><<<bgrid, 256, 0, st>>>(in, img_stride, pitch, w, h, S.blurred, bpitch);
        VSB_LAUNCHED(ctx);
        // (the selected-corner list is consumed by now: its space holds the key points' cos / sin, one pair per corner slot)
        float2* cs = (ctx->orb_impl & 2) ? nullptr : reinterpret_cast<float2*>(S.sel);
        if (cs) {
            orb_trig_kernel<<<dim3(vsb_div_up(min(cap, fcap), 256), zc), 256, 0, st>>>(kp_angle, n_kp, cap, fcap, cs);
            VSB_LAUNCHED(ctx);
        }
        orb_describe_kernel<<<dim3(min(vsb_div_up(cap, 8), max(4, vsb_div_up(nfeatures + 64, 8 * ORB_KP_PER_WARP))), zc), 256, 0, st>>>(
            S.blurred, w, h, bpitch, kp_xy, kp_angle, cs, fcap, n_kp, cap, desc);
        VSB_LAUNCHED(ctx);
    }
    return VSB_OK;
}

}  // namespace

// ORB key points + descriptors of `count` frames, one pyramid level (cv::ORB::create(nfeatures, 1.2f, 1, 31, 0, 2,
// HARRIS_SCORE, 31, fast_threshold)->detectAndCompute).  See include/vislam_b200.h.
// Scratch budget of one chunk of frames.  Every stage is one launch per chunk and pyramid level, and the upper levels of
// the scale pyramid are small, so few large chunks keep the machine filled where many small ones are launch-bound: 2000
// 752x480 frames need 7.6 GB in one chunk (an HBM3e part has 180 GB), which the default budget allows.
static size_t orb_scratch_budget(const vsb_ctx* ctx) { return (size_t)(ctx->orb_scratch_mb > 0 ? ctx->orb_scratch_mb : 32768) << 20; }

extern "C" int vsb_orb_detect_compute(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h, int count,
                                      int nfeatures, int fast_threshold, int cap, int32_t* kp_xy, float* kp_resp,
                                      float* kp_angle, uint8_t* desc, int32_t* n_kp, void* stream) {
    if (!ctx || !img || !kp_xy || !kp_resp || !kp_angle || !n_kp) return VSB_ERR_INVALID;
    if (w <= 2 * ORB_EDGE || h <= 2 * ORB_EDGE || pitch < w || count < 0 || cap <= 0 || nfeatures < 0) return VSB_ERR_INVALID;
    if (fast_threshold < 0 || fast_threshold > 255) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // FAST corners per frame: 3x3 non-maximum suppression leaves at most one corner per 2x2 block, so w*h/4 can never
    // overflow (dense noise does reach a quarter of that); the batch is processed in chunks that keep the scratch bounded
    const size_t per_frame = orb_scratch_bytes(1, w, h, desc != nullptr);
    const int chunk = (int)max((size_t)1, min((size_t)count, orb_scratch_budget(ctx) / per_frame));
    void* scratch = nullptr;
    int rc = vsb_scratch2_reserve(ctx, orb_scratch_bytes(chunk, w, h, desc != nullptr) + 256, &scratch);
    if (rc) return rc;
    uint8_t* p = static_cast<uint8_t*>(scratch);
    const OrbScratch S = orb_scratch_carve(p, chunk, w, h, desc != nullptr);
    void* fast_scratch = nullptr;
    if ((rc = vsb_scratch_reserve(ctx, vsb_fast_scratch_bytes(w, h, chunk), &fast_scratch))) return rc;
    for (int z0 = 0; z0 < count; z0 += chunk) {
        const int zc = min(chunk, count - z0);
        rc = orb_one_level(ctx, S, fast_scratch, img + (size_t)z0 * img_stride, img_stride, pitch, w, h, zc, nfeatures, fast_threshold, cap,
                           kp_xy + (size_t)z0 * cap * 2, kp_resp + (size_t)z0 * cap, kp_angle + (size_t)z0 * cap,
                           desc ? desc + (size_t)z0 * cap * 32 : nullptr, n_kp + z0, st);
        if (rc) return rc;
    }
    return VSB_OK;
}

// cv::ORB with its scale pyramid (the reference's ORB::create(n): nlevels 8, scale factor 1.2).  See include/vislam_b200.h.
// Workspace bytes per frame of vsb_orb_detect_compute_pyr_ws (its own scratch, two level images, per-level key-point lists, the
// FAST detector's scratch).
static size_t orb_pyr_frame_bytes(int w, int h, int cap, bool describe) {
    const size_t tmp_frame = orb_al((size_t)cap * 8) + 2 * orb_al((size_t)cap * 4) + (describe ? orb_al((size_t)cap * 32) : 0);
    const size_t wpad = (size_t)((w + 15) & ~15);            // level rows are padded to 16 bytes: a level can be up to wpad x h bytes (scale factors close to 1)
    return orb_scratch_bytes(1, w, h, describe) + 2 * orb_al(wpad * h) + tmp_frame + 256 + vsb_fast_scratch_bytes(w, h, 1);
}
constexpr int ORB_LP_LEVELS = 8;
// the levels cv::ORB builds: (int)lrintf(size / scale_l), until the 31-pixel border filter leaves nothing
static int orb_level_dims(int w, int h, float scale_factor, int nlevels, int* lw, int* lh) {
    int nl = 0;
    for (int l = 0; l < nlevels && l < 32; l++, nl++) {
        const float scale = (float)pow((double)scale_factor, (double)l);
        lw[l] = l ? (int)lrintf((float)w / scale) : w;
        lh[l] = l ? (int)lrintf((float)h / scale) : h;
        if (l > 0 && (lw[l] <= 2 * ORB_EDGE || lh[l] <= 2 * ORB_EDGE)) break;
    }
    return nl;
}
// workspace of the levels-in-parallel form: one block of scratch per level, sized for that level
static size_t orb_pyr_lp_frame_bytes(int w, int h, int cap, bool describe, float scale_factor, int nlevels) {
    int lw[32], lh[32];
    const int nl = orb_level_dims(w, h, scale_factor, nlevels, lw, lh);
    size_t sum = 0;
    for (int l = 0; l < nl; l++) sum += orb_pyr_frame_bytes(l ? (lw[l] + 15) & ~15 : w, lh[l], cap, describe);
    return sum;
}
// Workspace for `frames` frames at once: the levels-in-parallel form when it fits `budget` bytes (small and medium batches are
// bound by the chain of dependent launches, see below), else one block that the levels share.
size_t vsb_orb_pyr_ws_bytes(int w, int h, int frames, int cap, int describe, float scale_factor, int nlevels, size_t budget) {
    const size_t fixed = orb_al((size_t)(w + h) * sizeof(int4)) + 4096;
    const size_t lp = fixed + (size_t)frames * orb_pyr_lp_frame_bytes(w, h, cap, describe != 0, scale_factor, nlevels);
    if (nlevels > 1 && nlevels <= ORB_LP_LEVELS && lp <= budget) return lp;
    return fixed + (size_t)frames * orb_pyr_frame_bytes(w, h, cap, describe != 0);
}

namespace {
struct OrbBlock {                     // everything one run of the one-level pipeline needs besides the input image
    OrbScratch S;
    uint8_t* img[2];                  // level images (ping-pong in the one-block form)
    int32_t* lxy; float* lresp; float* langle; uint8_t* ldesc; int32_t* ln;
    void* fast_scratch;
};
OrbBlock orb_block_carve(uint8_t*& p, int chunk, int w, int h, int cap, bool describe) {
    OrbBlock b;
    const size_t wpad = (size_t)((w + 15) & ~15);
    const size_t b_img = orb_al((size_t)chunk * wpad * h);
    b.S = orb_scratch_carve(p, chunk, w, h, describe);
    b.img[0] = p; p += b_img;
    b.img[1] = p; p += b_img;
    b.lxy = reinterpret_cast<int32_t*>(p); p += orb_al((size_t)chunk * cap * 8);
    b.lresp = reinterpret_cast<float*>(p); p += orb_al((size_t)chunk * cap * 4);
    b.langle = reinterpret_cast<float*>(p); p += orb_al((size_t)chunk * cap * 4);
    b.ldesc = nullptr;
    if (describe) { b.ldesc = p; p += orb_al((size_t)chunk * cap * 32); }
    b.ln = reinterpret_cast<int32_t*>(p); p += orb_al((size_t)chunk * 4);
    b.fast_scratch = p; p += (vsb_fast_scratch_bytes(w, h, chunk) + 255) & ~(size_t)255;
    return b;
}
// one level image from the one below it (cv::resize, INTER_LINEAR_EXACT), rows padded to 16 bytes
int orb_resize_level(vsb_ctx_t* ctx, const uint8_t* cur, int64_t cur_stride, int cp, int cw, int ch, uint8_t* dst, int nw, int nh,
                      int np, int zc, int4* rz_tab, cudaStream_t st) {
    const bool words = ((reinterpret_cast<uintptr_t>(cur) | (uintptr_t)cur_stride | (uintptr_t)cp) & 3u) == 0 &&
                       cw >= 2 && ch >= 2 && 2 * nw >= cw && !(ctx->orb_impl & 1);
    if (words) {
        orb_resize_coeff_folded_kernel<<<vsb_div_up(nw + nh, 256), 256, 0, st>>>(cw, ch, nw, nh, reinterpret_cast<int2*>(rz_tab));
        VSB_LAUNCHED(ctx);
        // tile form: a tile row spans ceil(127 s) + 2 taps, up to 15 bytes of alignment and the 12-byte window
        const int nvec = ((127 * cw + nw - 1) / nw + 2 + 15 + 12 + 15) / 16;
        const int nrows_cap = ((8 * RZ_ROWS - 1) * ch + nh - 1) / nh + 3;
        const size_t tile_bytes = (size_t)nvec * 16 * nrows_cap;
        const bool tiles = ((reinterpret_cast<uintptr_t>(cur) | (uintptr_t)cur_stride | (uintptr_t)cp) & 15u) == 0 &&
                           tile_bytes <= 48 * 1024 && !(ctx->orb_impl & 8);
        const dim3 rgrid(vsb_div_up(nw, 128), vsb_div_up(nh, 8 * RZ_ROWS), zc);
        if (tiles)
            orb_resize_tile_kernel<<<rgrid, 256, tile_bytes, st>>>(cur, cur_stride, cp, dst, nw, nh, np,
                                                                   reinterpret_cast<const int2*>(rz_tab), nvec, nrows_cap);
        else
            orb_resize4_kernel<<<rgrid, 256, 0, st>>>(cur, cur_stride, cp, dst, nw, nh, np, reinterpret_cast<const int2*>(rz_tab));
    } else {
        orb_resize_coeff_kernel<<<vsb_div_up(nw + nh, 256), 256, 0, st>>>(cw, ch, nw, nh, rz_tab);
        VSB_LAUNCHED(ctx);
        orb_resize_kernel<<<dim3(vsb_div_up(nw, 128), vsb_div_up(nh, 8), zc), 256, 0, st>>>(cur, cur_stride, cp, cw, ch, dst, nw, nh, np, rz_tab);
    }
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}
}  // namespace

// The detector on a caller-provided workspace `ws` of `ws_bytes` (256-byte aligned): the batch is processed in chunks of as
// many frames as the workspace holds.  The tracker gives each of its two slots its own workspace, so that the chunks of a
// sequence can run on two streams at once; the public entry below takes the workspace from the context.
// A batch whose per-level scratch fits the workspace (a few hundred frames under the default budget) is bound by the chain of ~14 dependent launches per level:
// there the scale pyramid is built first, then every level's pipeline runs on its own stream with its own block of scratch,
// and the levels' key points are appended in order at the end — the chain is 7 + 14 launches deep instead of 8 x 15.
int vsb_orb_detect_compute_pyr_ws(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h,
                                  int count, int nfeatures, float scale_factor, int nlevels, int fast_threshold, int cap,
                                  float* kp_xy, int32_t* kp_octave, float* kp_resp, float* kp_angle, uint8_t* desc,
                                  int32_t* n_kp, void* ws, size_t ws_bytes, void* stream) {
    if (!ctx || !img || !kp_xy || !kp_resp || !kp_angle || !n_kp) return VSB_ERR_INVALID;
    if (w <= 2 * ORB_EDGE || h <= 2 * ORB_EDGE || pitch < w || count < 0 || cap <= 0 || nfeatures < 0) return VSB_ERR_INVALID;
    if (fast_threshold < 0 || fast_threshold > 255 || nlevels < 1 || nlevels > 32 || !(scale_factor > 1.f)) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // per-level budget and scale exactly as cv::ORB computes them (orb.cpp); the factor is a float stored in a double
    int budget[32];
    {
        const float factor = (float)(1.0 / (double)scale_factor);
        float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; l++) { budget[l] = (int)lrintf(nd); sum += budget[l]; nd *= factor; }
        budget[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    }
    const bool describe = desc != nullptr;
    if (!ws) return VSB_ERR_INVALID;
    const size_t fixed = orb_al((size_t)(w + h) * sizeof(int4)) + 4096;
    const size_t per = orb_pyr_frame_bytes(w, h, cap, describe);
    if (ws_bytes < fixed + per) return VSB_ERR_CAPACITY;
    // levels in parallel: the whole batch at once, one block per level (sized for that level)
    int lvw[32], lvh[32];
    const int nlv = orb_level_dims(w, h, scale_factor, nlevels, lvw, lvh);
    const bool lp = ctx->orb_lp && nlevels > 1 && nlevels <= ORB_LP_LEVELS &&
                    ws_bytes >= fixed + (size_t)count * orb_pyr_lp_frame_bytes(w, h, cap, describe, scale_factor, nlevels);
    const int chunk = lp ? count : (int)min((size_t)count, (ws_bytes - fixed) / per);
    int rc;
    uint8_t* p = static_cast<uint8_t*>(ws);
    int4* rz_tab = reinterpret_cast<int4*>(p); p += orb_al((size_t)(w + h) * sizeof(int4));
    OrbBlock blk[ORB_LP_LEVELS];
    if (lp) for (int l = 0; l < nlv; l++) blk[l] = orb_block_carve(p, chunk, l ? (lvw[l] + 15) & ~15 : w, lvh[l], cap, describe);
    else blk[0] = orb_block_carve(p, chunk, w, h, cap, describe);
    if (lp && (rc = vsb_orb_lp_streams(ctx))) return rc;
    for (int z0 = 0; z0 < count; z0 += chunk) {
        const int zc = min(chunk, count - z0);
        VSB_CUDA(ctx, cudaMemsetAsync(n_kp + z0, 0, (size_t)zc * sizeof(int32_t), st));
        const uint8_t* cur = img + (size_t)z0 * img_stride;
        int64_t cur_stride = img_stride;
        int cw = w, ch = h, cp = pitch;
        int nl = 0;                                               // levels that exist (the border filter empties the small ones)
        float scales[32];
        for (int l = 0; l < nlevels; l++, nl++) {
            const float scale = (float)pow((double)scale_factor, (double)l);
            scales[l] = scale;
            OrbBlock& B = blk[lp ? l : 0];
            if (l > 0) {
                const int nw = (int)lrintf((float)w / scale), nh = (int)lrintf((float)h / scale);
                if (nw <= 2 * ORB_EDGE || nh <= 2 * ORB_EDGE) break;       // the border filter leaves nothing from here on
                uint8_t* dst = B.img[lp ? 0 : (l & 1)];
                ProfScope ps(ctx, VSB_K_ORB, st);
                // level images get rows padded to 16 bytes (the workspace is sized for padded rows), so that every reader takes its aligned
                // word path (FAST's loader, the blur, the resize's packed stores) whatever nw is
                const int np = (nw + 15) & ~15;
                if ((rc = orb_resize_level(ctx, cur, cur_stride, cp, cw, ch, dst, nw, nh, np, zc, rz_tab, st))) return rc;
                cur = dst; cur_stride = (int64_t)np * nh; cw = nw; ch = nh; cp = np;
            }
            cudaStream_t sl = st;
            if (lp) {                                             // this level's image is ready: its pipeline forks off
                sl = ctx->orb_stream[l];
                VSB_CUDA(ctx, cudaEventRecord(ctx->orb_ev_ready[l], st));
                VSB_CUDA(ctx, cudaStreamWaitEvent(sl, ctx->orb_ev_ready[l], 0));
            }
            rc = orb_one_level(ctx, B.S, B.fast_scratch, cur, cur_stride, cp, cw, ch, zc, budget[l], fast_threshold, cap, B.lxy, B.lresp,
                               B.langle, B.ldesc, B.ln, sl, l > 0 ? cp : 0);
            if (rc) return rc;
            if (lp) {
                VSB_CUDA(ctx, cudaEventRecord(ctx->orb_ev_done[l], sl));
                continue;                                         // appended below, in level order
            }
            ProfScope ps(ctx, VSB_K_ORB, st);
            orb_append_kernel<<<dim3(vsb_div_up(cap, 256), zc), 256, 0, st>>>(
                B.lxy, B.lresp, B.langle, B.ldesc, B.ln, cap, scale, l, kp_xy + (size_t)z0 * cap * 2, kp_octave ? kp_octave + (size_t)z0 * cap : nullptr,
                kp_resp + (size_t)z0 * cap, kp_angle + (size_t)z0 * cap, desc ? desc + (size_t)z0 * cap * 32 : nullptr, n_kp + z0);
            VSB_LAUNCHED(ctx);
            orb_advance_kernel<<<vsb_div_up(zc, 256), 256, 0, st>>>(n_kp + z0, B.ln, zc);
            VSB_LAUNCHED(ctx);
        }
        if (lp) {
            for (int l = 0; l < nl; l++) {
                OrbBlock& B = blk[l];
                VSB_CUDA(ctx, cudaStreamWaitEvent(st, ctx->orb_ev_done[l], 0));
                ProfScope ps(ctx, VSB_K_ORB, st);
                orb_append_kernel<<<dim3(vsb_div_up(cap, 256), zc), 256, 0, st>>>(
                    B.lxy, B.lresp, B.langle, B.ldesc, B.ln, cap, scales[l], l, kp_xy + (size_t)z0 * cap * 2, kp_octave ? kp_octave + (size_t)z0 * cap : nullptr,
                    kp_resp + (size_t)z0 * cap, kp_angle + (size_t)z0 * cap, desc ? desc + (size_t)z0 * cap * 32 : nullptr, n_kp + z0);
                VSB_LAUNCHED(ctx);
                orb_advance_kernel<<<vsb_div_up(zc, 256), 256, 0, st>>>(n_kp + z0, B.ln, zc);
                VSB_LAUNCHED(ctx);
            }
        }
    }
    return VSB_OK;
}

extern "C" int vsb_orb_detect_compute_pyr(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h,
                                          int count, int nfeatures, float scale_factor, int nlevels, int fast_threshold, int cap,
                                          float* kp_xy, int32_t* kp_octave, float* kp_resp, float* kp_angle, uint8_t* desc,
                                          int32_t* n_kp, void* stream) {
    if (!ctx || !img || !kp_xy || !kp_resp || !kp_angle || !n_kp) return VSB_ERR_INVALID;
    if (w <= 2 * ORB_EDGE || h <= 2 * ORB_EDGE || pitch < w || count < 0 || cap <= 0 || nfeatures < 0) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    // Every stage is one launch per chunk of frames and pyramid level, and the upper levels are small, so few large chunks keep
    // the machine filled where many small ones are launch-bound: the budget (orb_scratch_mb) allows 2000 752x480 frames at once.
    if (nlevels < 1 || nlevels > 32 || !(scale_factor > 1.f)) return VSB_ERR_INVALID;
    const size_t budget = orb_scratch_budget(ctx);
    size_t bytes = vsb_orb_pyr_ws_bytes(w, h, count, cap, desc != nullptr, scale_factor, nlevels, budget);
    if (bytes > budget) {                                         // as many frames as the budget holds (at least one)
        const size_t one = vsb_orb_pyr_ws_bytes(w, h, 1, cap, desc != nullptr, scale_factor, nlevels, 0);
        const size_t per_frame = orb_pyr_frame_bytes(w, h, cap, desc != nullptr);
        const size_t frames = (budget > one ? (budget - one) / per_frame : 0) + 1;
        bytes = vsb_orb_pyr_ws_bytes(w, h, (int)min(frames, (size_t)count), cap, desc != nullptr, scale_factor, nlevels, 0);
    }
    void* ws = nullptr;
    int rc = vsb_scratch2_reserve(ctx, bytes, &ws);
    if (rc) return rc;
    return vsb_orb_detect_compute_pyr_ws(ctx, img, img_stride, pitch, w, h, count, nfeatures, scale_factor, nlevels, fast_threshold,
                                         cap, kp_xy, kp_octave, kp_resp, kp_angle, desc, n_kp, ws, bytes, stream);
}

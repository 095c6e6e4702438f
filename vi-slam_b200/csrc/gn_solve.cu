// gn_solve.cu — per-pyramid-level Gauss-Newton photometric SE3 pose solve, one thread block per frame pair,
// every level and iteration inside one kernel.  Replaces VISystem::EstimatePoseFeatures (reference
// src/VISystem.cpp:1113-1448) with WarpFunctionSE3 (:1495-1558), IdentityWeights (:1561-1565) and the Sophus
// update pose <- pose * exp(delta) (:1421).
//
// Data flow.  Everything a candidate point contributes that does NOT depend on the pose — its source
// intensity I_prev(y1,x1) and the previous-frame Scharr gradient at that pixel (:1320-1325) — is gathered
// once per level by gn_prepare_kernel into an 8-byte attribute record, so an iteration reads 24 coalesced
// bytes per point (the reference-layout float4 candidate + the record) and does ONE scattered load, the
// current-frame intensity at the warped position.  Points are processed in batches of GT*U (U = 2) so a thread has U
// independent gathers in flight.  Per point the 8-vector V = (J0..J5, r*w, r) is staged through shared memory
// (transposed: one row of V per warp lane group) and the warp accumulates the 8x8 Gram matrix V V^T of 4 points per
// instruction on the FP64 tensor cores; A = J^T J, b = -J^T r and the error are entries of that matrix.  69 registers
// per thread: six 128-thread blocks per SM, so one block's serial 6x6 solve overlaps the other blocks' point loops.
//
// Parity design: every float operation of the reference's per-point arithmetic is issued with an explicitly
// rounded intrinsic in the source order (no FMA contraction); cv::gemm's "float in, double accumulate" is
// reproduced with exact float*float products summed in FP64.  The reduction order is fixed (batch order, DMMA k
// order, warp order), so results are run-to-run deterministic and independent of the grid.
// (An FP32-partials variant — per-thread FMA sums, FP64 only across threads — was built in round 1 and removed: it was
// slower and ended 1.4e-5 from the reference pose, outside the 1e-5 tolerance.)  The 6x6 system is solved by warp 0 with OpenCV's LU elimination order (cv::solve), one column of [A | b] per lane.
#include "common.cuh"
#include "se3.cuh"
#include "gn_common.cuh"
#include <stdlib.h>

namespace {

using namespace gn;


struct GnParams {
    const uint8_t* prev_pyr;
    const uint8_t* cur_pyr;
    const int16_t* prev_gx;
    const int16_t* prev_gy;
    long long pair_stride;
    vsb_pyr_layout_t lay;
    const float4* cand;
    const double2* xy;           // tracker form (optional): back-projected (X, Y) of unit-depth points, replaces cand (z = w = 1)
    uint2* patt;                 // [count][levels][cand_cap] {gx | gy << 16, I_prev}
    float* resid;                // [count][cand_cap] residual scratch of the Tukey pre-pass (weight_mode 1 only)
    int cand_cap;
    const int32_t* n_cand;
    vsb_intr_t K[VSB_MAX_LEVELS];
    const float* pose_in;
    vsb_gn_opts_t o;
    float* pose_out;
    vsb_gn_trace_t* trace;
    int32_t* n_trace;
    unsigned long long* stats;   // optional work counters: pairs, iterations, point visits, pose updates
};

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    if (p < 0) return -p;
    if (p >= len) return 2 * len - 2 - p;
    return p;
}

// ---- per-level, pose-independent point attributes ------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_prepare_kernel(const GnParams P) {
    const int prob = blockIdx.z;
    const int lvl = P.o.first_lvl - (int)blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int ncand = min(P.n_cand[(size_t)prob * P.lay.levels + lvl], P.cand_cap);
    if (i >= ncand) return;
    const int cols = P.lay.w[lvl], rows = P.lay.h[lvl];
    const size_t off = (size_t)prob * P.pair_stride + P.lay.offset[lvl];
    const uint8_t* __restrict__ image1 = P.prev_pyr + off;
    const size_t slot = ((size_t)prob * P.lay.levels + lvl) * P.cand_cap + i;
    const float4 c = __ldg(P.cand + slot);
    int sx = (int)c.x, sy = (int)c.y;                      // image1.at<uchar>(y1, x1): float -> int truncation (:1320)
    sx = min(max(sx, 0), cols - 1);                        // candidates from Camera.cpp:393 are always inside
    sy = min(max(sy, 0), rows - 1);
    const size_t src = (size_t)sy * cols + sx;
    const uint32_t i1 = __ldg(image1 + src);
    int gx, gy;
    if (P.o.grad_mode == 0) {                              // Frame::gradientX/Y (Camera.cpp:171-172), :1324-1325
        gx = __ldg(P.prev_gx + off + src);
        gy = __ldg(P.prev_gy + off + src);
    } else {                                               // the same Scharr x3 evaluated at the point
        const int xm = reflect101(sx - 1, cols), xp = reflect101(sx + 1, cols);
        const int ym = reflect101(sy - 1, rows), yp = reflect101(sy + 1, rows);
        const uint8_t* q0 = image1 + (size_t)ym * cols;
        const uint8_t* q1 = image1 + (size_t)sy * cols;
        const uint8_t* q2 = image1 + (size_t)yp * cols;
        const int a00 = __ldg(q0 + xm), a01 = __ldg(q0 + sx), a02 = __ldg(q0 + xp);
        const int a10 = __ldg(q1 + xm), a12 = __ldg(q1 + xp);
        const int a20 = __ldg(q2 + xm), a21 = __ldg(q2 + sx), a22 = __ldg(q2 + xp);
        gx = 3 * (3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20));
        gy = 3 * (3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02));
    }
    P.patt[slot] = make_uint2(((uint32_t)gx & 0xFFFFu) | ((uint32_t)gy << 16), i1);
}

// ---- Tukey weights (optional mode; upstream it is commented out at VISystem.cpp:1344) ---------------------------
// TukeyFunctionWeights (:1797-1826) needs the scale 1.4826 * MAD of ALL residuals of the iteration before any weight
// exists, and MedianMat (:1846-1870) takes both medians from a 256-bin histogram of the values saturate-cast to u8.
// So the iteration gets a pre-pass: residuals -> histogram -> median -> histogram of |r - median| -> MAD.  The residual
// arithmetic below is the same sequence of rounded operations as the main loop's (warp :1519-1553, round() :1321).
struct LvlConst {
    float fx, fy, cx, cy, invfx, invfy, bx, by, frows, fcols;
    int cols, rows, npix;
};

__device__ __forceinline__ bool point_residual(const float4 c, uint32_t i_prev, const double* md, const LvlConst& L,
                                               const uint8_t* __restrict__ image2, int sample_mode, float& res) {
    const float X = F_MUL(F_ADD(F_MUL(c.x, L.invfx), L.bx), c.z);      // folded conversion, see se3.cuh backproj_offset
    const float Y = F_MUL(F_ADD(F_MUL(c.y, L.invfy), L.by), c.z);
    const double dX = X, dY = Y, dZ = c.z, dW = c.w;
    double s0 = md[0] * dX; s0 += md[1] * dY; s0 += md[2] * dZ; s0 += md[3] * dW;
    double s1 = md[4] * dX; s1 += md[5] * dY; s1 += md[6] * dZ; s1 += md[7] * dW;
    double s2 = md[8] * dX; s2 += md[9] * dY; s2 += md[10] * dZ; s2 += md[11] * dW;
    const float r0 = (float)s0, r1 = (float)s1, r2 = (float)s2, r3 = c.w;
    const float x2 = F_MUL(F_ADD(F_DIV(F_MUL(r0, L.fx), r2), L.cx), r3);
    const float y2 = F_MUL(F_ADD(F_DIV(F_MUL(r1, L.fy), r2), L.cy), r3);
    if (!((y2 > 0.f && y2 < L.frows && x2 > 0.f && x2 < L.fcols) && (r2 != 0.f))) return false;
    const int ix = __float2int_rz(x2), iy = __float2int_rz(y2);
    float i2;
    if (sample_mode == 0) {
        const int rx = ix + ((F_SUB(x2, (float)ix) >= 0.5f) ? 1 : 0);
        const int ry = iy + ((F_SUB(y2, (float)iy) >= 0.5f) ? 1 : 0);
        const int l = ry * L.cols + rx;
        if (l >= L.npix) return false;
        i2 = (float)__ldg(image2 + l);
    } else {
        if (ix + 1 >= L.cols || iy + 1 >= L.rows) return false;
        const uint8_t* p0 = image2 + iy * L.cols + ix;
        const float i00 = (float)__ldg(p0), i01 = (float)__ldg(p0 + 1);
        const float i10 = (float)__ldg(p0 + L.cols), i11 = (float)__ldg(p0 + L.cols + 1);
        const float ax = F_SUB(x2, floorf(x2)), ay = F_SUB(y2, floorf(y2));
        const float top = F_ADD(i00, F_MUL(ax, F_SUB(i01, i00)));
        const float bot = F_ADD(i10, F_MUL(ax, F_SUB(i11, i10)));
        i2 = F_ADD(top, F_MUL(ay, F_SUB(bot, top)));
    }
    res = F_SUB(i2, (float)i_prev);
    return true;
}

__device__ __forceinline__ int sat_u8(float v) {              // Mat::convertTo(CV_8U): saturate_cast<uchar>(cvRound(v))
    return min(max(__float2int_rn(v), 0), 255);
}

// MedianMat's scan (:1856-1866): first bin whose running count exceeds (float)(n / 2); -1 when there is none (n == 0).
__device__ float hist_median(const int* hist) {
    int n = 0;
    for (int i = 0; i < 256; i++) n += hist[i];
    const float m = (float)(n / 2);
    int bin = 0;
    for (int i = 0; i < 256; i++) {
        bin += hist[i];
        if ((float)bin > m) return (float)i;
    }
    return -1.0f;
}

// Returns 1 / (1.4826 * MAD) as TukeyFunctionWeights forms it (MAD == 0 -> 1).  Called by every thread of the block.
__device__ __noinline__ float tukey_inv_mad(const float4* __restrict__ cand, const uint2* __restrict__ patt,
                                            float* __restrict__ resid, int ncand, const double* s_md, LvlConst L,
                                            const uint8_t* __restrict__ image2, int sample_mode, int* s_hist,
                                            float* s_val) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 256; i += nt) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < ncand; i += nt) {
        float res = 0.f;
        const bool ok = point_residual(__ldg(cand + i), __ldg(patt + i).y, s_md, L, image2, sample_mode, res);
        resid[i] = ok ? res : __int_as_float(0x7fc00000);
        if (ok) atomicAdd(&s_hist[sat_u8(res)], 1);
    }
    __syncthreads();
    if (tid == 0) s_val[0] = hist_median(s_hist);
    __syncthreads();
    const float median = s_val[0];
    for (int i = tid; i < 256; i += nt) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < ncand; i += nt) {
        const float r = resid[i];
        if (r == r) atomicAdd(&s_hist[sat_u8(fabsf(F_SUB(r, median)))], 1);          // abs(_input - median), :1836
    }
    __syncthreads();
    if (tid == 0) {
        float mad = F_MUL(1.4826f, hist_median(s_hist));                               // c * MAD, :1841
        if (mad == 0.f) mad = 1.f;                                                     // :1807-1810
        s_val[1] = (float)(1.0 / (double)mad);                                         // :1811
    }
    __syncthreads();
    return s_val[1];
}

constexpr int VROW = 36;   // staging row stride (floats): conflict-free for the [g][4s+t] reads of the MMA feed

// U = points per thread per batch; TPS = resident threads per SM the register budget is sized for
// WM = weight mode (0 identity: the reference default, 1 Tukey, 2 Huber) at compile time: identity drops the seven
// multiplications by w = 1 per point
// UNITZW: the points come as pre-back-projected doubles (X, Y) with z = w = 1 (the tracker's fused candidate pass): the
// per-iteration back-projection, four conversions and the two multiplications by w vanish; same bits.
template <int GT, int WM, int U = 2, int TPS = 768, bool UNITZW = false>
__global__ void __launch_bounds__(GT, (TPS / GT) > 0 ? (TPS / GT) : 1)
gn_solve_kernel(const GnParams P) {
    constexpr int NW = GT / 32;
    const int prob = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, t4 = lane & 3;
    // the serial part of an iteration rotates over the warps with the block index: warp w runs on SM sub-partition w % 4,
    // and with warp 0 of EVERY block in that role sub-partition 0 would carry all of it while the other three wait
    const int swarp = (int)(blockIdx.x % NW);

    __shared__ float s_pose[7];
    __shared__ double s_md[12];
    constexpr int SV_FLOATS = U * 8 * VROW;     // one staging slot per point of a batch
    __shared__ __align__(16) float s_v[NW][SV_FLOATS];
    static_assert(SV_FLOATS * sizeof(float) >= 64 * sizeof(double), "a warp's staging area doubles as its 8x8 partial sum");
    __shared__ int s_cnt[NW];
    __shared__ double s_G[64];
    __shared__ int s_nv;
    __shared__ int s_stop;      // 1 = leave the level
    __shared__ int s_ntrace;
    __shared__ float s_last_err;
    __shared__ unsigned long long s_pts;
    __shared__ int s_upd;
    constexpr bool TUKEY = WM == 1;
    __shared__ int s_hist[TUKEY ? 256 : 1];
    __shared__ float s_tk[2];

    const vsb_gn_opts_t& o = P.o;
    if (tid < 7) s_pose[tid] = P.pose_in[(size_t)prob * 7 + tid];
    if (tid == 0) { s_ntrace = 0; s_pts = 0ull; s_upd = 0; }
    __syncthreads();

    const uint8_t* cur_base = P.cur_pyr + (size_t)prob * P.pair_stride;
    vsb_gn_trace_t* trace = P.trace ? P.trace + (size_t)prob * VSB_MAX_TRACE : nullptr;
    float* sv = s_v[warp];
    double* red = reinterpret_cast<double*>(sv);     // the warp's partial Gram matrix, written after its last staging read

    for (int lvl = o.first_lvl; lvl >= o.last_lvl; lvl--) {                       // VISystem.cpp:1181
        const int cols = P.lay.w[lvl], rows = P.lay.h[lvl];
        const uint8_t* __restrict__ image2 = cur_base + P.lay.offset[lvl];
        const size_t slot0 = ((size_t)prob * P.lay.levels + lvl) * P.cand_cap;
        const float4* __restrict__ cand = UNITZW ? nullptr : P.cand + slot0;
        const double2* __restrict__ xyp = UNITZW ? P.xy + slot0 : nullptr;
        const uint2* __restrict__ patt = P.patt + slot0;
        const int ncand = min(P.n_cand[(size_t)prob * P.lay.levels + lvl], P.cand_cap);
        const float fx = P.K[lvl].fx, fy = P.K[lvl].fy, cx = P.K[lvl].cx, cy = P.K[lvl].cy;
        const float invfx = P.K[lvl].invfx, invfy = P.K[lvl].invfy;
        const float bpx = vsb::backproj_offset(cx, invfx), bpy = vsb::backproj_offset(cy, invfy);
        const float zf = o.z_factor;
        const float frows = (float)rows, fcols = (float)cols;
        const int npix = rows * cols;
        if (tid == 0) s_last_err = 50000.0f;                                      // VISystem.cpp:1185

        for (int k = 0; k < o.max_iterations; k++) {                              // VISystem.cpp:1214
            if (tid == 0) {
                float m34[12];
                vsb::se3_matrix34(s_pose, m34);
                for (int i = 0; i < 12; i++) s_md[i] = (double)m34[i];
            }
            __syncthreads();
            float tk_inv_mad = 1.f;
            if (TUKEY) {
                LvlConst L;
                L.fx = fx; L.fy = fy; L.cx = cx; L.cy = cy; L.invfx = invfx; L.invfy = invfy; L.bx = bpx; L.by = bpy;
                L.frows = frows; L.fcols = fcols; L.cols = cols; L.rows = rows; L.npix = npix;
                tk_inv_mad = tukey_inv_mad(cand, patt, P.resid + (size_t)prob * P.cand_cap, ncand, s_md, L, image2,
                                           o.sample_mode, s_hist, s_tk);
            }
            double md[12];
#pragma unroll
            for (int i = 0; i < 12; i++) md[i] = s_md[i];
            int nv = 0;
            double acc0 = 0.0, acc1 = 0.0;      // this lane's two entries of the warp's 8x8 Gram matrix

            for (int base = 0; base < ncand; base += GT * U) {                    // VISystem.cpp:1281-1338
                // ---- phase 1: coalesced loads of U points per thread --------------------------------------
                float4 c[U];
                double2 pxy[U];
                bool live[U];
                uint2 at[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = base + u * GT + tid;
                    live[u] = i < ncand;
                    c[u] = make_float4(0.f, 0.f, 0.f, 0.f);                                        // z = 0 => invalid
                    pxy[u] = make_double2(0.0, 0.0);
                    at[u] = make_uint2(0u, 0u);
                    if (live[u]) {
                        if (UNITZW) pxy[u] = __ldg(xyp + i); else c[u] = __ldg(cand + i);
                        at[u] = __ldg(patt + i);
                    }
                }
                // ---- phase 2: warp (WarpFunctionSE3, :1519-1553), validity, address of the one gather ------
                float x2[U], y2[U], iz[U];
                int lin[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    double s0, s1, s2;
                    float r3;
                    if (UNITZW) {
                        // z = w = 1: m * 1.0 is m exactly, so the two last terms are plain additions; (...) * w is the identity
                        const double dX = pxy[u].x, dY = pxy[u].y;
                        s0 = md[0] * dX; s0 += md[1] * dY; s0 += md[2]; s0 += md[3];
                        s1 = md[4] * dX; s1 += md[5] * dY; s1 += md[6]; s1 += md[7];
                        s2 = md[8] * dX; s2 += md[9] * dY; s2 += md[10]; s2 += md[11];
                        r3 = 1.f;
                    } else {
                        const float X = F_MUL(F_ADD(F_MUL(c[u].x, invfx), bpx), c[u].z);
                        const float Y = F_MUL(F_ADD(F_MUL(c[u].y, invfy), bpy), c[u].z);
                        const double dX = X, dY = Y, dZ = c[u].z, dW = c[u].w;
                        s0 = md[0] * dX; s0 += md[1] * dY; s0 += md[2] * dZ; s0 += md[3] * dW;
                        s1 = md[4] * dX; s1 += md[5] * dY; s1 += md[6] * dZ; s1 += md[7] * dW;
                        s2 = md[8] * dX; s2 += md[9] * dY; s2 += md[10] * dZ; s2 += md[11] * dW;
                        r3 = c[u].w;                                                               // last row of T is (0,0,0,1)
                    }
                    const float r0 = (float)s0, r1 = (float)s1, r2 = (float)s2;
                    if (UNITZW) {
                        x2[u] = F_ADD(F_DIV(F_MUL(r0, fx), r2), cx);
                        y2[u] = F_ADD(F_DIV(F_MUL(r1, fy), r2), cy);
                    } else {
                        x2[u] = F_MUL(F_ADD(F_DIV(F_MUL(r0, fx), r2), cx), r3);
                        y2[u] = F_MUL(F_ADD(F_DIV(F_MUL(r1, fy), r2), cy), r3);
                    }
                    float izz = F_DIV(1.f, r2);
                    bool v = (y2[u] > 0.f && y2[u] < frows && x2[u] > 0.f && x2[u] < fcols) && (r2 != 0.f);  // :1299-1300
                    if (UNITZW) v = v && live[u];
                    if (izz < 0.f) izz = 0.f;                                                            // :1301
                    iz[u] = izz;
                    int l = 0;
                    if (v) {
                        // x2, y2 are positive here: truncation == floor, and the fraction x2 - floor(x2) is exact
                        const int ix = __float2int_rz(x2[u]), iy = __float2int_rz(y2[u]);
                        if (o.sample_mode == 0) {                                    // round(), :1321
                            const int rx = ix + ((F_SUB(x2[u], u23_to_float((uint32_t)ix)) >= 0.5f) ? 1 : 0);
                            const int ry = iy + ((F_SUB(y2[u], u23_to_float((uint32_t)iy)) >= 0.5f) ? 1 : 0);
                            l = ry * cols + rx;
                            if (l >= npix) v = false;
                        } else {                                                     // bilinear extension
                            if (ix + 1 >= cols || iy + 1 >= rows) v = false;
                            l = iy * cols + ix;
                        }
                    }
                    ok[u] = v;
                    lin[u] = v ? l : 0;
                }
                float i2[U];
                if (o.sample_mode == 0) {
#pragma unroll
                    for (int u = 0; u < U; u++) i2[u] = u23_to_float((uint32_t)__ldg(image2 + lin[u]));
                } else {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        float val = 0.f;
                        if (ok[u]) {
                            const uint8_t* p0 = image2 + lin[u];
                            const float i00 = (float)__ldg(p0), i01 = (float)__ldg(p0 + 1);
                            const float i10 = (float)__ldg(p0 + cols), i11 = (float)__ldg(p0 + cols + 1);
                            const float ax = F_SUB(x2[u], floorf(x2[u])), ay = F_SUB(y2[u], floorf(y2[u]));
                            const float top = F_ADD(i00, F_MUL(ax, F_SUB(i01, i00)));
                            const float bot = F_ADD(i10, F_MUL(ax, F_SUB(i11, i10)));
                            val = F_ADD(top, F_MUL(ay, F_SUB(bot, top)));
                        }
                        i2[u] = val;
                    }
                }
                // ---- phase 3+4: Jacobian row, residual, weight (:1304-1327, :1343, :1404-1405), then the Gram
                //      matrix of V = (J0..J5, r*w, r) of 32 points at a time on the FP64 tensor cores -------------
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const float X2 = x2[u], Y2 = y2[u], Z = iz[u];
                    const float fxx = F_MUL(fx, X2), fyy = F_MUL(fy, Y2);
                    const float iz2x = F_MUL(F_MUL(fxx, Z), Z);                       // fx*x2*iz*iz
                    const float iz2y = F_MUL(F_MUL(fyy, Z), Z);                       // fy*y2*iz*iz
                    const float jw00 = F_MUL(fx, Z);
                    const float jw02 = F_MUL(-iz2x, zf);
                    const float jw03 = -F_MUL(F_MUL(F_MUL(fxx, Y2), Z), Z);
                    const float jw04 = F_MUL(fx, F_ADD(1.f, F_MUL(F_MUL(F_MUL(X2, X2), Z), Z)));
                    const float jw05 = F_MUL(F_MUL(-fx, Y2), Z);
                    const float jw11 = F_MUL(fy, Z);
                    const float jw12 = F_MUL(-iz2y, zf);
                    const float jw13 = -F_MUL(fy, F_ADD(1.f, F_MUL(F_MUL(F_MUL(Y2, Y2), Z), Z)));
                    const float jw14 = F_MUL(F_MUL(F_MUL(F_MUL(fy, X2), Y2), Z), Z);
                    const float jw15 = F_MUL(F_MUL(-fy, X2), Z);
                    const int gxi = (int)(short)(at[u].x & 0xFFFFu);                  // gradientX1.at<short>(y1,x1), :1324
                    const int gyi = (int)(short)(at[u].x >> 16);                      // gradientY1, :1325
                    const float res = F_SUB(i2[u], u23_to_float(at[u].y));            // :1323
                    float wgt = 1.f;
                    if (TUKEY) {                                                      // Tukey, :1812-1823
                        const float tk_b = 4.6851f;
                        const float inv_b2 = (float)(1.0 / (double)F_MUL(tk_b, tk_b));
                        const float x = F_MUL(res, tk_inv_mad);
                        if (fabsf(x) <= tk_b) {
                            const float tukey = (float)(1.0 - (double)F_MUL(F_MUL(x, x), inv_b2));
                            wgt = F_MUL(tukey, tukey);
                        } else {
                            wgt = 0.f;
                        }
                    } else if (WM == 2) {                                             // Huber extension
                        const float a = fabsf(res);
                        wgt = (a <= o.huber_k) ? 1.f : F_DIV(o.huber_k, a);
                    }
                    // J = Jl * Jw (1x2 * 2x6 cv::gemm: products exact in double, one rounding to float).
                    // Columns 0 and 1 have a single non-zero product, so the float product is already that rounding.
                    const double dgx = i32_to_double(gxi), dgy = i32_to_double(gyi);
                    float V[8];
                    V[0] = F_MUL((float)gxi, jw00);
                    V[1] = F_MUL((float)gyi, jw11);
                    { double s = dgx * (double)jw02; s += dgy * (double)jw12; V[2] = (float)s; }
                    { double s = dgx * (double)jw03; s += dgy * (double)jw13; V[3] = (float)s; }
                    { double s = dgx * (double)jw04; s += dgy * (double)jw14; V[4] = (float)s; }
                    { double s = dgx * (double)jw05; s += dgy * (double)jw15; V[5] = (float)s; }
                    const bool good = ok[u];
#pragma unroll
                    for (int q = 0; q < 6; q++) V[q] = good ? (WM == 0 ? V[q] : F_MUL(wgt, V[q])) : 0.f;  // row *= w, :1404-1405
                    V[6] = good ? (WM == 0 ? res : F_MUL(res, wgt)) : 0.f;
                    V[7] = good ? res : 0.f;
                    nv += good ? 1 : 0;
                    // stage only: no barrier between the points of a batch, so their (long, dependent) Jacobian
                    // chains can be interleaved by the scheduler; the Gram update of the whole batch follows below
#pragma unroll
                    for (int q = 0; q < 8; q++) sv[(u * 8 + q) * VROW + lane] = V[q];
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < U; u++) {
#pragma unroll
                    for (int s = 0; s < 8; s++) {
                        const double d = (double)sv[(u * 8 + g8) * VROW + 4 * s + t4];   // V[g] of point 4s+t of slot u
                        dmma_8x8x4(acc0, acc1, d, d);
                    }
                }
                __syncwarp();                 // the next batch overwrites the slots
            }
            // ---- cross-warp reduction in warp order (deterministic) --------------------------------------------
            red[g8 * 8 + 2 * t4] = acc0;
            red[g8 * 8 + 2 * t4 + 1] = acc1;
            {
                int cnum = nv;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) cnum += __shfl_down_sync(0xffffffffu, cnum, off);
                if (lane == 0) s_cnt[warp] = cnum;
            }
            __syncthreads();
            if (tid < 64) {
                double v = reinterpret_cast<const double*>(s_v[0])[tid];
#pragma unroll
                for (int wv = 1; wv < NW; wv++) v += reinterpret_cast<const double*>(s_v[wv])[tid];
                s_G[tid] = v;
            }
            if (tid == GT - 1) {
                int cnum = 0;
                for (int wv = 0; wv < NW; wv++) cnum += s_cnt[wv];
                s_nv = cnum;
            }
            __syncthreads();
            // ---- error test, normal equations, pose update (warp 0; VISystem.cpp:1343-1421) ----------------
            if (warp == swarp) {
                int stop = 0, updated = 0;
                float err = 0.f;
                float delta[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                const int n_valid = s_nv;
                if (n_valid == 0) {
                    stop = 1;
                } else {
                    const float inv_n = (float)(1.0 / (double)n_valid);               // :1347
                    err = (float)((double)inv_n * s_G[7 * 8 + 6]);                    // :1349-1350
                    const float last = s_last_err;
                    if (err >= last || k == o.max_iterations - 1 || fabsf(F_SUB(err, last)) < o.epsilon) {  // :1357
                        stop = 1;
                    } else {
                        warp_solve6(s_G, lane, delta);                                // :1408-1412
                        updated = 1;
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (updated) {
                        s_last_err = err;                                             // :1377
                        float e[7], np[7], cur[7];
                        for (int i = 0; i < 7; i++) cur[i] = s_pose[i];
                        vsb::se3_exp(delta, e);
                        vsb::se3_mul(cur, e, np);                                     // :1421
                        for (int i = 0; i < 7; i++) s_pose[i] = np[i];
                    }
                    if (trace && s_ntrace < VSB_MAX_TRACE) {
                        vsb_gn_trace_t* tr = trace + s_ntrace;
                        tr->lvl = lvl; tr->iter = k; tr->n_valid = n_valid; tr->updated = updated; tr->error = err;
                        for (int i = 0; i < 7; i++) tr->pose[i] = s_pose[i];
                        for (int i = 0; i < 6; i++) tr->delta[i] = delta[i];
                    }
                    s_ntrace++;
                    s_pts += (unsigned long long)ncand;
                    s_upd += updated;
                    s_stop = stop;
                }
            }
            __syncthreads();
            if (s_stop) break;
        }
        __syncthreads();
    }
    if (tid < 7) P.pose_out[(size_t)prob * 7 + tid] = s_pose[tid];                    // :1445
    if (tid == 0 && P.n_trace) P.n_trace[prob] = min(s_ntrace, VSB_MAX_TRACE);
    if (tid == 0 && P.stats) {
        atomicAdd(P.stats + 0, 1ull);
        atomicAdd(P.stats + 1, (unsigned long long)s_ntrace);
        atomicAdd(P.stats + 2, s_pts);
        atomicAdd(P.stats + 3, (unsigned long long)s_upd);
    }
}

}  // namespace

extern "C" void vsb_gn_default_opts(vsb_gn_opts_t* o) {
    if (!o) return;
    o->first_lvl = 3; o->last_lvl = 0; o->max_iterations = 10;   // VISystem.cpp:1117-1120
    o->epsilon = 0.001f; o->z_factor = 0.002f;                   // :1115, :1121
    o->weight_mode = 0; o->sample_mode = 0; o->huber_k = 10.0f;
    o->grad_mode = 0; o->accum_mode = 0;
}

// Internal entry (tracker): `patt` is caller-owned scratch of count * levels * cand_cap uint2 (NULL = context scratch).
int vsb_gn_solve_stats(vsb_ctx_t* ctx, const uint8_t* prev_pyr, const uint8_t* cur_pyr, const int16_t* prev_gx,
                       const int16_t* prev_gy, int64_t pair_stride_pixels, const vsb_pyr_layout_t* layout,
                       const float* cand, int cand_cap, const int32_t* n_cand, const vsb_intr_t K[VSB_MAX_LEVELS],
                       const float* pose_in, const vsb_gn_opts_t* opts, int count, float* pose_out,
                       vsb_gn_trace_t* trace, int32_t* n_trace, unsigned long long* stats, void* patt_scratch,
                       int patt_ready, const void* xy_ready, void* stream) {
    // xy_ready (tracker): [count][levels][cand_cap] double2 (X, Y) of unit-depth points, written together with the
    // attribute records by the fused candidate pass; `cand` may then be NULL
    if (!ctx || !prev_pyr || !cur_pyr || !layout || (!cand && !xy_ready) || !n_cand || !K || !pose_in || !opts || !pose_out)
        return VSB_ERR_INVALID;
    if (opts->accum_mode != 0) return VSB_ERR_UNSUPPORTED;     // the FP32-partials variant of round 1 is gone (see the header)
    if (xy_ready && !(patt_ready && patt_scratch && opts->weight_mode == 0)) return VSB_ERR_INVALID;
    if (count < 0 || cand_cap < 0) return VSB_ERR_INVALID;
    if (opts->first_lvl >= layout->levels || opts->last_lvl < 0 || opts->first_lvl < opts->last_lvl)
        return VSB_ERR_INVALID;
    if (opts->weight_mode < 0 || opts->weight_mode > 2) return VSB_ERR_UNSUPPORTED;     // 0 identity, 1 Tukey, 2 Huber
    if (opts->sample_mode != 0 && opts->sample_mode != 1) return VSB_ERR_UNSUPPORTED;
    if (opts->grad_mode == 0 && (!prev_gx || !prev_gy)) return VSB_ERR_INVALID;
    if (trace && (opts->first_lvl - opts->last_lvl + 1) * opts->max_iterations > VSB_MAX_TRACE) return VSB_ERR_CAPACITY;
    if (count == 0) return VSB_OK;
    // point attributes: the caller's buffer, or the context scratch (stand-alone entry; see the header on concurrent use).
    // Tukey residuals: workspace of the launching STREAM — the tracker's host entry runs two chunks on two streams.
    cudaStream_t st = (cudaStream_t)stream;
    if (!patt_scratch) {
        const size_t patt_bytes = (size_t)count * layout->levels * cand_cap * sizeof(uint2) + 256;
        int rc = vsb_scratch_reserve(ctx, patt_bytes, &patt_scratch);
        if (rc) return rc;
    }
    float* resid = nullptr;
    if (opts->weight_mode == 1) {
        void* ws = nullptr;
        int rc = vsb_stream_ws_reserve(ctx, st, (size_t)count * cand_cap * sizeof(float) + 256, &ws);
        if (rc) return rc;
        resid = static_cast<float*>(ws);
    }
    GnParams P;
    P.resid = resid;
    P.prev_pyr = prev_pyr; P.cur_pyr = cur_pyr; P.prev_gx = prev_gx; P.prev_gy = prev_gy;
    P.pair_stride = pair_stride_pixels;
    P.lay = *layout;
    P.cand = reinterpret_cast<const float4*>(cand);
    P.xy = reinterpret_cast<const double2*>(xy_ready);
    P.patt = reinterpret_cast<uint2*>(patt_scratch);
    P.cand_cap = cand_cap;
    P.n_cand = n_cand;
    for (int l = 0; l < VSB_MAX_LEVELS; l++) P.K[l] = K[l];
    P.pose_in = pose_in; P.o = *opts; P.pose_out = pose_out; P.trace = trace; P.n_trace = n_trace;
    P.stats = stats;
    const int nlev = opts->first_lvl - opts->last_lvl + 1;
    if (cand_cap > 0 && !(patt_ready && patt_scratch)) {     // patt_ready: the caller's fused candidate pass wrote the records
        for (int z0 = 0; z0 < count; z0 += 65535) {
            GnParams Q = P;
            const int zc = count - z0 < 65535 ? count - z0 : 65535;
            // shift every per-pair base by z0 pairs
            Q.prev_pyr += (size_t)z0 * P.pair_stride;
            if (Q.prev_gx) Q.prev_gx += (size_t)z0 * P.pair_stride;
            if (Q.prev_gy) Q.prev_gy += (size_t)z0 * P.pair_stride;
            Q.cand += (size_t)z0 * layout->levels * cand_cap;
            Q.patt += (size_t)z0 * layout->levels * cand_cap;
            Q.n_cand += (size_t)z0 * layout->levels;
            dim3 grid(vsb_div_up(cand_cap, 256), nlev, zc);
            ProfScope ps(ctx, VSB_K_GN_PREPARE, st);
            gn_prepare_kernel<<<grid, 256, 0, st>>>(Q);
            VSB_LAUNCHED(ctx);
        }
    }
    ProfScope ps(ctx, VSB_K_GN_SOLVE, st);
    // Threads per frame pair.  A pair is one block, so a small batch cannot fill the machine with 128-thread blocks: the
    // fewer pairs there are, the more threads each one gets (a lone 752x480 pair: 0.77 ms at 128 threads, see
    // tools/latency_pairs.py).  The partition of points over warps changes with it, i.e. the order of the FP64 sums; the
    // float results are the same unless a sum sits within 2^-29 of a rounding boundary.
    int gt_env = ctx->gn_threads;
    if (gt_env == 0) {
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        gt_env = count >= 4 * sms ? 128 : count >= 2 * sms ? 256 : count >= sms / 2 ? 512 : 1024;
    }
#define GN_LAUNCH(T, U, TPS)                                                                          \
    do {                                                                                              \
        if (opts->weight_mode == 1) gn_solve_kernel<T, 1, U, TPS><<<count, T, 0, st>>>(P);            \
        else if (opts->weight_mode == 2) gn_solve_kernel<T, 2, U, TPS><<<count, T, 0, st>>>(P);       \
        else if (xy_ready) gn_solve_kernel<T, 0, U, TPS, true><<<count, T, 0, st>>>(P);               \
        else gn_solve_kernel<T, 0, U, TPS><<<count, T, 0, st>>>(P);                                   \
    } while (0)
    if (gt_env >= 1024) GN_LAUNCH(1024, 1, 1024);
    else if (gt_env >= 512) GN_LAUNCH(512, 2, 1024);
    else if (gt_env >= 256) GN_LAUNCH(256, 2, 768);
    else if (gt_env >= 128) GN_LAUNCH(128, 2, 768);
    else GN_LAUNCH(64, 2, 768);
#undef GN_LAUNCH
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_gn_solve(vsb_ctx_t* ctx, const uint8_t* prev_pyr, const uint8_t* cur_pyr, const int16_t* prev_gx,
                            const int16_t* prev_gy, int64_t pair_stride_pixels, const vsb_pyr_layout_t* layout,
                            const float* cand, int cand_cap, const int32_t* n_cand, const vsb_intr_t K[VSB_MAX_LEVELS],
                            const float* pose_in, const vsb_gn_opts_t* opts, int count, float* pose_out,
                            vsb_gn_trace_t* trace, int32_t* n_trace, void* stream) {
    return vsb_gn_solve_stats(ctx, prev_pyr, cur_pyr, prev_gx, prev_gy, pair_stride_pixels, layout, cand, cand_cap,
                              n_cand, K, pose_in, opts, count, pose_out, trace, n_trace, nullptr, nullptr, 0, nullptr, stream);
}

// gn_solve.cu — per-pyramid-level Gauss-Newton photometric SE3 pose solve, one thread block per frame pair,
// every level and iteration inside one kernel.  Replaces VISystem::EstimatePoseFeatures (reference
// src/VISystem.cpp:1113-1448) with WarpFunctionSE3 (:1495-1558), IdentityWeights (:1561-1565) and the Sophus
// update pose <- pose * exp(delta) (:1421).
//
// Parity design: every float operation of the reference's per-point arithmetic is issued with an explicitly
// rounded intrinsic in the source order (no FMA contraction); cv::gemm's "float in, double accumulate"
// is reproduced with exact float*float products summed in FP64.  The 6x6 J^T J, 6-vector J^T r and the
// residual energy are reduced with a fixed tree (thread partials -> warp shuffle tree -> 8 warp sums added in
// warp order), so results are run-to-run deterministic and independent of the grid.  accum_mode 1 keeps
// FP32 per-thread partials (FMA) and only the cross-thread part in FP64, the north-star's wording.
#include "common.cuh"
#include "se3.cuh"

namespace {

constexpr int GT = 256;          // threads per frame pair
constexpr int NW = GT / 32;
constexpr int NRED = 28;         // 21 (upper triangle of J^T J) + 6 (J^T r) + 1 (sum w r^2)

struct GnParams {
    const uint8_t* prev_pyr;
    const uint8_t* cur_pyr;
    const int16_t* prev_gx;
    const int16_t* prev_gy;
    long long pair_stride;
    vsb_pyr_layout_t lay;
    const float4* cand;
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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

template <bool FP32_PARTIALS>
__global__ void __launch_bounds__(GT)
gn_solve_kernel(const GnParams P) {
    const int prob = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    __shared__ float s_pose[7];
    __shared__ float s_m[12];
    __shared__ double s_red[NW][NRED];
    __shared__ int s_cnt[NW];
    __shared__ double s_sum[NRED];
    __shared__ int s_nv;
    __shared__ int s_stop;      // 1 = leave the level
    __shared__ int s_ntrace;
    __shared__ float s_last_err;
    __shared__ unsigned long long s_pts;
    __shared__ int s_upd;

    const vsb_gn_opts_t& o = P.o;
    if (tid < 7) s_pose[tid] = P.pose_in[(size_t)prob * 7 + tid];
    if (tid == 0) { s_ntrace = 0; s_pts = 0ull; s_upd = 0; }
    __syncthreads();

    const uint8_t* prev_base = P.prev_pyr + (size_t)prob * P.pair_stride;
    const uint8_t* cur_base = P.cur_pyr + (size_t)prob * P.pair_stride;
    const int16_t* gx_base = P.prev_gx ? P.prev_gx + (size_t)prob * P.pair_stride : nullptr;
    const int16_t* gy_base = P.prev_gy ? P.prev_gy + (size_t)prob * P.pair_stride : nullptr;
    vsb_gn_trace_t* trace = P.trace ? P.trace + (size_t)prob * VSB_MAX_TRACE : nullptr;

    for (int lvl = o.first_lvl; lvl >= o.last_lvl; lvl--) {                       // VISystem.cpp:1181
        const int cols = P.lay.w[lvl], rows = P.lay.h[lvl];
        const uint8_t* __restrict__ image1 = prev_base + P.lay.offset[lvl];
        const uint8_t* __restrict__ image2 = cur_base + P.lay.offset[lvl];
        const int16_t* __restrict__ gx1 = gx_base ? gx_base + P.lay.offset[lvl] : nullptr;
        const int16_t* __restrict__ gy1 = gy_base ? gy_base + P.lay.offset[lvl] : nullptr;
        const float4* __restrict__ cand = P.cand + ((size_t)prob * P.lay.levels + lvl) * P.cand_cap;
        const int ncand = min(P.n_cand[(size_t)prob * P.lay.levels + lvl], P.cand_cap);
        const float fx = P.K[lvl].fx, fy = P.K[lvl].fy, cx = P.K[lvl].cx, cy = P.K[lvl].cy;
        const float invfx = P.K[lvl].invfx, invfy = P.K[lvl].invfy;
        const float zf = o.z_factor;
        const float frows = (float)rows, fcols = (float)cols;
        if (tid == 0) s_last_err = 50000.0f;                                      // VISystem.cpp:1185

        for (int k = 0; k < o.max_iterations; k++) {                              // VISystem.cpp:1214
            if (tid == 0) vsb::se3_matrix34(s_pose, s_m);
            __syncthreads();
            float m[12];
#pragma unroll
            for (int i = 0; i < 12; i++) m[i] = s_m[i];

            double acc[NRED];
            float accf[NRED];
#pragma unroll
            for (int i = 0; i < NRED; i++) { acc[i] = 0.0; accf[i] = 0.f; }
            int nv = 0;

            for (int i = tid; i < ncand; i += GT) {                               // VISystem.cpp:1281-1338
                const float4 c = __ldg(cand + i);
                // WarpFunctionSE3, VISystem.cpp:1519-1553
                const float X = F_MUL(F_MUL(F_SUB(c.x, cx), invfx), c.z);
                const float Y = F_MUL(F_MUL(F_SUB(c.y, cy), invfy), c.z);
                const double dX = X, dY = Y, dZ = c.z, dW = c.w;
                double s0 = (double)m[0] * dX; s0 += (double)m[1] * dY; s0 += (double)m[2] * dZ; s0 += (double)m[3] * dW;
                double s1 = (double)m[4] * dX; s1 += (double)m[5] * dY; s1 += (double)m[6] * dZ; s1 += (double)m[7] * dW;
                double s2 = (double)m[8] * dX; s2 += (double)m[9] * dY; s2 += (double)m[10] * dZ; s2 += (double)m[11] * dW;
                const float r0 = (float)s0, r1 = (float)s1, r2 = (float)s2, r3 = c.w;  // last row of T is (0,0,0,1)
                const float x2 = F_MUL(F_ADD(F_DIV(F_MUL(r0, fx), r2), cx), r3);
                const float y2 = F_MUL(F_ADD(F_DIV(F_MUL(r1, fy), r2), cy), r3);
                const float z2 = r2;
                float iz = F_DIV(1.f, z2);
                if (!(y2 > 0.f && y2 < frows && x2 > 0.f && x2 < fcols)) continue;   // :1299
                if (!(z2 != 0.f)) continue;                                           // :1300
                if (iz < 0.f) iz = 0.f;                                               // :1301
                float i2;
                if (o.sample_mode == 0) {                                             // round(), :1321
                    const float fxr = floorf(x2), fyr = floorf(y2);
                    const int rx = (int)fxr + ((F_SUB(x2, fxr) >= 0.5f) ? 1 : 0);
                    const int ry = (int)fyr + ((F_SUB(y2, fyr) >= 0.5f) ? 1 : 0);
                    const long long lin = (long long)ry * cols + rx;
                    if (lin >= (long long)rows * cols) continue;
                    i2 = (float)__ldg(image2 + lin);
                } else {                                                              // bilinear extension
                    const float x0f = floorf(x2), y0f = floorf(y2);
                    const int ix = (int)x0f, iy = (int)y0f;
                    if (ix + 1 >= cols || iy + 1 >= rows) continue;
                    const float ax = F_SUB(x2, x0f), ay = F_SUB(y2, y0f);
                    const uint8_t* p0 = image2 + (size_t)iy * cols + ix;
                    const float i00 = (float)__ldg(p0), i01 = (float)__ldg(p0 + 1);
                    const float i10 = (float)__ldg(p0 + cols), i11 = (float)__ldg(p0 + cols + 1);
                    const float top = F_ADD(i00, F_MUL(ax, F_SUB(i01, i00)));
                    const float bot = F_ADD(i10, F_MUL(ax, F_SUB(i11, i10)));
                    i2 = F_ADD(top, F_MUL(ay, F_SUB(bot, top)));
                }
                // Jw, VISystem.cpp:1304-1316 (pixel coordinates and z_factor kept as the reference has them)
                const float iz2x = F_MUL(F_MUL(F_MUL(fx, x2), iz), iz);               // fx*x2*iz*iz
                const float iz2y = F_MUL(F_MUL(F_MUL(fy, y2), iz), iz);               // fy*y2*iz*iz
                float Jw0[6], Jw1[6];
                Jw0[0] = F_MUL(fx, iz);
                Jw0[1] = 0.f;
                Jw0[2] = F_MUL(-iz2x, zf);
                Jw0[3] = -F_MUL(F_MUL(F_MUL(F_MUL(fx, x2), y2), iz), iz);
                Jw0[4] = F_MUL(fx, F_ADD(1.f, F_MUL(F_MUL(F_MUL(x2, x2), iz), iz)));
                Jw0[5] = F_MUL(F_MUL(-fx, y2), iz);
                Jw1[0] = 0.f;
                Jw1[1] = F_MUL(fy, iz);
                Jw1[2] = F_MUL(-iz2y, zf);
                Jw1[3] = -F_MUL(fy, F_ADD(1.f, F_MUL(F_MUL(F_MUL(y2, y2), iz), iz)));
                Jw1[4] = F_MUL(F_MUL(F_MUL(F_MUL(fy, x2), y2), iz), iz);
                Jw1[5] = F_MUL(F_MUL(-fy, x2), iz);
                // source pixel and image gradient of the PREVIOUS frame, :1320-1325
                const int sx = (int)c.x, sy = (int)c.y;
                const size_t src = (size_t)sy * cols + sx;
                const float i1 = (float)__ldg(image1 + src);
                float jl0, jl1;
                if (o.grad_mode == 0) {
                    jl0 = (float)__ldg(gx1 + src);
                    jl1 = (float)__ldg(gy1 + src);
                } else {                                                              // Scharr x3 on the fly
                    const int xm = reflect101(sx - 1, cols), xp = reflect101(sx + 1, cols);
                    const int ym = reflect101(sy - 1, rows), yp = reflect101(sy + 1, rows);
                    const uint8_t* q0 = image1 + (size_t)ym * cols;
                    const uint8_t* q1 = image1 + (size_t)sy * cols;
                    const uint8_t* q2 = image1 + (size_t)yp * cols;
                    const int a00 = __ldg(q0 + xm), a01 = __ldg(q0 + sx), a02 = __ldg(q0 + xp);
                    const int a10 = __ldg(q1 + xm), a12 = __ldg(q1 + xp);
                    const int a20 = __ldg(q2 + xm), a21 = __ldg(q2 + sx), a22 = __ldg(q2 + xp);
                    jl0 = (float)(3 * (3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20)));
                    jl1 = (float)(3 * (3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02)));
                }
                const float r = F_SUB(i2, i1);                                        // :1323
                float wgt = 1.f;
                if (o.weight_mode == 2) {                                             // Huber extension
                    const float a = fabsf(r);
                    wgt = (a <= o.huber_k) ? 1.f : F_DIV(o.huber_k, a);
                }
                // J = Jl * Jw (1x2 * 2x6 gemm: double accumulate, one rounding), then row *= w (:1404-1405)
                float J[6];
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    double s = (double)jl0 * (double)Jw0[q];
                    s += (double)jl1 * (double)Jw1[q];
                    J[q] = F_MUL(wgt, (float)s);
                }
                const float rw = F_MUL(r, wgt);
                nv++;
                if (!FP32_PARTIALS) {
                    int t = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int b = a; b < 6; b++) acc[t++] += (double)J[a] * (double)J[b];
                    }
#pragma unroll
                    for (int a = 0; a < 6; a++) acc[21 + a] += (double)J[a] * (double)rw;
                    acc[27] += (double)r * (double)rw;
                } else {
                    int t = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int b = a; b < 6; b++) { accf[t] = fmaf(J[a], J[b], accf[t]); t++; }
                    }
#pragma unroll
                    for (int a = 0; a < 6; a++) accf[21 + a] = fmaf(J[a], rw, accf[21 + a]);
                    accf[27] = fmaf(r, rw, accf[27]);
                }
            }
            // ---- deterministic reduction: warp tree, then the 8 warp sums in warp order -----------------
#pragma unroll
            for (int i = 0; i < NRED; i++) {
                double v = FP32_PARTIALS ? (double)accf[i] : acc[i];
                v = warp_sum(v);
                if (lane == 0) s_red[warp][i] = v;
            }
            {
                int c = nv;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
                if (lane == 0) s_cnt[warp] = c;
            }
            __syncthreads();
            if (tid < NRED) {
                double v = s_red[0][tid];
#pragma unroll
                for (int wv = 1; wv < NW; wv++) v += s_red[wv][tid];
                s_sum[tid] = v;
            }
            if (tid == 32) {
                int c = 0;
                for (int wv = 0; wv < NW; wv++) c += s_cnt[wv];
                s_nv = c;
            }
            __syncthreads();
            // ---- error test, normal equations, pose update (one thread; VISystem.cpp:1343-1421) ---------
            if (tid == 0) {
                vsb_gn_trace_t tr;
                tr.lvl = lvl; tr.iter = k; tr.n_valid = s_nv; tr.updated = 0; tr.error = 0.f;
                for (int i = 0; i < 6; i++) tr.delta[i] = 0.f;
                int stop = 0;
                if (s_nv == 0) {
                    stop = 1;
                } else {
                    const float inv_n = (float)(1.0 / (double)s_nv);                  // :1347
                    const float err = (float)((double)inv_n * s_sum[27]);             // :1349-1350
                    tr.error = err;
                    const float last = s_last_err;
                    if (err >= last || k == o.max_iterations - 1 || fabsf(F_SUB(err, last)) < o.epsilon) {  // :1357
                        stop = 1;
                    } else {
                        s_last_err = err;                                             // :1377
                        float A[36], b[6], Ainv[36], delta[6];
                        int t = 0;
                        for (int a = 0; a < 6; a++)
                            for (int c = a; c < 6; c++) {
                                const float v = (float)s_sum[t++];                    // A = J^T J, :1408
                                A[6 * a + c] = v;
                                A[6 * c + a] = v;
                            }
                        for (int a = 0; a < 6; a++) b[a] = (float)(-1.0 * s_sum[21 + a]);   // :1409
                        vsb::inv6(A, Ainv);                                           // :1412
                        for (int a = 0; a < 6; a++) {
                            double s = 0.0;
                            for (int c = 0; c < 6; c++) s += (double)Ainv[6 * a + c] * (double)b[c];
                            delta[a] = (float)s;
                        }
                        float e[7], np[7], cur[7];
                        for (int i = 0; i < 7; i++) cur[i] = s_pose[i];
                        vsb::se3_exp(delta, e);
                        vsb::se3_mul(cur, e, np);                                     // :1421
                        for (int i = 0; i < 7; i++) s_pose[i] = np[i];
                        for (int i = 0; i < 6; i++) tr.delta[i] = delta[i];
                        tr.updated = 1;
                    }
                }
                for (int i = 0; i < 7; i++) tr.pose[i] = s_pose[i];
                if (trace && s_ntrace < VSB_MAX_TRACE) trace[s_ntrace] = tr;
                s_ntrace++;
                s_pts += (unsigned long long)ncand;
                s_upd += tr.updated;
                s_stop = stop;
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

int vsb_gn_solve_stats(vsb_ctx_t* ctx, const uint8_t* prev_pyr, const uint8_t* cur_pyr, const int16_t* prev_gx,
                       const int16_t* prev_gy, int64_t pair_stride_pixels, const vsb_pyr_layout_t* layout,
                       const float* cand, int cand_cap, const int32_t* n_cand, const vsb_intr_t K[VSB_MAX_LEVELS],
                       const float* pose_in, const vsb_gn_opts_t* opts, int count, float* pose_out,
                       vsb_gn_trace_t* trace, int32_t* n_trace, unsigned long long* stats, void* stream) {
    if (!ctx || !prev_pyr || !cur_pyr || !layout || !cand || !n_cand || !K || !pose_in || !opts || !pose_out)
        return VSB_ERR_INVALID;
    if (count < 0 || cand_cap < 0) return VSB_ERR_INVALID;
    if (opts->first_lvl >= layout->levels || opts->last_lvl < 0 || opts->first_lvl < opts->last_lvl)
        return VSB_ERR_INVALID;
    if (opts->weight_mode != 0 && opts->weight_mode != 2) return VSB_ERR_UNSUPPORTED;   // Tukey (:1797-1826) is dead code upstream
    if (opts->sample_mode != 0 && opts->sample_mode != 1) return VSB_ERR_UNSUPPORTED;
    if (opts->grad_mode == 0 && (!prev_gx || !prev_gy)) return VSB_ERR_INVALID;
    if (trace && (opts->first_lvl - opts->last_lvl + 1) * opts->max_iterations > VSB_MAX_TRACE) return VSB_ERR_CAPACITY;
    if (count == 0) return VSB_OK;
    GnParams P;
    P.prev_pyr = prev_pyr; P.cur_pyr = cur_pyr; P.prev_gx = prev_gx; P.prev_gy = prev_gy;
    P.pair_stride = pair_stride_pixels;
    P.lay = *layout;
    P.cand = reinterpret_cast<const float4*>(cand);
    P.cand_cap = cand_cap;
    P.n_cand = n_cand;
    for (int l = 0; l < VSB_MAX_LEVELS; l++) P.K[l] = K[l];
    P.pose_in = pose_in; P.o = *opts; P.pose_out = pose_out; P.trace = trace; P.n_trace = n_trace;
    P.stats = stats;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(ctx, VSB_K_GN_SOLVE, st);
    if (opts->accum_mode == 1) gn_solve_kernel<true><<<count, GT, 0, st>>>(P);
    else gn_solve_kernel<false><<<count, GT, 0, st>>>(P);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

extern "C" int vsb_gn_solve(vsb_ctx_t* ctx, const uint8_t* prev_pyr, const uint8_t* cur_pyr, const int16_t* prev_gx,
                            const int16_t* prev_gy, int64_t pair_stride_pixels, const vsb_pyr_layout_t* layout,
                            const float* cand, int cand_cap, const int32_t* n_cand, const vsb_intr_t K[VSB_MAX_LEVELS],
                            const float* pose_in, const vsb_gn_opts_t* opts, int count, float* pose_out,
                            vsb_gn_trace_t* trace, int32_t* n_trace, void* stream) {
    return vsb_gn_solve_stats(ctx, prev_pyr, cur_pyr, prev_gx, prev_gy, pair_stride_pixels, layout, cand, cand_cap,
                              n_cand, K, pose_in, opts, count, pose_out, trace, n_trace, nullptr, stream);
}

// host_api.cu — C-ABI helpers for a host that does not link the CUDA runtime itself (the C++ class mirrors in
// vi-slam_b200/host/ are plain g++ code): device / pinned memory, copies, streams, and the small public
// operations of VISystem that the classes expose one by one — WarpFunctionSE3 (reference
// src/VISystem.cpp:1495-1558) as a kernel over an N x 4 point list, and the Sophus SE3f exp / matrix helpers
// (thirdparty/sophus/se3.hpp:723-742, 253-268).
#include "common.cuh"
#include "se3.cuh"

extern "C" int vsb_malloc(vsb_ctx_t* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return VSB_ERR_INVALID;
    *dptr = nullptr;
    VSB_CUDA(ctx, cudaSetDevice(ctx->device));
    VSB_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
    return VSB_OK;
}

extern "C" int vsb_free(vsb_ctx_t* ctx, void* dptr) {
    if (!ctx) return VSB_ERR_INVALID;
    if (dptr) VSB_CUDA(ctx, cudaFree(dptr));
    return VSB_OK;
}

extern "C" int vsb_host_alloc(vsb_ctx_t* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return VSB_ERR_INVALID;
    *hptr = nullptr;
    VSB_CUDA(ctx, cudaMallocHost(hptr, bytes ? bytes : 1));
    return VSB_OK;
}

extern "C" int vsb_host_free(vsb_ctx_t* ctx, void* hptr) {
    if (!ctx) return VSB_ERR_INVALID;
    if (hptr) VSB_CUDA(ctx, cudaFreeHost(hptr));
    return VSB_OK;
}

extern "C" int vsb_upload(vsb_ctx_t* ctx, void* dst, const void* h_src, size_t bytes, void* stream) {
    if (!ctx || (bytes && (!dst || !h_src))) return VSB_ERR_INVALID;
    if (bytes) VSB_CUDA(ctx, cudaMemcpyAsync(dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return VSB_OK;
}

extern "C" int vsb_upload_2d(vsb_ctx_t* ctx, void* dst, size_t dst_pitch, const void* h_src, size_t src_pitch,
                             size_t width_bytes, size_t rows, void* stream) {
    if (!ctx || ((width_bytes && rows) && (!dst || !h_src))) return VSB_ERR_INVALID;
    if (width_bytes && rows)
        VSB_CUDA(ctx, cudaMemcpy2DAsync(dst, dst_pitch, h_src, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice,
                                        (cudaStream_t)stream));
    return VSB_OK;
}

extern "C" int vsb_download(vsb_ctx_t* ctx, void* h_dst, const void* src, size_t bytes, void* stream) {
    if (!ctx || (bytes && (!h_dst || !src))) return VSB_ERR_INVALID;
    if (bytes) VSB_CUDA(ctx, cudaMemcpyAsync(h_dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return VSB_OK;
}

extern "C" int vsb_copy(vsb_ctx_t* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    if (!ctx || (bytes && (!dst || !src))) return VSB_ERR_INVALID;
    if (bytes) VSB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return VSB_OK;
}

extern "C" int vsb_memset(vsb_ctx_t* ctx, void* dst, int value, size_t bytes, void* stream) {
    if (!ctx || (bytes && !dst)) return VSB_ERR_INVALID;
    if (bytes) VSB_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream));
    return VSB_OK;
}

extern "C" int vsb_stream_create(vsb_ctx_t* ctx, void** stream) {
    if (!ctx || !stream) return VSB_ERR_INVALID;
    cudaStream_t st = nullptr;
    VSB_CUDA(ctx, cudaSetDevice(ctx->device));
    VSB_CUDA(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = (void*)st;
    return VSB_OK;
}

extern "C" int vsb_stream_destroy(vsb_ctx_t* ctx, void* stream) {
    if (!ctx) return VSB_ERR_INVALID;
    if (stream) VSB_CUDA(ctx, cudaStreamDestroy((cudaStream_t)stream));
    return VSB_OK;
}

extern "C" int vsb_stream_sync(vsb_ctx_t* ctx, void* stream) {
    if (!ctx) return VSB_ERR_INVALID;
    VSB_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return VSB_OK;
}

// ---- Sophus helpers (host side) --------------------------------------------------------------------------
extern "C" int vsb_se3_exp(const float delta[6], float pose[7]) {
    if (!delta || !pose) return VSB_ERR_INVALID;
    float tmp[7];
    vsb::se3_exp(delta, tmp);
    for (int i = 0; i < 7; i++) pose[i] = tmp[i];
    return VSB_OK;
}

extern "C" int vsb_se3_matrix(const float pose[7], float m[16]) {
    if (!pose || !m) return VSB_ERR_INVALID;
    float m34[12];
    vsb::se3_matrix34(pose, m34);
    for (int i = 0; i < 12; i++) m[i] = m34[i];
    m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
    return VSB_OK;
}

// ---- the pose update of a Gauss-Newton iteration as a batched device operation -----------------------------
// out[i] = pose[i] * exp(delta[i])  (VISystem.cpp:1421; se3.hpp:723-742 exp, :285-321 operator*): the same device functions
// the solver kernels call, exposed so that they can be checked by the million against the host helpers / the oracle.
namespace {
__global__ void __launch_bounds__(256) se3_update_kernel(const float* __restrict__ pose, const float* __restrict__ delta, int n,
                                                          float* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float a[7], d[6], e[7], r[7];
    for (int k = 0; k < 7; k++) a[k] = pose[(size_t)i * 7 + k];
    for (int k = 0; k < 6; k++) d[k] = delta[(size_t)i * 6 + k];
    vsb::se3_exp(d, e);
    vsb::se3_mul(a, e, r);
    for (int k = 0; k < 7; k++) out[(size_t)i * 7 + k] = r[k];
}
}  // namespace

extern "C" int vsb_se3_update_batch(vsb_ctx_t* ctx, const float* pose, const float* delta, int n, float* out, void* stream) {
    if (!ctx || n < 0 || (n && (!pose || !delta || !out))) return VSB_ERR_INVALID;
    if (n == 0) return VSB_OK;
    se3_update_kernel<<<vsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(pose, delta, n, out);
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

// ---- WarpFunctionSE3 as a stand-alone operation ------------------------------------------------------------
namespace {
struct WarpParams {
    double m[16];
    float fx, fy, cx, cy, invfx, invfy, bx, by;
};

__global__ void __launch_bounds__(256) warp_se3_kernel(const float4* __restrict__ pts, int n, WarpParams P,
                                                        float4* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const float X = F_MUL(F_ADD(F_MUL(p.x, P.invfx), P.bx), p.z);               // VISystem.cpp:1519-1524 (folded, se3.cuh)
    const float Y = F_MUL(F_ADD(F_MUL(p.y, P.invfy), P.by), p.z);
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {                                               // :1536 cv::gemm, double accumulate
        double s = __dmul_rn(P.m[4 * k], (double)X);
        s = __dadd_rn(s, __dmul_rn(P.m[4 * k + 1], (double)Y));
        s = __dadd_rn(s, __dmul_rn(P.m[4 * k + 2], (double)p.z));
        s = __dadd_rn(s, __dmul_rn(P.m[4 * k + 3], (double)p.w));
        r[k] = (float)s;
    }
    const float x2 = F_ADD(F_DIV(F_MUL(r[0], P.fx), r[2]), P.cx);               // :1540-1547
    const float y2 = F_ADD(F_DIV(F_MUL(r[1], P.fy), r[2]), P.cy);
    out[i] = make_float4(F_MUL(x2, r[3]), F_MUL(y2, r[3]), r[2], r[3]);         // :1552-1553
}
}  // namespace

extern "C" int vsb_warp_se3(vsb_ctx_t* ctx, const float* pts, int n, const float pose[7], const vsb_intr_t* K,
                            float* out, void* stream) {
    if (!ctx || !pose || !K || n < 0 || (n > 0 && (!pts || !out))) return VSB_ERR_INVALID;
    if (n == 0) return VSB_OK;
    WarpParams P;
    float m34[12];
    vsb::se3_matrix34(pose, m34);
    for (int i = 0; i < 12; i++) P.m[i] = (double)m34[i];
    P.m[12] = 0.0; P.m[13] = 0.0; P.m[14] = 0.0; P.m[15] = 1.0;
    P.fx = K->fx; P.fy = K->fy; P.cx = K->cx; P.cy = K->cy; P.invfx = K->invfx; P.invfy = K->invfy;
    P.bx = vsb::backproj_offset(K->cx, K->invfx); P.by = vsb::backproj_offset(K->cy, K->invfy);
    ProfScope ps(ctx, VSB_K_WARP, (cudaStream_t)stream);
    warp_se3_kernel<<<vsb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(pts), n, P,
                                                                           reinterpret_cast<float4*>(out));
    VSB_LAUNCHED(ctx);
    return VSB_OK;
}

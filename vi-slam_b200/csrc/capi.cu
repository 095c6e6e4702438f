// capi.cu — context management and the host-side helpers of the C ABI (no device work here).
// vsb_init_pyramid restates VISystem::InitializePyramid (reference src/VISystem.cpp:1451-1493);
// vsb_initial_pose restates VISystem.cpp:1135-1168 with Plus.cpp:56-83 (rotationMatrix2RPY) and
// Plus.cpp:182-220 (RPY2rotationMatrix); vsb_se3_mul is Sophus SE3f::operator* (se3.hpp:285-321).
#include "common.cuh"
#include "se3.cuh"
#include <stdlib.h>

#define VSB_VERSION 100

extern "C" int vsb_version(void) { return VSB_VERSION; }

extern "C" const char* vsb_error_string(int status) {
    switch (status) {
        case VSB_OK: return "ok";
        case VSB_ERR_INVALID: return "invalid argument";
        case VSB_ERR_CUDA: return "CUDA runtime error";
        case VSB_ERR_UNSUPPORTED: return "unsupported mode";
        case VSB_ERR_CAPACITY: return "capacity exceeded";
        default: return "unknown status";
    }
}

extern "C" int vsb_ctx_create(int device, vsb_ctx_t** out) {
    if (!out) return VSB_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return VSB_ERR_CUDA;  // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return VSB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VSB_ERR_CUDA;
    vsb_ctx* c = new vsb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    c->last_error[0] = 0;
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    c->scratch2 = nullptr;
    c->scratch2_bytes = 0;
    c->ws_n = 0;
    c->l2_fallback_counter = nullptr;
    c->attr_knn_tc_done = 0;
    c->attr_knn_mx_done = 0;
    c->attr_l2_tc_smem = 0;
    c->prof_on = 0;
    for (int i = 0; i < VSB_K_COUNT; i++) { c->prof_ms[i] = 0.0; c->prof_n[i] = 0; }
    c->knn_impl = 6;     // auto: 4-bit persistent kernel from 768 descriptors per set up, int8 kernel below
    c->gn_threads = 0;      // auto: by batch size
    c->pyr_impl = 1;
    c->gn_variant = 0;
    c->gn_impl = 1;
    c->gn_stage_bytes = 8192;
    c->gn_tail = 1;
    c->gn_dedup = 1;
    c->gn_cluster = 1;
    c->gn_cluster_threads = 0;
    c->orb_scratch_mb = 32768;
    c->orb_impl = 0;
    c->orb_lp = 1;
    c->orb_lp_ready = 0;
    c->fast_impl = 0;
    c->knn_l2_impl = 1;
    if (const char* e = getenv("VSB_KNN_L2_IMPL")) c->knn_l2_impl = atoi(e) ? 1 : 0;
    if (const char* e = getenv("VSB_GN_VARIANT")) c->gn_variant = atoi(e);
    if (const char* e = getenv("VSB_KNN_IMPL")) vsb_ctx_option(c, "knn_impl", atoi(e));
    if (const char* e = getenv("VSB_GN_IMPL")) c->gn_impl = atoi(e) ? 1 : 0;
    if (const char* e = getenv("VSB_GN_STAGE_BYTES")) vsb_ctx_option(c, "gn_stage_bytes", atoi(e));
    if (const char* e = getenv("VSB_GN_TAIL")) c->gn_tail = atoi(e) ? 1 : 0;
    if (const char* e = getenv("VSB_GN_DEDUP")) c->gn_dedup = atoi(e) ? 1 : 0;
    if (const char* e = getenv("VSB_GN_CLUSTER")) vsb_ctx_option(c, "gn_cluster", atoi(e));
    if (const char* e = getenv("VSB_GN_CLUSTER_THREADS")) vsb_ctx_option(c, "gn_cluster_threads", atoi(e));
    if (const char* e = getenv("VSB_ORB_SCRATCH_MB")) vsb_ctx_option(c, "orb_scratch_mb", atoi(e));
    if (const char* e = getenv("VSB_ORB_IMPL")) vsb_ctx_option(c, "orb_impl", atoi(e));
    if (const char* e = getenv("VSB_ORB_LP")) vsb_ctx_option(c, "orb_lp", atoi(e));
    if (const char* e = getenv("VSB_FAST_IMPL")) vsb_ctx_option(c, "fast_impl", atoi(e));
    if (const char* e = getenv("VSB_GN_THREADS")) vsb_ctx_option(c, "gn_threads", atoi(e));
    *out = c;
    return VSB_OK;
}

extern "C" int vsb_ctx_option(vsb_ctx_t* ctx, const char* name, int value) {
    if (!ctx || !name) return VSB_ERR_INVALID;
    if (!strcmp(name, "knn_impl")) {
        if (value < 0 || value > 6) return VSB_ERR_INVALID;
        ctx->knn_impl = value;
        return VSB_OK;
    }
    if (!strcmp(name, "knn_l2_impl")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->knn_l2_impl = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_variant")) {
        if (value < 0 || value > 5) return VSB_ERR_INVALID;
        ctx->gn_variant = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_impl")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->gn_impl = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_dedup")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->gn_dedup = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_cluster")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return VSB_ERR_INVALID;
        ctx->gn_cluster = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_cluster_threads")) {
        if (value != 0 && value != 256 && value != 512) return VSB_ERR_INVALID;
        ctx->gn_cluster_threads = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_tail")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->gn_tail = value;
        return VSB_OK;
    }
    if (!strcmp(name, "orb_scratch_mb")) {
        if (value < 16 || value > 65536) return VSB_ERR_INVALID;
        ctx->orb_scratch_mb = value;
        return VSB_OK;
    }
    if (!strcmp(name, "fast_impl")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->fast_impl = value;
        return VSB_OK;
    }
    if (!strcmp(name, "orb_lp")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->orb_lp = value;
        return VSB_OK;
    }
    if (!strcmp(name, "orb_impl")) {
        if (value < 0 || value > 15) return VSB_ERR_INVALID;
        ctx->orb_impl = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_stage_bytes")) {
        if (value < 0 || value > 128 * 1024) return VSB_ERR_INVALID;
        ctx->gn_stage_bytes = value & ~15;
        return VSB_OK;
    }
    if (!strcmp(name, "pyr_impl")) {
        if (value < 0 || value > 1) return VSB_ERR_INVALID;
        ctx->pyr_impl = value;
        return VSB_OK;
    }
    if (!strcmp(name, "gn_threads")) {
        if (value != 0 && value != 64 && value != 128 && value != 256 && value != 512 && value != 1024) return VSB_ERR_INVALID;
        ctx->gn_threads = value;
        return VSB_OK;
    }
    return VSB_ERR_INVALID;
}

extern "C" int vsb_ctx_destroy(vsb_ctx_t* ctx) {
    if (!ctx) return VSB_ERR_INVALID;
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    for (int i = 0; i < ctx->ws_n; i++) if (ctx->ws_ptr[i]) cudaFree(ctx->ws_ptr[i]);
    if (ctx->orb_lp_ready)
        for (int i = 0; i < 8; i++) { cudaStreamDestroy(ctx->orb_stream[i]); cudaEventDestroy(ctx->orb_ev_ready[i]); cudaEventDestroy(ctx->orb_ev_done[i]); }
    for (auto& r : ctx->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->prof_free) cudaEventDestroy(e);
    delete ctx;
    return VSB_OK;
}

extern "C" const char* vsb_last_cuda_error(vsb_ctx_t* ctx) { return ctx ? ctx->last_error : ""; }
extern "C" int vsb_sm_count(vsb_ctx_t* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" long long vsb_launch_count(vsb_ctx_t* ctx) { return ctx ? ctx->launches : 0; }

int vsb_orb_lp_streams(vsb_ctx* ctx) {
    if (ctx->orb_lp_ready) return VSB_OK;
    for (int i = 0; i < 8; i++) {
        VSB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->orb_stream[i], cudaStreamNonBlocking));
        VSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->orb_ev_ready[i], cudaEventDisableTiming));
        VSB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->orb_ev_done[i], cudaEventDisableTiming));
    }
    ctx->orb_lp_ready = 1;
    return VSB_OK;
}

int vsb_scratch_reserve(vsb_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) VSB_CUDA(ctx, cudaFree(ctx->scratch));   // cudaFree synchronises: no kernel still uses it
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 4;
        VSB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return VSB_OK;
}

int vsb_scratch2_reserve(vsb_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch2_bytes) {
        if (ctx->scratch2) VSB_CUDA(ctx, cudaFree(ctx->scratch2));
        ctx->scratch2 = nullptr;
        ctx->scratch2_bytes = 0;
        ctx->l2_fallback_counter = nullptr;
        size_t want = bytes + bytes / 4;
        VSB_CUDA(ctx, cudaMalloc(&ctx->scratch2, want));
        ctx->scratch2_bytes = want;
    }
    *out = ctx->scratch2;
    return VSB_OK;
}

int vsb_stream_ws_reserve(vsb_ctx* ctx, cudaStream_t st, size_t bytes, void** out) {
    int slot = -1;
    for (int i = 0; i < ctx->ws_n; i++) if (ctx->ws_stream[i] == st) slot = i;
    if (slot < 0) {
        if (ctx->ws_n == 8) {                              // more streams than slots: recycle the oldest (cudaFree synchronises)
            if (ctx->ws_ptr[0]) VSB_CUDA(ctx, cudaFree(ctx->ws_ptr[0]));
            for (int i = 1; i < 8; i++) { ctx->ws_ptr[i - 1] = ctx->ws_ptr[i]; ctx->ws_bytes[i - 1] = ctx->ws_bytes[i]; ctx->ws_stream[i - 1] = ctx->ws_stream[i]; }
            ctx->ws_n = 7;
        }
        slot = ctx->ws_n++;
        ctx->ws_ptr[slot] = nullptr; ctx->ws_bytes[slot] = 0; ctx->ws_stream[slot] = st;
    }
    if (bytes > ctx->ws_bytes[slot]) {
        if (ctx->ws_ptr[slot]) VSB_CUDA(ctx, cudaFree(ctx->ws_ptr[slot]));   // cudaFree synchronises: no kernel still uses it
        ctx->ws_ptr[slot] = nullptr; ctx->ws_bytes[slot] = 0;
        const size_t want = bytes + bytes / 8;
        VSB_CUDA(ctx, cudaMalloc(&ctx->ws_ptr[slot], want));
        ctx->ws_bytes[slot] = want;
    }
    *out = ctx->ws_ptr[slot];
    return VSB_OK;
}

extern "C" int vsb_init_pyramid(int w, int h, float fx, float fy, float cx, float cy, vsb_intr_t out[VSB_MAX_LEVELS]) {
    if (!out || w <= 0 || h <= 0) return VSB_ERR_INVALID;
    out[0].w = w; out[0].h = h;
    out[0].fx = fx; out[0].fy = fy; out[0].cx = cx; out[0].cy = cy;
    out[0].invfx = F_DIV(1.f, fx);                                           // VISystem.cpp:1463
    out[0].invfy = F_DIV(1.f, fy);
    for (int lvl = 1; lvl < VSB_MAX_LEVELS; lvl++) {
        out[lvl].w = w >> lvl;                                               // :1468
        out[lvl].h = h >> lvl;
        out[lvl].fx = (float)((double)out[lvl - 1].fx * 0.5);                // :1470
        out[lvl].fy = (float)((double)out[lvl - 1].fy * 0.5);
        out[lvl].cx = (float)(((double)cx + 0.5) / (double)(1 << lvl) - 0.5);  // :1472
        out[lvl].cy = (float)(((double)cy + 0.5) / (double)(1 << lvl) - 0.5);
        out[lvl].invfx = F_DIV(1.f, out[lvl].fx);                            // :1481
        out[lvl].invfy = F_DIV(1.f, out[lvl].fy);
    }
    return VSB_OK;
}

static void mat33_mul(const float* a, const float* b, float* c) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            c[3 * i + j] = vsb::dot3(a[3 * i], b[j], a[3 * i + 1], b[3 + j], a[3 * i + 2], b[6 + j]);
}

static void rot_to_quat(const float* m, float* q) {   // Eigen Quaternion(Matrix3)
    float t = F_ADD(F_ADD(m[0], m[4]), m[8]);
    if (t > 0.0f) {
        t = F_SQRT(F_ADD(t, 1.0f));
        q[3] = F_MUL(0.5f, t);
        t = F_DIV(0.5f, t);
        q[0] = F_MUL(F_SUB(m[7], m[5]), t);
        q[1] = F_MUL(F_SUB(m[2], m[6]), t);
        q[2] = F_MUL(F_SUB(m[3], m[1]), t);
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = F_SQRT(F_ADD(F_SUB(F_SUB(m[4 * i], m[4 * j]), m[4 * k]), 1.0f));
        q[i] = F_MUL(0.5f, t);
        t = F_DIV(0.5f, t);
        q[3] = F_MUL(F_SUB(m[3 * k + j], m[3 * j + k]), t);
        q[j] = F_MUL(F_ADD(m[3 * j + i], m[3 * i + j]), t);
        q[k] = F_MUL(F_ADD(m[3 * k + i], m[3 * i + k]), t);
    }
}

extern "C" int vsb_initial_pose(const float imu2cam[9], const float r_imu_res[9], const float t_res[3], float pose[7]) {
    if (!imu2cam || !r_imu_res || !t_res || !pose) return VSB_ERR_INVALID;
    const float it[9] = {imu2cam[0], imu2cam[3], imu2cam[6], imu2cam[1], imu2cam[4], imu2cam[7],
                         imu2cam[2], imu2cam[5], imu2cam[8]};
    float tmp[9], rc[9], r0[9];
    mat33_mul(it, r_imu_res, tmp);      // imu2camRotation.t() * residual_rotationMatrix * imu2camRotation, :1135
    mat33_mul(tmp, imu2cam, rc);
    // rotationMatrix2RPY (Plus.cpp:56-83), negated, RPY2rotationMatrix (Plus.cpp:182-220, VISystem.cpp:1146)
    const double r11 = rc[0], r21 = rc[3], r31 = rc[6], r32 = rc[7], r33 = rc[8];
    const double yaw = -atan2(r21, r11);
    const double pitch = -atan2(-r31, sqrt(r32 * r32 + r33 * r33));
    const double roll = -atan2(r32, r33);
    const double c1 = cos(roll), s1 = sin(roll), c2 = cos(pitch), s2 = sin(pitch), c3 = cos(yaw), s3 = sin(yaw);
    r0[0] = (float)(c3 * c2); r0[1] = (float)(c3 * s2 * s1 - s3 * c1); r0[2] = (float)(c3 * s2 * c1 + s3 * s1);
    r0[3] = (float)(s3 * c2); r0[4] = (float)(s3 * s2 * s1 + c3 * c1); r0[5] = (float)(s3 * s2 * c1 - c3 * s1);
    r0[6] = (float)(-s2);     r0[7] = (float)(c2 * s1);                r0[8] = (float)(c2 * c1);
    rot_to_quat(r0, pose);              // SE3(rotationEigen, Point(-sx,-sy,-sz)), :1162
    pose[4] = -t_res[0]; pose[5] = -t_res[1]; pose[6] = -t_res[2];
    return VSB_OK;
}

extern "C" int vsb_se3_from_rt(const float r[9], const float t[3], float pose[7]) {
    if (!r || !t || !pose) return VSB_ERR_INVALID;
    float q[4];
    rot_to_quat(r, q);   // SE3(Matrix3, Point): Eigen Quaternion(Matrix3), so3.hpp:422-427
    pose[0] = q[0]; pose[1] = q[1]; pose[2] = q[2]; pose[3] = q[3];
    pose[4] = t[0]; pose[5] = t[1]; pose[6] = t[2];
    return VSB_OK;
}

extern "C" int vsb_se3_mul(const float a[7], const float b[7], float out[7]) {
    if (!a || !b || !out) return VSB_ERR_INVALID;
    float tmp[7];
    vsb::se3_mul(a, b, tmp);
    for (int i = 0; i < 7; i++) out[i] = tmp[i];
    return VSB_OK;
}

// ---- per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg) --------------
static cudaEvent_t prof_event(vsb_ctx* ctx) {
    if (!ctx->prof_free.empty()) {
        cudaEvent_t e = ctx->prof_free.back();
        ctx->prof_free.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

int vsb_prof_begin(vsb_ctx* ctx, int id, cudaStream_t st) {
    vsb_prof_rec r;
    r.id = id;
    r.a = prof_event(ctx);
    r.b = prof_event(ctx);
    cudaEventRecord(r.a, st);
    ctx->prof_pending.push_back(r);
    return (int)ctx->prof_pending.size() - 1;
}

void vsb_prof_end(vsb_ctx* ctx, int slot, cudaStream_t st) { cudaEventRecord(ctx->prof_pending[slot].b, st); }

static const char* kKernelNames[VSB_K_COUNT] = {"knn2_hamming", "knn_unpack", "match_filter", "gather_keypoints",
                                                "pyramid", "gradient", "candidates", "gn_solve", "knn2_l2",
                                                "knn2_l2_prep", "gn_prepare", "match_stage", "warp_se3",
                                                "knn2_l2_final", "fast_score", "fast_compact", "orb"};

extern "C" int vsb_kernel_count(void) { return VSB_K_COUNT; }
extern "C" const char* vsb_kernel_name(int id) { return (id >= 0 && id < VSB_K_COUNT) ? kKernelNames[id] : ""; }

extern "C" int vsb_profile_enable(vsb_ctx_t* ctx, int on) {
    if (!ctx) return VSB_ERR_INVALID;
    ctx->prof_on = on ? 1 : 0;
    return VSB_OK;
}

// Folds every finished event pair into the per-kernel totals (synchronises on the recorded events).
static int prof_collect(vsb_ctx* ctx) {
    for (auto& r : ctx->prof_pending) {
        VSB_CUDA(ctx, cudaEventSynchronize(r.b));
        float ms = 0.f;
        VSB_CUDA(ctx, cudaEventElapsedTime(&ms, r.a, r.b));
        ctx->prof_ms[r.id] += ms;
        ctx->prof_n[r.id] += 1;
        ctx->prof_free.push_back(r.a);
        ctx->prof_free.push_back(r.b);
    }
    ctx->prof_pending.clear();
    return VSB_OK;
}

extern "C" int vsb_profile_reset(vsb_ctx_t* ctx) {
    if (!ctx) return VSB_ERR_INVALID;
    int rc = prof_collect(ctx);
    for (int i = 0; i < VSB_K_COUNT; i++) { ctx->prof_ms[i] = 0.0; ctx->prof_n[i] = 0; }
    return rc;
}

extern "C" int vsb_profile_read(vsb_ctx_t* ctx, int kernel_id, double* total_ms, long long* launches) {
    if (!ctx || kernel_id < 0 || kernel_id >= VSB_K_COUNT) return VSB_ERR_INVALID;
    int rc = prof_collect(ctx);
    if (total_ms) *total_ms = ctx->prof_ms[kernel_id];
    if (launches) *launches = ctx->prof_n[kernel_id];
    return rc;
}

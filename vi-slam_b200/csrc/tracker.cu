// tracker.cu — the whole tracking step on the device: the stage sequence of VISystemGPU::AddFrameGPU
// (reference src/VISystemGPU.cpp:144-169) and CameraGPU::addGPUKeyframe (src/CameraGPU.cpp:138-163) —
// Update (pyramid) -> computeGPUGoodMatches -> computeGradient -> ObtainPatchesPointsPreviousFrame ->
// EstimatePoseFeatures — batched over independent frame pairs, with no host round trip between stages.
// Pairs of a sequence are independent given their priors, so a sequence is one batch.
#include <cmath>
#include <cstdlib>
#include "common.cuh"
#include "knn_keys.cuh"

int vsb_knn2_hamming_keys(vsb_ctx* ctx, const uint8_t* d1, int n1_max, const int32_t* n1, const uint8_t* d2,
                          int n2_max, const int32_t* n2, int count, uint32_t* key12, uint32_t* key21, cudaStream_t st);
int vsb_gn_solve_stats(vsb_ctx_t* ctx, const uint8_t* prev_pyr, const uint8_t* cur_pyr, const int16_t* prev_gx,
                       const int16_t* prev_gy, int64_t pair_stride_pixels, const vsb_pyr_layout_t* layout,
                       const float* cand, int cand_cap, const int32_t* n_cand, const vsb_intr_t K[VSB_MAX_LEVELS],
                       const float* pose_in, const vsb_gn_opts_t* opts, int count, float* pose_out,
                       vsb_gn_trace_t* trace, int32_t* n_trace, unsigned long long* stats, void* patt_scratch,
                       int patt_ready, const void* xy_ready, void* stream);
int vsb_match_filter_keys(vsb_ctx_t* ctx, const void* keys12, const void* keys21, int key_bytes, int n1_max,
                          const int32_t* n1, int n2_max, const int32_t* n2, const float* kp1_xy, int count, int w, int h,
                          int n_cells, float ratio, int sym_mode, int32_t* good_q, int32_t* good_t, float* good_d,
                          int good_cap, int32_t* n_good, int32_t* n_sym, float* good_xy, void* stream);
int vsb_candidates_prepare(vsb_ctx_t* ctx, const float* good_xy, int good_cap, const int32_t* n_good, int count, int levels,
                           const int* lw, const int* lh, float* cand, int cand_cap, int32_t* n_cand,
                           const uint8_t* prev_pyr, int64_t pair_stride, const vsb_pyr_layout_t* layout, int first_lvl,
                           int last_lvl, void* patt, void* xy, const vsb_intr_t* K, int rec_abs, uint32_t dedup_mask,
                           int32_t* n_pts, const uint8_t* prev_l0, int64_t l0_stride, void* stream);
int vsb_pyramid_build_levels(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int count,
                             const vsb_pyr_layout_t* layout, uint8_t* pyr, int copy_l0, void* stream);
int vsb_gn_track(vsb_ctx_t* ctx, const uint8_t* cur_pyr, int64_t pair_stride_pixels, const vsb_pyr_layout_t* layout,
                 const void* patt, const int32_t* n_cand, const int32_t* n_pts, uint32_t dedup_mask, int cand_cap,
                 const vsb_intr_t K[VSB_MAX_LEVELS], const float* pose_in, const vsb_gn_opts_t* opts, int pair0, int count,
                 int threads, float* pose_out, vsb_gn_trace_t* trace, int32_t* n_trace, unsigned long long* stats,
                 const uint8_t* cur_l0, int64_t l0_stride, void* stream);
size_t vsb_orb_pyr_ws_bytes(int w, int h, int frames, int cap, int describe, float scale_factor, int nlevels, size_t budget);
int vsb_orb_detect_compute_pyr_ws(vsb_ctx_t* ctx, const uint8_t* img, int64_t img_stride, int pitch, int w, int h,
                                  int count, int nfeatures, float scale_factor, int nlevels, int fast_threshold, int cap,
                                  float* kp_xy, int32_t* kp_octave, float* kp_resp, float* kp_angle, uint8_t* desc,
                                  int32_t* n_kp, void* ws, size_t ws_bytes, void* stream);
int vsb_knn_unpack(vsb_ctx* ctx, const uint32_t* keys, int n_max, const int32_t* n, int count, int32_t* idx,
                   float* dist, cudaStream_t st);
int vsb_knn2_l2_keys(vsb_ctx* ctx, const float* d1, int n1_max, const int32_t* n1, const float* d2, int n2_max,
                     const int32_t* n2, int dim, int count, unsigned long long* key12, unsigned long long* key21,
                     cudaStream_t st);
int vsb_knn_unpack64(vsb_ctx* ctx, const unsigned long long* keys, int n_max, const int32_t* n, int count,
                     int32_t* idx, float* dist, cudaStream_t st);

namespace {

struct Slot {
    // frames: up to max_pairs + 1 (sequence) or 2 * max_pairs (independent pairs)
    uint8_t* pyr = nullptr;
    int16_t* gx = nullptr;
    int16_t* gy = nullptr;
    uint8_t* desc = nullptr;      // staging for the host entry: [max_pairs + 1][n_feat][desc_bytes]
    float* kp = nullptr;          // [max_pairs + 1][n_feat][2]
    int32_t* n_feat = nullptr;    // [max_pairs + 1]
    float* prior = nullptr;       // [max_pairs][7]
    uint32_t* key12 = nullptr;
    uint32_t* key21 = nullptr;
    int32_t* good_q = nullptr;
    int32_t* good_t = nullptr;
    float* good_d = nullptr;
    int32_t* n_good = nullptr;
    int32_t* n_sym = nullptr;
    float* good_xy = nullptr;
    float* cand = nullptr;
    uint2* patt = nullptr;        // per-point attributes of the solver, [max_pairs][levels][cand_cap]
    int32_t* n_cand = nullptr;
    int32_t* n_pts = nullptr;     // candidate points per level before merging, [max_pairs][levels] (gn_track.cu)
    float* pose = nullptr;
    void* orb_ws = nullptr;       // the detector's workspace of this slot (vsb_orb_detect_compute_pyr_ws), allocated on first use
    size_t orb_ws_bytes = 0;
    float* orb_resp = nullptr;    // [max_pairs + 1][n_feat] each, allocated on the first vsb_track_sequence_orb call
    float* orb_angle = nullptr;
    uint8_t* desc2 = nullptr;     // second descriptor set of the pairs host entry, [max_pairs][n_feat][desc_bytes], allocated on first use
    int32_t* n_feat2 = nullptr;   // its counts, [max_pairs]
    uint8_t* stage = nullptr;     // [max_pairs + 1][w * h] contiguous frames of the host entry (VSB_HOST_STAGING=1), allocated on first use
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // the solver's tail launch (pairs of the last partial wave, more threads each) runs beside the main one
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t ready = nullptr;  // vsb_track_sequence_orb_host: the slot's frames have arrived (recorded on the copy stream)
};

}  // namespace

struct vsb_tracker {
    vsb_ctx* ctx;
    vsb_tracker_cfg_t cfg;
    vsb_pyr_layout_t lay;
    vsb_intr_t K[VSB_MAX_LEVELS];
    int lw[VSB_MAX_LEVELS], lh[VSB_MAX_LEVELS];
    int good_cap, cand_cap, feat_cap;
    bool principal_point_ok;     // |cx|, |cy| >= 2^-10 at every level (gn_track.cu's division, see div3)
    Slot slot[2];
    int n_slots;
    unsigned long long* stats;   // device, 4 counters shared by both slots
    vsb_gn_trace_t* trace = nullptr;   // optional per-iteration trace of the next device-entry call (vsb_tracker_set_trace)
    int32_t* n_trace = nullptr;
    long long host_h2d_bytes, host_d2h_bytes, host_chunks;   // what the last vsb_track_sequence_host call moved
    // whole-sequence side inputs of the host entry (descriptors, key points, counts, priors: a tenth of the bytes), uploaded with
    // one copy each before the first chunk instead of four small copies per chunk; grown on demand
    uint8_t* seq_side = nullptr;
    size_t seq_side_bytes = 0;
    cudaEvent_t seq_side_ready = nullptr;
};

namespace {

template <typename T>
int dev_alloc(vsb_ctx* ctx, T** p, size_t n) {
    VSB_CUDA(ctx, cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)));
    return VSB_OK;
}

int slot_alloc(vsb_tracker* t, Slot& s) {
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    const size_t P = (size_t)c.max_pairs, F = 2 * P, N = (size_t)c.n_feat_max;
    int rc;
#define A(ptr, n) if ((rc = dev_alloc(ctx, &s.ptr, (n)))) return rc
    A(pyr, F * t->lay.frame_stride);
    if (c.gn.grad_mode == 0) {
        A(gx, F * t->lay.frame_stride);
        A(gy, F * t->lay.frame_stride);
    }
    A(desc, (P + 1) * N * c.desc_bytes);
    A(kp, (P + 1) * N * 2);
    A(n_feat, P + 1);
    A(prior, P * 7);
    const size_t kw = c.norm == 1 ? 1 : 2;   // L2 keys are 64-bit (float bits << 32 | index)
    A(key12, P * N * 2 * kw); A(key21, P * N * 2 * kw);
    A(good_q, P * t->good_cap); A(good_t, P * t->good_cap); A(good_d, P * t->good_cap);
    A(n_good, P); A(n_sym, P);
    A(good_xy, P * t->good_cap * 2);
    A(cand, P * VSB_MAX_LEVELS * (size_t)t->cand_cap * 4);
    A(patt, P * VSB_MAX_LEVELS * (size_t)t->cand_cap);
    A(n_cand, P * VSB_MAX_LEVELS);
    A(n_pts, P * VSB_MAX_LEVELS);
    A(pose, P * 7);
#undef A
    VSB_CUDA(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    VSB_CUDA(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    VSB_CUDA(ctx, cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
    VSB_CUDA(ctx, cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    VSB_CUDA(ctx, cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    VSB_CUDA(ctx, cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
    return VSB_OK;
}

void slot_free(Slot& s) {
    void* ptrs[] = {s.pyr, s.gx, s.gy, s.desc, s.kp, s.n_feat, s.prior, s.key12, s.key21, s.good_q, s.good_t, s.good_d, s.n_good, s.n_sym, s.good_xy, s.cand, s.patt, s.n_cand, s.n_pts, s.pose, s.orb_resp, s.orb_angle, s.stage, s.desc2, s.n_feat2};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
    if (s.aux) cudaStreamDestroy(s.aux);
    if (s.fork) cudaEventDestroy(s.fork);
    if (s.join) cudaEventDestroy(s.join);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.orb_ws) cudaFree(s.orb_ws);
    s = Slot();
}

// Threads per frame pair of gn_track.cu.  A pair is one block, so a batch that cannot fill the machine with small blocks
// gets more threads per pair: the largest block size whose resident blocks still hold the whole batch.  Two regimes,
// both measured (tools/kbench_gn.py, tools/leg_once.py):
//   * num_cells 49 (<= 64 features, 8.6 KB of back-projection tables): 128-thread blocks, 6-7 per SM, are the fastest for
//     a large batch (1999 pairs: 1.85 ms; 256 threads 1.96, 512 threads 2.09);
//   * the 200-feature cap (35 KB of tables, four times the points per pair): shared memory leaves room for three
//     128-thread blocks per SM only, and 512-thread blocks win (1024 pairs x 5000 features: 5.4 ms against 7.2 ms), 1024
//     threads when the batch fits one block per SM (199 KITTI pairs: 1.63 against 1.85 ms).
// A large batch runs in waves; the pairs of the last, partial wave would leave most of the machine idle while they finish,
// so they go to a second launch with more threads each, on a second stream: its blocks fill the SMs the main launch
// drains (1999 pairs: 1.85 ms with the tail launch, 1.89 ms without; tail sizes 100..600 pairs and 256 / 512 threads all
// land within 2 % of each other, tools/sweep_gn_tail.sh).
void gn_plan(const vsb_ctx* ctx, int count, int feat_cap, int* threads, int* n_tail, int* threads_tail) {
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const bool big = feat_cap > 64;
    const int base = big ? 512 : 128;                       // block size of a large batch
    const int per_sm = big ? 2 : (ctx->gn_variant == 0 ? 4 : 6);      // resident blocks per SM the wave arithmetic assumes (128 registers: 4)
    auto fit = [&](int n) {
        const int t = n <= (big ? 2 : 1) * sms ? 1024 : n <= 2 * sms ? 512 : n <= 3 * sms ? 256 : 128;
        return t > base ? t : base;
    };
    *n_tail = 0; *threads_tail = base;
    if (ctx->gn_threads) { *threads = ctx->gn_threads < 128 ? 128 : ctx->gn_threads; return; }
    *threads = fit(count);
    if (const char* e = getenv("VSB_GN_TAIL_PAIRS")) {           // experiment knob: explicit tail size / threads
        const char* tt = getenv("VSB_GN_TAIL_THREADS");
        *n_tail = atoi(e) < count ? atoi(e) : 0; *threads_tail = tt ? atoi(tt) : 512;
        return;
    }
    if (ctx->gn_tail && count > per_sm * sms) {
        const int rem = count % (per_sm * sms);
        const int tt = fit(rem);
        if (rem > 0 && tt > base) { *n_tail = rem; *threads_tail = tt; }
    }
}

// Stages after the pyramids exist: match -> filter -> candidates -> GN.  prev pyramid of pair c is
// pyr_prev + c * frame_stride, current is pyr_cur + c * frame_stride.
// The solver's reference-mode path (gn_track.cu behind the fused candidate pass): the only one that can take level 0 from
// the caller's frames instead of a copy inside the packed pyramid.
bool tables_path(const vsb_tracker* t) {
    const vsb_tracker_cfg_t& c = t->cfg;
    return c.gn.grad_mode == 1 && t->ctx->gn_impl == 1 && c.gn.weight_mode == 0 && c.gn.sample_mode == 0 &&
           t->principal_point_ok && t->lay.w[0] <= 4095 && t->lay.h[0] <= 4095;
}

// prev_l0 / cur_l0 (optional, tables_path only): level 0 of pair c at prev_l0 / cur_l0 + c * w * h; the packed pyramids then
// hold levels 1.. only.
int run_pairs(vsb_tracker* t, Slot& s, const uint8_t* pyr_prev, const uint8_t* pyr_cur, const int16_t* gx_prev,
              const int16_t* gy_prev, const uint8_t* d1, const uint8_t* d2, const float* kp1, const int32_t* n1,
              const int32_t* n2, const float* prior, int count, float* pose_out, int32_t* n_good_out, cudaStream_t st,
              const uint8_t* prev_l0 = nullptr, const uint8_t* cur_l0 = nullptr) {
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    const int N = c.n_feat_max;
    int rc;
    // kNN (both directions) -> packed keys; the filter reads them in place and gathers the good key points itself
    int key_bytes;
    if (c.norm == 1) {
        if (c.desc_bytes != 32) return VSB_ERR_UNSUPPORTED;
        if ((rc = vsb_knn2_hamming_keys(ctx, d1, N, n1, d2, N, n2, count, s.key12, s.key21, st))) return rc;
        key_bytes = 4;
    } else {
        if (c.desc_bytes <= 0 || (c.desc_bytes & 3)) return VSB_ERR_INVALID;
        const int dim = c.desc_bytes / 4;
        unsigned long long* k12 = reinterpret_cast<unsigned long long*>(s.key12);
        unsigned long long* k21 = reinterpret_cast<unsigned long long*>(s.key21);
        if ((rc = vsb_knn2_l2_keys(ctx, reinterpret_cast<const float*>(d1), N, n1, reinterpret_cast<const float*>(d2), N,
                                   n2, dim, count, k12, k21, st)))
            return rc;
        key_bytes = 8;
    }
    if ((rc = vsb_match_filter_keys(ctx, s.key12, s.key21, key_bytes, N, n1, N, n2, kp1, count, c.w, c.h, c.n_cells, c.ratio,
                                    c.sym_mode, s.good_q, s.good_t, s.good_d, t->good_cap, s.n_good, s.n_sym, s.good_xy, st)))
        return rc;
    // candidate points and — when the gradients are evaluated at the points (grad_mode 1) — the solver's per-point
    // attribute records in the same pass
    const int fused = c.gn.grad_mode == 1 ? 1 : 0;
    // reference modes (identity weights, nearest-pixel lookup, FP64 Gram): gn_track.cu — 8-byte records that name their
    // slots in per-feature back-projection tables, nothing else per point
    const bool tables = tables_path(t);
    if (!tables && (prev_l0 || cur_l0)) return VSB_ERR_INVALID;
    const int64_t l0_stride = (int64_t)c.w * c.h;
    // levels whose candidate points are merged per distinct pixel: every level small enough for the candidate pass's byte map
    const uint32_t dedup_mask = (tables && ctx->gn_dedup && ctx->gn_variant != 1) ? 0xFFFFFFFFu : 0u;
    // ... otherwise gn_solve.cu; with identity weights the points are handed over already back-projected, as doubles,
    // in the candidate buffer itself (a double2 is as wide as the float4 row it replaces)
    const int unit = fused && !tables && c.gn.weight_mode == 0;
    if ((rc = vsb_candidates_prepare(ctx, s.good_xy, t->good_cap, s.n_good, count, t->lay.levels, t->lw, t->lh, s.cand,
                                     t->cand_cap, s.n_cand, fused ? pyr_prev : nullptr, t->lay.frame_stride, &t->lay,
                                     c.gn.first_lvl, c.gn.last_lvl, fused ? s.patt : nullptr, unit ? (void*)s.cand : nullptr,
                                     t->K, tables ? 1 : 0, dedup_mask, s.n_pts, prev_l0, l0_stride, st)))
        return rc;
    if (tables) {
        int threads, n_tail, threads_tail;
        gn_plan(ctx, count, t->feat_cap, &threads, &n_tail, &threads_tail);
        const int n_main = count - n_tail;
        ProfScope ps(ctx, VSB_K_GN_SOLVE, st);
        if (n_tail > 0) {
            VSB_CUDA(ctx, cudaEventRecord(s.fork, st));
            VSB_CUDA(ctx, cudaStreamWaitEvent(s.aux, s.fork, 0));
        }
        if ((rc = vsb_gn_track(ctx, pyr_cur, t->lay.frame_stride, &t->lay, s.patt, s.n_cand, s.n_pts, dedup_mask, t->cand_cap,
                               t->K, prior, &c.gn, 0, n_main, threads, pose_out, t->trace, t->n_trace, t->stats, cur_l0, l0_stride, st)))
            return rc;
        if (n_tail > 0) {
            if ((rc = vsb_gn_track(ctx, pyr_cur, t->lay.frame_stride, &t->lay, s.patt, s.n_cand, s.n_pts, dedup_mask, t->cand_cap,
                                   t->K, prior, &c.gn, n_main, n_tail, threads_tail, pose_out, t->trace,
                                   t->n_trace, t->stats, cur_l0, l0_stride, s.aux)))
                return rc;
            VSB_CUDA(ctx, cudaEventRecord(s.join, s.aux));
            VSB_CUDA(ctx, cudaStreamWaitEvent(st, s.join, 0));
        }
    } else if ((rc = vsb_gn_solve_stats(ctx, pyr_prev, pyr_cur, gx_prev, gy_prev, t->lay.frame_stride, &t->lay,
                                        unit ? nullptr : s.cand, t->cand_cap, s.n_cand, t->K, prior, &c.gn, count, pose_out,
                                        t->trace, t->n_trace, t->stats, s.patt, fused, unit ? (const void*)s.cand : nullptr, st)))
        return rc;
    if (n_good_out)
        VSB_CUDA(ctx, cudaMemcpyAsync(n_good_out, s.n_good, sizeof(int32_t) * count, cudaMemcpyDeviceToDevice, st));
    return VSB_OK;
}

}  // namespace

extern "C" int vsb_tracker_create(vsb_ctx_t* ctx, const vsb_tracker_cfg_t* cfg, vsb_tracker_t** out) {
    if (!ctx || !cfg || !out) return VSB_ERR_INVALID;
    *out = nullptr;
    if (cfg->w <= 0 || cfg->h <= 0 || cfg->n_feat_max <= 0 || cfg->max_pairs <= 0 || cfg->n_cells < 1)
        return VSB_ERR_INVALID;
    vsb_tracker* t = new vsb_tracker();
    t->ctx = ctx;
    t->cfg = *cfg;
    int rc = vsb_pyr_layout(cfg->w, cfg->h, VSB_MAX_LEVELS, &t->lay);
    if (rc) { delete t; return rc; }
    vsb_init_pyramid(cfg->w, cfg->h, cfg->fx, cfg->fy, cfg->cx, cfg->cy, t->K);
    for (int l = 0; l < VSB_MAX_LEVELS; l++) { t->lw[l] = cfg->w >> l; t->lh[l] = cfg->h >> l; }  // Camera.cpp:44-47
    const int root = (int)floor(sqrt((double)cfg->n_cells));
    t->good_cap = root * root;
    const int nf = t->good_cap < VSB_MAX_GN_FEATURES ? t->good_cap : VSB_MAX_GN_FEATURES;
    t->cand_cap = 121 * nf;
    t->feat_cap = nf;
    t->principal_point_ok = true;
    for (int l = 0; l < VSB_MAX_LEVELS; l++)
        if (!(fabsf(t->K[l].cx) >= 9.765625e-4f && fabsf(t->K[l].cy) >= 9.765625e-4f)) t->principal_point_ok = false;
    t->n_slots = 2;
    t->stats = nullptr;
    t->host_h2d_bytes = t->host_d2h_bytes = t->host_chunks = 0;
    if (cudaMalloc((void**)&t->stats, 4 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(t->stats, 0, 4 * sizeof(unsigned long long)) != cudaSuccess) {
        delete t;
        return VSB_ERR_CUDA;
    }
    for (int i = 0; i < t->n_slots; i++) {
        rc = slot_alloc(t, t->slot[i]);
        if (rc) { vsb_tracker_destroy(t); return rc; }
    }
    *out = t;
    return VSB_OK;
}

extern "C" int vsb_tracker_destroy(vsb_tracker_t* t) {
    if (!t) return VSB_ERR_INVALID;
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) slot_free(t->slot[i]);
    if (t->seq_side) cudaFree(t->seq_side);
    if (t->seq_side_ready) cudaEventDestroy(t->seq_side_ready);
    if (t->stats) cudaFree(t->stats);
    delete t;
    return VSB_OK;
}

static int track_sequence_slot(vsb_tracker* t, Slot& s, const uint8_t* frames, bool frames_in_place, const uint8_t* desc,
                               const float* kp_xy, const int32_t* n_feat, const float* prior, int n_frames,
                               float* pose, int32_t* n_good, cudaStream_t st) {
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    const int pairs = n_frames - 1;
    if (pairs > c.max_pairs) return VSB_ERR_CAPACITY;
    int rc;
    // Camera::Update for every frame.  Level 0: already in place when the frames were uploaded straight into the pyramid;
    // read from the caller's frames by the candidate pass and the solver on the reference-mode path (no copy at all);
    // copied otherwise.
    const bool ext_l0 = !frames_in_place && tables_path(t);
    if ((rc = vsb_pyramid_build_levels(ctx, frames_in_place ? nullptr : frames, (int64_t)c.w * c.h, c.w, n_frames, &t->lay, s.pyr,
                                       ext_l0 ? 0 : 1, st)))
        return rc;
    // Camera::computeGradient — only the previous frame of each pair is read by the solver
    if (c.gn.grad_mode == 0)
        if ((rc = vsb_gradient_build(ctx, s.pyr, pairs, &t->lay, s.gx, s.gy, nullptr, st))) return rc;
    const size_t dstride = (size_t)c.n_feat_max * c.desc_bytes;
    return run_pairs(t, s, s.pyr, s.pyr + t->lay.frame_stride, s.gx, s.gy, desc, desc + dstride, kp_xy, n_feat,
                     n_feat ? n_feat + 1 : nullptr, prior, pairs, pose, n_good, st, ext_l0 ? frames : nullptr,
                     ext_l0 ? frames + (size_t)c.w * c.h : nullptr);
}

extern "C" int vsb_track_sequence(vsb_tracker_t* t, const uint8_t* frames, const uint8_t* desc, const float* kp_xy,
                                  const int32_t* n_feat, const float* pose_prior, int n_frames, float* pose,
                                  int32_t* n_good, void* stream) {
    if (!t || !frames || !desc || !kp_xy || !pose_prior || !pose) return VSB_ERR_INVALID;
    if (n_frames < 2) return n_frames < 0 ? VSB_ERR_INVALID : VSB_OK;
    if (n_frames - 1 > t->cfg.max_pairs) return VSB_ERR_CAPACITY;
    return track_sequence_slot(t, t->slot[0], frames, false, desc, kp_xy, n_feat, pose_prior, n_frames, pose, n_good,
                               (cudaStream_t)stream);
}

namespace {
__global__ void clamp_counts_kernel(int32_t* n, int count, int cap) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < count) n[i] = min(n[i], cap);
}
}  // namespace

// The loop from images alone: every frame goes through cv::ORB::create(nfeatures) on the device (orb.cu), the key points and
// descriptors stay on the device and feed the matcher (Camera::detectAndComputeFeatures -> computeGoodMatches ->
// EstimatePoseFeatures, VISystemGPU.cpp:144-169).  Frames with more key points than cfg.n_feat_max keep the first n_feat_max.
static int track_sequence_orb_slot(vsb_tracker* t, Slot& s, const uint8_t* frames, const float* prior, int n_frames, int nfeatures,
                                   float* pose, int32_t* n_good, int32_t* n_feat_out, cudaStream_t st) {
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    int rc;
    const size_t per = (size_t)(c.max_pairs + 1) * c.n_feat_max;
    if (!s.orb_resp && (rc = dev_alloc(ctx, &s.orb_resp, per))) return rc;
    if (!s.orb_angle && (rc = dev_alloc(ctx, &s.orb_angle, per))) return rc;
    if (!s.orb_ws) {
        // as many frames as the slot holds, inside the context's budget for the detector ("orb_scratch_mb")
        const size_t budget = (size_t)(ctx->orb_scratch_mb > 0 ? ctx->orb_scratch_mb : 32768) << 20;
        size_t bytes = vsb_orb_pyr_ws_bytes(c.w, c.h, c.max_pairs + 1, c.n_feat_max, 1, 1.2f, 8, budget);    // (one block per level if that fits)
        const size_t floor_bytes = vsb_orb_pyr_ws_bytes(c.w, c.h, 1, c.n_feat_max, 1, 1.2f, 8, 0);
        if (bytes > budget) bytes = budget > floor_bytes ? budget : floor_bytes;
        VSB_CUDA(ctx, cudaMalloc(&s.orb_ws, bytes));
        s.orb_ws_bytes = bytes;
    }
    if ((rc = vsb_orb_detect_compute_pyr_ws(ctx, frames, (int64_t)c.w * c.h, c.w, c.w, c.h, n_frames, nfeatures, 1.2f, 8, 20,
                                            c.n_feat_max, s.kp, nullptr, s.orb_resp, s.orb_angle, s.desc, s.n_feat, s.orb_ws,
                                            s.orb_ws_bytes, (void*)st)))
        return rc;
    clamp_counts_kernel<<<vsb_div_up(n_frames, 256), 256, 0, st>>>(s.n_feat, n_frames, c.n_feat_max);
    VSB_LAUNCHED(ctx);
    if (n_feat_out) VSB_CUDA(ctx, cudaMemcpyAsync(n_feat_out, s.n_feat, (size_t)n_frames * sizeof(int32_t), cudaMemcpyDefault, st));
    return track_sequence_slot(t, s, frames, false, s.desc, s.kp, s.n_feat, prior, n_frames, pose, n_good, st);
}

extern "C" int vsb_track_sequence_orb(vsb_tracker_t* t, const uint8_t* frames, const float* pose_prior, int n_frames,
                                      int nfeatures, float* pose, int32_t* n_good, int32_t* n_feat_out, void* stream) {
    if (!t || !frames || !pose_prior || !pose || nfeatures < 0) return VSB_ERR_INVALID;
    if (t->cfg.norm != 1 || t->cfg.desc_bytes != 32) return VSB_ERR_INVALID;        // ORB descriptors: 32 bytes, Hamming
    if (n_frames < 2) return n_frames < 0 ? VSB_ERR_INVALID : VSB_OK;
    if (n_frames - 1 > t->cfg.max_pairs) return VSB_ERR_CAPACITY;
    return track_sequence_orb_slot(t, t->slot[0], frames, pose_prior, n_frames, nfeatures, pose, n_good, n_feat_out,
                                   (cudaStream_t)stream);
}

// The loop from images alone, from HOST buffers: what a caller of the reference's CameraGPU path has — frames in host memory,
// key points and descriptors made on the device (VISystemGPU.cpp:144-169).  Chunks of cfg.max_pairs pairs alternate between the
// two slots, each with its own stream, frame buffer and detector workspace: the frames of chunk i + 1 travel while chunk i is
// tracked, and the launch-latency-bound parts of a chunk (a dozen launches per pyramid level, the solver's tail: ~1.3 ms per chunk
// whatever its size) overlap the other chunk's kernels.  A chunk's last frame is the next chunk's first: it travels and is
// described twice (1 / max_pairs of the work), which keeps the chunks independent.
extern "C" int vsb_track_sequence_orb_host(vsb_tracker_t* t, const uint8_t* h_frames, const float* h_pose_prior, int n_frames,
                                           int nfeatures, float* h_pose, int32_t* h_n_good, int32_t* h_n_feat) {
    if (!t || !h_frames || !h_pose_prior || !h_pose || nfeatures < 0) return VSB_ERR_INVALID;
    if (t->cfg.norm != 1 || t->cfg.desc_bytes != 32) return VSB_ERR_INVALID;
    if (n_frames < 2) return n_frames < 0 ? VSB_ERR_INVALID : VSB_OK;
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    const size_t fbytes = (size_t)c.w * c.h;
    const int total_pairs = n_frames - 1;
    long long h2d = 0, d2h = 0;
    int chunk_idx = 0, rc = VSB_OK;
    for (int p0 = 0; p0 < total_pairs && rc == VSB_OK; chunk_idx++) {
        const int remaining = total_pairs - p0;
        int pairs = remaining < c.max_pairs ? remaining : c.max_pairs;
        // growing chunks: the first is an eighth of the capacity, so the kernels start after a short copy (small chunks run the
        // detector's levels on separate streams and are not much dearer per frame), every next one 1.4 x its predecessor — its
        // copy (6.5 us per 752x480 frame) still hides behind the predecessor's kernels (9-10 us per frame)
        if (total_pairs > c.max_pairs && c.max_pairs >= 64) {
            double sz = c.max_pairs / 8.0;
            for (int k = 0; k < chunk_idx && sz < c.max_pairs; k++) sz *= 1.4;
            if ((int)sz < pairs) pairs = (int)sz;
        }
        const int nf = pairs + 1;
        Slot& s = t->slot[chunk_idx & 1];
        cudaStream_t st = s.stream;
        if (!s.stage) { if (cudaMalloc((void**)&s.stage, (size_t)(c.max_pairs + 1) * fbytes) != cudaSuccess) { rc = VSB_ERR_CUDA; break; } }
        // (stream order: the slot's previous chunk has been tracked before this copy overwrites its frames)
        VSB_CUDA(ctx, cudaMemcpyAsync(s.stage, h_frames + (size_t)p0 * fbytes, (size_t)nf * fbytes, cudaMemcpyHostToDevice, st));
        VSB_CUDA(ctx, cudaMemcpyAsync(s.prior, h_pose_prior + (size_t)p0 * 7, (size_t)pairs * 28, cudaMemcpyHostToDevice, st));
        h2d += (long long)nf * fbytes + pairs * 28LL;
        d2h += pairs * 28LL + (h_n_good ? pairs * 4LL : 0) + (h_n_feat ? nf * 4LL : 0);
        if ((rc = track_sequence_orb_slot(t, s, s.stage, s.prior, nf, nfeatures, s.pose, nullptr, h_n_feat ? h_n_feat + p0 : nullptr, st)))
            break;
        VSB_CUDA(ctx, cudaMemcpyAsync(h_pose + (size_t)p0 * 7, s.pose, (size_t)pairs * 28, cudaMemcpyDeviceToHost, st));
        if (h_n_good) VSB_CUDA(ctx, cudaMemcpyAsync(h_n_good + p0, s.n_good, pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        p0 += pairs;
    }
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(t->slot[i].stream);      // also on the error path: copies may be in flight
    if (rc) return rc;
    t->host_h2d_bytes = h2d; t->host_d2h_bytes = d2h; t->host_chunks = chunk_idx;
    return VSB_OK;
}

extern "C" int vsb_track_pairs(vsb_tracker_t* t, const uint8_t* prev, const uint8_t* cur, const uint8_t* d1,
                               const uint8_t* d2, const float* kp1_xy, const int32_t* n1, const int32_t* n2,
                               const float* pose_prior, int count, float* pose, int32_t* n_good, void* stream) {
    if (!t || !prev || !cur || !d1 || !d2 || !kp1_xy || !pose_prior || !pose) return VSB_ERR_INVALID;
    if (count < 0) return VSB_ERR_INVALID;
    if (count == 0) return VSB_OK;
    if (count > t->cfg.max_pairs) return VSB_ERR_CAPACITY;
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    Slot& s = t->slot[0];
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* pyr_prev = s.pyr;
    uint8_t* pyr_cur = s.pyr + (size_t)c.max_pairs * t->lay.frame_stride;
    int rc;
    const bool ext_l0 = tables_path(t);      // level 0 is read from prev / cur themselves
    if ((rc = vsb_pyramid_build_levels(ctx, prev, (int64_t)c.w * c.h, c.w, count, &t->lay, pyr_prev, ext_l0 ? 0 : 1, st))) return rc;
    if ((rc = vsb_pyramid_build_levels(ctx, cur, (int64_t)c.w * c.h, c.w, count, &t->lay, pyr_cur, ext_l0 ? 0 : 1, st))) return rc;
    if (c.gn.grad_mode == 0)
        if ((rc = vsb_gradient_build(ctx, pyr_prev, count, &t->lay, s.gx, s.gy, nullptr, st))) return rc;
    return run_pairs(t, s, pyr_prev, pyr_cur, s.gx, s.gy, d1, d2, kp1_xy, n1, n2, pose_prior, count, pose, n_good, st,
                     ext_l0 ? prev : nullptr, ext_l0 ? cur : nullptr);
}

extern "C" int vsb_track_sequence_host(vsb_tracker_t* t, const uint8_t* h_frames, const uint8_t* h_desc,
                                       const float* h_kp_xy, const int32_t* h_n_feat, const float* h_pose_prior,
                                       int n_frames, float* h_pose, int32_t* h_n_good) {
    if (!t || !h_frames || !h_desc || !h_kp_xy || !h_pose_prior || !h_pose) return VSB_ERR_INVALID;
    if (n_frames < 2) return n_frames < 0 ? VSB_ERR_INVALID : VSB_OK;
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    const size_t fbytes = (size_t)c.w * c.h;
    const size_t dstride = (size_t)c.n_feat_max * c.desc_bytes;
    const size_t kstride = (size_t)c.n_feat_max * 2;
    const int total_pairs = n_frames - 1;
    int chunk_idx = 0;
    long long h2d = 0, d2h = 0;
    const char* stg = getenv("VSB_HOST_STAGING");
    const bool staged = stg && stg[0] == '1';
    // side inputs in one go (up to 1 GB of them; longer sequences fall back to per-chunk copies)
    const char* upf = getenv("VSB_HOST_UPFRONT");
    const size_t al = 256;
    const size_t sz_desc = ((size_t)n_frames * dstride + al - 1) / al * al, sz_kp = ((size_t)n_frames * kstride * 4 + al - 1) / al * al;
    const size_t sz_nf = ((size_t)n_frames * 4 + al - 1) / al * al, sz_pr = ((size_t)total_pairs * 28 + al - 1) / al * al;
    const size_t side_total = sz_desc + sz_kp + sz_nf + sz_pr;
    // ... when they are small next to the frames (1000 ORB descriptors: 40 KB against a 361 KB frame).  Large side inputs
    // (5000 descriptors, float descriptors) would hold back the first chunk's kernels behind one long copy: they travel with
    // their chunk instead (configs[2]: 6.74 -> 6.12 ms, configs[3]: 5.43 -> 5.14 ms; tools/sweep_cfg_host_chunk.py).
    const bool side_small = (dstride + kstride * 4) * 4 <= fbytes;
    const bool upfront = (upf ? upf[0] != '0' : side_small) && side_total <= ((size_t)1 << 30);
    uint8_t* a_desc = nullptr; float* a_kp = nullptr; int32_t* a_nf = nullptr; float* a_pr = nullptr;
    if (upfront) {
        if (t->seq_side_bytes < side_total) {
            VSB_CUDA(ctx, cudaDeviceSynchronize());
            if (t->seq_side) cudaFree(t->seq_side);
            t->seq_side = nullptr; t->seq_side_bytes = 0;
            VSB_CUDA(ctx, cudaMalloc((void**)&t->seq_side, side_total));
            t->seq_side_bytes = side_total;
        }
        if (!t->seq_side_ready) VSB_CUDA(ctx, cudaEventCreateWithFlags(&t->seq_side_ready, cudaEventDisableTiming));
        a_desc = t->seq_side;
        a_kp = reinterpret_cast<float*>(t->seq_side + sz_desc);
        a_nf = reinterpret_cast<int32_t*>(t->seq_side + sz_desc + sz_kp);
        a_pr = reinterpret_cast<float*>(t->seq_side + sz_desc + sz_kp + sz_nf);
        cudaStream_t s0 = t->slot[0].stream;
        VSB_CUDA(ctx, cudaMemcpyAsync(a_desc, h_desc, (size_t)n_frames * dstride, cudaMemcpyHostToDevice, s0));
        VSB_CUDA(ctx, cudaMemcpyAsync(a_kp, h_kp_xy, (size_t)n_frames * kstride * 4, cudaMemcpyHostToDevice, s0));
        if (h_n_feat) VSB_CUDA(ctx, cudaMemcpyAsync(a_nf, h_n_feat, (size_t)n_frames * 4, cudaMemcpyHostToDevice, s0));
        VSB_CUDA(ctx, cudaMemcpyAsync(a_pr, h_pose_prior, (size_t)total_pairs * 28, cudaMemcpyHostToDevice, s0));
        VSB_CUDA(ctx, cudaEventRecord(t->seq_side_ready, s0));
        VSB_CUDA(ctx, cudaStreamWaitEvent(t->slot[1].stream, t->seq_side_ready, 0));
        h2d += (long long)n_frames * (dstride + kstride * 4) + (h_n_feat ? n_frames * 4LL : 0) + total_pairs * 28LL;
    }
    for (int p0 = 0; p0 < total_pairs; chunk_idx++) {
        const int remaining = total_pairs - p0;
        int pairs = remaining < c.max_pairs ? remaining : c.max_pairs;
        if (remaining <= 2 * c.max_pairs && remaining > 32) pairs = (remaining + 1) / 2 < c.max_pairs ? (remaining + 1) / 2 : c.max_pairs;
        if (pairs < 32 && remaining >= 32) pairs = 32;
        if (pairs > c.max_pairs) pairs = c.max_pairs;      // the slot buffers hold max_pairs pairs, whatever the schedule prefers
        const int nf = pairs + 1;
        Slot& s = t->slot[chunk_idx & 1];
        cudaStream_t st = s.stream;
        if (chunk_idx >= 2) VSB_CUDA(ctx, cudaEventSynchronize(s.done));   // slot buffers are free again
        if (staged) {
            // experiment knob: one contiguous copy into a staging block, the pyramid kernel copies level 0 (one more pass over
            // the frames on the device, hidden behind the next upload)
            if (!s.stage) VSB_CUDA(ctx, cudaMalloc((void**)&s.stage, (size_t)(c.max_pairs + 1) * fbytes));
            VSB_CUDA(ctx, cudaMemcpyAsync(s.stage, h_frames + (size_t)p0 * fbytes, (size_t)nf * fbytes, cudaMemcpyHostToDevice, st));
        } else {
            // frames go straight into level 0 of the packed pyramid (one strided copy, no repack kernel)
            VSB_CUDA(ctx, cudaMemcpy2DAsync(s.pyr, (size_t)t->lay.frame_stride, h_frames + (size_t)p0 * fbytes, fbytes,
                                            fbytes, nf, cudaMemcpyHostToDevice, st));
        }
        if (!upfront) {
            VSB_CUDA(ctx, cudaMemcpyAsync(s.desc, h_desc + (size_t)p0 * dstride, nf * dstride, cudaMemcpyHostToDevice, st));
            VSB_CUDA(ctx, cudaMemcpyAsync(s.kp, h_kp_xy + (size_t)p0 * kstride, nf * kstride * sizeof(float),
                                          cudaMemcpyHostToDevice, st));
            if (h_n_feat)
                VSB_CUDA(ctx, cudaMemcpyAsync(s.n_feat, h_n_feat + p0, nf * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            VSB_CUDA(ctx, cudaMemcpyAsync(s.prior, h_pose_prior + (size_t)p0 * 7, (size_t)pairs * 7 * sizeof(float),
                                          cudaMemcpyHostToDevice, st));
            h2d += (long long)nf * (dstride + kstride * sizeof(float)) + (h_n_feat ? nf * 4LL : 0) + pairs * 28LL;
        }
        h2d += (long long)nf * fbytes;
        d2h += pairs * 28LL + (h_n_good ? pairs * 4LL : 0);
        int rc = upfront
            ? track_sequence_slot(t, s, staged ? s.stage : nullptr, !staged, a_desc + (size_t)p0 * dstride, a_kp + (size_t)p0 * kstride,
                                  h_n_feat ? a_nf + p0 : nullptr, a_pr + (size_t)p0 * 7, nf, s.pose, nullptr, st)
            : track_sequence_slot(t, s, staged ? s.stage : nullptr, !staged, s.desc, s.kp, h_n_feat ? s.n_feat : nullptr, s.prior, nf,
                                  s.pose, nullptr, st);
        if (rc) {
            // copies into the caller's host buffers may still be in flight: finish them before handing the buffers back
            for (int i = 0; i < 2; i++) cudaStreamSynchronize(t->slot[i].stream);
            return rc;
        }
        VSB_CUDA(ctx, cudaMemcpyAsync(h_pose + (size_t)p0 * 7, s.pose, (size_t)pairs * 7 * sizeof(float),
                                      cudaMemcpyDeviceToHost, st));
        if (h_n_good)
            VSB_CUDA(ctx, cudaMemcpyAsync(h_n_good + p0, s.n_good, pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        VSB_CUDA(ctx, cudaEventRecord(s.done, st));
        p0 += pairs;
    }
    for (int i = 0; i < 2; i++) VSB_CUDA(ctx, cudaStreamSynchronize(t->slot[i].stream));
    t->host_h2d_bytes = h2d; t->host_d2h_bytes = d2h; t->host_chunks = chunk_idx;
    return VSB_OK;
}

// Independent pairs from HOST buffers (BASELINE configs[4] end to end): chunks of cfg.max_pairs pairs alternate between the
// two slots / streams, so the upload of chunk i + 1 overlaps the kernels of chunk i.  Frames go straight into level 0 of the
// packed pyramids (one strided copy per frame set).
extern "C" int vsb_track_pairs_host(vsb_tracker_t* t, const uint8_t* h_prev, const uint8_t* h_cur, const uint8_t* h_d1,
                                    const uint8_t* h_d2, const float* h_kp1_xy, const int32_t* h_n1, const int32_t* h_n2,
                                    const float* h_pose_prior, int count, float* h_pose, int32_t* h_n_good) {
    if (!t || !h_prev || !h_cur || !h_d1 || !h_d2 || !h_kp1_xy || !h_pose_prior || !h_pose) return VSB_ERR_INVALID;
    if ((h_n1 == nullptr) != (h_n2 == nullptr)) return VSB_ERR_INVALID;
    if (count <= 0) return count < 0 ? VSB_ERR_INVALID : VSB_OK;
    vsb_ctx* ctx = t->ctx;
    const vsb_tracker_cfg_t& c = t->cfg;
    const size_t fbytes = (size_t)c.w * c.h;
    const size_t dstride = (size_t)c.n_feat_max * c.desc_bytes;
    const size_t kstride = (size_t)c.n_feat_max * 2;
    long long h2d = 0, d2h = 0;
    int chunk_idx = 0;
    int rc = VSB_OK;
    for (int p0 = 0; p0 < count && rc == VSB_OK; chunk_idx++) {
        const int pairs = count - p0 < c.max_pairs ? count - p0 : c.max_pairs;
        Slot& s = t->slot[chunk_idx & 1];
        cudaStream_t st = s.stream;
        if (chunk_idx >= 2) VSB_CUDA(ctx, cudaEventSynchronize(s.done));   // slot buffers are free again
        if (!s.desc2) VSB_CUDA(ctx, cudaMalloc((void**)&s.desc2, (size_t)c.max_pairs * dstride));
        if (!s.n_feat2) VSB_CUDA(ctx, cudaMalloc((void**)&s.n_feat2, (size_t)c.max_pairs * sizeof(int32_t)));
        uint8_t* pyr_prev = s.pyr;
        uint8_t* pyr_cur = s.pyr + (size_t)c.max_pairs * t->lay.frame_stride;
        VSB_CUDA(ctx, cudaMemcpy2DAsync(pyr_prev, (size_t)t->lay.frame_stride, h_prev + (size_t)p0 * fbytes, fbytes, fbytes, pairs,
                                        cudaMemcpyHostToDevice, st));
        VSB_CUDA(ctx, cudaMemcpy2DAsync(pyr_cur, (size_t)t->lay.frame_stride, h_cur + (size_t)p0 * fbytes, fbytes, fbytes, pairs,
                                        cudaMemcpyHostToDevice, st));
        VSB_CUDA(ctx, cudaMemcpyAsync(s.desc, h_d1 + (size_t)p0 * dstride, pairs * dstride, cudaMemcpyHostToDevice, st));
        VSB_CUDA(ctx, cudaMemcpyAsync(s.desc2, h_d2 + (size_t)p0 * dstride, pairs * dstride, cudaMemcpyHostToDevice, st));
        VSB_CUDA(ctx, cudaMemcpyAsync(s.kp, h_kp1_xy + (size_t)p0 * kstride, pairs * kstride * sizeof(float), cudaMemcpyHostToDevice, st));
        VSB_CUDA(ctx, cudaMemcpyAsync(s.prior, h_pose_prior + (size_t)p0 * 7, (size_t)pairs * 28, cudaMemcpyHostToDevice, st));
        if (h_n1) {
            VSB_CUDA(ctx, cudaMemcpyAsync(s.n_feat, h_n1 + p0, pairs * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            VSB_CUDA(ctx, cudaMemcpyAsync(s.n_feat2, h_n2 + p0, pairs * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        }
        h2d += (long long)pairs * (2 * fbytes + 2 * dstride + kstride * sizeof(float) + 28 + (h_n1 ? 8 : 0));
        d2h += pairs * 28LL + (h_n_good ? pairs * 4LL : 0);
        if ((rc = vsb_pyramid_build(ctx, nullptr, (int64_t)fbytes, c.w, pairs, &t->lay, pyr_prev, st))) break;
        if ((rc = vsb_pyramid_build(ctx, nullptr, (int64_t)fbytes, c.w, pairs, &t->lay, pyr_cur, st))) break;
        if (c.gn.grad_mode == 0 && (rc = vsb_gradient_build(ctx, pyr_prev, pairs, &t->lay, s.gx, s.gy, nullptr, st))) break;
        if ((rc = run_pairs(t, s, pyr_prev, pyr_cur, s.gx, s.gy, s.desc, s.desc2, s.kp, h_n1 ? s.n_feat : nullptr,
                            h_n1 ? s.n_feat2 : nullptr, s.prior, pairs, s.pose, nullptr, st)))
            break;
        VSB_CUDA(ctx, cudaMemcpyAsync(h_pose + (size_t)p0 * 7, s.pose, (size_t)pairs * 28, cudaMemcpyDeviceToHost, st));
        if (h_n_good)
            VSB_CUDA(ctx, cudaMemcpyAsync(h_n_good + p0, s.n_good, pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        VSB_CUDA(ctx, cudaEventRecord(s.done, st));
        p0 += pairs;
    }
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(t->slot[i].stream);      // also on the error path: copies may be in flight
    if (rc) return rc;
    t->host_h2d_bytes = h2d; t->host_d2h_bytes = d2h; t->host_chunks = chunk_idx;
    return VSB_OK;
}

extern "C" int vsb_tracker_set_trace(vsb_tracker_t* t, vsb_gn_trace_t* trace, int32_t* n_trace) {
    if (!t || ((trace == nullptr) != (n_trace == nullptr))) return VSB_ERR_INVALID;
    t->trace = trace; t->n_trace = n_trace;
    return VSB_OK;
}

extern "C" int vsb_tracker_stats(vsb_tracker_t* t, long long out[4]) {
    if (!t || !out) return VSB_ERR_INVALID;
    VSB_CUDA(t->ctx, cudaDeviceSynchronize());
    unsigned long long h[4];
    VSB_CUDA(t->ctx, cudaMemcpy(h, t->stats, sizeof(h), cudaMemcpyDeviceToHost));
    VSB_CUDA(t->ctx, cudaMemset(t->stats, 0, sizeof(h)));
    for (int i = 0; i < 4; i++) out[i] = (long long)h[i];
    return VSB_OK;
}

extern "C" int vsb_tracker_host_traffic(vsb_tracker_t* t, long long out[3]) {
    if (!t || !out) return VSB_ERR_INVALID;
    out[0] = t->host_h2d_bytes; out[1] = t->host_d2h_bytes; out[2] = t->host_chunks;
    return VSB_OK;
}

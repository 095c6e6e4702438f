// common.cuh — shared host-side plumbing of libvislam_b200.so (context, status codes, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vislam_b200.h"

#include <vector>

// pixels of a pyramid level whose candidate points the fused candidate pass merges per distinct pixel (gn_track.cu)
#define VSB_DEDUP_PIX 8192

enum vsb_kernel_id {
    VSB_K_KNN_HAMMING = 0, VSB_K_KNN_UNPACK, VSB_K_MATCH_FILTER, VSB_K_GATHER, VSB_K_PYRAMID, VSB_K_GRADIENT,
    VSB_K_CANDIDATES, VSB_K_GN_SOLVE, VSB_K_KNN_L2, VSB_K_KNN_L2_PREP, VSB_K_GN_PREPARE,
    VSB_K_MATCH_STAGE, VSB_K_WARP, VSB_K_KNN_L2_FINAL, VSB_K_FAST_SCORE, VSB_K_FAST_COMPACT, VSB_K_ORB, VSB_K_COUNT
};

struct vsb_prof_rec { int id; cudaEvent_t a, b; };

struct vsb_ctx {
    int device;
    int sm_count;
    long long launches;
    char last_error[256];
    // scratch owned by the context (grown on demand, freed in vsb_ctx_destroy)
    void* scratch;
    size_t scratch_bytes;
    void* scratch2;          // second area: internals of an entry whose caller already holds `scratch`
    size_t scratch2_bytes;
    // per-STREAM workspaces (the 4-bit kNN's expanded descriptors): the tracker's host entry runs two streams at once, so a
    // workspace that two in-flight launches could share is keyed by the stream it is used on
    void* ws_ptr[8];
    size_t ws_bytes[8];
    cudaStream_t ws_stream[8];
    int ws_n;
    // kernel attributes are per device: remembered per context, not per process
    int attr_knn_tc_done;
    int attr_knn_mx_done;
    int attr_l2_tc_smem;
    unsigned long long* l2_fallback_counter;   // device counter of the last tensor-core L2 kNN call (diagnostics)
    // optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg)
    int prof_on;
    std::vector<vsb_prof_rec> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[VSB_K_COUNT];
    long long prof_n[VSB_K_COUNT];
    // tuning knobs (vsb_ctx_option)
    int knn_impl;     // Hamming kNN: 0 = POPC kernel (INT pipe), 1 = tcgen05 tensor-core kernel, 32-bit epilogue,
                      //              2 = tcgen05 kernel with the packed 16x2 epilogue, 3 = 4-bit operands (kind::mxf4, knn_mx.cu), 4 = the same with descriptors pre-expanded once per call,
                      //              5 = persistent 4-bit kernel with bulk-copied tiles, 6 (default) = 5 from 768 descriptors per set up, else 2
    int gn_threads;   // threads per frame pair of the GN solver: 64 / 128 / 256 / 512 / 1024, 0 = chosen from the batch size
    int knn_l2_impl;  // float kNN: 0 = exact FP64 kernel, 1 = tensor-core GEMM + exact re-check (dim <= 64, dim % 8 == 0)
    int gn_variant;   // GN solver register/unroll variant (tuning experiments; 0 = default)
    int gn_impl;      // tracker GN kernel: 1 (default) = gn_track.cu (8-byte point records + back-projection tables, staged coarse levels) for the reference modes, 0 = always gn_solve.cu
    int gn_stage_bytes;   // gn_track.cu: shared-memory budget for the staged current-image level (0 = gather from global memory)
    int gn_dedup;     // gn_track.cu: 1 (default) = candidate points of small levels are merged per distinct pixel (multiplicity)
    int gn_cluster;   // gn_track.cu: 1 (default) = a batch of fewer pairs than 0.7 x the SMs gives each pair a cluster of 2 / 4 / 8 blocks (by batch size); 0 = never; 2 / 4 / 8 = that size always
    int gn_cluster_threads;   // threads per block of the cluster kernel: 0 (default) = by batch size, 256, 512
    int gn_tail;      // gn_track.cu: 1 (default) = pairs of the last partial wave run with more threads each, 0 = one launch
    int orb_scratch_mb;   // ORB: scratch budget of one workspace in MB (default 32768: 2000 752x480 frames with one block of scratch per pyramid level)
    int fast_impl;    // FAST compaction: 0 (default) = 16 pixels per lane, suppression once (bit per pixel); 1 = 4 pixels per lane, suppression in both passes
    int orb_lp;       // ORB: 1 (default) = small batches run the pyramid levels on separate streams, 0 = one stream always
    cudaStream_t orb_stream[8];   // per-level streams / events of that form, created on first use
    cudaEvent_t orb_ev_ready[8], orb_ev_done[8];
    int orb_lp_ready;
    int orb_impl;     // ORB tuning switch, bit mask of the PREVIOUS forms kept for comparison (default 0): 1 = per-pixel resize, 2 = per-warp sin / cos in the descriptor kernel
    int pyr_impl;     // pyramid: 0 = generic shared-memory tile kernel, 1 = register-blocked kernel when w, h are multiples of 16
};

int vsb_prof_begin(vsb_ctx* ctx, int id, cudaStream_t st);
void vsb_prof_end(vsb_ctx* ctx, int slot, cudaStream_t st);

struct ProfScope {
    vsb_ctx* c; int slot; cudaStream_t st;
    ProfScope(vsb_ctx* ctx, int id, cudaStream_t s) : c(ctx), slot(-1), st(s) { if (c && c->prof_on) slot = vsb_prof_begin(c, id, st); }
    ~ProfScope() { if (slot >= 0) vsb_prof_end(c, slot, st); }
};

static inline int vsb_cuda_fail(vsb_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) snprintf(ctx->last_error, sizeof(ctx->last_error), "%s: %s", what, cudaGetErrorString(e));
    return VSB_ERR_CUDA;
}

#define VSB_CUDA(ctx, call)                                              \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return vsb_cuda_fail((ctx), e__, #call); \
    } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches truthfully
#define VSB_LAUNCHED(ctx)                                                          \
    do {                                                                           \
        (ctx)->launches++;                                                         \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) return vsb_cuda_fail((ctx), e__, "kernel launch"); \
    } while (0)

int vsb_scratch_reserve(vsb_ctx* ctx, size_t bytes, void** out);
int vsb_scratch2_reserve(vsb_ctx* ctx, size_t bytes, void** out);
int vsb_stream_ws_reserve(vsb_ctx* ctx, cudaStream_t st, size_t bytes, void** out);

static inline int vsb_div_up(int a, int b) { return (a + b - 1) / b; }

// knn_l2.cu — float-descriptor kNN (BFMatcher NORM_L2, reference src/Matcher.cpp:55).  Placeholder entry:
// the tensor-core distance GEMM + FP32 re-check lands in a later milestone of this round.
#include "common.cuh"

extern "C" int vsb_knn2_l2(vsb_ctx_t* ctx, const float* d1, int n1_max, const int32_t* n1, const float* d2,
                           int n2_max, const int32_t* n2, int dim, int count, int32_t* idx12, float* dist12,
                           int32_t* idx21, float* dist21, void* stream) {
    (void)ctx; (void)d1; (void)n1_max; (void)n1; (void)d2; (void)n2_max; (void)n2; (void)dim; (void)count;
    (void)idx12; (void)dist12; (void)idx21; (void)dist21; (void)stream;
    return VSB_ERR_UNSUPPORTED;
}

// gn_common.cuh — device helpers shared by the two Gauss-Newton kernels (gn_solve.cu: every mode of the public
// vsb_gn_solve entry; gn_track.cu: the tracker's reference-mode kernel): exact small-integer conversions off the
// conversion pipe, the FP64 tensor-core Gram update, and the 6x6 solve as cv::solve(A, b, DECOMP_LU) performs it.
#pragma once
#include "common.cuh"
#include "se3.cuh"

namespace gn {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// exact small-integer conversions on the ALU/FP pipes (keeps the quarter-rate conversion pipe for the rest)
__device__ __forceinline__ float u23_to_float(uint32_t v) {          // 0 <= v < 2^23
    return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f);
}
__device__ __forceinline__ double i32_to_double(int v) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)), 4503601774854144.0);
}

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor cores.  With a == b (lane (g,t) supplies V[g] of point t)
// this accumulates the Gram matrix V V^T of 4 points; products of floats are exact in double.
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- 6x6 solve by one warp -----------------------------------------------------------------------------
// deltaMat = A.inv() * b (VISystem.cpp:1412).  cv::MatExpr never forms the inverse here: inverse-times-matrix is turned
// into cv::solve(A, b, DECOMP_LU) (matop.cpp MatOp_Invert::matmul -> MatOp_Solve), i.e. hal::LU32f on [A | b] with ONE
// right-hand column.  Lane c < 6 owns column c of A, lane 6 owns b; every arithmetic operation is the one LU32f performs
// on that element, in the same order, so the result is bit-identical to the sequential code in the oracle (vso_solve6).
// G is the 8x8 Gram matrix of V = (J0..J5, r*w, r): A = G[0..5][0..5], J^T(r w) = G[a][6], sum r (r w) = G[7][6].
__device__ __forceinline__ void warp_solve6(const double* G, int lane, float delta[6]) {
    const unsigned FULL = 0xffffffffu;
    float v[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        float x = 0.f;
        if (lane < 6) x = (float)G[r * 8 + lane];                 // A = J^T J rounded once to float (:1408)
        else if (lane == 6) x = (float)(-1.0 * G[r * 8 + 6]);     // b = -J^T (r w): gemm alpha = -1, rounded once (:1409)
        v[r] = x;
    }
    const float eps = 1.1920929e-07f * 10;
    bool singular = false;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float col[6];
#pragma unroll
        for (int j = i; j < 6; j++) col[j] = __shfl_sync(FULL, v[j], i);
        int k = i;
        float best = fabsf(col[i]);
#pragma unroll
        for (int j = i + 1; j < 6; j++)
            if (fabsf(col[j]) > best) { best = fabsf(col[j]); k = j; }
        if (best < eps) singular = true;
#pragma unroll
        for (int j = i + 1; j < 6; j++)
            if (k == j) {
                float t = v[i]; v[i] = v[j]; v[j] = t;
                t = col[i]; col[i] = col[j]; col[j] = t;
            }
        const float d = F_DIV(-1.f, col[i]);
#pragma unroll
        for (int j = i + 1; j < 6; j++) {
            const float alpha = F_MUL(col[j], d);
            v[j] = F_ADD(v[j], F_MUL(alpha, v[i]));
        }
    }
    float x[6];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        float s = v[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) {
            const float u = __shfl_sync(FULL, v[i], k);
            s = F_SUB(s, F_MUL(u, x[k]));
        }
        const float diag = __shfl_sync(FULL, v[i], i);
        x[i] = F_DIV(s, diag);
    }
#pragma unroll
    for (int a = 0; a < 6; a++) {
        const float xa = __shfl_sync(FULL, x[a], 6);               // the solution is lane 6's column
        delta[a] = singular ? 0.f : xa;                            // singular => cv::solve zeroes the result => delta = 0
    }
}

}  // namespace gn

// tools/probe_minmax.cu — throughput of the packed min/max candidates for the kNN top-2 epilogue (warp instructions per clock
// per SM): VIMNMX.U16x2 (__vmaxu2), the 3-input form (__vimax3_u16x2), HMNMX2 (__hmax2 on half2), 32-bit IMNMX (max on u32),
// FMNMX (fmaxf).  Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_minmax tools/probe_minmax.cu && /tmp/probe_minmax
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[8];
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    uint32_t x = seed ^ threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __vmaxu2(a[i], x);
            if (OP == 1) a[i] = __vimax3_u16x2(a[i], x, a[(i + 1) & 7]);
            if (OP == 2) { __half2 h = __hmax2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&x)); a[i] = *reinterpret_cast<uint32_t*>(&h); }
            if (OP == 3) a[i] = max(a[i], x);
            if (OP == 4) a[i] = __float_as_uint(fmaxf(__uint_as_float(a[i]), __uint_as_float(x)));
            if (OP == 5) a[i] = __vminu2(a[i], x);
            if (OP == 6) {      // half of the registers on VIMNMX.U16x2, half on HMNMX2: do the two pipes add up?
                if (i & 1) a[i] = __vmaxu2(a[i], x);
                else { __half2 h = __hmax2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&x)); a[i] = *reinterpret_cast<uint32_t*>(&h); }
            }
            if (OP == 7) { __nv_bfloat162 h = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a[i]), *reinterpret_cast<__nv_bfloat162*>(&x)); a[i] = *reinterpret_cast<uint32_t*>(&h); }
            if (OP == 8) {      // VIMNMX.U16x2 + FMNMX
                if (i & 1) a[i] = __vmaxu2(a[i], x);
                else a[i] = __float_as_uint(fmaxf(__uint_as_float(a[i]), __uint_as_float(x)));
            }
            if (OP == 9) { __half2 h = __hmin2(*reinterpret_cast<__half2*>(&a[i]), *reinterpret_cast<__half2*>(&x)); a[i] = *reinterpret_cast<uint32_t*>(&h); }
        }
        x += 0x01010101u;
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;
}

template <int OP>
void run(const char* name) {
    uint32_t* d;
    cudaMalloc(&d, 4);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 4096, blocks = sms * 2;
    k<OP><<<blocks, 1024>>>(d, 7, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 1024>>>(d, 7, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 32 * iters * 8;
    const double per_clk_sm = warp_instr / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms;
    printf("%-22s %8.3f ms  %6.2f warp-instr/clk/SM (%5.1f lanes/clk/SM)\n", name, ms, per_clk_sm, per_clk_sm * 32);
    cudaFree(d);
}

int main() {
    run<0>("VIMNMX.U16x2 max");
    run<5>("VIMNMX.U16x2 min");
    run<1>("VIMNMX3.U16x2");
    run<2>("HMNMX2 (half2 max)");
    run<3>("IMNMX.U32");
    run<4>("FMNMX");
    run<9>("HMNMX2 (half2 min)");
    run<7>("HMNMX2.BF16 max");
    run<6>("VIMNMX.U16x2 + HMNMX2");
    run<8>("VIMNMX.U16x2 + FMNMX");
    return 0;
}

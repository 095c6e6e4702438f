// tools/probe_mxf4.cu — stand-alone probe for the NEXT Hamming kNN design: one 128x128 tile of
// tcgen05.mma kind::mxf4.block_scale with +-1 e2m1 operands expanded from descriptor bits and CONSTANT block scales filled
// into TMEM with tcgen05.st.  Prints how many of the 128x128 FP32 accumulators equal sb * (256 - 2 * hamming).
// Build + run (on the GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_mxf4 tools/probe_mxf4.cu && /tmp/probe_mxf4
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../vi-slam_b200/csrc/umma.cuh"

namespace {
constexpr int ROWS = 128;

__device__ __forceinline__ void mma_mxf4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t accumulate,
                                         uint32_t sfa, uint32_t sfb) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(id), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 8 descriptor bits -> 8 e2m1 nibbles: bit 0 -> +1 (0x2), bit 1 -> -1 (0xA); element i of the byte in nibble i
__device__ __forceinline__ uint32_t expand8(uint32_t v) {
    uint32_t x = (v | (v << 4)) & 0x0F0Fu;          // 0000abcd 0000efgh
    x = (x | (x << 2)) & 0x3333u;                    // 00ab00cd 00ef00gh: one 2-bit selector per output byte
    // selector s = (bit 2i+1, bit 2i): low nibble from bit 2i, high nibble from bit 2i+1
    return __byte_perm(0xAAA22A22u, 0u, x);          // pool bytes: s=0 -> 0x22, s=1 -> 0x2A, s=2 -> 0xA2, s=3 -> 0xAA
}

__global__ void __launch_bounds__(160) probe_kernel(const uint8_t* __restrict__ a_bits, const uint8_t* __restrict__ b_bits,
                                                    float* __restrict__ out, uint32_t sfa_byte, uint32_t sfb_byte,
                                                    uint32_t idesc_v, long long* __restrict__ clocks, int rep) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { umma::mbar_init(umma::smem_u32(bar), 1); umma::mbar_init(umma::smem_u32(bar + 1), 1); umma::fence_mbar_init(); }
    if (warp == 4) umma::tmem_alloc<512>(umma::smem_u32(slot));
    if (tid < ROWS) {
        const int p = tid;
        for (int which = 0; which < 2; which++) {
            const uint8_t* src = (which ? b_bits : a_bits) + (size_t)p * 32;
            uint8_t* rowp = (which ? sB : sA) + (p >> 3) * 1024 + (p & 7) * 128;
            for (int c = 0; c < 8; c++) {                     // 16-byte chunk = 32 elements = 4 descriptor bytes
                uint32_t o[4];
                for (int j = 0; j < 4; j++) o[j] = expand8(src[4 * c + j]);
                *reinterpret_cast<uint4*>(rowp + ((c ^ (p & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *slot;
    if (warp < 4) {
        // constant scale factors: 16 columns each for A (at +256) and B (at +288), every byte the same
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 16; c += 4) {
            tmem_st4(lane_base + 256 + c, sfa_byte * 0x01010101u);
            tmem_st4(lane_base + 288 + c, sfb_byte * 0x01010101u);
        }
        tmem_wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (warp == 4 && lane == 0) {
        const uint32_t aA = umma::smem_u32(sA), aB = umma::smem_u32(sB);
        for (int k = 0; k < 4; k++) {                          // K = 64 elements = 32 bytes per instruction
            const uint64_t da = umma::smem_desc(aA + k * 32, 16, 1024, umma::LAYOUT_SW128);
            const uint64_t db = umma::smem_desc(aB + k * 32, 16, 1024, umma::LAYOUT_SW128);
            mma_mxf4(tmem_base, da, db, idesc_v, k > 0 ? 1u : 0u, tmem_base + 256, tmem_base + 288);
        }
        umma::commit(umma::smem_u32(bar));
        if (clocks && rep > 0) {
            // issue cost vs execution time of back-to-back UMMAs (alternating between two accumulator regions)
            umma::mbar_wait(umma::smem_u32(bar), 0);
            const long long t0 = clock64();
            for (int r = 0; r < rep; r++)
                for (int k = 0; k < 4; k++) {
                    const uint64_t da = umma::smem_desc(aA + k * 32, 16, 1024, umma::LAYOUT_SW128);
                    const uint64_t db = umma::smem_desc(aB + k * 32, 16, 1024, umma::LAYOUT_SW128);
                    mma_mxf4(tmem_base + 128 * (r & 1), da, db, idesc_v, 1u, tmem_base + 256, tmem_base + 288);
                }
            const long long t1 = clock64();
            umma::commit(umma::smem_u32(bar + 1));
            umma::mbar_wait(umma::smem_u32(bar + 1), 0);
            const long long t2 = clock64();
            clocks[0] = t1 - t0; clocks[1] = t2 - t0;
        }
    }
    if (warp < 4) {
        umma::mbar_wait(umma::smem_u32(bar), 0);
        umma::fence_after_sync();
        const int row = warp * 32 + lane;
        for (int chunk = 0; chunk < 4; chunk++) {
            uint32_t v[32];
            umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + chunk * 32, v);
            umma::tmem_wait_ld();
            for (int i = 0; i < 32; i++) out[(size_t)row * 128 + chunk * 32 + i] = __uint_as_float(v[i]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4) { umma::fence_after_sync(); umma::tmem_dealloc<512>(tmem_base); }
}
}  // namespace

int main() {
    std::vector<uint8_t> a(ROWS * 32), b(ROWS * 32);
    srand(1);
    for (auto& v : a) v = (uint8_t)(rand() & 255);
    for (auto& v : b) v = (uint8_t)(rand() & 255);
    for (int i = 0; i < 32; i++) { b[i] = a[i]; b[32 + i] = (uint8_t)~a[32 + i]; }      // row 0 == row 0, row 1 == ~row 1
    uint8_t *da, *db;
    float* dout;
    cudaMalloc(&da, a.size()); cudaMalloc(&db, b.size()); cudaMalloc(&dout, ROWS * 128 * 4);
    cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
    long long* dclk;
    cudaMalloc(&dclk, 16);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34000);
    // instruction descriptor (block scaled): a/b format E2M1 = 1 at [7,10) / [10,13), N >> 3 at [17,23), scale format UE8M0 = 1
    // at bit 23, M >> 4 at [24,29), K64 (bit 31 = 0)
    const uint32_t idesc = (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    for (int trial = 0; trial < 2; trial++) {
        const uint32_t sfa = 0x7F, sfb = trial ? 0x85 : 0x7F;       // 1 x 1, then 1 x 64
        cudaMemset(dout, 0, ROWS * 128 * 4);
        probe_kernel<<<1, 160, 34000>>>(da, db, dout, sfa, sfb, idesc, trial ? dclk : nullptr, 256);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> out(ROWS * 128);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        const float s = trial ? 64.f : 1.f;
        int ok = 0;
        for (int r = 0; r < ROWS; r++)
            for (int c = 0; c < 128; c++) {
                int h = 0;
                for (int k = 0; k < 32; k++) h += __builtin_popcount((unsigned)(a[r * 32 + k] ^ b[c * 32 + k]));
                ok += out[r * 128 + c] == s * (256 - 2 * h);
            }
        printf("trial %d (scale %g): %d / %d accumulators exact; out[0][0..3] = %g %g %g %g (expect %g ...), out[1][1] = %g (expect %g)\n",
               trial, s, ok, ROWS * 128, out[0], out[1], out[2], out[3], s * 256, out[129], -s * 256);
    }
    long long clk[2];
    cudaMemcpy(clk, dclk, 16, cudaMemcpyDeviceToHost);
    printf("1024 back-to-back UMMAs (M128 N128 K64 mxf4) from one thread: issue %.1f cycles each, issue + execution %.1f cycles each\n",
           clk[0] / 1024.0, clk[1] / 1024.0);
    return 0;
}

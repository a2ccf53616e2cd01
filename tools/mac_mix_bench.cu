// Compute-only micro-benchmark of the MAC inner loop (registers only, no memory): what rate does the SM sustain for
// this exact instruction mix, as a function of resident warps?  Variants: 0 = Karatsuba (production Fq3Acc::mac),
// 1 = same but carries dropped (no IADD3.X / predicates), 2 = products only into 64-bit (no carry-out at all).
#include <cstdio>
#include <cuda_runtime.h>
#include "../latticeum_b200/csrc/goldilocks.cuh"
using gl::u64; using gl::u32;

template <int VAR>
__device__ __forceinline__ void mac_var(gl::Fq3Acc &A, u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2, u64 b01, u64 b02, u64 b12) {
    if constexpr (VAR == 0) {
        A.mac(a0, a1, a2, b0, b1, b2, b01, b02, b12);
    } else {
        auto m = [](gl::WideAcc &W, u64 a, u64 b) {
            u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
            W.c0.acc += (u64)al * bl; W.c1.acc += (u64)al * bh; W.c1.acc += (u64)ah * bl; W.c2.acc += (u64)ah * bh;
        };
        u64 s01 = a0 + a1, s02 = a0 + a2, s12 = a1 + a2;
        m(A.p0, a0, b0); m(A.p1, a1, b1); m(A.p2, a2, b2); m(A.p01, s01, b01); m(A.p02, s02, b02); m(A.p12, s12, b12);
    }
}

template <int VAR>
__global__ void __launch_bounds__(256) k(u64 *out, int iters, u64 seed) {
    gl::Fq3Acc A; A.clear();
    u64 a0 = seed + threadIdx.x, a1 = seed * 3 + blockIdx.x, a2 = seed * 7 + 11;
    u64 b0 = seed ^ 0x1234567, b1 = b0 * 5, b2 = b0 * 9, b01 = b0 + b1, b02 = b0 + b2, b12 = b1 + b2;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mac_var<VAR>(A, a0, a1, a2, b0, b1, b2, b01, b02, b12);
            a0 += b1; a1 ^= b2; a2 += 0x9E3779B97F4A7C15ull;  // cheap operand churn (ALU)
        }
    }
    u64 c0, c1, c2; A.finish(c0, c1, c2);
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2;
}

template <int VAR>
void run(int ctas_per_sm, int sms, double clk_hz, u64 *out) {
    const int iters = 2000;
    int grid = sms * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<VAR><<<grid, 256>>>(out, iters, 12345); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); k<VAR><<<grid, 256>>>(out, iters, 12345 + r); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double macs = (double)grid * 8 * iters * 4;            // warp-level Fq3 MACs ("warp-columns")
    double cyc_per_warpcol_per_smsp = best * 1e-3 * clk_hz / (macs / (sms * 4.0));
    printf("var=%d warps/SM=%2d  %.3f ms  cycles per warp-column per SMSP = %.1f  (24 IMAD.WIDE each -> %.2f cyc/WIDE)\n",
           VAR, ctas_per_sm * 8, best, cyc_per_warpcol_per_smsp, cyc_per_warpcol_per_smsp / 24.0);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u64 *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    for (int c : {1, 2, 4}) run<0>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2, 4}) run<1>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    return 0;
}

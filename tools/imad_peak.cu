// Micro-benchmark: peak rate of the integer multiply pipe on sm_100a, measured with dependency-chained
// IMAD.WIDE.U32 (32x32+64 -> 64 with carry-out) -- the instruction the Ajtai MAC is made of -- and with plain
// 32-bit IMAD for comparison.  The IMAD peak is not in MEASURED_PEAKS.json; this tool measures it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_peak tools/imad_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int CHAINS, bool CARRY>
__global__ void __launch_bounds__(256) wide_kernel(u32 *out, int iters, u32 x0, u32 y0) {
    u32 lo[CHAINS], hi[CHAINS], ov[CHAINS];
    u32 x = x0 + threadIdx.x, y = y0 + blockIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { lo[c] = c; hi[c] = c * 3; ov[c] = 0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (CARRY)
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                             : "+r"(lo[c]), "+r"(hi[c]), "+r"(ov[c]) : "r"(x), "r"(y));
            else
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
        }
    }
    u32 r = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) r ^= lo[c] ^ hi[c] ^ ov[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) imad32_kernel(u32 *out, int iters, u32 x0, u32 y0) {
    u32 a[CHAINS];
    u32 x = x0 + threadIdx.x, y = y0 + blockIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(x), "r"(y));
    }
    u32 r = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) r ^= a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class F>
double time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u32 *out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    const int iters = 4096;
    const int grid = sms * 8;  // 8 CTAs x 256 thr = 64 warps per SM
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f", p.name, sms, clk_khz / 1000.0);
    {
        constexpr int CH = 8;
        double ms = time_ms([&] { wide_kernel<CH, true><<<grid, 256>>>(out, iters, 12345u, 777u); });
        double ops = (double)grid * 256 * iters * CH;
        printf(", \"imad_wide_carry_Tops\": %.3f, \"imad_wide_carry_per_clk_per_sm\": %.2f", ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    {
        constexpr int CH = 8;
        double ms = time_ms([&] { wide_kernel<CH, false><<<grid, 256>>>(out, iters, 12345u, 777u); });
        double ops = (double)grid * 256 * iters * CH;
        printf(", \"imad_wide_Tops\": %.3f, \"imad_wide_per_clk_per_sm\": %.2f", ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    {
        constexpr int CH = 8;
        double ms = time_ms([&] { imad32_kernel<CH><<<grid, 256>>>(out, iters, 12345u, 777u); });
        double ops = (double)grid * 256 * iters * CH;
        printf(", \"imad32_Tops\": %.3f, \"imad32_per_clk_per_sm\": %.2f", ops / ms / 1e9,
               ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    printf("}\n");
    cudaFree(out);
    return 0;
}

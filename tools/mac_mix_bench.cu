// Compute-only micro-benchmark of the MAC inner loop: what rate does an SM sustain for one Fq3 multiply-accumulate
// per thread ("warp-column"), as a function of the pre-addition scheme, of where the operands come from, and of
// resident warps?  All operands are per-thread (vector) values, so nothing migrates to the uniform datapath.
//   SUMS 0: no matrix-side pre-additions (sums passed in): the 24 IMAD.WIDE.U32 + carries alone
//   SUMS 1: production gl::Fq3Acc::mac (exact 65-bit sums, carry added into the 2^64 column)
//   SUMS 2: sums folded to 64 bits with a borrow-mask carry chain (add.cc / addc.cc / subc / add.cc / addc)
//   SMEM 0: operands live in registers and are churned every iteration
//   SMEM 1: operands are read from a shared-memory tile laid out like the production kernel's stage
//           ([jj][c][row][slot] matrix tile, [jj][slot][6] extended witness), no TMA, no barriers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mac_mix_bench tools/mac_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../latticeum_b200/csrc/goldilocks.cuh"
using gl::u64; using gl::u32;

// (a + b) folded to 64 bits: on carry add 2^64 = 2^32 - 1 (mod q); five carry-chain instructions, all ALU pipe
__device__ __forceinline__ u64 add_fold(u64 a, u64 b) {
    u64 s;
    asm("{\n\t.reg .u32 al, ah, bl, bh, c;\n\t.reg .pred p;\n\tmov.b64 {al, ah}, %1;\n\tmov.b64 {bl, bh}, %2;\n\t"
        "add.cc.u32 al, al, bl;\n\taddc.cc.u32 ah, ah, bh;\n\taddc.u32 c, 0, 0;\n\tsetp.ne.u32 p, c, 0;\n\t"
        "@p add.cc.u32 al, al, 0xffffffff;\n\t@p addc.u32 ah, ah, 0;\n\tmov.b64 %0, {al, ah};\n\t}"
        : "=l"(s) : "l"(a), "l"(b));
    return s;
}

template <int SUMS>
__device__ __forceinline__ void mac_var(gl::Fq3Acc &A, u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2, u64 b01, u64 b02, u64 b12) {
    if (SUMS == 1) {
        A.mac(a0, a1, a2, b0, b1, b2, b01, b02, b12);
    } else if (SUMS == 2) {
        u64 s01 = add_fold(a0, a1), s02 = add_fold(a0, a2), s12 = add_fold(a1, a2);
        A.p0.mac(a0, b0); A.p1.mac(a1, b1); A.p2.mac(a2, b2);
        A.p01.mac(s01, b01); A.p02.mac(s02, b02); A.p12.mac(s12, b12);
    } else {
        A.p0.mac(a0, b0); A.p1.mac(a1, b1); A.p2.mac(a2, b2);
        A.p01.mac(a0 ^ a1, b01); A.p02.mac(a0 ^ a2, b02); A.p12.mac(a1 ^ a2, b12);   // 3 LOP3 pairs stand in for "free" sums
    }
}

constexpr int TJ = 8, ROWS = 32;
template <int SUMS, int SMEM>
__global__ void __launch_bounds__(256, 2) k(u64 *out, int iters, u64 seed) {
    extern __shared__ __align__(16) u64 sm[];
    gl::Fq3Acc A; A.clear();
    const u64 t = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + blockIdx.x;
    if (SMEM) {
        constexpr int STAGE = TJ * 3 * ROWS * 8 + TJ * 48;   // u64 per stage: [jj][c][row][slot] tile + [jj][slot][6] witness
        for (int i = threadIdx.x; i < 2 * STAGE; i += blockDim.x) sm[i] = (seed + i) * 0x9E3779B97F4A7C15ull % gl::Q;
        __syncthreads();
        const int row = threadIdx.x >> 3, slot = threadIdx.x & 7;
        for (int i = 0; i < iters; ++i) {
            const u64 *At = sm + (i & 1) * STAGE + row * 8 + slot;
            const ulonglong2 *Ft = reinterpret_cast<const ulonglong2 *>(sm + (i & 1) * STAGE + TJ * 3 * ROWS * 8) + slot * 3;
#pragma unroll
            for (int jj = 0; jj < TJ; ++jj) {
                const u64 *a = At + jj * 3 * ROWS * 8;
                const ulonglong2 *f = Ft + jj * 24;
                ulonglong2 x = f[0], y = f[1], z = f[2];
                mac_var<SUMS>(A, a[0], a[ROWS * 8], a[2 * ROWS * 8], x.x, x.y, y.x, y.y, z.x, z.y);
            }
        }
    } else {
        u64 a0 = seed + t, a1 = seed * 3 + t * 5, a2 = seed * 7 + t * 11;
        u64 b0 = (seed ^ 0x1234567) + t * 13, b1 = b0 * 5 + t, b2 = b0 * 9 + t, b01 = b0 + b1, b02 = b0 + b2, b12 = b1 + b2;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < TJ; ++q) {
                mac_var<SUMS>(A, a0, a1, a2, b0, b1, b2, b01, b02, b12);
                a0 += b1; a1 ^= b2; a2 += 0x9E3779B97F4A7C15ull; b0 ^= a1; b01 += a2;  // per-thread operand churn
            }
        }
    }
    u64 c0, c1, c2; A.finish(c0, c1, c2);
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2;
}

template <int SUMS, int SMEM>
void run(int ctas_per_sm, int sms, double clk_hz, u64 *out) {
    const int iters = 1000;
    int grid = sms * ctas_per_sm;
    size_t smem = SMEM ? (size_t)2 * (TJ * 3 * ROWS * 8 + TJ * 48) * 8 : 0;
    if (SMEM) cudaFuncSetAttribute(k<SUMS, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<SUMS, SMEM><<<grid, 256, smem>>>(out, iters, 12345); 
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); k<SUMS, SMEM><<<grid, 256, smem>>>(out, iters, 12345 + r); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double macs = (double)grid * 8 * iters * TJ;            // warp-level Fq3 MACs ("warp-columns")
    double cyc = best * 1e-3 * clk_hz / (macs / (sms * 4.0));
    printf("sums=%d smem=%d warps/SM=%2d  %.3f ms  cycles per warp-column per SMSP = %.1f\n", SUMS, SMEM, ctas_per_sm * 8, best, cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u64 *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    for (int c : {1, 2}) run<0, 0>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2}) run<1, 0>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2}) run<2, 0>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2}) run<0, 1>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2}) run<1, 1>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2}) run<2, 1>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    return 0;
}

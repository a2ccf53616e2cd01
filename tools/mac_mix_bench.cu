// Compute-only micro-benchmark of the MAC inner loop: what rate does an SM sustain for one Fq3 multiply-accumulate
// per thread ("warp-column"), as a function of the carry-handling scheme and of resident warps?  All operands are
// per-thread (vector) values and change every iteration, so nothing migrates to the uniform datapath.
//   VAR 0: production gl::Fq3Acc::mac (IMAD.WIDE.U32 with carry-out predicate + paired IADD3.X)
//   VAR 1: non-accumulating IMAD.WIDE.U32, products added in pairs with 3-input IADD3 / IADD3.X chains
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mac_mix_bench tools/mac_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../latticeum_b200/csrc/goldilocks.cuh"
using gl::u64; using gl::u32;

// ---- VAR 1 building blocks -----------------------------------------------------------------------------------------
struct Col2 { u32 lo, hi, ov; };
// acc(96) += p + q  (two 64-bit products), all in one carry chain: lo = lo + p.lo + q.lo, hi = hi + p.hi + q.hi + c, ov += c
__device__ __forceinline__ void add2(Col2 &c, u64 p, u64 q) {
    asm("{\n\t.reg .u32 pl, ph, ql, qh;\n\tmov.b64 {pl, ph}, %3;\n\tmov.b64 {ql, qh}, %4;\n\t"
        "add.cc.u32 %0, %0, pl;\n\taddc.cc.u32 %1, %1, ph;\n\taddc.u32 %2, %2, 0;\n\t"
        "add.cc.u32 %0, %0, ql;\n\taddc.cc.u32 %1, %1, qh;\n\taddc.u32 %2, %2, 0;\n\t}"
        : "+r"(c.lo), "+r"(c.hi), "+r"(c.ov) : "l"(p), "l"(q));
}
__device__ __forceinline__ void add1(Col2 &c, u64 p) {
    asm("{\n\t.reg .u32 pl, ph;\n\tmov.b64 {pl, ph}, %3;\n\t"
        "add.cc.u32 %0, %0, pl;\n\taddc.cc.u32 %1, %1, ph;\n\taddc.u32 %2, %2, 0;\n\t}"
        : "+r"(c.lo), "+r"(c.hi), "+r"(c.ov) : "l"(p));
}
struct Wide2 {
    Col2 c0, c1, c2;
    __device__ __forceinline__ void clear() { c0 = c1 = c2 = Col2{0, 0, 0}; }
    __device__ __forceinline__ void mac(u64 a, u64 b) {
        u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
        u64 ll = (u64)al * bl, lh = (u64)al * bh, hl = (u64)ah * bl, hh = (u64)ah * bh;
        add1(c0, ll); add2(c1, lh, hl); add1(c2, hh);
    }
};

template <int VAR> struct Acc;
template <> struct Acc<0> {
    gl::Fq3Acc A;
    __device__ void clear() { A.clear(); }
    __device__ __forceinline__ void mac(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2, u64 b01, u64 b02, u64 b12) {
        A.mac(a0, a1, a2, b0, b1, b2, b01, b02, b12);
    }
    __device__ u64 fold() { u64 c0, c1, c2; A.finish(c0, c1, c2); return c0 ^ c1 ^ c2; }
};
template <> struct Acc<1> {
    Wide2 p0, p1, p2, p01, p02, p12;
    __device__ void clear() { p0.clear(); p1.clear(); p2.clear(); p01.clear(); p02.clear(); p12.clear(); }
    __device__ __forceinline__ void mac(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2, u64 b01, u64 b02, u64 b12) {
        u64 s01, s02, s12; u32 k01, k02, k12;
        gl::add65(a0, a1, s01, k01); gl::add65(a0, a2, s02, k02); gl::add65(a1, a2, s12, k12);
        p0.mac(a0, b0); p1.mac(a1, b1); p2.mac(a2, b2);
        p01.mac(s01, b01); add1(p01.c2, k01 ? b01 : 0ull);
        p02.mac(s02, b02); add1(p02.c2, k02 ? b02 : 0ull);
        p12.mac(s12, b12); add1(p12.c2, k12 ? b12 : 0ull);
    }
    __device__ u64 fold() {
        u64 r = 0;
        Wide2 *w[6] = {&p0, &p1, &p2, &p01, &p02, &p12};
        for (int i = 0; i < 6; ++i) r ^= ((u64)w[i]->c0.hi << 32 | w[i]->c0.lo) ^ w[i]->c0.ov ^ ((u64)w[i]->c1.hi << 32 | w[i]->c1.lo) ^ w[i]->c1.ov ^ ((u64)w[i]->c2.hi << 32 | w[i]->c2.lo) ^ w[i]->c2.ov;
        return r;
    }
};

template <int VAR>
__global__ void __launch_bounds__(256) k(u64 *out, int iters, u64 seed) {
    Acc<VAR> A; A.clear();
    const u64 t = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + blockIdx.x;
    u64 a0 = seed + t, a1 = seed * 3 + t * 5, a2 = seed * 7 + t * 11;
    u64 b0 = (seed ^ 0x1234567) + t * 13, b1 = b0 * 5 + t, b2 = b0 * 9 + t, b01 = b0 + b1, b02 = b0 + b2, b12 = b1 + b2;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            A.mac(a0, a1, a2, b0, b1, b2, b01, b02, b12);
            a0 += b1; a1 ^= b2; a2 += 0x9E3779B97F4A7C15ull; b0 ^= a1; b01 += a2;  // per-thread operand churn
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = A.fold();
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
    double cyc = best * 1e-3 * clk_hz / (macs / (sms * 4.0));
    printf("var=%d warps/SM=%2d  %.3f ms  cycles per warp-column per SMSP = %.1f  (24 IMAD.WIDE each -> %.2f cyc/WIDE)\n",
           VAR, ctas_per_sm * 8, best, cyc, cyc / 24.0);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u64 *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    for (int c : {1, 2, 4}) run<0>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    for (int c : {1, 2, 4}) run<1>(c, p.multiProcessorCount, clk_khz * 1e3, out);
    return 0;
}

// Batched ring transforms and gadget decompositions for the Goldilocks ring (d = 24), sm_100a.
//
// One thread owns one ring element (24 x u64 in registers; the three CRT layers have strides 12/6/3, so no
// shuffles or shared-memory butterflies are needed).  Global traffic is staged through shared memory so that
// every global load/store is a contiguous, fully-used run of bytes per warp (an element is 192 B = 1.5 lines);
// rows in shared memory are padded to 25 words so that the per-thread row reads are conflict-free.
// These kernels are HBM/LSU-bound: 192 B in + 192 B out per element, no general multiplies in the forward
// direction, 12 in the inverse.
//
// Reference functions replaced (paths relative to /root/reference/latticeum/crates/):
//   crt_kernel<false>      CRT::elementwise_crt           stark-rings/crates/ring/src/cyclotomic_ring/crt.rs:10-25
//   crt_kernel<true>       ICRT::elementwise_icrt         .../cyclotomic_ring/crt.rs:34-49
//   icrt_decompose_kernel  Witness::from_w_ccs, first two steps (iCRT, gadget_decompose(B, L))
//                          latticefold/src/arith.rs:232-235; .../ring/src/balanced_decomposition/mod.rs:163-175
//   crt_small_kernel       Witness::from_w_ccs third step (CRT of the limbs)           latticefold/src/arith.rs:238
//   planes_kernel          decompose_B_vec_into_k_vec + Witness::from_f_coeff's CRT
//                          latticefold/src/nifs/decomposition/utils.rs:45-49; latticefold/src/arith.rs:327
#include "kernels.h"
#include "ring24.cuh"

namespace lat {
using gl::u32;

constexpr int EPB = 128;    // ring elements per block (= threads per block)
constexpr int PITCH = 25;   // u64 per staged element row (24 + 1 pad)

// Cooperative, coalesced copy of up to EPB elements between global memory and the padded tile.
__device__ __forceinline__ void stage_in(const u64 *__restrict__ g, u64 e0, u64 count, u64 *s) {
    u32 nelem = (u32)min((u64)EPB, count - e0);
    u32 nwords = nelem * ring::D;
    const u64 *src = g + e0 * ring::D;
    for (u32 i = threadIdx.x; i < nwords; i += EPB) s[(i / ring::D) * PITCH + (i % ring::D)] = src[i];
}
__device__ __forceinline__ void stage_out(u64 *__restrict__ g, u64 e0, u64 count, const u64 *s) {
    u32 nelem = (u32)min((u64)EPB, count - e0);
    u32 nwords = nelem * ring::D;
    u64 *dst = g + e0 * ring::D;
    for (u32 i = threadIdx.x; i < nwords; i += EPB) dst[i] = s[(i / ring::D) * PITCH + (i % ring::D)];
}
__device__ __forceinline__ void row_store(u64 *s, const u64 (&c)[ring::D]);
// Extended witness layout for the MAC kernel (kernels.h FX_WORDS): per slot (f0, f1, f2, f0+f1, f0+f2, f1+f2).
// Staged in two halves of 4 slots (24 words each) through the same padded tile.
__device__ __forceinline__ void row_store_fx_half(u64 *s, const u64 (&c)[ring::D], int half) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sl = half * 4 + q;
        u64 f0 = c[3 * sl], f1 = c[3 * sl + 1], f2 = c[3 * sl + 2];
        u64 *r = s + threadIdx.x * PITCH + q * 6;
        r[0] = f0; r[1] = f1; r[2] = f2;
        r[3] = gl::add(f0, f1); r[4] = gl::add(f0, f2); r[5] = gl::add(f1, f2);
    }
}
__device__ __forceinline__ void stage_out_fx_half(u64 *__restrict__ fx, u64 e0, u64 count, const u64 *s, int half) {
    u32 nelem = (u32)min((u64)EPB, count - e0);
    u32 nwords = nelem * ring::D;
    u64 *dst = fx + e0 * FX_WORDS + half * ring::D;
    for (u32 i = threadIdx.x; i < nwords; i += EPB) dst[(i / ring::D) * FX_WORDS + (i % ring::D)] = s[(i / ring::D) * PITCH + (i % ring::D)];
}
// block-wide: write the CRT-form element held by each active thread to out (plain) and/or fx (extended)
__device__ __forceinline__ void emit_element(u64 *s, const u64 (&c)[ring::D], bool active, u64 e0, u64 count,
                                             u64 *__restrict__ out, u64 *__restrict__ fx) {
    if (out) {
        if (active) row_store(s, c);
        __syncthreads();
        stage_out(out, e0, count, s);
        __syncthreads();
    }
    if (fx) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            if (active) row_store_fx_half(s, c, half);
            __syncthreads();
            stage_out_fx_half(fx, e0, count, s, half);
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void row_load(const u64 *s, u64 (&c)[ring::D]) {
#pragma unroll
    for (int k = 0; k < ring::D; ++k) c[k] = s[threadIdx.x * PITCH + k];
}
__device__ __forceinline__ void row_store(u64 *s, const u64 (&c)[ring::D]) {
#pragma unroll
    for (int k = 0; k < ring::D; ++k) s[threadIdx.x * PITCH + k] = c[k];
}

template <bool INVERSE>
__global__ void __launch_bounds__(EPB) crt_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 count) {
    __shared__ u64 s[EPB * PITCH];
    u64 e0 = (u64)blockIdx.x * EPB;
    stage_in(in, e0, count, s);
    __syncthreads();
    if (e0 + threadIdx.x < count) {
        u64 c[ring::D];
        row_load(s, c);
        if constexpr (INVERSE) ring::icrt24(c); else ring::crt24(c);
        row_store(s, c);
    }
    __syncthreads();
    stage_out(out, e0, count, s);
}

void launch_crt(const u64 *in, u64 *out, u64 count, cudaStream_t stream) {
    if (!count) return;
    crt_kernel<false><<<(unsigned)((count + EPB - 1) / EPB), EPB, 0, stream>>>(in, out, count);
}
void launch_icrt(const u64 *in, u64 *out, u64 count, cudaStream_t stream) {
    if (!count) return;
    crt_kernel<true><<<(unsigned)((count + EPB - 1) / EPB), EPB, 0, stream>>>(in, out, count);
}

// ---- iCRT + base-2^log2b balanced digits -------------------------------------------------------------------
// Dynamic shared memory: max(EPB*PITCH*8, EPB*L*24*2) bytes; the digit tile reuses the input tile.
template <bool MONT>
__global__ void __launch_bounds__(EPB)
icrt_decompose_kernel(const u64 *__restrict__ w, u64 w_len, int log2b, int L, bool in_coeff,
                      int16_t *__restrict__ f16, u64 *__restrict__ f_coeff, int *__restrict__ flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s = reinterpret_cast<u64 *>(smem_raw);
    int16_t *s16 = reinterpret_cast<int16_t *>(smem_raw);
    u64 e0 = (u64)blockIdx.x * EPB;
    u64 e = e0 + threadIdx.x;
    bool active = e < w_len;
    stage_in(w, e0, w_len, s);
    __syncthreads();
    u64 c[ring::D];
    if (active) {
        row_load(s, c);
        if (!in_coeff) ring::icrt24(c);
        if constexpr (MONT) {
#pragma unroll
            for (int k = 0; k < ring::D; ++k) c[k] = gl::from_mont(c[k]);
        }
    }
    __syncthreads();  // everyone holds its element in registers; the tile can be reused for digits
    if (active) {
        const u64 B = 1ull << log2b, half = B >> 1;
        bool overflow = false;
#pragma unroll
        for (int k = 0; k < ring::D; ++k) {
            bool negative;
            u64 m;
            ring::signed_rep(c[k], negative, m);  // fq_convertible.rs:22-34
            for (int l = 0; l < L; ++l) {         // balanced_decomposition/mod.rs:76-97 on the magnitude
                u64 rem = m & (B - 1);
                m >>= log2b;
                int dg = (int)rem;
                if (rem > half) {                 // |rem| == b/2 is kept (mod.rs:79)
                    dg -= (int)B;
                    m += 1;
                }
                if (negative) dg = -dg;
                s16[(threadIdx.x * L + l) * ring::D + k] = (int16_t)dg;
                if (f_coeff) f_coeff[((e * L + l) * ring::D) + k] = gl::from_small<MONT>(dg);
            }
            overflow |= (m != 0);                 // the reference would index out of bounds (mod.rs:80)
        }
        if (overflow) atomicOr(flag, 1);
    }
    __syncthreads();
    // coalesced copy-out of the digit tile: nelem * L * 24 int16 = nelem * L * 6 u64 words
    u32 nelem = (u32)min((u64)EPB, w_len - e0);
    u32 nwords = nelem * (u32)L * 6;
    u64 *dst = reinterpret_cast<u64 *>(f16 + e0 * (u64)L * ring::D);
    for (u32 i = threadIdx.x; i < nwords; i += EPB) dst[i] = s[i];
}

void launch_icrt_decompose(const u64 *w, u64 w_len, int log2b, int L, bool mont, bool in_coeff, int16_t *f16,
                           u64 *f_coeff, int *flag, cudaStream_t stream) {
    if (!w_len) return;
    size_t smem = (size_t)EPB * PITCH * 8;
    size_t smem16 = (size_t)EPB * L * ring::D * 2;
    if (smem16 > smem) smem = smem16;
    unsigned grid = (unsigned)((w_len + EPB - 1) / EPB);
    if (smem > 48 * 1024) {  // only the standalone decomposition with many digits gets here (L <= 32 -> 192 KB)
        cudaFuncSetAttribute(icrt_decompose_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(icrt_decompose_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    if (mont)
        icrt_decompose_kernel<true><<<grid, EPB, smem, stream>>>(w, w_len, log2b, L, in_coeff, f16, f_coeff, flag);
    else
        icrt_decompose_kernel<false><<<grid, EPB, smem, stream>>>(w, w_len, log2b, L, in_coeff, f16, f_coeff, flag);
}

// ---- int16 digits -> CRT form ------------------------------------------------------------------------------
__device__ __forceinline__ void load_i16x24(const int16_t *__restrict__ p, int (&d)[ring::D]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);  // 48 B per element, 16-B aligned
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        uint4 x = __ldg(q + v);
        u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d[v * 8 + i * 2] = (int)(int16_t)(w[i] & 0xFFFFu);
            d[v * 8 + i * 2 + 1] = (int)(int16_t)(w[i] >> 16);
        }
    }
}

template <bool MONT>
__global__ void __launch_bounds__(EPB)
crt_small_kernel(const int16_t *__restrict__ f16, u64 count, u64 *__restrict__ out, u64 *__restrict__ fx) {
    __shared__ u64 s[EPB * PITCH];
    u64 e0 = (u64)blockIdx.x * EPB;
    u64 e = e0 + threadIdx.x;
    bool active = e < count;
    u64 c[ring::D];
    if (active) {
        int d[ring::D];
        load_i16x24(f16 + e * ring::D, d);
#pragma unroll
        for (int k = 0; k < ring::D; ++k) c[k] = gl::from_small<MONT>(d[k]);
        ring::crt24(c);
    }
    emit_element(s, c, active, e0, count, out, fx);
}

void launch_crt_small(const int16_t *f16, u64 count, bool mont, u64 *out, u64 *fx, cudaStream_t stream) {
    if (!count) return;
    unsigned grid = (unsigned)((count + EPB - 1) / EPB);
    if (mont) crt_small_kernel<true><<<grid, EPB, 0, stream>>>(f16, count, out, fx);
    else crt_small_kernel<false><<<grid, EPB, 0, stream>>>(f16, count, out, fx);
}

// ---- int16 coefficients -> K sign*bit planes, each CRT'd ---------------------------------------------------
template <bool MONT>
__global__ void __launch_bounds__(EPB)
planes_kernel(const int16_t *__restrict__ f16, u64 n, int K, u64 *__restrict__ planes_f, u64 *__restrict__ planes_fx,
              u64 *__restrict__ planes_coeff) {
    __shared__ u64 s[EPB * PITCH];
    u64 e0 = (u64)blockIdx.x * EPB;
    u64 e = e0 + threadIdx.x;
    bool active = e < n;
    int d[ring::D];
    if (active) load_i16x24(f16 + e * ring::D, d);
    for (int k = 0; k < K; ++k) {
        u64 c[ring::D];
        if (active) {
#pragma unroll
            for (int t = 0; t < ring::D; ++t) {
                int a = d[t] < 0 ? -d[t] : d[t];
                int bit = (a >> k) & 1;
                c[t] = gl::from_small<MONT>(d[t] < 0 ? -bit : bit);  // digit k base 2 = sign * bit_k(|c|)
            }
        }
        if (planes_coeff) emit_element(s, c, active, e0, n, planes_coeff + (u64)k * n * ring::D, nullptr);
        if (planes_f || planes_fx) {
            if (active) ring::crt24(c);
            emit_element(s, c, active, e0, n, planes_f ? planes_f + (u64)k * n * ring::D : nullptr,
                         planes_fx ? planes_fx + (u64)k * n * FX_WORDS : nullptr);
        }
    }
}

void launch_planes(const int16_t *f16, u64 n, int K, bool mont, u64 *planes_f, u64 *planes_fx, u64 *planes_coeff,
                   cudaStream_t stream) {
    if (!n) return;
    unsigned grid = (unsigned)((n + EPB - 1) / EPB);
    if (mont) planes_kernel<true><<<grid, EPB, 0, stream>>>(f16, n, K, planes_f, planes_fx, planes_coeff);
    else planes_kernel<false><<<grid, EPB, 0, stream>>>(f16, n, K, planes_f, planes_fx, planes_coeff);
}

// ---- u64 coefficients -> int16 with range check ---------------------------------------------------------------
template <bool MONT>
__global__ void __launch_bounds__(256)
pack_coeff_kernel(const u64 *__restrict__ f_coeff, u64 nwords, int bits, int16_t *__restrict__ f16,
                  int *__restrict__ flag) {
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= nwords) return;
    u64 v = f_coeff[i];
    if constexpr (MONT) v = gl::from_mont(v);
    else v = gl::reduce128(v, 0);
    bool negative;
    u64 m;
    ring::signed_rep(v, negative, m);
    if (m >> bits) {
        atomicOr(flag, 1);
        m = 0;
    }
    f16[i] = (int16_t)(negative ? -(int)m : (int)m);
}

void launch_pack_coeff(const u64 *f_coeff, u64 count, bool mont, int bits, int16_t *f16, int *flag,
                       cudaStream_t stream) {
    u64 nwords = count * ring::D;
    if (!nwords) return;
    unsigned grid = (unsigned)((nwords + 255) / 256);
    if (mont) pack_coeff_kernel<true><<<grid, 256, 0, stream>>>(f_coeff, nwords, bits, f16, flag);
    else pack_coeff_kernel<false><<<grid, 256, 0, stream>>>(f_coeff, nwords, bits, f16, flag);
}

}  // namespace lat

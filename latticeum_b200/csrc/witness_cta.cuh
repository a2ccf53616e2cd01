// Witness::from_w_ccs for the columns of ONE CTA of the matrix-vector kernel (the fused launch of mac_kernels.cu):
// iCRT -> balanced digits base 2^log2b, L limbs -> CRT of every limb, written as the resident int16 digits and as the
// extended witness layout the tile loop streams.  Same arithmetic as witness_kernel (ring_kernels.cu), organised for a
// block that owns a contiguous column range instead of a grid that owns the whole vector:
//   phase A  eight lanes per w_ccs element (ring8.cuh): inverse transform by warp shuffles, digit loop on the lane's three
//            coefficients, digits into a shared-memory tile (rows in limb-element order);
//   phase B  one thread per limb element: forward transform in Z/(2^96 + 1) (ring96.cuh) straight from the tile, the
//            extended row (384 B) written with 256-bit stores (whole 32-byte sectors: no read-merge in L2).
// Reference: latticefold/src/arith.rs:230-248; stark-rings/crates/ring/src/balanced_decomposition/mod.rs:62-103,163-175.
#pragma once
#include "ring24.cuh"
#include "ring8.cuh"
#include "ring96.cuh"

namespace lat {

// 24 int16 (48 B, 16-B aligned) -> ints; works for global and shared pointers
__device__ __forceinline__ void load_i16x24_cta(const int16_t *p, int (&d)[ring::D]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        uint4 x = q[v];
        gl::u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d[v * 8 + i * 2] = (int)(int16_t)(w[i] & 0xFFFFu);
            d[v * 8 + i * 2 + 1] = (int)(int16_t)(w[i] >> 16);
        }
    }
}

__device__ __forceinline__ void st256(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

constexpr int CTA_WIT_ROWS = 256;  // limb rows per chunk (one per thread in phase B); the tile is CTA_WIT_ROWS x 48 B = 12 KB

// Columns [col0, col1) of the decomposed witness (limb elements), produced by all `nthreads` (= 256) threads of the block.
// tile: CTA_WIT_ROWS x 24 int16 of shared memory.  w: w_len ring elements (device memory, or page-locked host memory
// mapped into the device: the lanes then read it in place over PCIe).
template <bool MONT>
__device__ __forceinline__ void cta_witness(const u64 *__restrict__ w, u64 w_len, int log2b, int L, u64 col0, u64 col1,
                                            int16_t *__restrict__ f16, u64 *__restrict__ fx, int *__restrict__ flag,
                                            int16_t *tile) {
    const gl::u32 sl = threadIdx.x & 7, oct = threadIdx.x >> 3, nthreads = blockDim.x, nocts = nthreads >> 3;
    const ring8::Twiddles tw = ring8::make_twiddles(sl);
    const u64 i_begin = col0 / (u64)L, i_end = (col1 + (u64)L - 1) / (u64)L;  // w_ccs elements that touch the range
    const gl::u32 ec = CTA_WIT_ROWS / (gl::u32)L;                             // elements per chunk
    const u64 B = 1ull << log2b, half = B >> 1;
    for (u64 e0 = i_begin; e0 < i_end; e0 += ec) {
        const gl::u32 ne = (gl::u32)min((u64)ec, i_end - e0);
        // ---- phase A: ne elements, one octet each (all lanes of a warp take part in the shuffles) ----
        for (gl::u32 base = 0; base < ne; base += nocts) {
            const gl::u32 el = base + oct;
            const bool valid = el < ne;
            const u64 e = valid ? e0 + el : e0;  // padding octets recompute element e0 and store nothing
            u64 c[3];
            const u64 *p = w + e * ring::D + 3 * sl;
            c[0] = p[0]; c[1] = p[1]; c[2] = p[2];
            ring8::icrt8(c, tw);
            bool negative[3];
            u64 m[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if constexpr (MONT) c[k] = gl::from_mont(c[k]);
                ring::signed_rep(c[k], negative[k], m[k]);  // fq_convertible.rs:22-34
            }
            int16_t *trow = tile + (valid ? el : 0) * (L * ring::D) + 3 * sl;
            for (int l = 0; l < L; ++l) {  // balanced_decomposition/mod.rs:76-97 on the magnitudes, limb by limb
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    u64 rem = m[k] & (B - 1);
                    m[k] >>= log2b;
                    int dg = (int)rem;
                    if (rem > half) {  // |rem| == b/2 is kept (mod.rs:79)
                        dg -= (int)B;
                        m[k] += 1;
                    }
                    if (negative[k]) dg = -dg;
                    if (valid) trow[l * ring::D + k] = (int16_t)dg;
                }
            }
            if (valid && (m[0] | m[1] | m[2])) atomicOr(flag, 1);  // the reference would index out of bounds (mod.rs:80)
        }
        __syncthreads();
        // ---- rows of this chunk: tile row r is limb element e0 * L + r; only [col0, col1) belongs to this block ----
        const u64 row0 = e0 * (u64)L;
        const gl::u32 nrows = ne * (gl::u32)L;
        {   // the int16 witness, 16 bytes at a time
            const uint4 *src = reinterpret_cast<const uint4 *>(tile);
            uint4 *dst = reinterpret_cast<uint4 *>(f16 + row0 * ring::D);
            for (gl::u32 u = threadIdx.x; u < nrows * 3; u += nthreads) {
                const u64 row = row0 + u / 3;
                if (row >= col0 && row < col1) dst[u] = src[u];
            }
        }
        // ---- phase B: one thread per limb element ----
        for (gl::u32 r = threadIdx.x; r < nrows; r += nthreads) {
            const u64 row = row0 + r;
            if (row < col0 || row >= col1) continue;
            int d[ring::D];
            load_i16x24_cta(tile + r * ring::D, d);
            u64 x[ring::D];
            r96::crt24_small<MONT>(d, x);
            u64 *o = fx + row * 48;  // [slot][f0, f1, f2, f0+f1, f0+f2, f1+f2]: two slots = three 32-byte stores
#pragma unroll
            for (int s = 0; s < ring::NSLOT; s += 2) {
                const u64 a0 = x[3 * s], a1 = x[3 * s + 1], a2 = x[3 * s + 2];
                const u64 b0 = x[3 * s + 3], b1 = x[3 * s + 4], b2 = x[3 * s + 5];
                st256(o + s * 6, a0, a1, a2, gl::add_lazy(a0, a1));
                st256(o + s * 6 + 4, gl::add_lazy(a0, a2), gl::add_lazy(a1, a2), b0, b1);
                st256(o + s * 6 + 8, b2, gl::add_lazy(b0, b1), gl::add_lazy(b0, b2), gl::add_lazy(b1, b2));
            }
        }
        __syncthreads();  // the tile is rewritten by the next chunk
    }
}

}  // namespace lat

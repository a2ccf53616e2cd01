// Standalone negacyclic NTT over Z_q[X]/(X^d + 1), d = 2^k, q = 2^64 - 2^32 + 1 (BASELINE configs[3]).
//
// NOT part of the drop-in path and absent from the reference (its ring is Z_q[X]/(X^24 - X^12 + 1), see ring24.cuh):
// there is nothing in /root/reference to be identical to, so the definition below IS the specification, and parity is
// pinned only against the O(d^2) restatement used by the tests and the transform's algebraic properties.
//
//   psi    = 7^((q - 1) / 2d)            a primitive 2d-th root of unity (7 generates Z_q^*)
//   forward:  A[i] = sum_j a[j] psi^((2 i + 1) j)          evaluation at the d roots of X^d + 1, natural order in and out
//   inverse:  a[j] = d^-1 sum_i A[i] psi^(-(2 i + 1) j)
//   so NTT(a * b mod X^d + 1) = NTT(a) (.) NTT(b).
//
// Kernel (template on log2 d and direction): a block holds C = max(d, 2048) coefficients in padded shared memory and
// runs the stages four at a time, the 16 coefficients of a radix-16 unit in the registers of one thread (one last pass of
// 1-3 stages for the remainder), with the psi powers merged into the butterflies (Cooley-Tukey forward, natural ->
// bit-reversed; Gentleman-Sande inverse, bit-reversed -> natural), so there is no separate twist pass.  The forward
// transform's first pass reads global memory directly and the inverse's last pass writes it directly; the bit reversal is
// the shared-memory side of the forward's final store / the inverse's first load (16-byte global accesses).
// Arithmetic is lazy (any 64-bit representative; one operand of every add/sub canonical, see below), a general butterfly
// is 5 wide multiplies + about 25 ALU instructions, and the four stages with the fewest blocks use shift twiddles (their
// roots of unity are powers of two for every d), as does the final scaling by d^-1 = -2^(96 - log2 d).  The transform is
// bound by the integer ALU pipe (ncu: 78 % active, issue slots 66 %), not by the 16 bytes per coefficient it moves.
// Twiddle tables (psi^bitrev(k), psi^-bitrev(k); d words each) are built on the device once per (device, log2 d).
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <utility>

#include "goldilocks.cuh"
#include "kernels.h"

namespace lat {
namespace {

using gl::u32;
using gl::u64;

__host__ u64 host_mulmod(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % gl::Q); }
__host__ u64 host_powmod(u64 base, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = host_mulmod(r, base);
        base = host_mulmod(base, base);
        e >>= 1;
    }
    return r;
}

__device__ __forceinline__ u64 dev_pow(u64 base, u32 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl::mul(r, base);
        base = gl::mul(base, base);
        e >>= 1;
    }
    return r;
}

// table[k] = root^bitrev_{logd}(k), k < d
__global__ void __launch_bounds__(256) ntt_table_kernel(u64 root, u32 logd, u64 *__restrict__ table) {
    const u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k >= (1u << logd)) return;
    table[k] = dev_pow(root, __brev(k) >> (32 - logd));
}


// ---- lazy arithmetic: values are ANY 64-bit representative unless the name says canonical -------------------------------
// A butterfly is one product (4 IMAD.WIDE.U32), the special-form fold (1 IMAD.WIDE.U32 + carry chains) and two
// single-correction add/sub.  A single correction suffices when ONE operand is canonical (u + t < 2^64 + q, so the
// wrapped sum is < q - 1 and adding 2^32 - 1 cannot carry again; likewise for the borrow), so the product (forward) or
// the subtrahend (inverse) is canonicalised and everything else stays lazy until the final store.

// l + 2^64 h0 + 2^96 h1 -> some representative in [0, 2^64):  l - h1 (+ q on borrow) + h0 (2^32 - 1) (+ 2^32 - 1 on
// carry), because 2^64 = 2^32 - 1 and 2^96 = -1 (mod q).  Adding c (2^32 - 1) for a carry bit c is written as
// (hi:lo) - c + (c << 32) = sub.cc lo, c; subc hi, -c, which costs one instruction less than the add form.
__device__ __forceinline__ u64 fold_lazy(u64 l, u32 h0, u32 h1) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 w0, w1, c, m;\n\t"
        ".reg .u64 t, x;\n\t"
        "mov.b64 {w0, w1}, %1;\n\t"
        "sub.cc.u32 w0, w0, %3;\n\t"
        "subc.cc.u32 w1, w1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"              // 0xFFFFFFFF on borrow
        "sub.cc.u32 w0, w0, m;\n\t"
        "subc.u32 w1, w1, 0;\n\t"
        "mov.b64 t, {w0, w1};\n\t"
        "mul.wide.u32 x, %2, 0xFFFFFFFF;\n\t"
        "add.cc.u64 t, t, x;\n\t"
        "addc.u32 c, 0, 0;\n\t"              // (a subc here would read the carry flag as "no borrow")
        "neg.s32 m, c;\n\t"
        "mov.b64 {w0, w1}, t;\n\t"
        "sub.cc.u32 w0, w0, c;\n\t"
        "subc.u32 w1, w1, m;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "}"
        : "=l"(r)
        : "l"(l), "r"(h0), "r"(h1));
    return r;
}
// l + 2^64 h0, h0 < 2^31
__device__ __forceinline__ u64 fold_lazy(u64 l, u32 h0) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 w0, w1, c, m;\n\t"
        ".reg .u64 t, x;\n\t"
        "mul.wide.u32 x, %2, 0xFFFFFFFF;\n\t"
        "add.cc.u64 t, %1, x;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 m, c;\n\t"
        "mov.b64 {w0, w1}, t;\n\t"
        "sub.cc.u32 w0, w0, c;\n\t"
        "subc.u32 w1, w1, m;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "}"
        : "=l"(r)
        : "l"(l), "r"(h0));
    return r;
}
// a * b mod q for any 64-bit a, b; result is some representative in [0, 2^64).
// Product: a0 b0, a1 b1 and the 65-bit middle term a0 b1 + a1 b0 (one multiply-add with carry-out), joined by a single
// three-word carry chain into (w3 w2 w1 w0), then folded.
__device__ __forceinline__ u64 mul_lazy(u64 a, u64 b) {
    u64 l;
    u32 h0, h1;
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, w0, w1, ml, mh, c;\n\t"
        ".reg .u64 p0, p3, t, x, mid;\n\t"
        "mov.b64 {a0, a1}, %3;\n\t"
        "mov.b64 {b0, b1}, %4;\n\t"
        "mul.wide.u32 p0, a0, b0;\n\t"
        "mul.wide.u32 p3, a1, b1;\n\t"
        "mul.wide.u32 t, a0, b1;\n\t"
        "mul.wide.u32 x, a1, b0;\n\t"
        "add.cc.u64 mid, t, x;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mov.b64 {w0, w1}, p0;\n\t"
        "mov.b64 {%1, %2}, p3;\n\t"
        "mov.b64 {ml, mh}, mid;\n\t"
        "add.cc.u32 w1, w1, ml;\n\t"
        "addc.cc.u32 %1, %1, mh;\n\t"
        "addc.u32 %2, %2, c;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "}"
        : "=l"(l), "=r"(h0), "=r"(h1)
        : "l"(a), "l"(b));
    return fold_lazy(l, h0, h1);
}
// a * 2^E mod q for a compile-time 0 < E < 96, E not a multiple of 32; any 64-bit a, lazy result.
template <int E>
__device__ __forceinline__ u64 mul_pow2_lazy(u64 a) {
    static_assert(E > 0 && E < 96 && E % 32 != 0, "shift amount");
    if constexpr (E < 32) {
        return fold_lazy(a << E, (u32)(a >> (64 - E)));
    } else if constexpr (E < 64) {
        const u64 h = a >> (64 - E);
        return fold_lazy(a << E, (u32)h, (u32)(h >> 32));
    } else {
        // a 2^E = (y2 y1 y0) 2^64 with the 96-bit y = a << (E - 64);  2^64 = 2^32 - 1, 2^96 = -1, 2^128 = -2^32:
        // y0 (2^32 - 1) - (y2 : y1), one borrow correction (the subtrahend is < 2^63)
        constexpr int S = E - 64;
        const u64 y = a << S;
        const u64 p = (u64)(u32)y * 0xFFFFFFFFull;
        const u64 b = (y >> 32) | ((a >> (64 - S)) << 32);
        u64 r;
        asm("{\n\t"
            ".reg .u32 al, ah, bl, bh, m;\n\t"
            "mov.b64 {al, ah}, %1;\n\t"
            "mov.b64 {bl, bh}, %2;\n\t"
            "sub.cc.u32 al, al, bl;\n\t"
            "subc.cc.u32 ah, ah, bh;\n\t"
            "subc.u32 m, 0, 0;\n\t"
            "sub.cc.u32 al, al, m;\n\t"
            "subc.u32 ah, ah, 0;\n\t"
            "mov.b64 %0, {al, ah};\n\t"
            "}"
            : "=l"(r)
            : "l"(p), "l"(b));
        return r;
    }
}
// any representative -> [0, q)
__device__ __forceinline__ u64 canon(u64 x) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1;\n\t"
        ".reg .pred p;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "setp.eq.u32 p, l1, 0xFFFFFFFF;\n\t"
        "setp.ne.and.u32 p, l0, 0, p;\n\t"      // x >= q  <=>  high word all ones and low word >= 1
        "@p add.u32 l0, l0, 0xFFFFFFFF;\n\t"    // x - q = (0 : l0 - 1)
        "@p mov.u32 l1, 0;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}"
        : "=l"(r)
        : "l"(x));
    return r;
}
// u any representative, t canonical
__device__ __forceinline__ u64 add_lazy(u64 u, u64 t) {
    u64 d;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, m;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "add.cc.u32 al, al, bl;\n\t"
        "addc.cc.u32 ah, ah, bh;\n\t"
        "addc.u32 bl, 0, 0;\n\t"           // carry c: add c (2^32 - 1) = (hi:lo) - c + (c << 32)
        "neg.s32 m, bl;\n\t"
        "sub.cc.u32 al, al, bl;\n\t"
        "subc.u32 ah, ah, m;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(d)
        : "l"(u), "l"(t));
    return d;
}
__device__ __forceinline__ u64 sub_lazy(u64 u, u64 t) {
    u64 d;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, m;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "sub.cc.u32 al, al, bl;\n\t"
        "subc.cc.u32 ah, ah, bh;\n\t"
        "subc.u32 m, 0, 0;\n\t"            // 0xFFFFFFFF on borrow
        "sub.cc.u32 al, al, m;\n\t"
        "subc.u32 ah, ah, 0;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(d)
        : "l"(u), "l"(t));
    return d;
}

// Cooley-Tukey / Gentleman-Sande butterflies with the psi power merged in
__device__ __forceinline__ void ct(u64 &u, u64 &v, u64 s) {
    const u64 t = canon(mul_lazy(v, s));
    v = sub_lazy(u, t);
    u = add_lazy(u, t);
}
__device__ __forceinline__ void gs(u64 &u, u64 &v, u64 s) {
    const u64 c = canon(v);
    const u64 t = sub_lazy(u, c);
    u = add_lazy(u, c);
    v = mul_lazy(t, s);
}

// The twiddles of the four stages with the fewest blocks are powers of two for EVERY d (psi^(d / 2^(s+1)) is a
// primitive 2^(s+2)-th root of unity, and 2 has order 192 = 3 * 64): tw[k] = 2^POW2_FWD[k], inverse table 2^POW2_INV[k],
// 1 <= k < 16 (checked against the tables when they are built).  A multiplication by 2^E is shifts + one fold (one wide
// multiply instead of five); an exponent >= 96 is a negated twiddle (2^96 = -1), which swaps the butterfly's outputs.
__device__ constexpr int POW2_FWD[16] = {0, 48, 120, 168, 156, 12, 84, 132, 78, 126, 6, 54, 42, 90, 162, 18};
__device__ constexpr int POW2_INV[16] = {0, 144, 72, 24, 36, 180, 108, 60, 114, 66, 186, 138, 150, 102, 30, 174};
constexpr int HOST_POW2_FWD[16] = {0, 48, 120, 168, 156, 12, 84, 132, 78, 126, 6, 54, 42, 90, 162, 18};
constexpr int HOST_POW2_INV[16] = {0, 144, 72, 24, 36, 180, 108, 60, 114, 66, 186, 138, 150, 102, 30, 174};

template <int E>
__device__ __forceinline__ void ct_pow2(u64 &u, u64 &v) {
    const u64 t = canon(mul_pow2_lazy<E % 96>(v));
    const u64 s = add_lazy(u, t), d = sub_lazy(u, t);
    u = E < 96 ? s : d;
    v = E < 96 ? d : s;
}
template <int E>
__device__ __forceinline__ void gs_pow2(u64 &u, u64 &v) {
    const u64 cv = canon(v);
    u64 t;
    if constexpr (E < 96) {
        t = sub_lazy(u, cv);
    } else {
        t = sub_lazy(v, canon(u));   // (u - v) * -2^(E - 96)
    }
    u = add_lazy(u, cv);
    v = mul_pow2_lazy<E % 96>(t);
}
// stage r of a pass whose twiddles are the first ones of the table: butterflies B... of the unit
template <int R, int r, int... B>
__device__ __forceinline__ void fwd_stage_pow2(u64 (&x)[1 << R], std::integer_sequence<int, B...>) {
    constexpr int span = (1 << R) >> (r + 1);
    (ct_pow2<POW2_FWD[(1 << r) + B / span]>(x[(B / span) * 2 * span + B % span], x[(B / span) * 2 * span + B % span + span]), ...);
}
template <int R, int r, int... B>
__device__ __forceinline__ void inv_stage_pow2(u64 (&x)[1 << R], std::integer_sequence<int, B...>) {
    constexpr int span = 1 << r, nsb = (1 << R) >> (r + 1);
    (gs_pow2<POW2_INV[nsb + (B >> r)]>(x[((B >> r) << (r + 1)) + (B & (span - 1))], x[((B >> r) << (r + 1)) + (B & (span - 1)) + span]), ...);
}
template <int R, int... Rs>
__device__ __forceinline__ void fwd_unit_pow2(u64 (&x)[1 << R], std::integer_sequence<int, Rs...>) {
    (fwd_stage_pow2<R, Rs>(x, std::make_integer_sequence<int, (1 << R) / 2>{}), ...);
}
template <int R, int... Rs>
__device__ __forceinline__ void inv_unit_pow2(u64 (&x)[1 << R], std::integer_sequence<int, Rs...>) {
    (inv_stage_pow2<R, Rs>(x, std::make_integer_sequence<int, (1 << R) / 2>{}), ...);
}

// ---- geometry ------------------------------------------------------------------------------------------------------------
// A block holds C = max(d, 4096) coefficients (several polynomials when d < 4096) in padded shared memory and runs the
// log2(d) stages four at a time: a thread keeps the 16 coefficients of one radix-16 unit in registers (a last pass of
// 1-3 stages takes the remainder).  The forward transform's first pass reads global memory directly and the inverse
// transform's last pass writes it directly, so the tile is crossed once per pass boundary only.
#ifndef LAT_NTT_MIN_LOGC
#define LAT_NTT_MIN_LOGC 11   // 2048 coefficients = 128 threads per block for d <= 2048: 3-5 % faster inverse than 4096, forward equal
#endif
template <int LOGD>
struct Geo {
    static constexpr int LOGC = LOGD > LAT_NTT_MIN_LOGC ? LOGD : LAT_NTT_MIN_LOGC;
    static constexpr u32 C = 1u << LOGC;
    static constexpr int T = (int)(C / 16);   // one radix-16 unit per thread and pass
    static constexpr u32 PPB = C >> LOGD;   // polynomials per block
};
// Shared-memory index with padding: one extra word per 16, per 256 and per 4096 words, so that the strided accesses of
// the late passes and the bit-reversed gather spread over the banks.  pad(a + b) = pad(a) + pad(b) when a and b occupy
// disjoint bit ranges, which is what makes the per-coefficient offsets of a unit compile-time constants.
__host__ __device__ constexpr u32 pad(u32 i) { return i + (i >> 4) + (i >> 8) + (i >> 12); }

template <int LOGD, int LQ, int R>
struct Unit {
    u32 p, blk, base;
    __device__ __forceinline__ explicit Unit(u32 u) {
        constexpr int UPL = LOGD - R;   // log2(units per polynomial)
        p = u >> UPL;
        const u32 uu = u & ((1u << UPL) - 1);
        blk = uu >> LQ;
        base = (p << LOGD) + (blk << (LQ + R)) + (uu & ((1u << LQ) - 1));
    }
};

// Forward pass over stages S .. S+R-1 (stage s: 2^s blocks per polynomial, twiddle of block i is tw[2^s + i]).
template <int LOGD, int S, int R, bool FROM_GLOBAL>
__device__ __forceinline__ void fwd_pass(u64 *a, const u64 *__restrict__ in, u32 polys_here, const u64 *__restrict__ tw) {
    using G = Geo<LOGD>;
    constexpr int N = 1 << R;
    constexpr int LQ = LOGD - S - R;   // log2 of the distance between the coefficients a thread holds
#pragma unroll 1
    for (u32 u = threadIdx.x; u < (G::C >> R); u += G::T) {
        const Unit<LOGD, LQ, R> un(u);
        u64 x[N];
        u64 *sm = a + pad(un.base);
        if constexpr (FROM_GLOBAL) {
            const bool valid = un.p < polys_here;
#pragma unroll
            for (int k = 0; k < N; ++k) x[k] = valid ? in[un.base + ((u32)k << LQ)] : 0ull;
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) x[k] = sm[pad((u32)k << LQ)];
        }
        if constexpr (S == 0) {
            fwd_unit_pow2<R>(x, std::make_integer_sequence<int, R>{});   // stages 0 .. R-1 <= 3: shift twiddles
        } else {
            const u32 mb = (1u << S) + un.blk;
#pragma unroll
            for (int r = 0; r < R; ++r) {      // stage S + r: 2^r sub-blocks of the unit, pairs (i, i + N / 2^(r+1))
                u64 w[N / 2];
#pragma unroll
                for (int sb = 0; sb < N / 2; ++sb)
                    if (sb < (1 << r)) w[sb] = __ldg(tw + (mb << r) + sb);
#pragma unroll
                for (int b = 0; b < N / 2; ++b) {
                    const int span = N >> (r + 1), sb = b / span, i = sb * 2 * span + b % span;
                    ct(x[i], x[i + span], w[sb]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) sm[pad((u32)k << LQ)] = x[k];
    }
    __syncthreads();
}

// Inverse pass over the stages with distances 2^LQ .. 2^(LQ+R-1) (a stage with distance q has d / 2q blocks per polynomial;
// the twiddle of block i is tw[d / 2q + i]).
enum { OUT_TILE = 0, OUT_GLOBAL = 1, OUT_TILE_SCALED = 2 };
template <int LOGD, int LQ, int R, int MODE>
__device__ __forceinline__ void inv_pass(u64 *a, u64 *__restrict__ out, u32 polys_here, const u64 *__restrict__ tw) {
    using G = Geo<LOGD>;
    constexpr int N = 1 << R;
    constexpr u32 NB = 1u << (LOGD - LQ - R);   // units of this pass per polynomial and offset = blocks of its last stage
#pragma unroll 1
    for (u32 u = threadIdx.x; u < (G::C >> R); u += G::T) {
        const Unit<LOGD, LQ, R> un(u);
        u64 x[N];
        u64 *sm = a + pad(un.base);
#pragma unroll
        for (int k = 0; k < N; ++k) x[k] = sm[pad((u32)k << LQ)];
        if constexpr (NB == 1) {
            inv_unit_pow2<R>(x, std::make_integer_sequence<int, R>{});   // the last R <= 4 stages: shift twiddles
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {      // distance 2^r among the held coefficients, N / 2^(r+1) sub-blocks
                u64 w[N / 2];
#pragma unroll
                for (int sb = 0; sb < N / 2; ++sb)
                    if (sb < (N >> (r + 1))) w[sb] = __ldg(tw + (u32)(N >> (r + 1)) * (NB + un.blk) + sb);
#pragma unroll
                for (int b = 0; b < N / 2; ++b) {
                    const int span = 1 << r, sb = b >> r, i = (sb << (r + 1)) + (b & (span - 1));
                    gs(x[i], x[i + span], w[sb]);
                }
            }
        }
        if constexpr (MODE != OUT_TILE) {   // last pass: times d^-1 = 2^(192 - LOGD) = -2^(96 - LOGD), canonical
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const u64 y = canon(mul_pow2_lazy<96 - LOGD>(x[k]));
                x[k] = y ? gl::Q - y : 0ull;
            }
        }
        if constexpr (MODE == OUT_GLOBAL) {
            if (un.p < polys_here) {
#pragma unroll
                for (int k = 0; k < N; ++k) out[un.base + ((u32)k << LQ)] = x[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) sm[pad((u32)k << LQ)] = x[k];
        }
    }
    if constexpr (MODE != OUT_GLOBAL) __syncthreads();
}

template <int LOGD>
__device__ __forceinline__ u32 brev(u32 k) { return __brev(k) >> (32 - LOGD); }

// Block tile <-> global memory, two coefficients (16 bytes when aligned) per thread and step.  BITREV: coefficient k of a
// polynomial lives at tile position bitrev(k) (k and k + 1, k even, at bitrev(k) and bitrev(k) + d/2).
template <int LOGD, bool BITREV>
__device__ __forceinline__ void tile_load(u64 *a, const u64 *__restrict__ in, u32 polys_here, int vec16) {
    using G = Geo<LOGD>;
    constexpr u32 D = 1u << LOGD;
    for (u32 i = threadIdx.x; i < G::C / 2; i += G::T) {
        const u32 k0 = 2 * i, p = k0 >> LOGD, k = k0 & (D - 1);
        u64 v0 = 0, v1 = 0;
        if (p < polys_here) {
            if (vec16) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(in + k0);
                v0 = v.x;
                v1 = v.y;
            } else {
                v0 = in[k0];
                v1 = in[k0 + 1];
            }
        }
        const u32 pos = BITREV ? (p << LOGD) + brev<LOGD>(k) : k0;
        a[pad(pos)] = v0;
        a[pad(pos + (BITREV ? D / 2 : 1))] = v1;
    }
    __syncthreads();
}
template <int LOGD, bool BITREV>
__device__ __forceinline__ void tile_store(const u64 *a, u64 *__restrict__ out, u32 polys_here, int vec16) {
    using G = Geo<LOGD>;
    constexpr u32 D = 1u << LOGD;
    for (u32 i = threadIdx.x; i < G::C / 2; i += G::T) {
        const u32 k0 = 2 * i, p = k0 >> LOGD, k = k0 & (D - 1);
        if (p >= polys_here) break;
        const u32 pos = BITREV ? (p << LOGD) + brev<LOGD>(k) : k0;
        const u64 v0 = canon(a[pad(pos)]), v1 = canon(a[pad(pos + (BITREV ? D / 2 : 1))]);
        if (vec16) {
            *reinterpret_cast<ulonglong2 *>(out + k0) = make_ulonglong2(v0, v1);
        } else {
            out[k0] = v0;
            out[k0 + 1] = v1;
        }
    }
}

// For 8 <= d <= 32 the inverse transform's last pass would write global memory in runs of 1-2 coefficients per lane
// group; those sizes scale into the tile and leave through coalesced 16-byte stores (123 -> 95 us at d = 16).  Staging
// the forward transform's input the same way was slower at every size (d = 64: 80 -> 97 us) and is only kept as a switch.
#ifndef LAT_NTT_STAGE_FWD_BELOW_LOG2
#define LAT_NTT_STAGE_FWD_BELOW_LOG2 0
#endif
template <int LOGD, bool INVERSE>
__host__ __device__ constexpr bool staged() { return INVERSE ? (LOGD >= 3 && LOGD <= 5) : (LOGD < LAT_NTT_STAGE_FWD_BELOW_LOG2); }

template <int LOGD, bool INVERSE>
__global__ void __launch_bounds__(Geo<LOGD>::T)
ntt_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 batch, const u64 *__restrict__ tw, int vec16) {
    using G = Geo<LOGD>;
    extern __shared__ __align__(16) u64 a[];
    constexpr int A = LOGD / 4, REM = LOGD % 4;   // radix-16 passes + one pass of REM stages
    constexpr bool STAGED = staged<LOGD, INVERSE>();
    const u64 poly0 = (u64)blockIdx.x * G::PPB;
    const u32 polys_here = (u32)min((u64)G::PPB, batch - poly0);
    in += poly0 << LOGD;
    out += poly0 << LOGD;
    if constexpr (!INVERSE) {
        // natural order in; the passes leave position p of a polynomial holding output bitrev(p)
        if constexpr (STAGED) tile_load<LOGD, false>(a, in, polys_here, vec16);
        fwd_pass<LOGD, 0, (A > 0 ? 4 : REM), !STAGED>(a, in, polys_here, tw);
        if constexpr (A >= 2) fwd_pass<LOGD, 4, 4, false>(a, in, polys_here, tw);
        if constexpr (A >= 3) fwd_pass<LOGD, 8, 4, false>(a, in, polys_here, tw);
        if constexpr (A > 0 && REM > 0) fwd_pass<LOGD, 4 * A, REM, false>(a, in, polys_here, tw);
        tile_store<LOGD, true>(a, out, polys_here, vec16);
    } else {
        tile_load<LOGD, true>(a, in, polys_here, vec16);
        constexpr int LAST = STAGED ? OUT_TILE_SCALED : OUT_GLOBAL;
        if constexpr (A == 0) {
            inv_pass<LOGD, 0, REM, LAST>(a, out, polys_here, tw);
        } else {
            if constexpr (REM > 0) inv_pass<LOGD, 0, REM, OUT_TILE>(a, out, polys_here, tw);
            if constexpr (A >= 3) inv_pass<LOGD, REM, 4, OUT_TILE>(a, out, polys_here, tw);
            if constexpr (A >= 2) inv_pass<LOGD, LOGD - 8, 4, OUT_TILE>(a, out, polys_here, tw);
            inv_pass<LOGD, LOGD - 4, 4, LAST>(a, out, polys_here, tw);
        }
        if constexpr (STAGED) tile_store<LOGD, false>(a, out, polys_here, vec16);
    }
}

template <int LOGD, bool INVERSE>
cudaError_t launch_one(const u64 *in, u64 *out, u64 batch, const u64 *tw, int vec16, int dev, cudaStream_t stream) {
    using G = Geo<LOGD>;
    constexpr size_t smem = (size_t)(pad(G::C) + 4) * sizeof(u64);
    if (smem > 48 * 1024) {
        static bool set_on[64] = {};
        if (!set_on[dev & 63]) {
            cudaError_t e = cudaFuncSetAttribute(ntt_kernel<LOGD, INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            set_on[dev & 63] = true;
        }
    }
    const unsigned grid = (unsigned)((batch + G::PPB - 1) / G::PPB);
    ntt_kernel<LOGD, INVERSE><<<grid, G::T, smem, stream>>>(in, out, batch, tw, vec16);
    return cudaGetLastError();
}

struct Tables {
    u64 *fwd = nullptr, *inv = nullptr;
    u64 d_inv = 0;
};
std::mutex g_mu;
std::map<std::pair<int, u32>, Tables> g_tables;

}  // namespace

// 0 = ok, otherwise a cudaError_t (as int) from the table set-up or the launch
int launch_ntt_pow2(const u64 *in, u64 *out, u64 batch, u32 logd, bool inverse, cudaStream_t stream) {
    if (!batch) return 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    Tables t;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_tables.find({dev, logd});
        if (it == g_tables.end()) {
            const u64 d = 1ull << logd;
            const u64 psi = host_powmod(7, (gl::Q - 1) / (2 * d));
            const u64 psi_inv = host_powmod(psi, gl::Q - 2);
            if ((e = cudaMalloc(&t.fwd, d * sizeof(u64))) != cudaSuccess) return (int)e;
            if ((e = cudaMalloc(&t.inv, d * sizeof(u64))) != cudaSuccess) return (int)e;
            const unsigned grid = (unsigned)((d + 255) / 256);
            ntt_table_kernel<<<grid, 256, 0, stream>>>(psi, logd, t.fwd);
            ntt_table_kernel<<<grid, 256, 0, stream>>>(psi_inv, logd, t.inv);
            t.d_inv = host_powmod(d % gl::Q, gl::Q - 2);
            // the kernels hard-code the first 16 twiddles and d^-1 as powers of two: make sure they are what the tables hold
            for (u32 k = 1; k < 16 && k < d; ++k) {
                u32 rev = 0;
                for (u32 b = 0; b < logd; ++b) rev |= ((k >> b) & 1u) << (logd - 1 - b);
                if (host_powmod(psi, rev) != host_powmod(2, (u64)HOST_POW2_FWD[k]) ||
                    host_powmod(psi_inv, rev) != host_powmod(2, (u64)HOST_POW2_INV[k]))
                    return (int)cudaErrorAssert;
            }
            if (t.d_inv != gl::Q - host_powmod(2, 96 - logd)) return (int)cudaErrorAssert;
            if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
            // the tables are shared by every stream of this device from now on
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
            g_tables[{dev, logd}] = t;
        } else {
            t = it->second;
        }
    }
    const int vec16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 ? 1 : 0;
    e = cudaErrorInvalidValue;
    switch (logd) {
#define LAT_NTT_CASE(LOGD)                                                                                             \
    case LOGD:                                                                                                         \
        e = inverse ? launch_one<LOGD, true>(in, out, batch, t.inv, vec16, dev, stream)                               \
                    : launch_one<LOGD, false>(in, out, batch, t.fwd, vec16, dev, stream);                             \
        break;
        LAT_NTT_CASE(1) LAT_NTT_CASE(2) LAT_NTT_CASE(3) LAT_NTT_CASE(4) LAT_NTT_CASE(5) LAT_NTT_CASE(6) LAT_NTT_CASE(7)
        LAT_NTT_CASE(8) LAT_NTT_CASE(9) LAT_NTT_CASE(10) LAT_NTT_CASE(11) LAT_NTT_CASE(12) LAT_NTT_CASE(13) LAT_NTT_CASE(14)
#undef LAT_NTT_CASE
    default:
        break;
    }
    return (int)e;
}

}  // namespace lat

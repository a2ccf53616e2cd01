// Goldilocks field Z_q, q = 2^64 - 2^32 + 1, and the cubic extension Fq3 = Fq[u]/(u^3 - 2^40), for sm_100a.
//
// Reference semantics: crates/stark-rings/crates/ring/src/cyclotomic_ring/models/goldilocks/mod.rs:16-54
// (Fq = Fp64<MontBackend<FqConfig,1>>, Fq3 with NONRESIDUE = 2^40).  The reference's arithmetic lives in
// ark-ff 0.5.0; it is exact arithmetic mod q, so any exact reduction reproduces it bit for bit.  Here the
// reduction is the special-form fold 2^64 = 2^32 - 1, 2^96 = -1 (mod q), built from 32-bit IMAD/IADD3.
//
// Conventions: "canonical" = value in [0, q).  Every function takes and returns canonical values unless
// its name says otherwise.  2 has order 192 mod q, so multiplication by any power of two is a shift plus
// the fold; all CRT twiddles of the ring are such powers.
#pragma once
#include <cstdint>

namespace gl {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 Q = 0xFFFFFFFF00000001ull;
constexpr u64 EPS = 0xFFFFFFFFull;  // 2^64 mod q
constexpr u64 Q_HALF = (Q - 1) / 2;

// Canonical add / sub.  On the device they are carry chains (5 and 7 instructions; the compare-and-select forms the
// compiler derives from the C versions cost 8 and 11, and these two are most of what the ring transforms execute):
//   sub: d = a - b; on borrow add q, i.e. subtract 2^32 - 1 (mod 2^64)
//   add: a + b = a - (q - b), and q - b is a 2-instruction borrow chain (b = 0 gives q, which sub handles)
__host__ __device__ __forceinline__ u64 sub(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    u64 d;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, m;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "sub.cc.u32 al, al, bl;\n\t"
        "subc.cc.u32 ah, ah, bh;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // 0xFFFFFFFF on borrow, else 0
        "sub.cc.u32 al, al, m;\n\t"
        "subc.u32 ah, ah, 0;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(d)
        : "l"(a), "l"(b));
    return d;
#else
    u64 d = a - b;
    return (a < b) ? d - EPS : d;  // + q (mod 2^64)
#endif
}
__host__ __device__ __forceinline__ u64 add(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    u64 d;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, m;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "sub.cc.u32 bl, 1, bl;\n\t"          // q - b = (0xFFFFFFFF : 1) - (bh : bl)
        "subc.u32 bh, 0xFFFFFFFF, bh;\n\t"
        "sub.cc.u32 al, al, bl;\n\t"
        "subc.cc.u32 ah, ah, bh;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 al, al, m;\n\t"
        "subc.u32 ah, ah, 0;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(d)
        : "l"(a), "l"(b));
    return d;
#else
    u64 s = a + b;
    u64 t = s + EPS;  // s - q (mod 2^64); overflows iff s >= q
    return ((s < a) | (t < s)) ? t : s;
#endif
}
__host__ __device__ __forceinline__ u64 neg(u64 a) { return a ? Q - a : 0; }

// x = lo + 2^64 * hi  ->  canonical.   x = lo - hi_hi + hi_lo * (2^32 - 1)  (mod q)
__host__ __device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
#ifdef __CUDA_ARCH__
    // carry chains instead of compare/select (17 instructions for a shift-multiply against 25): lo - hi_hi (+ q on
    // borrow), + hi_lo (2^32 - 1) as one wide multiply-add (+ 2^32 - 1 on carry, written as (hi:lo) - c + (c << 32)),
    // then x >= q  <=>  high word all ones and low word >= 1
    u64 r;
    asm("{\n\t"
        ".reg .u32 w0, w1, h0, h1, c, m;\n\t"
        ".reg .u64 t, x;\n\t"
        ".reg .pred p;\n\t"
        "mov.b64 {w0, w1}, %1;\n\t"
        "mov.b64 {h0, h1}, %2;\n\t"
        "sub.cc.u32 w0, w0, h1;\n\t"
        "subc.cc.u32 w1, w1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 w0, w0, m;\n\t"
        "subc.u32 w1, w1, 0;\n\t"
        "mov.b64 t, {w0, w1};\n\t"
        "mul.wide.u32 x, h0, 0xFFFFFFFF;\n\t"
        "add.cc.u64 t, t, x;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 m, c;\n\t"
        "mov.b64 {w0, w1}, t;\n\t"
        "sub.cc.u32 w0, w0, c;\n\t"
        "subc.u32 w1, w1, m;\n\t"
        "setp.eq.u32 p, w1, 0xFFFFFFFF;\n\t"
        "setp.ne.and.u32 p, w0, 0, p;\n\t"
        "@p add.u32 w0, w0, 0xFFFFFFFF;\n\t"
        "@p mov.u32 w1, 0;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "}"
        : "=l"(r)
        : "l"(lo), "l"(hi));
    return r;
#else
    u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= EPS;
    u64 t1 = (hi_lo << 32) - hi_lo;
    u64 r = t0 + t1;
    if (r < t1) r += EPS;
    u64 c = r + EPS;  // canonicalise
    return (c < r) ? c : r;
#endif
}

// lo + 2^64 hi for hi < 2^32 (the shift-multiplies by less than 2^32): the fold and the canonicalisation only
__device__ __forceinline__ u64 reduce96(u64 lo, u32 hi) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 w0, w1, c, m;\n\t"
        ".reg .u64 t, x;\n\t"
        ".reg .pred p;\n\t"
        "mul.wide.u32 x, %2, 0xFFFFFFFF;\n\t"
        "add.cc.u64 t, %1, x;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "neg.s32 m, c;\n\t"
        "mov.b64 {w0, w1}, t;\n\t"
        "sub.cc.u32 w0, w0, c;\n\t"
        "subc.u32 w1, w1, m;\n\t"
        "setp.eq.u32 p, w1, 0xFFFFFFFF;\n\t"
        "setp.ne.and.u32 p, w0, 0, p;\n\t"
        "@p add.u32 w0, w0, 0xFFFFFFFF;\n\t"
        "@p mov.u32 w1, 0;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "}"
        : "=l"(r)
        : "l"(lo), "r"(hi));
    return r;
}

// 128-bit product: a0 b0, a1 b1 and the 65-bit middle term a0 b1 + a1 b0 (one multiply-add with carry-out), joined by one
// three-word carry chain -- 4 wide multiplies + 4 instructions (a * b with __umul64hi compiles to 5-6 wide multiplies + 6)
__device__ __forceinline__ void mul_wide(u64 a, u64 b, u64 &lo, u64 &hi) {
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, w0, w1, w2, w3, ml, mh, c;\n\t"
        ".reg .u64 p0, p3, t, x, mid;\n\t"
        "mov.b64 {a0, a1}, %2;\n\t"
        "mov.b64 {b0, b1}, %3;\n\t"
        "mul.wide.u32 p0, a0, b0;\n\t"
        "mul.wide.u32 p3, a1, b1;\n\t"
        "mul.wide.u32 t, a0, b1;\n\t"
        "mul.wide.u32 x, a1, b0;\n\t"
        "add.cc.u64 mid, t, x;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mov.b64 {w0, w1}, p0;\n\t"
        "mov.b64 {w2, w3}, p3;\n\t"
        "mov.b64 {ml, mh}, mid;\n\t"
        "add.cc.u32 w1, w1, ml;\n\t"
        "addc.cc.u32 w2, w2, mh;\n\t"
        "addc.u32 w3, w3, c;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "mov.b64 %1, {w2, w3};\n\t"
        "}"
        : "=l"(lo), "=l"(hi)
        : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 mul(u64 a, u64 b) {
    u64 lo, hi;
    mul_wide(a, b, lo, hi);
    return reduce128(lo, hi);
}

// a * 2^K mod q for a compile-time K in [0, 192).
template <int K>
__device__ __forceinline__ u64 mul_pow2(u64 a) {
    static_assert(K >= 0 && K < 192, "2 has order 192 mod q");
    if constexpr (K == 0) {
        return a;
    } else if constexpr (K >= 96) {
        return neg(mul_pow2<K - 96>(a));  // 2^96 = -1
    } else if constexpr (K <= 32) {
        return reduce96(a << K, (u32)(a >> (64 - K)));
    } else if constexpr (K < 64) {
        return reduce128(a << K, a >> (64 - K));
    } else if constexpr (K == 64) {
        return reduce128(0, a);
    } else {
        // a * 2^K = v * 2^64 with v = a << (K-64) = v_lo + 2^64 * v_hi, v_hi < 2^32; 2^128 = -2^32
        constexpr int S = K - 64;
        u64 v_lo = a << S, v_hi = a >> (64 - S);
        return sub(reduce128(0, v_lo), v_hi << 32);
    }
}

// ROOTS_OF_UNITY_24[I] = (2^40)^I = 2^(8 * (5 I mod 24))   (goldilocks/ntt.rs:15-40)
template <int I>
__device__ __forceinline__ u64 mul_w(u64 a) {
    return mul_pow2<8 * ((5 * (I % 24)) % 24)>(a);
}

// Montgomery form used by the reference's host memory (R = 2^64): mont(x) = x * 2^64, x = mont * 2^128.
__device__ __forceinline__ u64 to_mont(u64 x) { return mul_pow2<64>(x); }
__device__ __forceinline__ u64 from_mont(u64 m) { return mul_pow2<128>(m); }

// A small signed integer d (|d| < 2^31) as a field element, canonical or Montgomery.
template <bool MONT>
__device__ __forceinline__ u64 from_small(int d) {
    u64 m = (u64)(d < 0 ? -d : d);
    if constexpr (MONT) m = (m << 32) - m;  // * (2^32 - 1) = * 2^64 mod q; < q because m < 2^31
    return (d < 0 && m) ? Q - m : m;
}

// ---------------------------------------------------------------------------------------------------------
// Lazy multiply-accumulate: a sum of up to 2^31 64x64-bit products kept as three 32-bit-column accumulators
// (weights 2^0, 2^32, 2^64), each a 64-bit register pair + a 32-bit overflow word.  One 32x32 partial product is
// `mul.wide.u32 + add.cc.u64 + addc.u32`, which ptxas fuses into ONE IMAD.WIDE.U32 with carry-out predicate
// plus half an IADD3.X (two carries are absorbed per IADD3.X) -- checked in SASS; keeping the accumulator a
// 64-bit PTX register is what avoids per-iteration IMAD.MOV copies of loop-carried halves.
// IMAD.WIDE.U32 issues at half the IMAD rate on sm_100a (measured 31.5 /clk/SM, tools/imad_peak.cu), so the
// multiply count is what bounds the MAC: hence Karatsuba in Fq3 below (6 base products instead of 9) for the single
// witness, and Toom-3 (5) where several witnesses share a matrix entry.
// ---------------------------------------------------------------------------------------------------------
struct Col {
    u64 acc;
    u32 ov;
};
__device__ __forceinline__ void col_mac(Col &c, u32 x, u32 y) {
    asm("{\n\t"
        ".reg .u64 p;\n\t"
        "mul.wide.u32 p, %2, %3;\n\t"
        "add.cc.u64 %0, %0, p;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "+l"(c.acc), "+r"(c.ov)
        : "r"(x), "r"(y));
}
struct WideAcc {
    Col c0, c1, c2;
    __device__ __forceinline__ void clear() { c0 = c1 = c2 = Col{0ull, 0u}; }
    __device__ __forceinline__ void mac(u64 a, u64 b) {
        u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
        col_mac(c0, al, bl);
        col_mac(c1, al, bh);
        col_mac(c1, ah, bl);
        col_mac(c2, ah, bh);
    }
    // value = c0 + c1 * 2^32 + c2 * 2^64 (each column < 2^96)  ->  canonical
    __device__ __forceinline__ u64 reduce() const {
        u64 t;
        u32 w0 = (u32)c0.acc;
        t = (c0.acc >> 32) + (u32)c1.acc;
        u32 w1 = (u32)t;
        t = (t >> 32) + c0.ov + (c1.acc >> 32) + (u32)c2.acc;
        u32 w2 = (u32)t;
        t = (t >> 32) + c1.ov + (c2.acc >> 32);
        u32 w3 = (u32)t;
        t = (t >> 32) + c2.ov;
        u32 w4 = (u32)t;  // no sixth word: every column is < 2^96
        // 2^128 = -2^32 :  x = (w0 + w1 2^32 + w2 2^64 + w3 2^96) - w4 * 2^32
        u64 lo = ((u64)w1 << 32) | w0, hi = ((u64)w3 << 32) | w2;
        return sub(reduce128(lo, hi), (u64)w4 << 32);
    }
};

// a + b as SOME 64-bit representative of (a + b) mod q, for canonical a, b: on carry add 2^64 = 2^32 - 1 (mod q);
// a, b < q makes a second carry impossible.  Used for the witness-side Karatsuba sums of the extended layout: the
// MAC only multiplies by them, and products tolerate non-canonical operands.
__device__ __forceinline__ u64 add_lazy(u64 a, u64 b) {
    u64 s = a + b;
    return s + ((s < a) ? EPS : 0ull);
}

// Karatsuba pre-addition on the matrix side: the exact 65-bit sum a + b = s + c * 2^64 (c = carry).  The carry is NOT
// folded back into s (that costs IMAD-pipe instructions, measured); instead mac65 below adds c * y * 2^64 straight
// into the accumulator's weight-2^64 column with carry-chain adds, which ptxas must keep on the ALU pipe.
__device__ __forceinline__ void add65(u64 a, u64 b, u64 &s, u32 &c) {
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh;\n\t"
        "mov.b64 {al, ah}, %2;\n\t"
        "mov.b64 {bl, bh}, %3;\n\t"
        "add.cc.u32 al, al, bl;\n\t"
        "addc.cc.u32 ah, ah, bh;\n\t"
        "addc.u32 %1, 0, 0;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(s), "=r"(c)
        : "l"(a), "l"(b));
}
// acc += (s + c * 2^64) * y,  c in {0, 1}.  (A predicated add.cc/addc pair instead of select-then-add looks cheaper in
// PTX, but ptxas then merges the guarded results with extra IMAD.MOV / IMAD.X: 1019 instead of 951 instructions in the
// unrolled tile loop -- -DLAT_MAC65_PREDICATED keeps that variant for the record.)
__device__ __forceinline__ void mac65(WideAcc &A, u64 s, u32 c, u64 y) {
    A.mac(s, y);
#ifndef LAT_MAC65_PREDICATED
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u64 ym;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "selp.b64 ym, %3, 0, p;\n\t"
        "add.cc.u64 %0, %0, ym;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "+l"(A.c2.acc), "+r"(A.c2.ov)
        : "r"(c), "l"(y));
#else
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p add.cc.u64 %0, %0, %3;\n\t"
        "@p addc.u32 %1, %1, 0;\n\t"
        "}"
        : "+l"(A.c2.acc), "+r"(A.c2.ov)
        : "r"(c), "l"(y));
#endif
}

// Accumulators of one Fq3 output, Karatsuba form: for x = (x0,x1,x2), y = (y0,y1,y2) in Fq[u]/(u^3 - 2^40),
// with x01 = x0+x1 etc.:
//   P0 = sum x0 y0, P1 = sum x1 y1, P2 = sum x2 y2, P01 = sum x01 y01, P02 = sum x02 y02, P12 = sum x12 y12
//   c0 = P0 + NR (P12 - P1 - P2),  c1 = (P01 - P0 - P1) + NR P2,  c2 = (P02 - P0 - P2) + P1,  NR = 2^40.
// The x side (matrix) is summed on the fly as exact 65-bit values; the y side (witness) arrives pre-summed mod q.
// Column bounds: c2 gets at most 2 additions < 2^64 per MAC, so 2^31 MACs still fit its 96 bits.
struct Fq3Acc {
    WideAcc p0, p1, p2, p01, p02, p12;
    __device__ __forceinline__ void clear() {
        p0.clear(); p1.clear(); p2.clear(); p01.clear(); p02.clear(); p12.clear();
    }
    __device__ __forceinline__ void mac(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2, u64 b01, u64 b02, u64 b12) {
        u64 s01, s02, s12;
        u32 k01, k02, k12;
        add65(a0, a1, s01, k01);
        add65(a0, a2, s02, k02);
        add65(a1, a2, s12, k12);
        p0.mac(a0, b0);
        p1.mac(a1, b1);
        p2.mac(a2, b2);
        mac65(p01, s01, k01, b01);
        mac65(p02, s02, k02, b02);
        mac65(p12, s12, k12, b12);
    }
    __device__ __forceinline__ void finish(u64 &c0, u64 &c1, u64 &c2) const {
        u64 r0 = p0.reduce(), r1 = p1.reduce(), r2 = p2.reduce();
        u64 r01 = p01.reduce(), r02 = p02.reduce(), r12 = p12.reduce();
        c0 = add(r0, mul_pow2<40>(sub(sub(r12, r1), r2)));
        c1 = add(sub(sub(r01, r0), r1), mul_pow2<40>(r2));
        c2 = add(sub(sub(r02, r0), r2), r1);
    }
};

// ---------------------------------------------------------------------------------------------------------
// Toom-3 form of the same accumulation, for launches where SEVERAL witnesses meet the same matrix entry (the K-1 planes
// of a decomposition, commit batches): a degree-2 by degree-2 product has 5 coefficients d0..d4, so 5 point products
// determine it.  Both operands are evaluated at t = 0, infinity, 1, -1, 2,
//     e = (x0, x2, x0+x1+x2, x0-x1+x2, x0+2 x1+4 x2)    (mod q, any 64-bit representative),
// the five products are summed lazily over the columns -- sum_j a_j(t) y_j(t) is the evaluation of sum_j a_j y_j, the
// interpolation is linear -- and interpolated ONCE per output:
//     d0 = V0, d4 = Vinf, d2 = (V1+Vm1)/2 - d0 - d4, S = (V1-Vm1)/2 = d1+d3, T = (V2 - d0 - 4 d2 - 16 d4)/2 = d1+4 d3,
//     d3 = (T-S)/3, d1 = S-d3;   c0 = d0 + NR d3, c1 = d1 + NR d4, c2 = d2.
// 5 x 4 = 20 wide multiplies per Fq3 MAC instead of Karatsuba's 24 and no pre-additions in the loop at all: the matrix
// side is stored evaluated (5 words per entry instead of 3: lat::derive_toom_kernel), the witness side arrives evaluated
// from its producer (toom_eval below).  The single-witness kernel stays Karatsuba on the 3-word matrix: it is bound by
// the bytes of the matrix, and 5 words per entry would cost more HBM time than the four multiplies save.
// ---------------------------------------------------------------------------------------------------------
constexpr u64 INV3 = 0xAAAAAAAA00000001ull;  // 3 * INV3 = 2 q + 1

// x / 2 mod q for canonical x: (x + q) / 2 when x is odd, and (q + 1) / 2 = 2^63 - 2^31 + 1
__host__ __device__ __forceinline__ u64 half(u64 x) { return (x >> 1) + ((x & 1) ? 0x7FFFFFFF80000001ull : 0ull); }

// canonical (x0, x1, x2) -> the evaluations at 1, -1 and 2, canonical (the values at 0 and infinity are x0 and x2)
__device__ __forceinline__ void toom_eval(u64 x0, u64 x1, u64 x2, u64 &t1, u64 &tm, u64 &t2) {
    const u64 e = add(x0, x2);
    t1 = add(e, x1);
    tm = sub(e, x1);
    u64 w = add(add(x2, x2), x1);  // x1 + 2 x2
    t2 = add(add(w, w), x0);
}

struct ToomAcc {
    WideAcc v0, vinf, v1, vm1, v2;
    __device__ __forceinline__ void clear() {
        v0.clear(); vinf.clear(); v1.clear(); vm1.clear(); v2.clear();
    }
    // a*: the matrix entry's evaluations in the order (0, infinity, 1, -1, 2); y*: the witness slot's, same order
    __device__ __forceinline__ void mac(u64 a0, u64 ainf, u64 a1, u64 am1, u64 a2, u64 y0, u64 yinf, u64 y1, u64 ym1, u64 y2) {
        v0.mac(a0, y0);
        vinf.mac(ainf, yinf);
        v1.mac(a1, y1);
        vm1.mac(am1, ym1);
        v2.mac(a2, y2);
    }
    __device__ __forceinline__ void finish(u64 &c0, u64 &c1, u64 &c2) const {
        const u64 d0 = v0.reduce(), d4 = vinf.reduce(), r1 = v1.reduce(), rm = vm1.reduce(), r2 = v2.reduce();
        const u64 d2 = sub(sub(half(add(r1, rm)), d0), d4);
        const u64 s = half(sub(r1, rm));
        const u64 t = half(sub(sub(sub(r2, d0), mul_pow2<2>(d2)), mul_pow2<4>(d4)));
        const u64 d3 = mul(sub(t, s), INV3);
        const u64 d1 = sub(s, d3);
        c0 = add(d0, mul_pow2<40>(d3));
        c1 = add(d1, mul_pow2<40>(d4));
        c2 = d2;
    }
};

// Plain (eager) Fq3 product, for the small kernels.
__device__ __forceinline__ void fq3_mul(const u64 a[3], const u64 b[3], u64 c[3]) {
    u64 t12 = add(mul(a[1], b[2]), mul(a[2], b[1]));
    u64 c0 = add(mul(a[0], b[0]), mul_pow2<40>(t12));
    u64 c1 = add(add(mul(a[0], b[1]), mul(a[1], b[0])), mul_pow2<40>(mul(a[2], b[2])));
    u64 c2 = add(add(mul(a[0], b[2]), mul(a[1], b[1])), mul(a[2], b[0]));
    c[0] = c0; c[1] = c1; c[2] = c2;
}

}  // namespace gl

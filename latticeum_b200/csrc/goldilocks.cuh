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

__host__ __device__ __forceinline__ u64 add(u64 a, u64 b) {
    u64 s = a + b;
    u64 t = s + EPS;  // s - q (mod 2^64); overflows iff s >= q
    return ((s < a) | (t < s)) ? t : s;
}
__host__ __device__ __forceinline__ u64 sub(u64 a, u64 b) {
    u64 d = a - b;
    return (a < b) ? d - EPS : d;  // + q (mod 2^64)
}
__host__ __device__ __forceinline__ u64 neg(u64 a) { return a ? Q - a : 0; }

// x = lo + 2^64 * hi  ->  canonical.   x = lo - hi_hi + hi_lo * (2^32 - 1)  (mod q)
__host__ __device__ __forceinline__ u64 reduce128(u64 lo, u64 hi) {
    u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= EPS;
    u64 t1 = (hi_lo << 32) - hi_lo;
    u64 r = t0 + t1;
    if (r < t1) r += EPS;
    u64 c = r + EPS;  // canonicalise
    return (c < r) ? c : r;
}

__device__ __forceinline__ u64 mul(u64 a, u64 b) { return reduce128(a * b, __umul64hi(a, b)); }

// a * 2^K mod q for a compile-time K in [0, 192).
template <int K>
__device__ __forceinline__ u64 mul_pow2(u64 a) {
    static_assert(K >= 0 && K < 192, "2 has order 192 mod q");
    if constexpr (K == 0) {
        return a;
    } else if constexpr (K >= 96) {
        return neg(mul_pow2<K - 96>(a));  // 2^96 = -1
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
// Lazy multiply-accumulate: sum of up to 2^20 64x64-bit products kept as three 32-bit-column accumulators
// (weights 2^0, 2^32, 2^64), each 64 bits + a 32-bit overflow word.  One product = 4 IMAD.WIDE.U32 with
// carry-out, the carries absorbed pairwise by IADD3.X (checked in SASS).  Reduced once at the end.
// ---------------------------------------------------------------------------------------------------------
struct Col {
    u32 lo, hi, ov;
};
__device__ __forceinline__ void col_mac(Col &c, u32 x, u32 y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c.lo), "+r"(c.hi), "+r"(c.ov)
        : "r"(x), "r"(y));
}
struct WideAcc {
    Col c0, c1, c2;
    __device__ __forceinline__ void clear() { c0 = c1 = c2 = Col{0u, 0u, 0u}; }
    __device__ __forceinline__ void mac(u64 a, u64 b) {
        u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
        col_mac(c0, al, bl);
        col_mac(c1, al, bh);
        col_mac(c1, ah, bl);
        col_mac(c2, ah, bh);
    }
    // value = c0 + c1 * 2^32 + c2 * 2^64 (each column < 2^96)  ->  canonical
    __device__ __forceinline__ u64 reduce() const {
        // words w0..w5 of the 192-bit sum
        u64 t;
        u32 w0 = c0.lo;
        t = (u64)c0.hi + c1.lo;
        u32 w1 = (u32)t;
        t = (t >> 32) + c0.ov + c1.hi + c2.lo;
        u32 w2 = (u32)t;
        t = (t >> 32) + c1.ov + c2.hi;
        u32 w3 = (u32)t;
        t = (t >> 32) + c2.ov;
        u32 w4 = (u32)t;  // t < 2^33 only if a column overflowed 2^96, which n <= 2^20 excludes; w5 = 0
        // 2^128 = -2^32 :  x = (w0 + w1 2^32 + w2 2^64 + w3 2^96) - w4 * 2^32
        u64 lo = ((u64)w1 << 32) | w0, hi = ((u64)w3 << 32) | w2;
        return sub(reduce128(lo, hi), (u64)w4 << 32);
    }
};

// Accumulators of one Fq3 output: sum_j a_j * b_j in Fq[u]/(u^3 - 2^40)
//   c0 = S00 + NR * S12,  c1 = S01 + NR * S22,  c2 = S02
//   S00 = sum a0 b0, S12 = sum a1 b2 + a2 b1, S01 = sum a0 b1 + a1 b0, S22 = sum a2 b2,
//   S02 = sum a0 b2 + a1 b1 + a2 b0.
struct Fq3Acc {
    WideAcc s00, s12, s01, s22, s02;
    __device__ __forceinline__ void clear() {
        s00.clear(); s12.clear(); s01.clear(); s22.clear(); s02.clear();
    }
    __device__ __forceinline__ void mac(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2) {
        s00.mac(a0, b0);
        s12.mac(a1, b2); s12.mac(a2, b1);
        s01.mac(a0, b1); s01.mac(a1, b0);
        s22.mac(a2, b2);
        s02.mac(a0, b2); s02.mac(a1, b1); s02.mac(a2, b0);
    }
    __device__ __forceinline__ void finish(u64 &c0, u64 &c1, u64 &c2) const {
        c0 = add(s00.reduce(), mul_pow2<40>(s12.reduce()));
        c1 = add(s01.reduce(), mul_pow2<40>(s22.reduce()));
        c2 = s02.reduce();
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

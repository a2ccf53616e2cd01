// CRT / iCRT of the Goldilocks ring (d = 24) computed in Z/(2^96 + 1), one reduction mod q per output.
//
// q = 2^64 - 2^32 + 1 divides 2^96 + 1 = q (2^32 + 1), and every twiddle of the reference's transform is a power of two
// (omega = 2^40; GOLD/ntt.rs:15-47), so the whole butterfly network can run in the ring of integers mod 2^96 + 1 --
// where multiplying by 2^R is a 96-bit rotation whose wrapped-around part changes sign -- and only the 24 results are
// reduced mod q.  An element is the SIGNED integer
//        V = w0 + w1 2^32 + w2 2^64 + c 2^96          (three 32-bit words and a small signed overflow count c),
// read mod 2^96 + 1 (so 2^96 = -1: V = W - c).  Costs, in instructions, against the canonical 64-bit arithmetic of
// ring24.cuh (add 7, sub 5, shift-multiply about 20, each with a canonicalisation):
//        add / sub : one 4-word carry chain (4)
//        times 2^R : funnel shifts for the two halves of the rotated value + one chain (about 10; 5 when R is a
//                    multiple of 32); a negated twiddle (2^(96 + R)) swaps the operands of that chain, for free
//        to Fq     : once per output (about 23)
// which brings the forward transform from about 2000 to about 1250 instructions per element and the inverse from about
// 3000 to about 1800 (the inverse keeps its 12 general multiplications by KAPPA; GOLD/ntt.rs:43,314).
// Same values as ring24.cuh, bit for bit: all maps are exact.
//
// The header is self-contained and also compiles as plain C++ (the carry chains fall back to 64-bit arithmetic), which
// is how tests/test_ring96_cpu.py checks the network on the CPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define R96_FN __host__ __device__ __forceinline__
#else
#define R96_FN inline
#endif

namespace r96 {

typedef uint32_t u32;
typedef unsigned long long u64;

constexpr int D = 24;
constexpr u64 Q = 0xFFFFFFFF00000001ull;
constexpr u64 KAPPA = 12297829382473034411ull;  // GOLD/ntt.rs:43: (2 zeta - 1)^-1, zeta = omega^4

struct T {
    u32 w0, w1, w2;
    int c;
};

R96_FN T from_u64(u64 x) { return T{(u32)x, (u32)(x >> 32), 0u, 0}; }
// a small signed integer (|x| < 2^63), e.g. the exact first layer of a digit vector's transform
R96_FN T from_i64(long long x) {
    const u32 s = (u32)(x >> 63);
    return T{(u32)x, (u32)((u64)x >> 32), s, (int)s};
}

R96_FN T add(const T &a, const T &b) {
    T r;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, %9;\n\t"
        "addc.cc.u32 %2, %6, %10;\n\t"
        "addc.u32 %3, %7, %11;"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.c)
        : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(a.c), "r"(b.w0), "r"(b.w1), "r"(b.w2), "r"(b.c));
#else
    u64 s = (u64)a.w0 + b.w0;
    r.w0 = (u32)s;
    s = (u64)a.w1 + b.w1 + (s >> 32);
    r.w1 = (u32)s;
    s = (u64)a.w2 + b.w2 + (s >> 32);
    r.w2 = (u32)s;
    r.c = (int)((u32)a.c + (u32)b.c + (u32)(s >> 32));
#endif
    return r;
}
R96_FN T sub(const T &a, const T &b) {
    T r;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %4, %8;\n\t"
        "subc.cc.u32 %1, %5, %9;\n\t"
        "subc.cc.u32 %2, %6, %10;\n\t"
        "subc.u32 %3, %7, %11;"
        : "=r"(r.w0), "=r"(r.w1), "=r"(r.w2), "=r"(r.c)
        : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(a.c), "r"(b.w0), "r"(b.w1), "r"(b.w2), "r"(b.c));
#else
    u64 s = (u64)a.w0 - b.w0;
    r.w0 = (u32)s;
    s = (u64)a.w1 - b.w1 - ((s >> 32) & 1);
    r.w1 = (u32)s;
    s = (u64)a.w2 - b.w2 - ((s >> 32) & 1);
    r.w2 = (u32)s;
    r.c = (int)((u32)a.c - (u32)b.c - (u32)((s >> 32) & 1));
#endif
    return r;
}

// upper word of (hi:lo) << s, and lower word of (hi:lo) >> s, 0 < s < 32  (SHF.L / SHF.R, or PRMT for whole bytes)
R96_FN u32 fsl(u32 lo, u32 hi, int s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return (u32)((((u64)hi << 32) | lo) << s >> 32);
#endif
}
R96_FN u32 fsr(u32 lo, u32 hi, int s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    return (u32)((((u64)hi << 32) | lo) >> s);
#endif
}

// word i of the signed 128-bit integer (w0, w1, w2, c) extended by its sign; i is a compile-time constant
template <int I>
R96_FN u32 vword(const T &a) {
    if constexpr (I < 0) return 0u;
    else if constexpr (I == 0) return a.w0;
    else if constexpr (I == 1) return a.w1;
    else if constexpr (I == 2) return a.w2;
    else if constexpr (I == 3) return (u32)a.c;
    else return (u32)(a.c >> 31);
}
template <int I>
R96_FN u32 wword(const T &a) {  // word i of W = (w0, w1, w2) alone, zero outside
    if constexpr (I < 0 || I > 2) return 0u;
    else return vword<I>(a);
}

// V * 2^R (NEG: times -2^R) for 0 < R < 96.  V 2^R = (W << R) + c 2^(96 + R); with Lo = (W << R) mod 2^96 and
// H = V >> (96 - R) (arithmetic, on the signed 128-bit V) this is Lo - H mod 2^96 + 1.
template <int R, bool NEG>
R96_FN T rot(const T &a) {
    static_assert(R > 0 && R < 96, "rotation amount");
    constexpr int m = R / 32, r = R % 32;
    constexpr int S = 96 - R, i0 = S / 32, sh = S % 32;
    T lo, hi;
    if constexpr (r == 0) {
        lo.w0 = wword<0 - m>(a); lo.w1 = wword<1 - m>(a); lo.w2 = wword<2 - m>(a);
    } else {
        lo.w0 = fsl(wword<-1 - m>(a), wword<0 - m>(a), r);
        lo.w1 = fsl(wword<0 - m>(a), wword<1 - m>(a), r);
        lo.w2 = fsl(wword<1 - m>(a), wword<2 - m>(a), r);
    }
    lo.c = 0;
    if constexpr (sh == 0) {
        hi.w0 = vword<i0>(a); hi.w1 = vword<i0 + 1>(a); hi.w2 = vword<i0 + 2>(a); hi.c = (int)vword<i0 + 3>(a);
    } else {
        hi.w0 = fsr(vword<i0>(a), vword<i0 + 1>(a), sh);
        hi.w1 = fsr(vword<i0 + 1>(a), vword<i0 + 2>(a), sh);
        hi.w2 = fsr(vword<i0 + 2>(a), vword<i0 + 3>(a), sh);
        hi.c = (int)fsr(vword<i0 + 3>(a), vword<i0 + 4>(a), sh);
    }
    return NEG ? sub(hi, lo) : sub(lo, hi);
}

// V * 2^E for a compile-time E in [0, 192): 2^96 = -1
template <int E>
R96_FN T mul_pow2(const T &a) {
    static_assert(E >= 0 && E < 192, "2 has order 192 mod q");
    if constexpr (E == 0) return a;
    else if constexpr (E == 96) return sub(T{0u, 0u, 0u, 0}, a);
    else if constexpr (E < 96) return rot<E, false>(a);
    else return rot<E - 96, true>(a);
}
// ROOTS_OF_UNITY_24[I] = (2^40)^I = 2^(8 (5 I mod 24))      (GOLD/ntt.rs:15-40)
template <int I>
R96_FN T mul_w(const T &a) { return mul_pow2<8 * ((5 * (I % 24)) % 24)>(a); }

// (a, b) -> (a + w^I b, a - w^I b); a negated twiddle swaps the two outputs instead of negating the product
template <int I>
R96_FN void bf_fwd(T &a, T &b) {
    constexpr int E = 8 * ((5 * (I % 24)) % 24);
    if constexpr (E < 96) {
        const T t = mul_pow2<E>(b);
        b = sub(a, t);
        a = add(a, t);
    } else {
        const T t = mul_pow2<E - 96>(b);
        b = add(a, t);
        a = sub(a, t);
    }
}
// (a, b) -> (a + b, w^I (a - b))
template <int I>
R96_FN void bf_inv(T &a, T &b) {
    const T d = sub(a, b);
    a = add(a, b);
    b = mul_w<I>(d);
}
template <int I1, int I2>
R96_FN void swap_scale(T &c1, T &c2) {  // c1' = w^I1 c2, c2' = w^I2 c1
    const T t = c1;
    c1 = mul_w<I1>(c2);
    c2 = mul_w<I2>(t);
}

// V mod q as SOME 64-bit representative (lazy) or the canonical one.
//   W = (w1:w0) + w2 2^64 = (w1:w0) - w2 + w2 2^32 (mod q), then minus c; every wrap of 2^64 is worth 2^32 - 1.
template <bool CANONICAL>
R96_FN u64 to_fq(const T &a) {
#ifdef __CUDA_ARCH__
    u32 l, h;
    asm("{\n\t"
        ".reg .u32 m, k, s, t, nt, ht;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"      // (w1:w0) - w2
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"           // 0xFFFFFFFF on borrow: add q = subtract 2^32 - 1
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "add.cc.u32 %1, %1, %4;\n\t"      // + w2 2^32
        "addc.u32 k, 0, 0;\n\t"
        "sub.u32 k, 0, k;\n\t"            // 0xFFFFFFFF on carry: add 2^32 - 1 (cannot carry again)
        "add.cc.u32 %0, %0, k;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "shr.s32 s, %5, 31;\n\t"          // - c, c sign-extended to 96 bits: the top word t is -1, 0 or 1
        "sub.cc.u32 %0, %0, %5;\n\t"
        "subc.cc.u32 %1, %1, s;\n\t"
        "subc.u32 t, 0, s;\n\t"
        "sub.u32 nt, 0, t;\n\t"           // + t (2^32 - 1) = (t >> 31 : -t)
        "shr.s32 ht, t, 31;\n\t"
        "add.cc.u32 %0, %0, nt;\n\t"
        "addc.u32 %1, %1, ht;\n\t"
        "}"
        : "=&r"(l), "=&r"(h)
        : "r"(a.w0), "r"(a.w1), "r"(a.w2), "r"(a.c));
    if constexpr (CANONICAL) {
        asm("{\n\t"
            ".reg .u32 k;\n\t"
            "add.cc.u32 k, %0, 0xFFFFFFFF;\n\t"   // z + (2^32 - 1) = z - q (mod 2^64) carries iff z >= q
            "addc.cc.u32 k, %1, 0;\n\t"
            "addc.u32 k, 0, 0;\n\t"
            "sub.u32 k, 0, k;\n\t"
            "add.cc.u32 %0, %0, k;\n\t"
            "addc.u32 %1, %1, 0;\n\t"
            "}"
            : "+r"(l), "+r"(h));
    }
    return ((u64)h << 32) | l;
#else
    // portable: exact arithmetic on a signed 128-bit integer
    __int128 v = (__int128)(((u64)a.w1 << 32) | a.w0) - (__int128)a.w2 + ((__int128)a.w2 << 32) - (__int128)a.c;
    v %= (__int128)Q;
    if (v < 0) v += (__int128)Q;
    (void)CANONICAL;
    return (u64)v;
#endif
}

// the 128-bit product a b of two 64-bit representatives, as a ring element: p0 + p1 2^32 + p2 2^64 + p3 2^96 = (p0, p1, p2) - p3
R96_FN T mul_u64(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    const u64 lo = a * b, hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    const u64 lo = (u64)p, hi = (u64)(p >> 64);
#endif
    return sub(T{(u32)lo, (u32)(lo >> 32), (u32)hi, 0}, T{(u32)(hi >> 32), 0u, 0u, 0});
}

// ---- the transforms (same networks as ring24.cuh) -----------------------------------------------------------------------
R96_FN void homogenize(T (&c)[D]) {  // GOLD/ntt.rs:326-334, 349-430
    c[4] = mul_pow2<96>(c[4]);
    c[7] = mul_w<2>(c[7]);   c[8] = mul_w<4>(c[8]);
    c[10] = mul_w<6>(c[10]); c[11] = mul_w<12>(c[11]);
    swap_scale<3, 1>(c[13], c[14]);
    swap_scale<11, 5>(c[16], c[17]);
    swap_scale<7, 3>(c[19], c[20]);
    swap_scale<15, 7>(c[22], c[23]);
}
R96_FN void dehomogenize(T (&c)[D]) {  // GOLD/ntt.rs:337-346, 355-437
    c[4] = mul_pow2<96>(c[4]);
    c[7] = mul_w<22>(c[7]);   c[8] = mul_w<20>(c[8]);
    c[10] = mul_w<18>(c[10]); c[11] = mul_w<12>(c[11]);
    swap_scale<23, 21>(c[13], c[14]);
    swap_scale<19, 13>(c[16], c[17]);
    swap_scale<21, 17>(c[19], c[20]);
    swap_scale<17, 9>(c[22], c[23]);
}
// layers 2 and 3 and the twist of the forward transform (GOLD/ntt.rs:160-225, 326-334); layer 1 is the caller's
R96_FN void crt_tail(T (&c)[D]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        bf_fwd<2>(c[i], c[6 + i]);
        bf_fwd<10>(c[12 + i], c[18 + i]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        bf_fwd<1>(c[i], c[3 + i]);
        bf_fwd<7>(c[6 + i], c[9 + i]);
        bf_fwd<5>(c[12 + i], c[15 + i]);
        bf_fwd<11>(c[18 + i], c[21 + i]);
    }
    homogenize(c);
}
// 24 coefficients (any 64-bit representatives) -> 8 x Fq3 (index slot*3 + component), canonical.  GOLD/ntt.rs:135-228
R96_FN void crt24(u64 (&x)[D]) {
    T c[D];
#pragma unroll
    for (int i = 0; i < 12; ++i) {  // mod X^12 - zeta, X^12 - zeta^5 = X^12 - (1 - zeta); zeta = w^4 = -2^64   :146-152
        const T a = from_u64(x[i]), b = from_u64(x[12 + i]);
        const T zb = mul_pow2<64>(b);  // word-aligned: no funnel shifts
        c[i] = sub(a, zb);
        c[12 + i] = add(add(a, b), zb);
    }
    crt_tail(c);
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = to_fq<true>(c[i]);
}
// The same for small signed coefficients (|d| < 2^15: base-B limbs, bit planes): layer 1 is plain int64 arithmetic
// (zeta b = b - (b << 32) exactly), the results enter the ring as signed integers.  MONT: results times 2^64.
template <bool MONT>
R96_FN void crt24_small(const int (&d)[D], u64 (&x)[D]) {
    T c[D];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const long long a = d[i], b = d[12 + i];
        const long long zb = b - (long long)((u64)b << 32);
        c[i] = from_i64(a + zb);
        c[12 + i] = from_i64(a + b - zb);
    }
    crt_tail(c);
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = to_fq<true>(MONT ? mul_pow2<64>(c[i]) : c[i]);
}
// 8 x Fq3 -> 24 coefficients, canonical.  GOLD/ntt.rs:240-319
R96_FN void icrt24(u64 (&x)[D]) {
    T c[D];
#pragma unroll
    for (int i = 0; i < D; ++i) c[i] = from_u64(x[i]);
    dehomogenize(c);
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // :250-283
        bf_inv<23>(c[i], c[3 + i]);
        bf_inv<17>(c[6 + i], c[9 + i]);
        bf_inv<19>(c[12 + i], c[15 + i]);
        bf_inv<13>(c[18 + i], c[21 + i]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // :289-307
        bf_inv<22>(c[i], c[6 + i]);
        bf_inv<14>(c[12 + i], c[18 + i]);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {  // :310-317; 1/8 = 2^189, 1/4 = 2^190; the one general multiplication per pair
        const T a = c[i], b = c[12 + i];
        const T kd = mul_u64(KAPPA, to_fq<false>(sub(a, b)));
        x[i] = to_fq<true>(mul_pow2<189>(sub(add(a, b), kd)));
        x[12 + i] = to_fq<true>(mul_pow2<190>(kd));
    }
}

}  // namespace r96

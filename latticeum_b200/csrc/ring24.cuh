// The Goldilocks cyclotomic ring R_q = Z_q[X]/(X^24 - X^12 + 1) (72nd cyclotomic, d = 24) and its CRT form
// (8 slots x Fq3), one element per thread, entirely in registers.
//
// Follows crates/stark-rings/crates/ring/src/cyclotomic_ring/models/goldilocks/ntt.rs:
//   crt24   <- serial_goldilock_crt_in_place   :135-228  (+ homogenize_fq3   :326-334, 349-430)
//   icrt24  <- serial_goldilock_icrt_in_place  :240-319  (+ dehomogenize_fq3 :337-346, 355-437)
// Every twiddle is a power of two (omega = 2^40), so the forward transform has no general multiply; the
// inverse has 12 (KAPPA).  Both maps are Fq-linear with canonical constants, so they commute with the
// Montgomery scaling x -> x * 2^64: feeding Montgomery-form limbs yields Montgomery-form results.
//
// Balanced digit decomposition follows crates/stark-rings/crates/ring/src/balanced_decomposition/mod.rs:62-103
// with fq_convertible.rs:22-49 (signed representative) and linear_algebra/src/ops.rs:64-80 (rounded_div).
#pragma once
#include "goldilocks.cuh"

namespace ring {
using gl::u32;
using gl::u64;

constexpr int D = 24;      // ntt.rs:9
constexpr int NSLOT = 8;   // ntt.rs:12
constexpr u64 KAPPA = 12297829382473034411ull;  // ntt.rs:43: (2*zeta - 1)^-1, zeta = omega^4

template <int I>
__device__ __forceinline__ void bf_fwd(u64 &a, u64 &b) {  // (a, b) -> (a + w^I b, a - w^I b)
    u64 t = gl::mul_w<I>(b);
    b = gl::sub(a, t);
    a = gl::add(a, t);
}
template <int I>
__device__ __forceinline__ void bf_inv(u64 &a, u64 &b) {  // (a, b) -> (a + b, w^I (a - b))
    u64 d = gl::sub(a, b);
    a = gl::add(a, b);
    b = gl::mul_w<I>(d);
}
template <int I1, int I2>
__device__ __forceinline__ void swap_scale(u64 &c1, u64 &c2) {  // c1' = w^I1 c2, c2' = w^I2 c1
    u64 t = c1;
    c1 = gl::mul_w<I1>(c2);
    c2 = gl::mul_w<I2>(t);
}

__device__ __forceinline__ void homogenize(u64 (&c)[D]) {  // ntt.rs:326-334
    c[4] = gl::neg(c[4]);                                   // slot 1 (NR^13)   :350-352
    c[7] = gl::mul_w<2>(c[7]);   c[8] = gl::mul_w<4>(c[8]);   // slot 2 (NR^7)    :360-363
    c[10] = gl::mul_w<6>(c[10]); c[11] = gl::mul_w<12>(c[11]); // slot 3 (NR^19)   :372-375
    swap_scale<3, 1>(c[13], c[14]);                          // slot 4 (NR^5)    :384-388
    swap_scale<11, 5>(c[16], c[17]);                         // slot 5 (NR^17)   :398-402
    swap_scale<7, 3>(c[19], c[20]);                          // slot 6 (NR^11)   :412-416
    swap_scale<15, 7>(c[22], c[23]);                         // slot 7 (NR^23)   :426-430
}
__device__ __forceinline__ void dehomogenize(u64 (&c)[D]) {  // ntt.rs:337-346
    c[4] = gl::neg(c[4]);
    c[7] = gl::mul_w<22>(c[7]);   c[8] = gl::mul_w<20>(c[8]);
    c[10] = gl::mul_w<18>(c[10]); c[11] = gl::mul_w<12>(c[11]);
    swap_scale<23, 21>(c[13], c[14]);
    swap_scale<19, 13>(c[16], c[17]);
    swap_scale<21, 17>(c[19], c[20]);
    swap_scale<17, 9>(c[22], c[23]);
}

// 24 coefficients -> 8 x Fq3 (index slot*3 + component), in place.
__device__ __forceinline__ void crt24(u64 (&c)[D]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {  // mod X^12 - zeta, X^12 - zeta^5 with zeta^5 = 1 - zeta   ntt.rs:146-152
        u64 a = c[i], b = c[12 + i];
        u64 zb = gl::mul_w<4>(b);
        c[i] = gl::add(a, zb);
        c[12 + i] = gl::sub(gl::add(a, b), zb);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // ntt.rs:160-179
        bf_fwd<2>(c[i], c[6 + i]);
        bf_fwd<10>(c[12 + i], c[18 + i]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // ntt.rs:186-225
        bf_fwd<1>(c[i], c[3 + i]);
        bf_fwd<7>(c[6 + i], c[9 + i]);
        bf_fwd<5>(c[12 + i], c[15 + i]);
        bf_fwd<11>(c[18 + i], c[21 + i]);
    }
    homogenize(c);
}

// Forward CRT of an element whose coefficients are SMALL signed integers (|d| < 2^15: base-B limbs, bit planes).
// Layer 1 uses zeta = w^4 = 2^160 = -2^64 = -(2^32 - 1) (mod q), so zeta*b = b - (b << 32) is an exact 48-bit
// signed integer and the whole first layer is plain int64 arithmetic (no modular reduction); the results
// (|x| < 2^48) are mapped into the field once, then layers 2 and 3 and the twist run as in crt24.
// Same values as crt24(from_small(d)): the maps are Fq-linear and the integers are exact.
template <bool MONT>
__device__ __forceinline__ void crt24_small(const int (&d)[D], u64 (&c)[D]) {
    long long x[D];
#pragma unroll
    for (int i = 0; i < 12; ++i) {  // ntt.rs:146-152 with zb = zeta * b
        long long a = d[i], b = d[12 + i];
        long long zb = b - (b << 32);
        x[i] = a + zb;
        x[12 + i] = a + b - zb;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        // signed |x| < 2^48 -> field element (canonical), then to Montgomery form if the caller uses it
        u64 m = (u64)(x[i] < 0 ? -x[i] : x[i]);
        u64 v = (x[i] < 0 && m) ? gl::Q - m : m;
        c[i] = MONT ? gl::to_mont(v) : v;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // ntt.rs:160-179
        bf_fwd<2>(c[i], c[6 + i]);
        bf_fwd<10>(c[12 + i], c[18 + i]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // ntt.rs:186-225
        bf_fwd<1>(c[i], c[3 + i]);
        bf_fwd<7>(c[6 + i], c[9 + i]);
        bf_fwd<5>(c[12 + i], c[15 + i]);
        bf_fwd<11>(c[18 + i], c[21 + i]);
    }
    homogenize(c);
}

// 8 x Fq3 -> 24 coefficients, in place.
__device__ __forceinline__ void icrt24(u64 (&c)[D]) {
    dehomogenize(c);
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // ntt.rs:250-283
        bf_inv<23>(c[i], c[3 + i]);
        bf_inv<17>(c[6 + i], c[9 + i]);
        bf_inv<19>(c[12 + i], c[15 + i]);
        bf_inv<13>(c[18 + i], c[21 + i]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // ntt.rs:289-307
        bf_inv<22>(c[i], c[6 + i]);
        bf_inv<14>(c[12 + i], c[18 + i]);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {  // ntt.rs:310-317; 1/8 = 2^189, 1/4 = 2^190
        u64 a = c[i], b = c[12 + i];
        u64 kd = gl::mul(KAPPA, gl::sub(a, b));
        c[i] = gl::mul_pow2<189>(gl::sub(gl::add(a, b), kd));
        c[12 + i] = gl::mul_pow2<190>(kd);
    }
}

// ---- balanced digits -------------------------------------------------------------------------------------
// Signed representative in [-(q-1)/2, (q-1)/2] as (negative?, magnitude).  fq_convertible.rs:22-34
__device__ __forceinline__ void signed_rep(u64 v, bool &negative, u64 &mag) {
    negative = v > gl::Q_HALF;
    mag = negative ? gl::Q - v : v;
}

// Digits base b = 2^LOG2B of the signed representative, L of them, as small signed ints.
// The reference works on the signed value with truncating % and /; the digit sequence of -m is the negated
// digit sequence of m (mod.rs:76-92: rem and carry both flip sign), so we decompose the magnitude and re-sign.
// Tie rule: |rem| == b/2 is kept, not carried (mod.rs:79).  Returns false if the value needs more than L
// digits (the reference panics on out[current_i], mod.rs:80).
template <int LOG2B, int L>
__device__ __forceinline__ bool balanced_digits(u64 v, int (&digit)[L]) {
    bool negative;
    u64 m;
    signed_rep(v, negative, m);
    constexpr u64 B = 1ull << LOG2B, HALF = B >> 1;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        u64 rem = m & (B - 1);
        m >>= LOG2B;
        int dg;
        if (rem > HALF) {
            dg = (int)rem - (int)B;
            m += 1;
        } else {
            dg = (int)rem;
        }
        digit[l] = negative ? -dg : dg;
    }
    return m == 0;
}

}  // namespace ring

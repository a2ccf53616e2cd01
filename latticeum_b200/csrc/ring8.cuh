// The Goldilocks ring transforms with EIGHT LANES PER RING ELEMENT ("octet"): lane sl of an octet holds array
// positions 3*sl .. 3*sl+2, i.e. one CRT slot (an Fq3) in CRT form, or three consecutive coefficients in coefficient
// form.  The three butterfly layers of the reference's CRT have strides 12 / 6 / 3 in array positions, which are
// lane distances 4 / 2 / 1: each layer is one __shfl_xor per value, and the per-slot "homogenize" twist is lane
// local.  Compared with one thread per element (ring24.cuh) this gives 8x the parallelism for the small vectors of
// the fold step (19 763 and 98 815 elements) and makes every global access a contiguous 24/48-byte piece per lane.
//
// Same maths, same order of operations as
//   crates/stark-rings/crates/ring/src/cyclotomic_ring/models/goldilocks/ntt.rs:135-228 (CRT), :240-319 (iCRT),
//   :326-437 (homogenize / dehomogenize);
// twiddles are lane dependent here, so they are ordinary field multiplications by constants 2^k mod q held in
// registers (still no table of roots in memory: the constants are generated at compile time from omega = 2^40).
#pragma once
#include "goldilocks.cuh"

namespace ring8 {
using gl::u32;
using gl::u64;

constexpr u64 KAPPA = 12297829382473034411ull;  // ntt.rs:43: (2*zeta - 1)^-1

// 2^k mod q at compile time (k < 192)
constexpr u64 dbl_mod(u64 x) {
    // 2x mod q for x < q without 128-bit arithmetic
    u64 hi = x >> 63, lo = x << 1;           // 2x = hi * 2^64 + lo,  2^64 = 2^32 - 1 (mod q)
    u64 r = lo;
    if (hi) {
        u64 t = r + gl::EPS;                 // cannot wrap twice: lo <= 2^64 - 2, handled below
        r = (t < r) ? t + gl::EPS : t;
    }
    return r >= gl::Q ? r - gl::Q : r;
}
constexpr u64 pow2mod(int k) {
    u64 x = 1;
    for (int i = 0; i < k; ++i) x = dbl_mod(x);
    return x;
}
// ROOTS_OF_UNITY_24[i] = 2^(8 * (5 i mod 24))   (ntt.rs:15-40)
constexpr u64 Wc(int i) { return pow2mod(8 * ((5 * (i % 24)) % 24)); }
template <int I>
struct WConst {
    static constexpr u64 value = Wc(I);  // forced compile-time evaluation (usable in device code)
};
template <int K>
struct P2Const {
    static constexpr u64 value = pow2mod(K);
};
#define W(i) (ring8::WConst<(i)>::value)

struct Twiddles {
    // forward
    u64 f2, f3, h1, h2;
    // inverse
    u64 i3, i2, i1, d1, d2;
    bool swap, hi4, hi2, hi1;
};

__device__ __forceinline__ u64 sel8(u32 sl, u64 a0, u64 a1, u64 a2, u64 a3, u64 a4, u64 a5, u64 a6, u64 a7) {
    u64 lo = (sl & 1) ? ((sl & 2) ? a3 : a1) : ((sl & 2) ? a2 : a0);
    u64 hi = (sl & 1) ? ((sl & 2) ? a7 : a5) : ((sl & 2) ? a6 : a4);
    return (sl & 4) ? hi : lo;
}

__device__ __forceinline__ Twiddles make_twiddles(u32 sl) {
    Twiddles t;
    t.hi4 = sl & 4;
    t.hi2 = sl & 2;
    t.hi1 = sl & 1;
    t.swap = sl >= 4;
    // forward layer 2: blocks of 12 positions (lanes 0-3 / 4-7): sigma = W2 / W10          ntt.rs:160-179
    t.f2 = t.hi4 ? W(10) : W(2);
    // forward layer 3: blocks of 6 positions (lane pairs): W1, W7, W5, W11                   ntt.rs:186-225
    t.f3 = sel8(sl, W(1), W(1), W(7), W(7), W(5), W(5), W(11), W(11));
    // homogenize: c1' = H1 * (swap ? c2 : c1), c2' = H2 * (swap ? c1 : c2)                   ntt.rs:349-430
    t.h1 = sel8(sl, 1, W(12), W(2), W(6), W(3), W(11), W(7), W(15));
    t.h2 = sel8(sl, 1, 1, W(4), W(12), W(1), W(5), W(3), W(7));
    // dehomogenize                                                                            ntt.rs:355-437
    t.d1 = sel8(sl, 1, W(12), W(22), W(18), W(23), W(19), W(21), W(17));
    t.d2 = sel8(sl, 1, 1, W(20), W(12), W(21), W(13), W(17), W(9));
    // inverse layer 3 / 2 / 1                                                                 ntt.rs:250-317
    t.i3 = sel8(sl, W(23), W(23), W(17), W(17), W(19), W(19), W(13), W(13));
    t.i2 = t.hi4 ? W(14) : W(22);
    t.i1 = t.hi4 ? P2Const<190>::value : P2Const<189>::value;  // 1/4 for the upper half, 1/8 for the lower half
    return t;
}

__device__ __forceinline__ u64 shx(u64 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Forward CRT of the element spread over an octet: c[0..2] = positions 3*sl..3*sl+2, in place.
__device__ __forceinline__ void crt8(u64 (&c)[3], const Twiddles &t) {
    // layer 1 (lane distance 4): lower gets a + zeta b, upper gets a + b - zeta b, zeta = W4     ntt.rs:146-152
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u64 zb = gl::mul_w<4>(c[k]);             // meaningful on the upper lanes (they hold b)
        u64 r = shx(t.hi4 ? zb : c[k], 4);       // lower receives zeta*b, upper receives a
        c[k] = t.hi4 ? gl::sub(gl::add(r, c[k]), zb) : gl::add(c[k], r);
    }
    // layer 2 (lane distance 2): (a, b) -> (a + s b, a - s b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u64 tb = gl::mul(c[k], t.f2);
        u64 r = shx(t.hi2 ? tb : c[k], 2);
        c[k] = t.hi2 ? gl::sub(r, tb) : gl::add(c[k], r);
    }
    // layer 3 (lane distance 1)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u64 tb = gl::mul(c[k], t.f3);
        u64 r = shx(t.hi1 ? tb : c[k], 1);
        c[k] = t.hi1 ? gl::sub(r, tb) : gl::add(c[k], r);
    }
    // homogenize (lane local)
    u64 x1 = t.swap ? c[2] : c[1], x2 = t.swap ? c[1] : c[2];
    c[1] = gl::mul(x1, t.h1);
    c[2] = gl::mul(x2, t.h2);
}

// Inverse CRT, in place.
__device__ __forceinline__ void icrt8(u64 (&c)[3], const Twiddles &t) {
    // dehomogenize (lane local)
    u64 x1 = t.swap ? c[2] : c[1], x2 = t.swap ? c[1] : c[2];
    c[1] = gl::mul(x1, t.d1);
    c[2] = gl::mul(x2, t.d2);
    // layer 3 (lane distance 1): (a, b) -> (a + b, w (a - b))                                     ntt.rs:250-283
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u64 r = shx(c[k], 1);
        c[k] = t.hi1 ? gl::mul(gl::sub(r, c[k]), t.i3) : gl::add(c[k], r);
    }
    // layer 2 (lane distance 2)                                                                   ntt.rs:289-307
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u64 r = shx(c[k], 2);
        c[k] = t.hi2 ? gl::mul(gl::sub(r, c[k]), t.i2) : gl::add(c[k], r);
    }
    // layer 1 (lane distance 4): kd = KAPPA (a - b); lower = (a + b - kd)/8, upper = kd/4         ntt.rs:310-317
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        u64 r = shx(c[k], 4);
        u64 a = t.hi4 ? r : c[k], b = t.hi4 ? c[k] : r;
        u64 kd = gl::mul(KAPPA, gl::sub(a, b));
        u64 v = t.hi4 ? kd : gl::sub(gl::add(a, b), kd);
        c[k] = gl::mul(v, t.i1);
    }
}

#undef W
}  // namespace ring8

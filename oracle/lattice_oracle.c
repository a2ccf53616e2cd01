/*
 * CPU oracle (plain C) for the Ajtai-commitment hot path of Nesquiko/Latticeum.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under latticeum_b200/ links, loads or calls this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may, and there
 * only as the checker / the CPU arm -- never as the thing shipped.
 *
 * It restates (does not copy) the reference's algorithm, keeping the reference's parallel structure so
 * that it is also an honest CPU baseline: rayon-over-rows mat-vec (LINALG/matrix.rs:168-178), rayon over
 * elements in the decompositions (RING/balanced_decomposition/mod.rs:136,169), SERIAL CRT/iCRT per
 * vector (RING/cyclotomic_ring/crt.rs:10-49 has no parallel path) -- plus `_par` variants reported
 * separately.  Each function cites the reference file:line it follows; paths relative to
 * /root/reference/latticeum/, with
 *   GOLD   = crates/stark-rings/crates/ring/src/cyclotomic_ring/models/goldilocks
 *   RING   = crates/stark-rings/crates/ring/src
 *   LINALG = crates/stark-rings/crates/linear_algebra/src
 *   LF     = crates/latticefold/src
 *
 * Parity pinning: tests/test_oracle_c.py checks every exported function against the reference's own
 * known-answer vectors (tests/golden/reference_kats.json) and against oracle/lattice_oracle.py.
 * The field arithmetic of the reference lives in ark-ff 0.5.0 (Cargo.lock:279-282, not vendored): exact
 * arithmetic mod q, restated from the maths.  Sampler parity (AjtaiCommitmentScheme::rand) is unpinned.
 *
 * All values are canonical u64 in [0, q).  A ring element is 24 contiguous u64: coefficient form index =
 * degree; CRT form index = slot*3 + component (RING/cyclotomic_ring/flatten.rs:10-17).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define Q 0xFFFFFFFF00000001ULL /* GOLD/mod.rs:21 */
#define Q_HALF ((Q - 1) / 2)
#define D 24   /* GOLD/ntt.rs:9  */
#define NS 8   /* GOLD/ntt.rs:12 */

/* status codes shared with include/lattice_ajtai.h */
#define LO_OK 0
#define LO_E_WRONG_WITNESS_LENGTH 1
#define LO_E_DIGIT_OVERFLOW 4
#define LO_E_INVALID_ARG 5

/* ---- a1. Z_q --------------------------------------------------------------------------------- */
static inline u64 fadd(u64 a, u64 b) { u128 s = (u128)a + b; return (u64)(s >= Q ? s - Q : s); }
static inline u64 fsub(u64 a, u64 b) { return a >= b ? a - b : a + (Q - b); }
static inline u64 fneg(u64 a) { return a ? Q - a : 0; }
/* 2^64 = 2^32 - 1 and 2^96 = -1 (mod q): x = lo + 2^64*(hi_lo + 2^32*hi_hi) = lo - hi_hi + hi_lo*(2^32-1).
 * (Any exact reduction gives the reference's value; ark-ff uses Montgomery reduction instead.) */
static inline u64 reduce128(u128 x) {
    const u64 EPS = 0xFFFFFFFFULL;
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= EPS;
    u64 t1 = hi_lo * EPS;
    u64 r = t0 + t1;
    if (r < t1) r += EPS;
    return r >= Q ? r - Q : r;
}
static inline u64 fmul(u64 a, u64 b) { return reduce128((u128)a * b); }

/* ROOTS_OF_UNITY_24[i] = (2^40)^i, GOLD/ntt.rs:15-40 (computed, checked against the literals in tests) */
static u64 W[24];
static const u64 KAPPA = 12297829382473034411ULL;     /* GOLD/ntt.rs:43 (inverse of 2*zeta-1) */
static const u64 EIGHT_INV = 16140901060737761281ULL; /* GOLD/ntt.rs:45 */
static const u64 FOUR_INV = 13835058052060938241ULL;  /* GOLD/ntt.rs:47 */
static const u64 NONRESIDUE = 1ULL << 40;             /* GOLD/mod.rs:42 */
static int g_init = 0;

static void lo_init(void) {
    if (g_init) return;
    W[0] = 1;
    for (int i = 1; i < 24; ++i) W[i] = fmul(W[i - 1], NONRESIDUE);
    g_init = 1;
}
__attribute__((constructor)) static void lo_ctor(void) { lo_init(); }

void lo_roots(u64 *out24) { lo_init(); memcpy(out24, W, sizeof(W)); }
int lo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm of the benchmark asks for all cores explicitly */
void lo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}


/* ---- a1. Fq3 = Fq[u]/(u^3 - 2^40): value fixed by the maths (ark-ff Fp3) ------------------------ */
static inline void fq3_mul(const u64 *a, const u64 *b, u64 *c) {
    u64 a0 = a[0], a1 = a[1], a2 = a[2], b0 = b[0], b1 = b[1], b2 = b[2];
    u64 t12 = fadd(fmul(a1, b2), fmul(a2, b1));
    u64 c0 = fadd(fmul(a0, b0), fmul(NONRESIDUE, t12));
    u64 c1 = fadd(fadd(fmul(a0, b1), fmul(a1, b0)), fmul(NONRESIDUE, fmul(a2, b2)));
    u64 c2 = fadd(fadd(fmul(a0, b2), fmul(a1, b1)), fmul(a2, b0));
    c[0] = c0; c[1] = c1; c[2] = c2;
}

/* ---- a4. CRT of one element, in place ------------------------------------------------------------ */
/* homogenize_fq3: GOLD/ntt.rs:326-334 with the per-slot maps of :349-430 */
static void homogenize(u64 *c) {
    c[4] = fneg(c[4]);
    c[7] = fmul(c[7], W[2]);   c[8] = fmul(c[8], W[4]);
    c[10] = fmul(c[10], W[6]); c[11] = fmul(c[11], W[12]);
    static const int base[4] = {12, 15, 18, 21};
    static const int m1[4] = {3, 11, 7, 15}, m2[4] = {1, 5, 3, 7};
    for (int k = 0; k < 4; ++k) {
        u64 c1 = c[base[k] + 1];
        c[base[k] + 1] = fmul(c[base[k] + 2], W[m1[k]]);
        c[base[k] + 2] = fmul(c1, W[m2[k]]);
    }
}
/* dehomogenize_fq3: GOLD/ntt.rs:337-346 with :355-437 */
static void dehomogenize(u64 *c) {
    c[4] = fneg(c[4]);
    c[7] = fmul(c[7], W[22]);   c[8] = fmul(c[8], W[20]);
    c[10] = fmul(c[10], W[18]); c[11] = fmul(c[11], W[12]);
    static const int base[4] = {12, 15, 18, 21};
    static const int m1[4] = {23, 19, 21, 17}, m2[4] = {21, 13, 17, 9};
    for (int k = 0; k < 4; ++k) {
        u64 c1 = c[base[k] + 1];
        c[base[k] + 1] = fmul(c[base[k] + 2], W[m1[k]]);
        c[base[k] + 2] = fmul(c1, W[m2[k]]);
    }
}
/* serial_goldilock_crt_in_place: GOLD/ntt.rs:135-228 */
static void crt_one(u64 *c) {
    for (int i = 0; i < 12; ++i) { /* :146-152, zeta = W[4], zeta^5 = 1 - zeta */
        u64 a = c[i], b = c[12 + i], zb = fmul(W[4], b);
        c[i] = fadd(a, zb);
        c[12 + i] = fsub(fadd(a, b), zb);
    }
    for (int i = 0; i < 6; ++i) { /* :160-179 */
        u64 a = c[i], t = fmul(W[2], c[6 + i]);
        c[i] = fadd(a, t); c[6 + i] = fsub(a, t);
        a = c[12 + i]; t = fmul(W[10], c[18 + i]);
        c[12 + i] = fadd(a, t); c[18 + i] = fsub(a, t);
    }
    static const int base[4] = {0, 6, 12, 18}, tw[4] = {1, 7, 5, 11};
    for (int i = 0; i < 3; ++i) /* :186-225 */
        for (int k = 0; k < 4; ++k) {
            u64 a = c[base[k] + i], t = fmul(W[tw[k]], c[base[k] + 3 + i]);
            c[base[k] + i] = fadd(a, t); c[base[k] + 3 + i] = fsub(a, t);
        }
    homogenize(c);
}
/* serial_goldilock_icrt_in_place: GOLD/ntt.rs:240-319 */
static void icrt_one(u64 *c) {
    dehomogenize(c);
    static const int base[4] = {0, 6, 12, 18}, tw[4] = {23, 17, 19, 13};
    for (int i = 0; i < 3; ++i) /* :250-283 */
        for (int k = 0; k < 4; ++k) {
            u64 a = c[base[k] + i], b = c[base[k] + 3 + i];
            c[base[k] + i] = fadd(a, b);
            c[base[k] + 3 + i] = fmul(W[tw[k]], fsub(a, b));
        }
    for (int i = 0; i < 6; ++i) { /* :289-307 */
        u64 a = c[i], b = c[6 + i];
        c[i] = fadd(a, b); c[6 + i] = fmul(W[22], fsub(a, b));
        a = c[12 + i]; b = c[18 + i];
        c[12 + i] = fadd(a, b); c[18 + i] = fmul(W[14], fsub(a, b));
    }
    for (int i = 0; i < 12; ++i) { /* :310-317 */
        u64 a = c[i], b = c[12 + i], kd = fmul(KAPPA, fsub(a, b));
        c[i] = fmul(EIGHT_INV, fsub(fadd(a, b), kd));
        c[12 + i] = fmul(FOUR_INV, kd);
    }
}

/* test hooks for the pre-homogenize KAT layout of GOLD/ntt.rs:563-787 */
void lo_homogenize(u64 *c) { lo_init(); homogenize(c); }
void lo_dehomogenize(u64 *c) { lo_init(); dehomogenize(c); }

/* ---- a6. elementwise CRT / iCRT: serial like RING/cyclotomic_ring/crt.rs:10-49 --------------------- */
void lo_crt(const u64 *in, u64 count, u64 *out) {
    lo_init();
    if (in != out) memcpy(out, in, count * D * sizeof(u64));
    for (u64 e = 0; e < count; ++e) crt_one(out + e * D);
}
void lo_icrt(const u64 *in, u64 count, u64 *out) {
    lo_init();
    if (in != out) memcpy(out, in, count * D * sizeof(u64));
    for (u64 e = 0; e < count; ++e) icrt_one(out + e * D);
}
/* all-core variants (the reference has none; reported separately so the GPU is not flattered) */
void lo_crt_par(const u64 *in, u64 count, u64 *out) {
    lo_init();
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)count; ++e) {
        if (in != out) memcpy(out + e * D, in + e * D, D * sizeof(u64));
        crt_one(out + e * D);
    }
}
void lo_icrt_par(const u64 *in, u64 count, u64 *out) {
    lo_init();
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)count; ++e) {
        if (in != out) memcpy(out + e * D, in + e * D, D * sizeof(u64));
        icrt_one(out + e * D);
    }
}

/* ---- a7/a8. balanced digits ------------------------------------------------------------------------ */
/* RING/balanced_decomposition/fq_convertible.rs:22-34 then mod.rs:62-103 with LINALG/ops.rs:64-80.
 * Works on i128 like the reference; `%` and `/` in C truncate toward zero like Rust's. */
static int decompose_balanced(u64 v, u64 b_, int padding, u64 *out, int stride) {
    __int128 cur = v > Q_HALF ? (__int128)v - (__int128)Q : (__int128)v;
    const __int128 b = (__int128)b_, half = b / 2;
    int i = 0;
    for (;;) {
        __int128 rem = cur % b, digit;
        __int128 arem = rem < 0 ? -rem : rem;
        if (arem <= half) {
            digit = rem;
            cur /= b;
        } else {
            digit = rem < 0 ? rem + b : rem - b;
            /* rounded_div(rem, b): same sign -> (rem + b/2)/b else (rem - b/2)/b, LINALG/ops.rs:75-79 */
            __int128 carry = ((rem ^ b) >= 0) ? (rem + half) / b : (rem - half) / b;
            cur = cur / b + carry;
        }
        if (i >= padding) return LO_E_DIGIT_OVERFLOW; /* the reference panics: out[current_i] unchecked */
        out[(size_t)i * stride] = digit < 0 ? (u64)((__int128)Q + digit) : (u64)digit; /* fq_convertible.rs:38-49 */
        ++i;
        if (cur == 0) break;
    }
    for (; i < padding; ++i) out[(size_t)i * stride] = 0;
    return LO_OK;
}

/* a9+a10. GadgetDecompose for &[R]: out[i*L + l] = limb l of v[i]; coefficient-wise.
 * RING/balanced_decomposition/mod.rs:163-175; RING/cyclotomic_ring/coeff_form.rs:588-606.  rayon over elements. */
int lo_gadget_decompose(const u64 *in, u64 count, u64 b, int L, u64 *out) {
    if (b < 2 || (b & 1)) return LO_E_INVALID_ARG;
    int status = LO_OK;
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)count; ++e)
        for (int c = 0; c < D; ++c) {
            int st = decompose_balanced(in[e * D + c], b, L, out + (size_t)e * L * D + c, D);
            if (st != LO_OK) {
#pragma omp atomic write
                status = st;
            }
        }
    return status;
}

/* a11. decompose_B_vec_into_k_vec = decompose_to_vec(b, K).transpose(): planes[k][j] = digit k of x[j].
 * LF/nifs/decomposition/utils.rs:45-49; RING/balanced_decomposition/mod.rs:119-140; LINALG/ops.rs:13-34 */
int lo_decompose_planes(const u64 *f_coeff, u64 n, u64 b, int K, u64 *planes) {
    if (b < 2 || (b & 1)) return LO_E_INVALID_ARG;
    int status = LO_OK;
#pragma omp parallel for schedule(static)
    for (long long j = 0; j < (long long)n; ++j)
        for (int c = 0; c < D; ++c) {
            int st = decompose_balanced(f_coeff[j * D + c], b, K, planes + (size_t)j * D + c, (int)(n * D));
            if (st != LO_OK) {
#pragma omp atomic write
                status = st;
            }
        }
    return status;
}

/* a12. gadget_recompose in CRT form: Horner sum b^l * v[l] over chunks of L with b = scalar in all slots.
 * RING/balanced_decomposition/mod.rs:105-117,177-190 */
void lo_gadget_recompose_ntt(const u64 *f, u64 n, u64 b, int L, u64 *out) {
    u64 count = n / (u64)L;
    u64 bq = b % Q;
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)count; ++e) {
        u64 acc[D];
        memset(acc, 0, sizeof(acc));
        for (int l = L - 1; l >= 0; --l)
            for (int t = 0; t < D; ++t) /* (b,0,0) * (x0,x1,x2) = (b x0, b x1, b x2) */
                acc[t] = fadd(fmul(acc[t], bq), f[((size_t)e * L + l) * D + t]);
        memcpy(out + (size_t)e * D, acc, sizeof(acc));
    }
}

/* ---- a13. Matrix::checked_mul_vec: rayon over rows, serial left fold over columns -------------------- */
/* LINALG/matrix.rs:168-178 with RqNTT Mul/Sum (RING/cyclotomic_ring/ntt_form.rs:159-175,521-536,640-646).
 * A is kappa x n x 24 contiguous, row-major. */
int lo_commit(const u64 *A, u64 kappa, u64 n, const u64 *f, u64 f_len, u64 *cm) {
    if (f_len != n) return LO_E_WRONG_WITNESS_LENGTH; /* LF/commitment/commitment_scheme.rs:64-77 */
#pragma omp parallel for schedule(dynamic, 1)
    for (long long i = 0; i < (long long)kappa; ++i) {
        u64 acc[D];
        memset(acc, 0, sizeof(acc));
        const u64 *row = A + (size_t)i * n * D;
        for (u64 j = 0; j < n; ++j) {
            const u64 *a = row + j * D, *v = f + j * D;
            for (int s = 0; s < NS; ++s) {
                u64 p[3];
                fq3_mul(a + 3 * s, v + 3 * s, p);
                acc[3 * s] = fadd(acc[3 * s], p[0]);
                acc[3 * s + 1] = fadd(acc[3 * s + 1], p[1]);
                acc[3 * s + 2] = fadd(acc[3 * s + 2], p[2]);
            }
        }
        memcpy(cm + (size_t)i * D, acc, sizeof(acc));
    }
    return LO_OK;
}

/* ---- a17. Witness::from_w_ccs: iCRT -> gadget_decompose(B,L) -> CRT ; LF/arith.rs:230-248 ------------- */
int lo_witness_from_w_ccs(const u64 *w_ccs, u64 w_len, u64 B, int L, u64 *f_coeff, u64 *f) {
    u64 *w_coeff = (u64 *)malloc((size_t)w_len * D * sizeof(u64));
    if (!w_coeff) return LO_E_INVALID_ARG;
    lo_icrt(w_ccs, w_len, w_coeff);                       /* serial, as the reference */
    int st = lo_gadget_decompose(w_coeff, w_len, B, L, f_coeff);
    free(w_coeff);
    if (st != LO_OK) return st;
    lo_crt(f_coeff, w_len * (u64)L, f);                   /* serial, as the reference */
    return LO_OK;
}

/* ---- a18. decompose_witness + commit_witnesses: LF/nifs/decomposition.rs:162-201 ------------------------
 * planes_coeff / planes_f: K x n x 24 (either may be NULL -> scratch); cms: K x kappa x 24 where
 * cms[1..] are matrix commits of planes 1..K-1 and cms[0] = cm - fold_rev((acc + y_i) * b). */
int lo_decompose_commit(const u64 *A, u64 kappa, u64 n, const u64 *f_coeff, const u64 *cm, u64 b, int K,
                        u64 *planes_coeff, u64 *planes_f, u64 *cms) {
    size_t plane_sz = (size_t)n * D;
    u64 *pc = planes_coeff ? planes_coeff : (u64 *)malloc(plane_sz * K * sizeof(u64));
    u64 *pf = planes_f ? planes_f : (u64 *)malloc(plane_sz * K * sizeof(u64));
    int st = lo_decompose_planes(f_coeff, n, b, K, pc);
    if (st == LO_OK) {
        /* rayon over planes, serial CRT inside each: LF/nifs/decomposition.rs:164-166, LF/arith.rs:327 */
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < K; ++k) lo_crt(pc + k * plane_sz, n, pf + k * plane_sz);
        for (int k = 1; k < K; ++k) /* nested rayon in the reference; rows-parallel here */
            lo_commit(A, kappa, n, pf + k * plane_sz, n, cms + (size_t)k * kappa * D);
        /* y_0 by homomorphism, LF/nifs/decomposition.rs:189-197: b_sum = fold over y_{K-1}..y_1 of (acc + y)*b */
        u64 bq = b % Q;
        for (u64 t = 0; t < kappa * D; ++t) {
            u64 acc = 0;
            for (int k = K - 1; k >= 1; --k) acc = fmul(fadd(acc, cms[(size_t)k * kappa * D + t]), bq);
            cms[t] = fsub(cm[t], acc);
        }
    }
    if (!planes_coeff) free(pc);
    if (!planes_f) free(pf);
    return st;
}

/* ---- a19. compute_f_0: f_0[j] = sum_i rho_i * f_i[j] ; LF/nifs/folding.rs:258-268 ------------------------ */
void lo_compute_f0(const u64 *rho, const u64 *const *f_s, int count, u64 n, u64 *f0) {
#pragma omp parallel for schedule(static)
    for (long long j = 0; j < (long long)n; ++j) {
        u64 acc[D];
        memset(acc, 0, sizeof(acc));
        for (int i = 0; i < count; ++i)
            for (int s = 0; s < NS; ++s) {
                u64 p[3];
                fq3_mul(rho + (size_t)i * D + 3 * s, f_s[i] + (size_t)j * D + 3 * s, p);
                acc[3 * s] = fadd(acc[3 * s], p[0]);
                acc[3 * s + 1] = fadd(acc[3 * s + 1], p[1]);
                acc[3 * s + 2] = fadd(acc[3 * s + 2], p[2]);
            }
        memcpy(f0 + (size_t)j * D, acc, sizeof(acc));
    }
}

/* ---- F8. Montgomery (R = 2^64) <-> canonical: host limbs are x*2^64 mod q (GOLD/mod.rs:20-24) ------------ */
void lo_to_mont(const u64 *in, u64 count, u64 *out) {
    const u64 R = 0xFFFFFFFFULL; /* 2^64 mod q */
    for (u64 i = 0; i < count; ++i) out[i] = fmul(in[i], R);
}
void lo_from_mont(const u64 *in, u64 count, u64 *out) {
    const u64 RINV = 18446744065119617025ULL; /* 2^128 mod q = 2^-64 since 2^192 = 1 */
    for (u64 i = 0; i < count; ++i) out[i] = fmul(in[i], RINV);
}

/* ---- seeded synthetic inputs shared by oracle, tests and bench (SplitMix64) -------------------- */
static inline u64 splitmix64(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* uniform in [0,q) by rejection; element i of stream `seed` does not depend on count or threading */
void lo_fill_uniform(u64 *out, u64 count, u64 seed) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)count; ++i) {
        u64 s = seed * 0xD1342543DE82EF95ULL + (u64)i * 0x9E3779B97F4A7C15ULL;
        u64 v;
        do { v = splitmix64(&s); } while (v >= Q);
        out[i] = v;
    }
}

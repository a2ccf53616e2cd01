// Batched ring transforms and gadget decompositions for the Goldilocks ring (d = 24), sm_100a.
//
// Eight lanes own one ring element (ring8.cuh): lane sl holds CRT slot sl / coefficients 3sl..3sl+2, the three
// butterfly layers are warp shuffles at lane distance 4/2/1, and every global access is a contiguous 24-byte
// (plain layout) or 48-byte (extended layout) piece per lane, i.e. a fully used run of bytes per warp, with no
// shared-memory staging.  The vectors of a fold step are small (19 763 and 98 815 elements), so the 8x
// parallelism over one-thread-per-element is what keeps these kernels from being latency bound.
//
// Reference functions replaced (paths relative to /root/reference/latticeum/crates/):
//   crt_kernel<false>      CRT::elementwise_crt           stark-rings/crates/ring/src/cyclotomic_ring/crt.rs:10-25
//   crt_kernel<true>       ICRT::elementwise_icrt         .../cyclotomic_ring/crt.rs:34-49
//   witness_kernel         Witness::from_w_ccs: iCRT, gadget_decompose(B, L), CRT of the limbs, fused
//                          latticefold/src/arith.rs:230-248; .../ring/src/balanced_decomposition/mod.rs:163-175
//   planes_kernel          decompose_B_vec_into_k_vec + Witness::from_f_coeff's CRT
//                          latticefold/src/nifs/decomposition/utils.rs:45-49; latticefold/src/arith.rs:327
#include <cstdint>

#include "kernels.h"
#include "ring24.cuh"
#include "ring8.cuh"
#include "ring96.cuh"
#include "spin.cuh"
#include "tma.cuh"

namespace lat {
using gl::u32;

// One-thread-per-element transforms: the network in Z/(2^96 + 1) of ring96.cuh (about 40 % fewer instructions than the
// canonical 64-bit arithmetic of ring24.cuh, which -DLAT_RING24 brings back for A/B timing; same values bit for bit).
#ifdef LAT_RING24
namespace xf = ring;
#else
namespace xf = r96;
#endif

constexpr int OPB = 16;            // octets (ring elements) per block
constexpr int THREADS = OPB * 8;   // 128

struct Octet {
    u64 e;       // element index (clamped into range so that every lane can take part in the shuffles)
    u32 sl;      // slot / lane within the octet
    bool valid;  // false for the padding octets of the last block: they compute but never store
};
__device__ __forceinline__ Octet octet_of(u64 count) {
    Octet o;
    u64 e = (u64)blockIdx.x * OPB + (threadIdx.x >> 3);
    o.valid = e < count;
    o.e = o.valid ? e : count - 1;
    o.sl = threadIdx.x & 7;
    return o;
}
__device__ __forceinline__ void load3(const u64 *__restrict__ base, u64 elem, u32 sl, u64 (&c)[3]) {
    const u64 *p = base + elem * ring::D + 3 * sl;
    c[0] = p[0]; c[1] = p[1]; c[2] = p[2];
}
__device__ __forceinline__ void store3(u64 *__restrict__ base, u64 elem, u32 sl, const u64 (&c)[3]) {
    u64 *p = base + elem * ring::D + 3 * sl;
    p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
}
// one slot of the MAC kernel's extended layout: (f0, f1, f2, f0+f1, f0+f2, f1+f2), 48 B, 16-B aligned
__device__ __forceinline__ void store_fx(u64 *__restrict__ fx, u64 elem, u32 sl, const u64 (&c)[3]) {
    ulonglong2 *p = reinterpret_cast<ulonglong2 *>(fx + (elem * ring::NSLOT + sl) * 6);
    p[0] = make_ulonglong2(c[0], c[1]);
    p[1] = make_ulonglong2(c[2], gl::add_lazy(c[0], c[1]));
    p[2] = make_ulonglong2(gl::add_lazy(c[0], c[2]), gl::add_lazy(c[1], c[2]));
}

// ---- coalesced row output for the one-thread-per-element phases -------------------------------------------------
// A thread that owns a whole element would write one private 192 B / 384 B run: every store instruction of the warp
// then touches 32 different sectors with 16 B each, and L2 has to read-merge the half-written sectors (measured on
// the first planes kernel: 206 MB of DRAM reads and 699 MB of writes for 569 MB of payload, 547 us).  Instead each
// thread drops its row into a padded shared-memory tile and the block copies the tile out as one contiguous run.
// Rows are in 16-byte units; the pitch is ROW_UNITS + 1 units so that quarter-warps hit distinct banks.
constexpr int PLAIN_UNITS = ring::D / 2;  // 12 x 16 B = 192 B
constexpr int FX_UNITS = FX_WORDS / 2;    // 24 x 16 B = 384 B
__device__ __forceinline__ void row_put(ulonglong2 *tile, int row, const u64 (&c)[ring::D]) {
    ulonglong2 *p = tile + row * (PLAIN_UNITS + 1);
#pragma unroll
    for (int k = 0; k < PLAIN_UNITS; ++k) p[k] = make_ulonglong2(c[2 * k], c[2 * k + 1]);
}
__device__ __forceinline__ void row_put_fx(ulonglong2 *tile, int row, const u64 (&c)[ring::D]) {
    ulonglong2 *p = tile + row * (FX_UNITS + 1);
#pragma unroll
    for (int sl = 0; sl < ring::NSLOT; ++sl) {
        u64 f0 = c[3 * sl], f1 = c[3 * sl + 1], f2 = c[3 * sl + 2];
        p[3 * sl] = make_ulonglong2(f0, f1);
        p[3 * sl + 1] = make_ulonglong2(f2, gl::add_lazy(f0, f1));
        p[3 * sl + 2] = make_ulonglong2(gl::add_lazy(f0, f2), gl::add_lazy(f1, f2));
    }
}
// One thread writes its own extended row (384 B) with twelve 256-bit stores: every store covers a whole 32-byte sector and
// the row's three 128-byte lines are complete before they leave L2, so there is nothing to read-merge and no staging tile.
__device__ __forceinline__ void st256(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
// TOOM: the slot's derived words are its values at 1, -1, 2 (planes, read by the several-witness kernels and fold_kernel)
// instead of the three Karatsuba sums (the single witness) -- kernels.h, FX_WORDS.
template <bool TOOM = false>
__device__ __forceinline__ void store_fx_row(u64 *__restrict__ o, const u64 (&x)[ring::D]) {
#pragma unroll
    for (int s = 0; s < ring::NSLOT; s += 2) {
        const u64 a0 = x[3 * s], a1 = x[3 * s + 1], a2 = x[3 * s + 2];
        const u64 b0 = x[3 * s + 3], b1 = x[3 * s + 4], b2 = x[3 * s + 5];
        u64 p0, p1, p2, q0, q1, q2;
        if constexpr (TOOM) {
            gl::toom_eval(a0, a1, a2, p0, p1, p2);
            gl::toom_eval(b0, b1, b2, q0, q1, q2);
        } else {
            p0 = gl::add_lazy(a0, a1); p1 = gl::add_lazy(a0, a2); p2 = gl::add_lazy(a1, a2);
            q0 = gl::add_lazy(b0, b1); q1 = gl::add_lazy(b0, b2); q2 = gl::add_lazy(b1, b2);
        }
        st256(o + s * 6, a0, a1, a2, p0);
        st256(o + s * 6 + 4, p1, p2, b0, b1);
        st256(o + s * 6 + 8, b2, q0, q1, q2);
    }
}
// block-wide: copy `nrows` rows of ROW_UNITS units from the padded tile to the contiguous global run at `dst`
template <int ROW_UNITS>
__device__ __forceinline__ void rows_out(const ulonglong2 *tile, u64 *__restrict__ dst, u32 nrows) {
    __syncthreads();
    ulonglong2 *g = reinterpret_cast<ulonglong2 *>(dst);
    for (u32 u = threadIdx.x; u < nrows * ROW_UNITS; u += blockDim.x) {
        u32 r = u / ROW_UNITS, c = u - r * ROW_UNITS;
        g[u] = tile[r * (ROW_UNITS + 1) + c];
    }
    __syncthreads();
}

// ---- batched CRT / iCRT ------------------------------------------------------------------------------------
template <bool INVERSE>
__global__ void __launch_bounds__(THREADS) crt_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 count) {
    const Octet o = octet_of(count);
    const ring8::Twiddles tw = ring8::make_twiddles(o.sl);
    u64 c[3];
    load3(in, o.e, o.sl, c);
    if constexpr (INVERSE) ring8::icrt8(c, tw); else ring8::crt8(c, tw);
    if (o.valid) store3(out, o.e, o.sl, c);
}

// Large batches: one thread per element with the compile-time shift twiddles of ring24.cuh (about a fifth of the
// instructions of the 8-lane form), rows staged through shared memory so that the stores are contiguous.
constexpr int BIG_THREADS = 64;
constexpr u64 BIG_BATCH = 1u << 14;  // from here on there are enough elements to fill the machine one per thread
template <bool INVERSE>
__global__ void __launch_bounds__(BIG_THREADS) crt_big_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 count) {
    __shared__ __align__(16) ulonglong2 otile[BIG_THREADS * (ring::D / 2 + 1)];
    const u64 e0 = (u64)blockIdx.x * BIG_THREADS;
    const u64 e = e0 + threadIdx.x;
    const bool active = e < count;
    u64 c[ring::D];
    if (active) {
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(in + e * ring::D);
#pragma unroll
        for (int k = 0; k < ring::D / 2; ++k) {
            ulonglong2 v = p[k];
            c[2 * k] = v.x;
            c[2 * k + 1] = v.y;
        }
        if constexpr (INVERSE) xf::icrt24(c); else xf::crt24(c);
        ulonglong2 *r = otile + threadIdx.x * (ring::D / 2 + 1);
#pragma unroll
        for (int k = 0; k < ring::D / 2; ++k) r[k] = make_ulonglong2(c[2 * k], c[2 * k + 1]);
    }
    rows_out<PLAIN_UNITS>(otile, out + e0 * ring::D, (u32)min((u64)BIG_THREADS, count - e0));
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
void launch_crt(const u64 *in, u64 *out, u64 count, cudaStream_t stream) {
    if (!count) return;
    if (count >= BIG_BATCH && aligned16(in) && aligned16(out))
        crt_big_kernel<false><<<(unsigned)((count + BIG_THREADS - 1) / BIG_THREADS), BIG_THREADS, 0, stream>>>(in, out, count);
    else
        crt_kernel<false><<<(unsigned)((count + OPB - 1) / OPB), THREADS, 0, stream>>>(in, out, count);
}
void launch_icrt(const u64 *in, u64 *out, u64 count, cudaStream_t stream) {
    if (!count) return;
    if (count >= BIG_BATCH && aligned16(in) && aligned16(out))
        crt_big_kernel<true><<<(unsigned)((count + BIG_THREADS - 1) / BIG_THREADS), BIG_THREADS, 0, stream>>>(in, out, count);
    else
        crt_kernel<true><<<(unsigned)((count + OPB - 1) / OPB), THREADS, 0, stream>>>(in, out, count);
}

// 24 int16 (48 B, 16-B aligned) -> ints; works for global and shared pointers
__device__ __forceinline__ void load_i16x24(const int16_t *p, int (&d)[ring::D]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        uint4 x = q[v];
        u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d[v * 8 + i * 2] = (int)(int16_t)(w[i] & 0xFFFFu);
            d[v * 8 + i * 2 + 1] = (int)(int16_t)(w[i] >> 16);
        }
    }
}

// ---- Witness::from_w_ccs in one kernel: iCRT -> balanced digits base 2^log2b -> CRT of every limb -------------
// Phase A (8 lanes per w_ccs element, OPB elements per block): iCRT by warp shuffles, then the digit loop on the
// lane's three coefficients; digits go to the device-resident int16 witness and to a shared-memory tile.
// Phase B (one thread per LIMB element, OPB*L of them per block): forward CRT with compile-time shift twiddles
// (ring24.cuh: no general multiplies) straight from the tile, written in the plain and/or the MAC kernel's
// extended layout through the coalescing tile.  The small vector (19 763 elements) gets the 8x parallelism where
// it needs it, the large one (98 815 limb elements) gets the cheap transform.
// Dynamic shared memory: OPB*L rows x (PLAIN_UNITS + 1) x 16 B, only when the plain CRT-form output is wanted (the extended
// rows go out with 256-bit stores straight from registers).
constexpr int WIT_MAX_L = 8;
template <bool MONT>
__device__ __forceinline__ void witness_body(const u64 *__restrict__ w, u64 w_len, int log2b, int L, bool in_coeff,
                                             int16_t *__restrict__ f16, u64 *__restrict__ f_coeff, u64 *__restrict__ f_plain,
                                             u64 *__restrict__ fx, int *__restrict__ flag, bool stage_input) {
    __shared__ __align__(16) int16_t tile[OPB * WIT_MAX_L * ring::D];  // [octet][limb][24] = 6 KB
    __shared__ __align__(128) u64 wstage[OPB * ring::D];               // the block's input elements, 3 KB
    __shared__ __align__(8) u64 wbar;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    ulonglong2 *otile = reinterpret_cast<ulonglong2 *>(dyn_smem);
    const Octet o = octet_of(w_len);
    const bool tiled = L <= WIT_MAX_L;  // the engine's L is <= 8; only lat_ring_gadget_decompose allows more
    if (stage_input) {
        // Input in page-locked HOST memory: one bulk copy per block fetches its 16 elements (3 KB) over PCIe as a few large
        // read requests instead of 128 lanes' 8-byte loads; the lanes then read shared memory.
        const u64 e0 = (u64)blockIdx.x * OPB;
        if (threadIdx.x == 0) {
            mbar_init(&wbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const u32 bytes = (u32)min((u64)OPB, w_len - e0) * ring::D * 8;
            mbar_arrive_expect_tx(&wbar, bytes);
            tma_bulk_g2s(wstage, w + e0 * ring::D, bytes, &wbar);
        }
        __syncthreads();
        mbar_wait(&wbar, 0);
    }
    {
        const ring8::Twiddles tw = ring8::make_twiddles(o.sl);
        u64 c[3];
        if (stage_input) load3(wstage, o.e - (u64)blockIdx.x * OPB, o.sl, c);
        else load3(w, o.e, o.sl, c);
        if (!in_coeff) ring8::icrt8(c, tw);
        u32 sgn[3];   // 0 or 0xFFFFFFFF: the digit sequence of -m is the negated digit sequence of m (mod.rs:76-92)
        u64 m[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if constexpr (MONT) c[k] = gl::from_mont(c[k]);
            bool negative;
            ring::signed_rep(c[k], negative, m[k]);  // fq_convertible.rs:22-34
            sgn[k] = negative ? 0xFFFFFFFFu : 0u;
        }
        // balanced_decomposition/mod.rs:76-97 on the magnitudes, limb by limb; out[i*L + l] = limb l of element i
        // (mod.rs:163-175).  log2b <= 15, so a remainder is a 32-bit quantity: and, 64-bit shift, compare, conditional
        // subtract of B, carry into the magnitude, re-sign, store -- about a dozen instructions per digit.
        const u32 Bv = 1u << log2b, mask = Bv - 1, half = Bv >> 1;
        auto next_digit = [&](int k) -> int {
            const u32 rem = (u32)m[k] & mask;
            const u32 up = rem > half ? 1u : 0u;   // |rem| == b/2 is kept (mod.rs:79)
            m[k] = (m[k] >> log2b) + up;
            const u32 dg = rem - (up ? Bv : 0u);
            return (int)((dg ^ sgn[k]) - sgn[k]);
        };
        if (tiled) {
            int16_t *trow = tile + (threadIdx.x >> 3) * (L * ring::D) + 3 * o.sl;
            u64 *gcoeff = (o.valid && f_coeff) ? f_coeff + (o.e * (u64)L) * ring::D + 3 * o.sl : nullptr;
#pragma unroll
            for (int l = 0; l < WIT_MAX_L; ++l) {
                if (l < L) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int dg = next_digit(k);
                        trow[l * ring::D + k] = (int16_t)dg;
                        if (gcoeff) gcoeff[l * ring::D + k] = gl::from_small<MONT>(dg);
                    }
                }
            }
        } else {   // standalone decompositions with L > 8 (lat_ring_gadget_decompose only)
            for (int l = 0; l < L; ++l) {
                const u64 at = (o.e * (u64)L + l) * ring::D + 3 * o.sl;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int dg = next_digit(k);
                    if (o.valid) {
                        f16[at + k] = (int16_t)dg;
                        if (f_coeff) f_coeff[at + k] = gl::from_small<MONT>(dg);
                    }
                }
            }
        }
        // the reference would index out of bounds (mod.rs:80) if a value needed more than L digits
        if (o.valid && (m[0] | m[1] | m[2])) atomicOr(flag, 1);
    }
    if (!tiled) return;
    __syncthreads();
    const u64 e0 = (u64)blockIdx.x * OPB;
    const u32 nvalid = (u32)min((u64)OPB, w_len - e0);
    const u32 nrows = nvalid * (u32)L;  // tile rows are already in limb-element order
    {   // the int16 witness: one contiguous run per block, copied with 16-byte stores
        const uint4 *src = reinterpret_cast<const uint4 *>(tile);
        uint4 *dst = reinterpret_cast<uint4 *>(f16 + e0 * (u64)L * ring::D);
        for (u32 u = threadIdx.x; u < nrows * 3; u += THREADS) dst[u] = src[u];
    }
    if (!f_plain && !fx) return;
    u64 c[ring::D];
    const bool active = threadIdx.x < nrows;  // L <= 8 and OPB = 16: at most 128 rows = one per thread
    if (active) {
        int d[ring::D];
        load_i16x24(tile + threadIdx.x * ring::D, d);
        xf::crt24_small<MONT>(d, c);
    }
    const u64 elem0 = e0 * (u64)L;
    if (f_plain) {
        if (active) row_put(otile, threadIdx.x, c);
        rows_out<PLAIN_UNITS>(otile, f_plain + elem0 * ring::D, nrows);
    }
    if (fx && active) store_fx_row(fx + (elem0 + threadIdx.x) * FX_WORDS, c);
}

// Programmatic dependent launches on both sides of this kernel (see mac_kernel):
//  * the matrix-vector kernel launched BEHIND it may start its prologue (barriers, first matrix tiles) on SMs as they
//    drain; it still waits for this grid to complete before touching the witness;
//  * with `chained` (lat_ajtai_set_step_overlap) this grid was itself launched while the PREVIOUS step's
//    matrix-vector kernel is still draining: it writes the other witness buffer, and block 0 does not retire before
//    that kernel has completed, so "this grid complete" implies "previous commitment complete" for everything
//    ordered after it (the buffer two steps back, the workspace, the output).
#ifndef LAT_WITNESS_BLOCKS
#define LAT_WITNESS_BLOCKS 6  // 80 registers, no spills: 133.3 us per step against 135.5 at 4 (121 registers)
#endif
template <bool MONT>
__global__ void __launch_bounds__(THREADS, LAT_WITNESS_BLOCKS)
witness_kernel(const u64 *__restrict__ w, u64 w_len, int log2b, int L, bool in_coeff, int16_t *__restrict__ f16,
               u64 *__restrict__ f_coeff, u64 *__restrict__ f_plain, u64 *__restrict__ fx, int *__restrict__ flag, int chained,
               const unsigned long long *__restrict__ ready_flag, unsigned long long ready_value, SpinGuard guard,
               int stage_input) {
    asm volatile("griddepcontrol.launch_dependents;");
    if (ready_flag) {
        // pipelined host-buffer steps: the input is uploaded by a copy engine on another stream, followed by a copy of
        // the step's ticket into *ready_flag; waiting for it here keeps event waits out of the kernel chain.  The
        // wait is bounded (SpinGuard): on expiry the host gets LAT_E_CUDA from lat_ajtai_wait, not a hung GPU.
        if (threadIdx.x == 0) spin_until_equals(ready_flag, ready_value, guard, SPIN_UPLOAD_TICKET, ready_value);
        __syncthreads();
    }
    witness_body<MONT>(w, w_len, log2b, L, in_coeff, f16, f_coeff, f_plain, fx, flag, stage_input != 0);
    if (chained && blockIdx.x == 0 && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

void launch_witness(const u64 *w, u64 w_len, int log2b, int L, bool mont, bool in_coeff, int16_t *f16, u64 *f_coeff,
                    u64 *f_plain, u64 *fx, int *flag, cudaStream_t stream, bool overlap_previous,
                    const unsigned long long *ready_flag, unsigned long long ready_value, const SpinGuard &guard, bool stage_input) {
    if (!w_len) return;
    unsigned grid = (unsigned)((w_len + OPB - 1) / OPB);
    const int stage = stage_input ? 1 : 0;
    size_t smem = f_plain ? (size_t)OPB * L * (PLAIN_UNITS + 1) * 16 : 0;  // staging tile of the plain output only
    if (smem + sizeof(int16_t) * OPB * WIT_MAX_L * ring::D > 48 * 1024) {
        cudaFuncSetAttribute(witness_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(witness_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap_previous ? 1 : 0;
    const int chained = overlap_previous ? 1 : 0;
    if (mont) cudaLaunchKernelEx(&cfg, witness_kernel<true>, w, w_len, log2b, L, in_coeff, f16, f_coeff, f_plain, fx, flag, chained, ready_flag, ready_value, guard, stage);
    else cudaLaunchKernelEx(&cfg, witness_kernel<false>, w, w_len, log2b, L, in_coeff, f16, f_coeff, f_plain, fx, flag, chained, ready_flag, ready_value, guard, stage);
}

// ---- int16 coefficients -> K sign*bit planes, each CRT'd ---------------------------------------------------
// decompose_B_vec_into_k_vec (latticefold/src/nifs/decomposition/utils.rs:45-49) + the CRT of Witness::from_f_coeff
// (latticefold/src/arith.rs:327).  One thread per element, K planes of work each: plane k of it = sign * bit_k(|c|).
//
// The CRT is Fq-linear and a plane's coefficients are -1, 0, 1, so a plane's CRT is a difference of two subset sums of the
// 24 basis images CRT(X^t).  Coefficient X^(3i+c) only reaches one word of every slot (the ring is Fq[X]/(X^3 - zeta_s)
// per slot: X^(3i+c) = zeta_s^i X^c), so the 24 coefficients fall into three classes of eight, and a 256-row table per
// class -- row `pat` = sum of the images of the coefficients whose bit is set in pat, 8 slots each -- turns the whole
// transform into 6 row reads and 24 modular subtractions:
//     out[word(c, s)] = T[c][bits of the positive coefficients of class c][s] - T[c][bits of the negative ones][s].
// The table (3 x 256 x 8 u64 = 48 KB, in the caller's representation) is built once per handle from crt24_small of the
// unit vectors (planes_lut_kernel: the same arithmetic as every other transform here, so the results are the same bit
// for bit) and copied into shared memory by every block.  About 600 instructions per element and plane instead of 2 100.
// Table layout: [class c][slot pair h][pat] x 16 B, so that the 16-byte reads of a warp (one random row per lane) spread
// over all banks.  Two kernels: planes_fx_kernel writes only the extended (Toom-3) rows the matrix-vector and fold kernels
// read -- a slot pair at a time, so nothing but six table entries is live -- with the planes split over blockIdx.y;
// planes_kernel also serves callers that want the plain CRT-form or coefficient-form planes.
constexpr int PLANE_THREADS = 128;
constexpr int LUT_ROWS = 3 * 256;
constexpr int LUT_WORDS = LUT_ROWS * ring::NSLOT;  // 6144 u64
// word of slot s that class c (coefficients 3i + c) lands in: homogenize_fq3 swaps components 1 and 2 in slots 4..7
__host__ __device__ constexpr int plane_word(int c, int s) { return 3 * s + (c == 0 ? 0 : (s < 4 ? c : 3 - c)); }

template <bool MONT>
__global__ void __launch_bounds__(LUT_ROWS) planes_lut_kernel(u64 *__restrict__ lut) {
    __shared__ u64 basis[ring::D][ring::NSLOT];  // basis[t][s] = the one nonzero word of CRT(X^t) in slot s
    if (threadIdx.x < ring::D) {
        int d[ring::D];
        u64 x[ring::D];
#pragma unroll
        for (int t = 0; t < ring::D; ++t) d[t] = (t == (int)threadIdx.x) ? 1 : 0;
        xf::crt24_small<MONT>(d, x);
        const int c = threadIdx.x % 3;
#pragma unroll
        for (int s = 0; s < ring::NSLOT; ++s) {
            const int want = plane_word(c, s);  // select without dynamic indexing of x
            u64 v = 0;
#pragma unroll
            for (int j = 3 * s; j < 3 * s + 3; ++j)
                if (j == want) v = x[j];
            basis[threadIdx.x][s] = v;
        }
    }
    __syncthreads();
    const int c = threadIdx.x >> 8, pat = threadIdx.x & 255;
    u64 acc[ring::NSLOT];
#pragma unroll
    for (int s = 0; s < ring::NSLOT; ++s) acc[s] = 0;
    for (int i = 0; i < 8; ++i)
        if ((pat >> i) & 1) {
#pragma unroll
            for (int s = 0; s < ring::NSLOT; ++s) acc[s] = gl::add(acc[s], basis[3 * i + c][s]);
        }
#pragma unroll
    for (int s = 0; s < ring::NSLOT; ++s) lut[(((size_t)c * 4 + (s >> 1)) * 256 + pat) * 2 + (s & 1)] = acc[s];
}
void launch_planes_lut(bool mont, u64 *lut, cudaStream_t stream) {
    if (mont) planes_lut_kernel<true><<<1, LUT_ROWS, 0, stream>>>(lut);
    else planes_lut_kernel<false><<<1, LUT_ROWS, 0, stream>>>(lut);
}

template <bool MONT>
__global__ void __launch_bounds__(PLANE_THREADS)
planes_kernel(const int16_t *__restrict__ f16, u64 n, int K, const u64 *__restrict__ lut, u64 *__restrict__ planes_f,
              u64 *__restrict__ planes_fx, u64 *__restrict__ planes_fx0, u64 *__restrict__ planes_coeff) {
    asm volatile("griddepcontrol.launch_dependents;");  // the MAC behind it may start its prologue (see mac_kernel)
    // dynamic shared memory: the table, then (only when a plain output is wanted) the staging tile of the plain rows
    extern __shared__ __align__(16) unsigned char planes_smem[];
    ulonglong2 *s_lut = reinterpret_cast<ulonglong2 *>(planes_smem);
    ulonglong2 *otile = s_lut + LUT_WORDS / 2;
    {
        const ulonglong2 *g = reinterpret_cast<const ulonglong2 *>(lut);
        for (int u = threadIdx.x; u < LUT_WORDS / 2; u += PLANE_THREADS) s_lut[u] = g[u];
    }
    __syncthreads();
    const u64 e0 = (u64)blockIdx.x * PLANE_THREADS;
    const u64 e = e0 + threadIdx.x;
    const bool active = e < n;
    const u32 nrows = (u32)min((u64)PLANE_THREADS, n - e0);
    int d[ring::D];
    u32 sgn[3] = {0, 0, 0};  // bit i of sgn[c]: coefficient 3i + c is negative
    u32 mag[ring::D];
    if (active) {
        load_i16x24(f16 + e * ring::D, d);
#pragma unroll
        for (int t = 0; t < ring::D; ++t) {
            mag[t] = (u32)(d[t] < 0 ? -d[t] : d[t]);
            sgn[t % 3] |= (d[t] < 0 ? 1u : 0u) << (t / 3);
        }
    }
    for (int k = 0; k < K; ++k) {
        u64 c[ring::D];
        const u64 elem0 = (u64)k * n + e0;
        if (planes_coeff) {
            if (active) {
#pragma unroll
                for (int t = 0; t < ring::D; ++t) {
                    const int bit = (int)((mag[t] >> k) & 1u);
                    c[t] = gl::from_small<MONT>(d[t] < 0 ? -bit : bit);  // digit k base 2 = sign * bit_k(|c|)
                }
                row_put(otile, threadIdx.x, c);
            }
            rows_out<PLAIN_UNITS>(otile, planes_coeff + elem0 * ring::D, nrows);
        }
        if (planes_f || planes_fx) {
            if (active) {
#pragma unroll
                for (int cl = 0; cl < 3; ++cl) {
                    u32 m = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) m |= ((mag[3 * i + cl] >> k) & 1u) << i;
                    const ulonglong2 *rp = s_lut + cl * 1024 + (m & ~sgn[cl]);
                    const ulonglong2 *rn = s_lut + cl * 1024 + (m & sgn[cl]);
#pragma unroll
                    for (int h = 0; h < ring::NSLOT / 2; ++h) {
                        const ulonglong2 a = rp[h * 256], b = rn[h * 256];
                        c[plane_word(cl, 2 * h)] = gl::sub(a.x, b.x);
                        c[plane_word(cl, 2 * h + 1)] = gl::sub(a.y, b.y);
                    }
                }
            }
            if (planes_f) {
                if (active) row_put(otile, threadIdx.x, c);
                rows_out<PLAIN_UNITS>(otile, planes_f + elem0 * ring::D, nrows);
            }
            if (planes_fx && active)
                store_fx_row<true>((k == 0 ? planes_fx0 : planes_fx + (u64)(k - 1) * n * FX_WORDS) + e * FX_WORDS, c);
        }
    }
}

// The common case: only the extended rows.  One thread per (element, slot pair): the four lanes of an element write its
// 384-byte row as four adjacent 96-byte pieces, so a warp's three 256-bit stores cover 3 KB without gaps (the mapping of
// fext_kernel; one thread per row reaches 3.6 TB/s of stores, this one is bound by the table reads instead).  Persistent
// blocks (3 per SM, the table is copied once per block) stride over the work items (64 elements, a third of the planes).
constexpr int PLANE_FX_THREADS = 256;
constexpr int PLANE_FX_ELEMS = PLANE_FX_THREADS / 4;
#ifndef LAT_PLANES_FX_BLOCKS
#define LAT_PLANES_FX_BLOCKS 3  // 71 registers; 4 blocks per SM (63 registers) measured 138.5 us against 133.8 for pack + planes;
                                // requesting the next item's digits one item ahead changed nothing (133.8): the store path is the limit;
                                // staging an item's 64 rows in shared memory and writing them with one 24 KB bulk store
                                // (cp.async.bulk.global.shared::cta, single-buffered) was slower: 169.6 us
#endif
__global__ void __launch_bounds__(PLANE_FX_THREADS, LAT_PLANES_FX_BLOCKS)
planes_fx_kernel(const int16_t *__restrict__ f16, u64 n, int K, const u64 *__restrict__ lut, u64 *__restrict__ planes_fx,
                 u64 *__restrict__ planes_fx0) {
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ __align__(16) unsigned char planes_smem[];
    ulonglong2 *s_lut = reinterpret_cast<ulonglong2 *>(planes_smem);
    {
        const ulonglong2 *g = reinterpret_cast<const ulonglong2 *>(lut);
        for (int u = threadIdx.x; u < LUT_WORDS / 2; u += PLANE_FX_THREADS) s_lut[u] = g[u];
    }
    __syncthreads();
    const u64 neb = (n + PLANE_FX_ELEMS - 1) / PLANE_FX_ELEMS;
    const int ppi = (K + 2) / 3, ng = (K + ppi - 1) / ppi;  // planes per item, plane groups
    const u32 h = threadIdx.x & 3;                          // slot pair (slots 2h, 2h+1)
    for (u64 item = blockIdx.x; item < neb * (u64)ng; item += gridDim.x) {
        const int g = (int)(item / neb);
        const u64 e = (item - (u64)g * neb) * PLANE_FX_ELEMS + (threadIdx.x >> 2);
        if (e >= n) continue;
        int d[ring::D];
        load_i16x24(f16 + e * ring::D, d);
        u32 sgn[3] = {0, 0, 0};
        u32 mag[ring::D];
#pragma unroll
        for (int t = 0; t < ring::D; ++t) {
            mag[t] = (u32)(d[t] < 0 ? -d[t] : d[t]);
            sgn[t % 3] |= (d[t] < 0 ? 1u : 0u) << (t / 3);
        }
        const int k1 = min(K, (g + 1) * ppi);
        for (int k = g * ppi; k < k1; ++k) {
            // words of the slot: class c lands in component c for slots 0..3, classes 1 and 2 swap in slots 4..7 (plane_word)
            u64 a[3], b[3];
#pragma unroll
            for (int cl = 0; cl < 3; ++cl) {
                u32 m = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) m |= ((mag[3 * i + cl] >> k) & 1u) << i;
                const ulonglong2 p = s_lut[cl * 1024 + h * 256 + (m & ~sgn[cl])];
                const ulonglong2 q = s_lut[cl * 1024 + h * 256 + (m & sgn[cl])];
                const u64 x = gl::sub(p.x, q.x), y = gl::sub(p.y, q.y);
                if (cl == 0) {
                    a[0] = x; b[0] = y;
                } else {
                    const bool swap = h >= 2;
                    if (cl == 1) { a[1] = x; b[1] = y; }
                    else {  // cl == 2: after this both classes are known, put them in place
                        const u64 a1 = a[1], b1 = b[1];
                        a[1] = swap ? x : a1; b[1] = swap ? y : b1;
                        a[2] = swap ? a1 : x; b[2] = swap ? b1 : y;
                    }
                }
            }
            u64 p0, p1, p2, q0, q1, q2;
            gl::toom_eval(a[0], a[1], a[2], p0, p1, p2);
            gl::toom_eval(b[0], b[1], b[2], q0, q1, q2);
            u64 *o = (k == 0 ? planes_fx0 : planes_fx + (u64)(k - 1) * n * FX_WORDS) + e * FX_WORDS + h * 12;
            st256(o, a[0], a[1], a[2], p0);
            st256(o + 4, p1, p2, b[0], b[1]);
            st256(o + 8, b[2], q0, q1, q2);
        }
    }
}

void launch_planes(const int16_t *f16, u64 n, int K, bool mont, const u64 *lut, u64 *planes_f, u64 *planes_fx,
                   u64 *planes_fx0, u64 *planes_coeff, cudaStream_t stream) {
    if (!n) return;
    if (!planes_f && !planes_coeff) {
        if (!planes_fx) return;
        static int sm_count_on[64] = {};  // 0 = not asked yet; the attribute is set at the same time
        int dev = 0;
        cudaGetDevice(&dev);
        int &sms = sm_count_on[dev & 63];
        if (!sms) {
            cudaFuncSetAttribute(planes_fx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_WORDS * 8);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (sms <= 0) sms = 148;
        }
        const u64 items = ((n + PLANE_FX_ELEMS - 1) / PLANE_FX_ELEMS) * 3;
        const unsigned grid = (unsigned)min(items, (u64)sms * LAT_PLANES_FX_BLOCKS);
        planes_fx_kernel<<<grid, PLANE_FX_THREADS, LUT_WORDS * 8, stream>>>(f16, n, K, lut, planes_fx, planes_fx0);
        return;
    }
    unsigned grid = (unsigned)((n + PLANE_THREADS - 1) / PLANE_THREADS);
    const bool plain = planes_f || planes_coeff;
    const size_t smem = (size_t)LUT_WORDS * 8 + (plain ? (size_t)PLANE_THREADS * (PLAIN_UNITS + 1) * 16 : 0);
    static bool attr_set_on[64] = {};  // per device, as for mac_kernel
    int dev = 0;
    cudaGetDevice(&dev);
    bool &attr_set = attr_set_on[dev & 63];
    if (!attr_set) {
        const int cap = LUT_WORDS * 8 + PLANE_THREADS * (PLAIN_UNITS + 1) * 16;
        cudaFuncSetAttribute(planes_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(planes_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        attr_set = true;
    }
    if (mont) planes_kernel<true><<<grid, PLANE_THREADS, smem, stream>>>(f16, n, K, lut, planes_f, planes_fx, planes_fx0, planes_coeff);
    else planes_kernel<false><<<grid, PLANE_THREADS, smem, stream>>>(f16, n, K, lut, planes_f, planes_fx, planes_fx0, planes_coeff);
}

// ---- Witness::get_fhat (latticefold/src/arith.rs:273-297) from the resident digits ---------------------------------------------
template <bool MONT>
__global__ void __launch_bounds__(256) fhat_kernel(const int16_t *__restrict__ f16, u64 n, u64 *__restrict__ fhat) {
    const u64 idx = (u64)blockIdx.x * 256 + threadIdx.x;  // (table j, element i, slot s), s fastest
    if (idx >= 3 * n * ring::NSLOT) return;
    const u64 s = idx & 7, ji = idx >> 3, j = ji / n, i = ji - j * n;
    u64 *o = fhat + ji * ring::D + 3 * s;
    o[0] = gl::from_small<MONT>((int)f16[i * ring::D + 8 * j + s]);
    o[1] = 0;
    o[2] = 0;
}
void launch_fhat(const int16_t *f16, u64 n, bool mont, u64 *fhat, cudaStream_t stream) {
    if (!n) return;
    const unsigned grid = (unsigned)((3 * n * ring::NSLOT + 255) / 256);
    if (mont) fhat_kernel<true><<<grid, 256, 0, stream>>>(f16, n, fhat);
    else fhat_kernel<false><<<grid, 256, 0, stream>>>(f16, n, fhat);
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

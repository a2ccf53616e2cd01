// The Ajtai matrix-vector product in CRT form on sm_100a:  cm[p][i] = sum_j A[i][j] (*) F[p][j]  with (*) the
// slot-wise Fq3 product.  Replaces Matrix::checked_mul_vec (stark-rings/crates/linear_algebra/src/matrix.rs:168-178,
// called from AjtaiCommitmentScheme::commit, latticefold/src/commitment/commitment_scheme.rs:63-80) and the loop of
// K-1 such commits in LFDecompositionProver::commit_witnesses (latticefold/src/nifs/decomposition.rs:185-187).
//
// Design (DESIGN.md "mac kernel"):
//  * The matrix is re-laid out ONCE at upload into column tiles  [row block][tile][column jj][component c][row i][slot s]
//    (canonical form), so that a tile is one contiguous run of bytes in HBM and a warp's shared-memory read of
//    (jj, c) is 32 consecutive u64 (conflict-free).  Tiles are streamed with TMA bulk copies (cp.async.bulk ->
//    UBLKCP) into a multi-stage shared-memory ring guarded by mbarriers; RG*CG warps, all of them consumers -- the
//    last warp to release a stage issues its refill (no producer warp: a 9th warp halves the occupancy).
//  * The witness side comes in the "extended" layout [element][slot][6] = (f0, f1, f2, f0+f1, f0+f2, f1+f2): the
//    Karatsuba pre-additions of the witness are done once by whoever produces it (the CRT kernels, or fext_kernel
//    for a caller-supplied witness), not once per matrix row.  The matrix-side pre-additions are 3 exact 65-bit sums
//    per thread per column.
//  * Split-K over columns: every CTA owns a contiguous range of tiles and keeps, per thread, the UNREDUCED
//    accumulators of one output (row, slot) for PT witnesses ("planes"): 6 sums x 3 columns x (64+32) bits
//    (gl::Fq3Acc).  One 64x64 product = 4 IMAD.WIDE.U32 with carry-out + 2 IADD3.X; 6 products per Fq3 MAC
//    (Karatsuba) = 24 IMAD.WIDE.U32; a single special-form reduction per output at the end.  IMAD.WIDE.U32 runs at
//    31.5 /clk/SM (measured), so this multiply count, not HBM, is what the batched (PT > 1) case is bound by.
//    No tensor cores: this is exact 64-bit modular integer work.
//  * Launches with SEVERAL witnesses (the K-1 planes of a decomposition, commit batches) run in Toom-3 form instead
//    (TOOM = true; gl::ToomAcc): both operands evaluated at 0, infinity, 1, -1, 2 -- the witness rows by their producer,
//    the matrix as a second device copy with 5 words per entry (derive_toom_kernel) -- so one Fq3 MAC is 5 products =
//    20 IMAD.WIDE.U32 and no pre-additions, interpolated once per output.  4 witnesses per thread (250 registers), one
//    CTA per SM, two stages of 80 KB tiles.  The single witness stays Karatsuba: it is bound by the matrix bytes.
//  * Programmatic dependent launch: the prologue and the first matrix tiles do not wait for the kernel that
//    produces the witness, and the next call's kernels may start while this one drains.
//  * Cross-CTA sum inside the same kernel: canonical partials are added as 32-bit halves with 64-bit REDs into a
//    (zeroed, self-cleaning) workspace and the last CTA folds them mod q into the commitment.
//  * A is canonical and F is in the caller's representation, so canonical(A) * repr(F) = repr(A*F): the
//    commitment comes out in the caller's representation without any conversion (the map is Fq-linear).
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "kernels.h"
#include "ring24.cuh"
#include "spin.cuh"
#include "tma.cuh"

namespace lat {
using gl::u32;

constexpr int SM_RESERVED_SMEM = 1024;
#ifndef LAT_TJ_BYTES
#define LAT_TJ_BYTES 256
#endif
constexpr int FX = 48;  // u64 per element in the extended witness layout (8 slots x 6)

// Compile-time geometry per row-group count RG (rows per block RB = 4 RG <= 32).
__host__ __device__ constexpr int geo_cg(int rg) { return rg == 1 ? 8 : rg == 2 ? 4 : rg <= 4 ? 2 : 1; }
__host__ __device__ constexpr int geo_tj(int rg) { return ((LAT_TJ_BYTES / (4 * rg)) / geo_cg(rg)) * geo_cg(rg) < geo_cg(rg) ? geo_cg(rg) : ((LAT_TJ_BYTES / (4 * rg)) / geo_cg(rg)) * geo_cg(rg); }
// the 5-word (Toom-3) matrix: as many columns per tile as let TWO stages of (tile + the witness rows of 4 witnesses) fit
// the SM's shared memory -- 8 columns = 80 KB tiles at kappa = 32.  The several-witness kernels run one CTA of 8 warps per
// SM, all of which reach a tile boundary together (barrier poll, release atomic, refill), and that boundary costs more
// than a shallower ring: 4 / 6 / 8 columns per tile at kappa = 32 gave 1.09 / 1.03 / 1.01 ms for 14 commits, and 2 stages
// ran as fast as 4 at every size (tools/ab_mac.py).
#ifndef LAT_TJ5_STAGE_BYTES
#define LAT_TJ5_STAGE_BYTES 98304
#endif
__host__ __device__ constexpr int geo_tj5(int rg) {
    // bytes per column of a stage: 5 words x 4 rg rows x 8 slots x 8 B of matrix + 4 witnesses x 384 B
    return (LAT_TJ5_STAGE_BYTES / (1280 * rg + 1536)) / geo_cg(rg) * geo_cg(rg) < geo_cg(rg)
               ? geo_cg(rg)
               : (LAT_TJ5_STAGE_BYTES / (1280 * rg + 1536)) / geo_cg(rg) * geo_cg(rg);
}

MatLayout make_layout(uint32_t kappa, u64 n, bool toom) {
    MatLayout l{};
    l.nc = toom ? 5 : 3;
    l.kappa = kappa;
    l.n = n;
    uint32_t k4 = (kappa + 3) / 4 * 4;
    if (k4 <= 32) {
        l.kappa_pad = k4;
        l.rb = k4;
    } else {
        l.kappa_pad = (kappa + 31) / 32 * 32;
        l.rb = 32;
    }
    l.nrb = l.kappa_pad / l.rb;
    l.rg = l.rb / 4;
    l.cg = geo_cg((int)l.rg);
    l.tj = toom ? geo_tj5((int)l.rg) : geo_tj((int)l.rg);  // ~48 KB tiles (tj * rb * 192 B): per-tile barrier/refill costs favour large tiles
    l.ntiles = (n + l.tj - 1) / l.tj;
    if (l.ntiles == 0) l.ntiles = 1;
    l.n_pad = l.ntiles * l.tj;
    return l;
}

// ---- upload-time re-layout ------------------------------------------------------------------------------------
// One thread per destination u64 of the rows being uploaded.  dst index inside a tile: ((jj*3 + c)*rb + il)*8 + s.
template <bool MONT>
__global__ void __launch_bounds__(256)
relayout_kernel(const u64 *__restrict__ rows, uint32_t row0, uint32_t nrows, u64 row_stride, MatLayout lay,
                u64 *__restrict__ A_dev) {
    // thread -> (r, j, c, s) with s fastest, then c... we iterate in SOURCE order for coalesced reads:
    u64 idx = (u64)blockIdx.x * 256 + threadIdx.x;
    u64 per_row = lay.n * ring::D;
    if (idx >= (u64)nrows * per_row) return;
    uint32_t r = (uint32_t)(idx / per_row);
    u64 rem = idx - (u64)r * per_row;
    u64 j = rem / ring::D;
    uint32_t t = (uint32_t)(rem - j * ring::D);
    uint32_t s = t / 3, c = t - 3 * s;
    u64 v = rows[((u64)r * row_stride + j) * ring::D + t];
    if constexpr (MONT) v = gl::from_mont(v);
    else v = gl::reduce128(v, 0);
    uint32_t i = row0 + r;
    uint32_t rbk = i / lay.rb, il = i - rbk * lay.rb;
    u64 tile = j / lay.tj;
    uint32_t jj = (uint32_t)(j - tile * lay.tj);
    u64 dst = ((u64)rbk * lay.ntiles + tile) * lay.tile_elems() + ((u64)(jj * 3 + c) * lay.rb + il) * 8 + s;
    A_dev[dst] = v;
}

void launch_relayout(const u64 *rows, uint32_t row0, uint32_t nrows, u64 row_stride, bool mont, const MatLayout &lay,
                     u64 *A_dev, cudaStream_t stream) {
    u64 total = (u64)nrows * lay.n * ring::D;
    if (!total) return;
    unsigned grid = (unsigned)((total + 255) / 256);
    if (mont) relayout_kernel<true><<<grid, 256, 0, stream>>>(rows, row0, nrows, row_stride, lay, A_dev);
    else relayout_kernel<false><<<grid, 256, 0, stream>>>(rows, row0, nrows, row_stride, lay, A_dev);
}

// ---- 3-word matrix -> 5-word (Toom-3 evaluations) matrix, on the device ------------------------------------------------
// One thread per (row block, column, row, slot); (row, slot) fastest, so a warp reads and writes 32 consecutive u64 of a
// component / an evaluation.  Padding rows are zero in A and stay zero (every evaluation of 0 is 0).
__global__ void __launch_bounds__(256)
derive_toom_kernel(const u64 *__restrict__ A3, MatLayout lay, u64 *__restrict__ A5, MatLayout lay5) {
    const u64 per_col = (u64)lay.rb * 8;
    const u64 idx = (u64)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (u64)lay.nrb * lay.n * per_col) return;
    const u64 w = idx % per_col;  // il * 8 + s
    const u64 cj = idx / per_col;
    const u64 j = cj % lay.n;
    const u64 rbk = cj / lay.n;
    const u64 t3 = j / lay.tj, jj3 = j - t3 * lay.tj;
    const u64 *src = A3 + (rbk * lay.ntiles + t3) * lay.tile_elems() + jj3 * 3 * per_col + w;
    const u64 a0 = src[0], a1 = src[per_col], a2 = src[2 * per_col];
    u64 t1, tm, t2;
    gl::toom_eval(a0, a1, a2, t1, tm, t2);
    const u64 t5 = j / lay5.tj, jj5 = j - t5 * lay5.tj;
    u64 *dst = A5 + (rbk * lay5.ntiles + t5) * lay5.tile_elems() + jj5 * 5 * per_col + w;
    dst[0] = a0;
    dst[per_col] = a2;
    dst[2 * per_col] = t1;
    dst[3 * per_col] = tm;
    dst[4 * per_col] = t2;
}
void launch_derive_toom(const u64 *A_dev, const MatLayout &lay, u64 *A5_dev, const MatLayout &lay5, cudaStream_t stream) {
    const u64 total = (u64)lay.nrb * lay.n * lay.rb * 8;
    if (!total) return;
    derive_toom_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(A_dev, lay, A5_dev, lay5);
}

// ---- witness -> extended layout (only for caller-supplied CRT-form witnesses; the CRT kernels emit it directly) -----
// One thread per PAIR of slots: 48 bytes in (three 16-byte loads), 96 bytes out as three 256-bit stores -- whole 32-byte
// sectors, so L2 never has to read-merge a half-written one (the 16-byte stores of round 1 ran at 3.2 TB/s).
// the three derived words of a slot: Karatsuba sums (any representative) or Toom-3 evaluations (see kernels.h FX_WORDS)
template <bool TOOM>
__device__ __forceinline__ void fx_derived(u64 f0, u64 f1, u64 f2, u64 &d0, u64 &d1, u64 &d2) {
    if constexpr (TOOM) {
        gl::toom_eval(f0, f1, f2, d0, d1, d2);
    } else {
        d0 = gl::add_lazy(f0, f1);
        d1 = gl::add_lazy(f0, f2);
        d2 = gl::add_lazy(f1, f2);
    }
}
template <bool TOOM>
__global__ void __launch_bounds__(256)
fext_kernel(const u64 *__restrict__ f, u64 count_pairs, u64 *__restrict__ fx) {
    asm volatile("griddepcontrol.launch_dependents;");  // the MAC behind it may start its prologue (see mac_kernel)
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;  // (element, slot pair)
    if (i >= count_pairs) return;
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(f + i * 6);
    const ulonglong2 v0 = p[0], v1 = p[1], v2 = p[2];
    const u64 a0 = gl::reduce128(v0.x, 0), a1 = gl::reduce128(v0.y, 0), a2 = gl::reduce128(v1.x, 0);
    const u64 b0 = gl::reduce128(v1.y, 0), b1 = gl::reduce128(v2.x, 0), b2 = gl::reduce128(v2.y, 0);
    u64 *o = fx + i * 12;
    u64 x0, x1, x2, y0, y1, y2;
    fx_derived<TOOM>(a0, a1, a2, x0, x1, x2);
    fx_derived<TOOM>(b0, b1, b2, y0, y1, y2);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o), "l"(a0), "l"(a1), "l"(a2), "l"(x0) : "memory");
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o + 4), "l"(x1), "l"(x2), "l"(b0), "l"(b1) : "memory");
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o + 8), "l"(b2), "l"(y0), "l"(y1), "l"(y2) : "memory");
}
// unaligned callers (a witness pointer that is not 16-byte aligned): one thread per slot, 8-byte accesses
template <bool TOOM>
__global__ void __launch_bounds__(256)
fext_kernel_unaligned(const u64 *__restrict__ f, u64 count_slots, u64 *__restrict__ fx) {
    asm volatile("griddepcontrol.launch_dependents;");
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;  // (element, slot)
    if (i >= count_slots) return;
    u64 f0 = f[i * 3], f1 = f[i * 3 + 1], f2 = f[i * 3 + 2];
    f0 = gl::reduce128(f0, 0); f1 = gl::reduce128(f1, 0); f2 = gl::reduce128(f2, 0);
    u64 d0, d1, d2;
    fx_derived<TOOM>(f0, f1, f2, d0, d1, d2);
    ulonglong2 *o = reinterpret_cast<ulonglong2 *>(fx + i * 6);
    o[0] = make_ulonglong2(f0, f1);
    o[1] = make_ulonglong2(f2, d0);
    o[2] = make_ulonglong2(d1, d2);
}
void launch_fext(const u64 *f, u64 count, u64 *fx, cudaStream_t stream, bool toom) {
    if (!count) return;
    if ((reinterpret_cast<uintptr_t>(f) & 15) == 0 && (reinterpret_cast<uintptr_t>(fx) & 31) == 0) {
        const u64 pairs = count * (ring::NSLOT / 2);
        const unsigned grid = (unsigned)((pairs + 255) / 256);
        if (toom) fext_kernel<true><<<grid, 256, 0, stream>>>(f, pairs, fx);
        else fext_kernel<false><<<grid, 256, 0, stream>>>(f, pairs, fx);
    } else {
        const u64 slots = count * ring::NSLOT;
        const unsigned grid = (unsigned)((slots + 255) / 256);
        if (toom) fext_kernel_unaligned<true><<<grid, 256, 0, stream>>>(f, slots, fx);
        else fext_kernel_unaligned<false><<<grid, 256, 0, stream>>>(f, slots, fx);
    }
}

// ---- the MAC kernel --------------------------------------------------------------------------------------------
// grid = (column chunks, row blocks, plane groups); block = RG*CG warps.
// Shared memory: STAGES x { A tile | PT x TJ x 48 u64 of extended witness } + 2*STAGES mbarriers.
// Workspace: ws[2*i], ws[2*i+1] = sums of the low / high 32-bit halves of output i = (p * kappa + row) * 24 + s*3 + c;
// ws[2 * nout] = finished-CTA counter.  Must be zero before the first launch; every launch leaves it zero again.
template <int PT, int RG, bool TOOM>
struct MacGeo {
    static constexpr int RB = 4 * RG, CG = geo_cg(RG), TJ = TOOM ? geo_tj5(RG) : geo_tj(RG);
    static constexpr int NC = TOOM ? 5 : 3;  // u64 per matrix entry
    // RG*CG warps, all consumers; lane 0 of warp 0 also issues the TMA copies.  (A dedicated 9th producer warp
    // halves the occupancy: warp slots are handed out four at a time -- measured with the occupancy API.)
    static constexpr int NCONS = RG * CG, THREADS = NCONS * 32;
    static constexpr u32 TILE_ELEMS = TJ * NC * RB * 8;
    static constexpr u32 TILE_BYTES = TILE_ELEMS * 8;
    static constexpr u32 F_BYTES = TJ * FX * 8;  // per plane per tile
    static constexpr u32 STAGE_BYTES = TILE_BYTES + PT * F_BYTES;
    static constexpr int MIN_CTAS = (PT == 1 && !TOOM) ? 2 : 1;
};

#ifdef LAT_MAC_TRACE
// Tuning builds only (latticeum_b200.build.build_variant + tools/trace_mac.py): per-CTA globaltimer stamps of the last
// launch -- entry, first tile landed, loop done, partials published, exit, SM id -- to size the fixed cost of a launch.
__device__ unsigned long long g_mac_trace[8192 * 8];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TRACE(k)                                                                                               \
    do {                                                                                                       \
        if (threadIdx.x == 0) {                                                                                \
            const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);               \
            if (cta < 8192) g_mac_trace[cta * 8 + (k)] = gtime();                                              \
        }                                                                                                      \
    } while (0)
#else
#define TRACE(k)
#endif

template <int PT, int RG, bool TOOM>
__global__ void __launch_bounds__(MacGeo<PT, RG, TOOM>::THREADS, MacGeo<PT, RG, TOOM>::MIN_CTAS)
mac_kernel(const u64 *__restrict__ A_dev, MatLayout lay, const u64 *__restrict__ Fx, u64 f_stride, uint32_t planes,
           uint32_t stages, u64 *__restrict__ ws, u64 *__restrict__ cms, uint32_t dependent_launch, MacReport report) {
    using G = MacGeo<PT, RG, TOOM>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // after the stages: [full mbarrier x stages][release counter x stages]
    u64 *bars = reinterpret_cast<u64 *>(smem_raw + (size_t)stages * G::STAGE_BYTES);
    u32 *released = reinterpret_cast<u32 *>(bars + stages);

    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 rbk = blockIdx.y;
    const u32 p0 = blockIdx.z * PT;
    TRACE(0);
#ifdef LAT_MAC_TRACE
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (cta < 8192) g_mac_trace[cta * 8 + 5] = smid;
    }
#endif

    // contiguous tile range of this CTA
    const u64 t_begin = lay.ntiles * blockIdx.x / gridDim.x;
    const u64 t_end = lay.ntiles * (blockIdx.x + 1) / gridDim.x;
    const u32 my_tiles = (u32)(t_end - t_begin);
    const u64 *a_src = A_dev + ((u64)rbk * lay.ntiles + t_begin) * G::TILE_ELEMS;

    // Stream tile `nt` into stage `st` with TMA bulk copies (one thread).  There is no producer role: the warp that
    // is LAST to release a stage refills it at once (release counters below), so no warp ever waits for another
    // except through the data itself, and `stages` tiles are in flight or being consumed at all times.
    // The last tile may hang over the end of F (columns >= n): copy only the valid columns.  The matching matrix
    // columns are zero padding, so whatever the stale tail of the stage holds contributes 0 (exact integer
    // arithmetic, no NaNs to worry about).
    auto witness_bytes = [&](u32 nt) { return (u32)min((u64)G::TJ, lay.n - (t_begin + nt) * G::TJ) * FX * 8; };
    auto issue_matrix = [&](u32 nt, u32 st) {
        mbar_arrive_expect_tx(&bars[st], G::TILE_BYTES + PT * witness_bytes(nt));
#ifndef LAT_NO_L2_HINT
        if constexpr (PT == 1 && !TOOM)  // streamed once; with several plane groups the other groups' CTAs re-read the tile from L2
            tma_bulk_g2s_hint(smem_raw + (size_t)st * G::STAGE_BYTES, a_src + (u64)nt * G::TILE_ELEMS, G::TILE_BYTES, &bars[st],
                              L2_EVICT_FIRST);
        else
#endif
        tma_bulk_g2s(smem_raw + (size_t)st * G::STAGE_BYTES, a_src + (u64)nt * G::TILE_ELEMS, G::TILE_BYTES, &bars[st]);
    };
    auto issue_witness = [&](u32 nt, u32 st) {
        unsigned char *dst = smem_raw + (size_t)st * G::STAGE_BYTES;
        const u64 col0 = (t_begin + nt) * G::TJ;
        const u32 fb = witness_bytes(nt);
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            const u64 *f_src = Fx + ((u64)(p0 + p) * f_stride + col0) * FX;
#ifndef LAT_NO_L2_HINT
            tma_bulk_g2s_hint(dst + G::TILE_BYTES + p * G::F_BYTES, f_src, fb, &bars[st], L2_EVICT_LAST);
#else
            tma_bulk_g2s(dst + G::TILE_BYTES + p * G::F_BYTES, f_src, fb, &bars[st]);
#endif
        }
    };
    auto issue_tile = [&](u32 nt, u32 st) {
        issue_matrix(nt, st);
        issue_witness(nt, st);
    };

    if (threadIdx.x == 0) {
        for (u32 st = 0; st < stages; ++st) {
            mbar_init(&bars[st], 1);  // full: one arrive (whoever issues the copies) + tx bytes
            released[st] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // Programmatic dependent launch: this grid may start while the kernel that produces the witness (Fx) is
        // still draining.  The matrix does not depend on it, so the first tiles' matrix halves are requested at
        // once; only the witness halves wait for the producer grid to complete.
        const u32 pre = min(stages, my_tiles);
        for (u32 nt = 0; nt < pre; ++nt) issue_matrix(nt, nt);
        if (dependent_launch) asm volatile("griddepcontrol.wait;" ::: "memory");
        // Only now may the kernel behind this one start (the next step's witness kernel, if the caller allows the
        // overlap): once every CTA has passed its wait the producer grid -- and through its block 0 the previous
        // commitment -- is complete, so that kernel can reuse the buffers of two steps back.
        asm volatile("griddepcontrol.launch_dependents;");
        for (u32 nt = 0; nt < pre; ++nt) issue_witness(nt, nt);
    }
    __syncthreads();

    const u32 rgi = warp / G::CG, cgi = warp % G::CG;
    const u32 il = rgi * 4 + (lane >> 3), s = lane & 7;
    typename std::conditional<TOOM, gl::ToomAcc, gl::Fq3Acc>::type acc[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) acc[p].clear();

    u32 st = 0, ph = 0;  // stage and phase parity of tile t (no runtime division in the loop)
    bool ready = false;  // result of the early, non-blocking poll of this tile's barrier
    for (u32 t = 0; t < my_tiles; ++t) {
        if (!ready) mbar_wait(&bars[st], ph);
#ifdef LAT_MAC_TRACE
        if (t == 0) TRACE(1);
#endif
        const u64 *sa = reinterpret_cast<const u64 *>(smem_raw + (size_t)st * G::STAGE_BYTES) + il * 8 + s;
        const ulonglong2 *sf =
            reinterpret_cast<const ulonglong2 *>(smem_raw + (size_t)st * G::STAGE_BYTES + G::TILE_BYTES) + s * 3;
        // poll the NEXT tile's barrier now, so that its latency hides under this tile's arithmetic
        u32 st_n = st + 1, ph_n = ph;
        if (st_n == stages) {
            st_n = 0;
            ph_n ^= 1;
        }
        ready = (t + 1 < my_tiles) && mbar_test(&bars[st_n], ph_n);
#pragma unroll
        for (int q = 0; q < G::TJ / G::CG; ++q) {
            const u32 jj = cgi + q * G::CG;
            const u64 *pa = sa + jj * (G::NC * G::RB * 8);
            u64 a0 = pa[0], a1 = pa[G::RB * 8], a2 = pa[2 * G::RB * 8];
            if constexpr (TOOM) {
                // the entry's evaluations at (0, infinity, 1, -1, 2) against each witness slot's (f0, f1, f2, f(1), f(-1), f(2))
                const u64 a3 = pa[3 * G::RB * 8], a4 = pa[4 * G::RB * 8];
#pragma unroll
                for (int p = 0; p < PT; ++p) {
                    const ulonglong2 *pf = sf + (p * G::TJ + jj) * (FX / 2);
                    const u64 y0 = reinterpret_cast<const u64 *>(pf)[0];
                    const ulonglong2 y = pf[1], z = pf[2];
                    acc[p].mac(a0, a1, a2, a3, a4, y0, y.x, y.y, z.x, z.y);
                }
            } else if constexpr (PT == 1) {
                // one witness: exact 65-bit sums, the carry goes straight into the accumulator (fewest instructions)
                ulonglong2 x = sf[jj * (FX / 2)], y = sf[jj * (FX / 2) + 1], z = sf[jj * (FX / 2) + 2];
                acc[0].mac(a0, a1, a2, x.x, x.y, y.x, y.y, z.x, z.y);
            } else {
                static_assert(TOOM || PT == 1, "launches with several witnesses take the Toom-3 instance");
            }
        }
        __syncwarp();
        // release the stage; the last warp to do so refills it with tile t + stages
        if (lane == 0) {
            if (atomicAdd(&released[st], 1u) == G::NCONS - 1) {
                released[st] = 0;
                if (t + stages < my_tiles) issue_tile(t + stages, st);
            }
        }
        st = st_n;
        ph = ph_n;
    }

    TRACE(2);
    // ===== epilogue ============================================================================================
    // One special-form reduction per output, then the cross-CTA sum: every partial (canonical, < 2^64) is split into
    // its 32-bit halves and added with two 64-bit REDs into ws[2*idx], ws[2*idx+1] (a few hundred addends cannot
    // overflow); the last CTA to finish folds lo + 2^32 hi mod q into cms and leaves the workspace zeroed for the
    // next launch.  No second kernel, no partials round trip.
    const u32 row = rbk * G::RB + il;
    if (row < lay.kappa) {
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            u64 c[3];
            acc[p].finish(c[0], c[1], c[2]);
            u64 *dst = ws + 2 * ((((u64)(p0 + p)) * lay.kappa + row) * ring::D + s * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                atomicAdd(reinterpret_cast<unsigned long long *>(dst + 2 * k), c[k] & 0xFFFFFFFFull);
                atomicAdd(reinterpret_cast<unsigned long long *>(dst + 2 * k + 1), c[k] >> 32);
            }
        }
    }
    __shared__ u32 s_last;
    const u64 nout = (u64)planes * lay.kappa * ring::D;
    u64 *counter = ws + 2 * nout;
    __threadfence();
    __syncthreads();
    TRACE(3);
    if (threadIdx.x == 0) {
        const u32 total = gridDim.x * gridDim.y * gridDim.z;
        s_last = (atomicAdd(reinterpret_cast<unsigned long long *>(counter), 1ull) == total - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (u64 i = threadIdx.x; i < nout; i += blockDim.x) {
            u64 lo = __ldcg(ws + 2 * i), hi = __ldcg(ws + 2 * i + 1);   // sums of low / high halves
            u64 v_lo = lo + (hi << 32);
            u64 v_hi = (hi >> 32) + (v_lo < lo ? 1ull : 0ull);
            const u64 v = gl::reduce128(v_lo, v_hi);
            cms[i] = v;
            if (report.cm_host) report.cm_host[i] = v;
            ws[2 * i] = 0;
            ws[2 * i + 1] = 0;
        }
        if (threadIdx.x == 0) *counter = 0;
        if (threadIdx.x == 0 && report.flag_dev) {  // the step's overflow flag moves to the host, the device copy is cleared
            *report.flag_host = *reinterpret_cast<volatile int *>(report.flag_dev);
            *report.flag_dev = 0;
        }
        if (report.done_host) {  // results and flag first, then the ticket, each with system scope (the host polls it)
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(report.done_host), "l"(report.done_value) : "memory");
            }
        }
    }
    TRACE(4);
}

#ifdef LAT_MAC_TRACE
extern "C" int lat_debug_mac_trace(unsigned long long *out, int ctas) {
    return (int)cudaMemcpyFromSymbol(out, g_mac_trace, (size_t)ctas * 8 * sizeof(unsigned long long));
}
#endif

static size_t stage_bytes_for(uint32_t pt, const MatLayout &lay) {
    return (size_t)lay.tile_elems() * 8 + (size_t)pt * lay.tj * FX * 8;
}

MacPlan plan_mac(const MatLayout &lay, uint32_t planes, int sm_count) {
    MacPlan m{};
    // witnesses per thread: the Toom-3 accumulators of 4 witnesses still fit the register file (250 registers, no spills),
    // and every doubling halves the matrix bytes an SM pulls from L2 per witness -- at 2 per thread the 5-word matrix
    // runs into the L2 -> SM port (about 44 GB/s per SM measured) before the multiply pipe is full
    m.pt = lay.nc != 5 ? 1 : planes % 4 == 0 ? 4 : planes % 2 == 0 ? 2 : 1;
    size_t stage_bytes = stage_bytes_for(m.pt, lay);
    // resident CTAs per SM the grid is sized for: two for the single-witness (3-word) kernel; the 5-word tiles leave room
    // for one CTA only, whatever the number of witnesses per thread
    uint32_t occ_cap = (m.pt == 1 && lay.nc != 5) ? 2 : 1;
    // as many stages as fit next to occ_cap resident CTAs (227 KB usable, 1 KB reserved per CTA), at most 6
    size_t per_cta = (227 * 1024) / occ_cap - SM_RESERVED_SMEM - 1024;  // 16 B of sync state per stage
    uint32_t stages = (uint32_t)(per_cta / stage_bytes);
    if (stages > 6) stages = 6;
    if (stages < 2) stages = 2;
    if (const char *e = getenv("LAT_MAC_STAGES")) stages = (uint32_t)atoi(e);  // tuning hooks (tools/tune_mac.py)
    if (const char *e = getenv("LAT_MAC_OCC")) occ_cap = (uint32_t)atoi(e);
    m.stages = stages;
    m.smem_bytes = m.stages * stage_bytes + 2 * m.stages * sizeof(u64);
    uint32_t groups = planes / m.pt;
    u64 want = (u64)sm_count * occ_cap;
    // plane groups and row blocks multiply the grid; keep the whole grid near one resident wave
    u64 gx = want / ((u64)groups * lay.nrb);
    if (gx < 1) gx = 1;
    if (const char *e = getenv("LAT_MAC_GRIDX")) gx = (u64)atoll(e);
    if (gx > lay.ntiles) gx = lay.ntiles;
    m.grid_x = (uint32_t)gx;
    m.nslots = m.grid_x * lay.cg;
    m.ws_elems = (size_t)planes * lay.kappa * ring::D * 2 + 2;  // (lo, hi) sums per output + the CTA counter
    return m;
}

template <int PT, int RG, bool TOOM>
static void launch_mac_t(dim3 grid, const u64 *A_dev, const MatLayout &lay, const u64 *Fx, u64 f_stride, uint32_t planes,
                         const MacPlan &plan, u64 *workspace, u64 *cms, cudaStream_t stream, bool pdl, const MacReport &report) {
    // function attributes are per device: a process may hold handles on several GPUs
    static bool attr_set_on[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool &attr_set = attr_set_on[dev & 63];
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mac_kernel<PT, RG, TOOM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);  // minus the static bytes
        if (e != cudaSuccess) fprintf(stderr, "lattice_ajtai: cudaFuncSetAttribute(mac_kernel<%d,%d,%d>): %s\n", PT, RG, (int)TOOM, cudaGetErrorString(e));
        attr_set = true;
    }
    if (getenv("LAT_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, mac_kernel<PT, RG, TOOM>);
        int occ = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mac_kernel<PT, RG, TOOM>, MacGeo<PT, RG, TOOM>::THREADS, plan.smem_bytes);
        fprintf(stderr, "mac_kernel<%d,%d,%d>: grid=(%u,%u,%u) block=%d smem=%zu stages=%u regs=%d maxDyn=%d static=%zu occ=%d maxThreads=%d\n",
                PT, RG, (int)TOOM, grid.x, grid.y, grid.z, MacGeo<PT, RG, TOOM>::THREADS, plan.smem_bytes, plan.stages, fa.numRegs,
                fa.maxDynamicSharedSizeBytes, fa.sharedSizeBytes, occ, fa.maxThreadsPerBlock);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(MacGeo<PT, RG, TOOM>::THREADS);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, mac_kernel<PT, RG, TOOM>, A_dev, lay, Fx, f_stride, planes, plan.stages, workspace, cms, (uint32_t)(pdl ? 1 : 0), report);
}

template <int PT, bool TOOM>
static void launch_mac_pt(dim3 grid, const u64 *A_dev, const MatLayout &lay, const u64 *Fx, u64 f_stride, uint32_t planes,
                          const MacPlan &plan, u64 *workspace, u64 *cms, cudaStream_t stream, bool pdl, const MacReport &report) {
    switch (lay.rg) {
        case 1: launch_mac_t<PT, 1, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        case 2: launch_mac_t<PT, 2, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        case 3: launch_mac_t<PT, 3, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        case 4: launch_mac_t<PT, 4, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        case 5: launch_mac_t<PT, 5, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        case 6: launch_mac_t<PT, 6, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        case 7: launch_mac_t<PT, 7, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
        default: launch_mac_t<PT, 8, TOOM>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report); break;
    }
}

void launch_mac(const u64 *A_dev, const MatLayout &lay, const u64 *Fx, u64 f_stride, uint32_t planes, const MacPlan &plan,
                u64 *workspace, u64 *cms, cudaStream_t stream, cudaEvent_t ev_begin, cudaEvent_t ev_end,
                const MacReport &report) {
    dim3 grid(plan.grid_x, lay.nrb, planes / plan.pt);
    // overlap with the producer kernel's tail unless events bracket the launch (they would serialise it anyway)
    static const bool pdl_off = getenv("LAT_NO_PDL") != nullptr;
    const bool pdl = !ev_begin && !pdl_off;
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    if (lay.nc == 5) {  // several witnesses: Toom-3 evaluations on both sides
        if (plan.pt == 1) launch_mac_pt<1, true>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report);
        else if (plan.pt == 4) launch_mac_pt<4, true>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report);
        else launch_mac_pt<2, true>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report);
    } else {            // one witness: Karatsuba on the 3-word matrix (bound by the bytes of the matrix)
        launch_mac_pt<1, false>(grid, A_dev, lay, Fx, f_stride, planes, plan, workspace, cms, stream, pdl, report);
    }
    if (ev_end) cudaEventRecord(ev_end, stream);
}

// cms[0] = cm - sum_{k>=1} 2^k cms[k]: Horner from the top plane, (acc + y_k) * 2.  decomposition.rs:189-197
__global__ void __launch_bounds__(256)
y0_kernel(const u64 *__restrict__ cm, u64 *__restrict__ cms, uint32_t K, uint32_t nwords) {
    u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nwords) return;
    u64 acc = 0;
    for (uint32_t k = K - 1; k >= 1; --k) acc = gl::mul_pow2<1>(gl::add(acc, cms[(u64)k * nwords + i]));
    cms[i] = gl::sub(cm[i], acc);
}

void launch_y0(const u64 *cm, u64 *cms, uint32_t K, uint32_t kappa, cudaStream_t stream) {
    uint32_t nwords = kappa * ring::D;
    y0_kernel<<<(nwords + 255) / 256, 256, 0, stream>>>(cm, cms, K, nwords);
}

// ---- compute_f_0: f0[j] = sum_i rho_i (*) f_i[j] over the 2K resident planes (LF/nifs/folding.rs:258-268) --------------
// One thread per (element, slot): it walks the planes (a 48-byte extended-layout slot each, consecutive threads read
// consecutive slots; Toom-3 form, as planes_kernel writes them), multiplies by rho_i's slot -- evaluated at the same five
// points once per block, in shared memory -- and accumulates lazily with the same Toom-3 accumulators as the MAC; one
// interpolation and reduction at the end.  HBM-bound on reading the planes once (2K x n x 384 B).
constexpr int FOLD_MAX_PLANES = 64;
constexpr int FOLD_THREADS = 256, FOLD_STAGES = 4;
constexpr u32 FOLD_STAGE_BYTES = FOLD_THREADS * 6 * 8;  // one plane's 48-byte slots of the block's 256 (element, slot) items
template <bool MONT>
__global__ void __launch_bounds__(FOLD_THREADS)
fold_kernel(const u64 *__restrict__ s0, const u64 *__restrict__ s1, const u64 *__restrict__ z0, const u64 *__restrict__ z1,
            int nsides, int pps, u64 n, u64 plane_stride, const u64 *__restrict__ rho, u64 *__restrict__ f0) {
    // The block's items are consecutive (element, slot) pairs, so plane p's share of them is ONE contiguous run of
    // 256 x 48 B: it is streamed with a TMA bulk copy into a ring of stages, FOLD_STAGES planes ahead of the
    // arithmetic (plain loads left the kernel latency-bound at 4.1 TB/s: two 256-thread blocks per SM cannot keep
    // enough 16-byte loads in flight).
    extern __shared__ __align__(128) unsigned char fold_smem[];
    __shared__ u64 s_rho[FOLD_MAX_PLANES * ring::NSLOT * 5];  // per (plane, slot): rho's values at 0, infinity, 1, -1, 2
    __shared__ __align__(8) u64 bars[FOLD_STAGES];
    const int nplanes = nsides * pps;
    const u64 total = n * ring::NSLOT;
    const u64 item0 = (u64)blockIdx.x * FOLD_THREADS;
    const u32 nitems = (u32)min((u64)FOLD_THREADS, total - item0);
    const u32 bytes = nitems * 48;
    // n elements starting at s* / z* / f0 (the caller offsets them for a sub-range).  Side s: plane 0 at z_s, planes 1 .. pps-1
    // at s_s, plane_stride elements apart (launch_planes keeps the committed planes of both sides back to back)
    auto plane_src = [&](int p) {
        const int side = p >= pps ? 1 : 0, k = p - side * pps;
        const u64 *base = k == 0 ? (side ? z1 : z0) : (side ? s1 : s0) + (u64)(k - 1) * plane_stride * FX;
        return base + item0 * 6;
    };
    if (threadIdx.x == 0) {
        for (int st = 0; st < FOLD_STAGES; ++st) mbar_init(&bars[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int p = 0; p < FOLD_STAGES && p < nplanes; ++p) {
            mbar_arrive_expect_tx(&bars[p], bytes);
            tma_bulk_g2s(fold_smem + (size_t)p * FOLD_STAGE_BYTES, plane_src(p), bytes, &bars[p]);
        }
    }
    for (int i = threadIdx.x; i < nplanes * ring::NSLOT; i += blockDim.x) {  // (plane, slot)
        u64 r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const u64 v = rho[i * 3 + c];
            r[c] = MONT ? gl::from_mont(v) : gl::reduce128(v, 0);  // canonical(rho) * repr(f) = repr(rho * f)
        }
        u64 *o = s_rho + i * 5;
        o[0] = r[0];
        o[1] = r[2];
        gl::toom_eval(r[0], r[1], r[2], o[2], o[3], o[4]);
    }
    __syncthreads();
    const u64 idx = item0 + threadIdx.x;  // (element, slot)
    const bool active = threadIdx.x < nitems;
    const u32 sl = (u32)(idx & 7);
    gl::ToomAcc acc;
    acc.clear();
    int st = 0;
    u32 ph = 0;
    for (int p = 0; p < nplanes; ++p) {
        mbar_wait(&bars[st], ph);
        if (active) {
            const ulonglong2 *pf = reinterpret_cast<const ulonglong2 *>(fold_smem + (size_t)st * FOLD_STAGE_BYTES) + threadIdx.x * 3;
            const ulonglong2 x = pf[0], y = pf[1], z = pf[2];
            const u64 *r = s_rho + (p * ring::NSLOT + sl) * 5;
            acc.mac(r[0], r[1], r[2], r[3], r[4], x.x, y.x, y.y, z.x, z.y);
        }
        __syncthreads();  // everyone has read the stage
        if (threadIdx.x == 0 && p + FOLD_STAGES < nplanes) {
            mbar_arrive_expect_tx(&bars[st], bytes);
            tma_bulk_g2s(fold_smem + (size_t)st * FOLD_STAGE_BYTES, plane_src(p + FOLD_STAGES), bytes, &bars[st]);
        }
        if (++st == FOLD_STAGES) {
            st = 0;
            ph ^= 1;
        }
    }
    if (!active) return;
    u64 c0, c1, c2;
    acc.finish(c0, c1, c2);
    u64 *o = f0 + idx * 3;  // element * 24 + slot * 3
    o[0] = c0; o[1] = c1; o[2] = c2;
}
void launch_fold(const u64 *const *sides_fx, const u64 *const *sides_fx0, int nsides, int planes_per_side, u64 n, const u64 *rho,
                 bool mont, u64 *f0, cudaStream_t stream, u64 elem0, u64 count) {
    const u64 plane_stride = n;
    if (count == ~0ull) count = n - elem0;
    n = count;
    if (!n) return;
    unsigned grid = (unsigned)((n * ring::NSLOT + FOLD_THREADS - 1) / FOLD_THREADS);
    const u64 *a = sides_fx[0] + elem0 * FX, *b = nsides > 1 ? sides_fx[1] + elem0 * FX : nullptr;
    const u64 *za = sides_fx0[0] + elem0 * FX, *zb = nsides > 1 ? sides_fx0[1] + elem0 * FX : nullptr;
    f0 += elem0 * ring::D;
    const size_t smem = (size_t)FOLD_STAGES * FOLD_STAGE_BYTES;
    static bool attr_set_on[64] = {};  // per device, as for mac_kernel
    int dev = 0;
    cudaGetDevice(&dev);
    bool &attr_set = attr_set_on[dev & 63];
    if (!attr_set) {
        cudaFuncSetAttribute(fold_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(fold_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (mont) fold_kernel<true><<<grid, FOLD_THREADS, smem, stream>>>(a, b, za, zb, nsides, planes_per_side, n, plane_stride, rho, f0);
    else fold_kernel<false><<<grid, FOLD_THREADS, smem, stream>>>(a, b, za, zb, nsides, planes_per_side, n, plane_stride, rho, f0);
}

// ---- cm_0 = sum_i rho_i (*) cm_i over the 2K commitments of a fold step (LF/nifs/folding/utils.rs:466-472) -----------------
// One WARP per (row, slot) of the kappa x 8 outputs, one lane per commitment: each lane's Fq3 product is independent (one
// round of loads instead of 2K dependent ones -- this kernel sits alone on the critical path of lat_ajtai_fold_step_finish),
// then a shuffle tree of modular additions.  Exact arithmetic mod q: the order of the sum does not matter.
template <bool MONT>
__global__ void __launch_bounds__(256)
lincomb_kernel(const u64 *__restrict__ rho, const u64 *__restrict__ cms0, const u64 *__restrict__ cms1, int K, uint32_t kappa,
               u64 *__restrict__ out) {
    const u32 lane = threadIdx.x & 31;
    const u32 idx = blockIdx.x * 8 + (threadIdx.x >> 5);  // row * 8 + slot
    if (idx >= kappa * ring::NSLOT) return;               // the whole warp leaves together
    const u32 sl = idx & 7;
    u64 acc[3] = {0, 0, 0};
    for (int p = lane; p < 2 * K; p += 32) {
        const u64 *c = (p < K ? cms0 + (u64)p * kappa * ring::D : cms1 + (u64)(p - K) * kappa * ring::D) + (u64)idx * 3;
        const u64 *r = rho + p * ring::D + 3 * sl;
        u64 a[3], y[3], t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            a[k] = MONT ? gl::from_mont(r[k]) : gl::reduce128(r[k], 0);  // canonical(rho) * repr(cm) = repr(rho * cm)
            y[k] = gl::reduce128(c[k], 0);
        }
        gl::fq3_mul(a, y, t);
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = gl::add(acc[k], t[k]);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = gl::add(acc[k], __shfl_down_sync(0xFFFFFFFFu, acc[k], off));
    }
    if (lane == 0) {
        u64 *o = out + (u64)idx * 3;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    }
}
void launch_lincomb(const u64 *rho, const u64 *cms0, const u64 *cms1, int K, uint32_t kappa, bool mont, u64 *out,
                    cudaStream_t stream) {
    const unsigned grid = (kappa * ring::NSLOT + 7) / 8;
    if (mont) lincomb_kernel<true><<<grid, 256, 0, stream>>>(rho, cms0, cms1, K, kappa, out);
    else lincomb_kernel<false><<<grid, 256, 0, stream>>>(rho, cms0, cms1, K, kappa, out);
}

// ---- gadget_recompose in CRT form: Horner from the top limb, result = result * B + v_l ---------------------------------
__global__ void __launch_bounds__(256)
recompose_kernel(const u64 *__restrict__ f, u64 nwords, int L, u64 bq, u64 *__restrict__ out) {
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;  // (element, word)
    if (i >= nwords) return;
    const u64 e = i / ring::D, t = i - e * ring::D;
    u64 acc = 0;
    for (int l = L - 1; l >= 0; --l) acc = gl::add(gl::mul(acc, bq), gl::reduce128(f[(e * L + l) * ring::D + t], 0));
    out[i] = acc;
}
void launch_recompose(const u64 *f, u64 count, int log2b, int L, u64 *out, cudaStream_t stream) {
    u64 nwords = count * ring::D;
    if (!nwords) return;
    recompose_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, stream>>>(f, nwords, L, 1ull << log2b, out);
}

// ---- fused exchange + fold of the column-shard partial commitments over NVLink peer memory --------------------------------
// One block per rank.  Push: plain stores of the 6 KB partial into every rank's receive slot (peer addresses go out over
// NVLink), a system-scope fence, then one release store per peer flag.  Pull: spin on the local flags with acquire loads
// until all ranks have delivered, then sum the `world` partials mod q from the local buffer (volatile loads: the data
// was written by other GPUs).
__global__ void __launch_bounds__(256)
exchange_kernel(const u64 *__restrict__ partial, u64 words, int rank, int world, PeerPtrs peers, u64 epoch,
                u64 *__restrict__ out, u64 *__restrict__ report_cm, unsigned long long *report_done,
                unsigned long long done_value, SpinGuard guard) {
    // Programmatic dependent launch on both sides: this block may become resident while the matrix-vector kernel that
    // produces `partial` is still running (it waits for it here), and the NEXT step's witness kernel may start behind
    // it at once -- so the whole exchange, NVLink latency included, hides under that kernel (lat_ajtai_set_step_overlap).
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const u64 slot = epoch & 1;
    for (int r = 0; r < world; ++r) {
        u64 *dst = peers.recv[r] + (slot * world + rank) * words;
        for (u64 i = threadIdx.x; i < words; i += blockDim.x) dst[i] = partial[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        u64 *f = peers.flags[threadIdx.x] + slot * world + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
        // bounded (SpinGuard): a peer that never delivers ends in LAT_E_CUDA on the host, not in a hung GPU
        const u64 *mine = peers.flags[rank] + slot * world + threadIdx.x;
        spin_until_equals(mine, epoch, guard, SPIN_PEER_FLAG, (u64)threadIdx.x | (epoch << 8));
    }
    __syncthreads();
    __threadfence_system();
    const volatile u64 *rb = peers.recv[rank] + slot * world * words;
    for (u64 i = threadIdx.x; i < words; i += blockDim.x) {
        u64 lo = 0, hi = 0;
        for (int r = 0; r < world; ++r) {
            u64 v = rb[(u64)r * words + i];
            lo += v;
            hi += (lo < v);
        }
        const u64 r = gl::reduce128(lo, hi);
        out[i] = r;
        if (report_cm) report_cm[i] = r;  // mapped host memory
    }
    if (report_done) {  // the host polls this word (see MacReport)
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(report_done), "l"(done_value) : "memory");
    }
}
void launch_exchange(const u64 *partial, u64 words, int rank, int world, const PeerPtrs &peers, u64 epoch, u64 *out,
                     cudaStream_t stream, u64 *report_cm, unsigned long long *report_done, unsigned long long done_value,
                     const SpinGuard &guard) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, exchange_kernel, partial, words, rank, world, peers, epoch, out, report_cm, report_done, done_value, guard);
}

// Sum of `count` partial commitments mod q (column-sharded multi-GPU exchange).
__global__ void __launch_bounds__(256)
commitment_sum_kernel(const u64 *__restrict__ parts, uint32_t count, u64 words, u64 *__restrict__ out) {
    u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= words) return;
    u64 lo = 0, hi = 0;
    for (uint32_t p = 0; p < count; ++p) {
        u64 v = parts[(u64)p * words + i];
        lo += v;
        hi += (lo < v);
    }
    out[i] = gl::reduce128(lo, hi);
}
void launch_commitment_sum(const u64 *parts, uint32_t count, u64 words, u64 *out, cudaStream_t stream) {
    if (!words) return;
    commitment_sum_kernel<<<(unsigned)((words + 255) / 256), 256, 0, stream>>>(parts, count, words, out);
}

}  // namespace lat

// Witness::from_w_ccs + Witness::commit as ONE cooperative launch with two kinds of CTA (the zkVM's shape: kappa in 29..32,
// i.e. one row block of 32 rows, L <= 8).  Reference: latticefold/src/arith.rs:230-248,357-362; zkvm/src/main.rs:348-367.
//
// The two-kernel chain (witness_kernel, then mac_kernel<1,8>) has a grid-wide dependency in the middle: the matrix stream
// cannot start before the LAST witness element is transformed, and for a host-buffer call that means before the last byte
// of w_ccs has crossed PCIe (76 us for 3.8 MB).  Here the step is cut into C column CHUNKS and both stages live in the same
// grid, two CTAs per SM:
//   * TRANSFORM CTAs walk the chunks in order; each transforms its 1/Xth of a chunk's w_ccs elements -- fetched one chunk
//     ahead by a bulk copy, straight from page-locked host memory when that is where they live -- writes the int16 digits
//     and the extended witness rows, and counts itself done for the chunk (release);
//   * MAC CTAs run the tile loop of mac_kernel<1,8> over their 1/Mth of every chunk, chunk by chunk; the lane that refills
//     a stage requests the matrix half at once and the witness half as soon as the chunk's counter is complete (acquire).
// The matrix stream therefore starts when chunk 0 is transformed and the upload of the later chunks runs under it.  The
// loop is issue-bound per SM, and a lone 8-warp CTA sustains an SM's full tile rate, so giving the
// second CTA slot of every SM to the transform costs the matrix stream only the issue slots the transform really uses
// (about a tenth).  Which CTA plays which role is decided on arrival, per SM (%smid), so that every SM gets one of each.
// Spin waits between CTAs of one grid need all of them resident: the launch is cooperative (it fails instead of
// deadlocking when the grid does not fit) and every wait is bounded (lat::SpinGuard).
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ring24.cuh"
#include "ring8.cuh"
#include "ring96.cuh"
#include "spin.cuh"
#include "tma.cuh"

namespace lat {

constexpr int ST_TJ = 8, ST_RB = 32, ST_THREADS = 256, ST_STAGES = 2, ST_FX = 48;
constexpr u32 ST_TILE_ELEMS = ST_TJ * 3 * ST_RB * 8, ST_TILE_BYTES = ST_TILE_ELEMS * 8, ST_F_BYTES = ST_TJ * ST_FX * 8;
constexpr u32 ST_STAGE_BYTES = ST_TILE_BYTES + ST_F_BYTES;
constexpr u32 ST_HDR_BYTES = 256;
constexpr u32 ST_SMEM = ST_HDR_BYTES + ST_STAGES * ST_STAGE_BYTES;  // the transform role lives inside the stage area
constexpr u32 ST_MAX_CHUNKS = 32;
constexpr u32 ST_SYNC_WORDS = 256 + 2 + ST_MAX_CHUNKS;  // per-SM arrival counters, role counters, chunk counters

size_t step_sync_bytes() { return ST_SYNC_WORDS * sizeof(u32); }

struct StepShape {   // how the columns are cut; the same arithmetic on both sides
    u64 ntiles, n, w_len;
    u32 L, C, M, X;  // limbs per element, chunks, MAC CTAs, transform CTAs
    __device__ __forceinline__ u64 chunk_tile(u32 c) const { return ntiles * c / C; }                 // first tile of chunk c
    __device__ __forceinline__ u64 chunk_elem(u32 c) const {                                           // first element
        const u64 col = min(n, chunk_tile(c) * ST_TJ);
        return c >= C ? w_len : (col + L - 1) / L;  // elements that END inside the earlier chunks' tiles belong to those
    }
    // MAC CTA m's tiles of chunk c
    __device__ __forceinline__ void mac_share(u32 c, u32 m, u64 &lo, u64 &hi) const {
        const u64 t0 = chunk_tile(c), cnt = chunk_tile(c + 1) - t0;
        lo = t0 + cnt * m / M;
        hi = t0 + cnt * (m + 1) / M;
    }
    // transform CTA x's elements of chunk c
    __device__ __forceinline__ void tr_share(u32 c, u32 x, u64 &lo, u64 &hi) const {
        const u64 e0 = chunk_elem(c), cnt = chunk_elem(c + 1) - e0;
        lo = e0 + cnt * x / X;
        hi = e0 + cnt * (x + 1) / X;
    }
};

struct StepHdr {   // shared-memory header of a MAC CTA (offset 48 behind the barriers)
    StepShape sh;
    const u64 *A;
    const u64 *fx;
    const u32 *chunk_done;
    SpinGuard guard;
    u32 m;            // this CTA's index among the MAC CTAs
    u32 iss_c;        // refill cursor: chunk and tile of the next request
    u64 iss_t;
    u32 ready_chunks; // chunks known complete
    u32 my_tiles;
};
static_assert(sizeof(StepHdr) + 48 <= ST_HDR_BYTES, "shared-memory header");

__device__ __forceinline__ void st256g(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void load_digits24(const int16_t *p, int (&d)[ring::D]) {
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

// ---- transform role ----------------------------------------------------------------------------------------------------------
// Elements [e0, e1) (at most ST_PIECE elements, already in shared memory at `piece`): iCRT, digits, CRT of the limbs.
constexpr u32 ST_PIECE = 32;  // elements per pass: 8 lanes each = the whole block in phase A
template <bool MONT>
__device__ __forceinline__ void transform_piece(const u64 *piece, u64 e0, u32 ne, u32 L, u32 log2b, int16_t *dtile, int16_t *__restrict__ f16,
                                                u64 *__restrict__ fx, int *__restrict__ flag) {
    const u32 sl = threadIdx.x & 7, oct = threadIdx.x >> 3;
    if ((oct & ~3u) < ne) {  // warp-uniform: the shuffles of an octet stay inside its warp
        const ring8::Twiddles tw = ring8::make_twiddles(sl);
        const u64 Bd = 1ull << log2b, halfB = Bd >> 1;
        const bool valid = oct < ne;
        const u64 *p = piece + (valid ? oct : 0) * ring::D + 3 * sl;
        u64 c[3] = {p[0], p[1], p[2]};
        ring8::icrt8(c, tw);
        bool negative[3];
        u64 mg[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if constexpr (MONT) c[q] = gl::from_mont(c[q]);
            ring::signed_rep(c[q], negative[q], mg[q]);  // fq_convertible.rs:22-34
        }
        int16_t *trow = dtile + (valid ? oct : 0) * (L * ring::D) + 3 * sl;
        for (u32 l = 0; l < L; ++l) {  // balanced_decomposition/mod.rs:76-97 on the magnitudes, limb by limb
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                u64 rem = mg[q] & (Bd - 1);
                mg[q] >>= log2b;
                int dg = (int)rem;
                if (rem > halfB) {  // |rem| == b/2 is kept (mod.rs:79)
                    dg -= (int)Bd;
                    mg[q] += 1;
                }
                if (negative[q]) dg = -dg;
                if (valid) trow[l * ring::D + q] = (int16_t)dg;
            }
        }
        if (valid && (mg[0] | mg[1] | mg[2])) atomicOr(flag, 1);  // the reference would index out of bounds (mod.rs:80)
    }
    __syncthreads();
    const u64 row0 = e0 * L;
    const u32 nrows = ne * L;
    {   // the resident int16 digits, 16 bytes at a time
        const uint4 *src = reinterpret_cast<const uint4 *>(dtile);
        uint4 *dst = reinterpret_cast<uint4 *>(f16 + row0 * ring::D);
        for (u32 u = threadIdx.x; u < nrows * 3; u += ST_THREADS) dst[u] = src[u];
    }
    for (u32 r = threadIdx.x; r < nrows; r += ST_THREADS) {  // one thread per limb element
        int d[ring::D];
        load_digits24(dtile + r * ring::D, d);
        u64 x[ring::D];
        r96::crt24_small<MONT>(d, x);
        u64 *o = fx + (row0 + r) * ST_FX;  // [slot][f0, f1, f2, f0+f1, f0+f2, f1+f2]: two slots = three 32-byte stores
#pragma unroll
        for (int s = 0; s < ring::NSLOT; s += 2) {
            const u64 a0 = x[3 * s], a1 = x[3 * s + 1], a2 = x[3 * s + 2];
            const u64 b0 = x[3 * s + 3], b1 = x[3 * s + 4], b2 = x[3 * s + 5];
            st256g(o + s * 6, a0, a1, a2, gl::add_lazy(a0, a1));
            st256g(o + s * 6 + 4, gl::add_lazy(a0, a2), gl::add_lazy(a1, a2), b0, b1);
            st256g(o + s * 6 + 8, b2, gl::add_lazy(b0, b1), gl::add_lazy(b0, b2), gl::add_lazy(b1, b2));
        }
    }
    __threadfence();  // the rows, before whoever counts this block done for the chunk
    __syncthreads();  // the digit tile and the piece are rewritten by the next pass
}

template <bool MONT>
__device__ __noinline__ void transform_role(unsigned char *smem, const StepShape sh, u32 x, const FusedWitness fw, u32 *chunk_done) {
    // shared memory of this role: two pieces of ST_PIECE elements (bulk-copy ring), one digit tile, two barriers
    u64 *piece = reinterpret_cast<u64 *>(smem + ST_HDR_BYTES);                                  // 2 x 6 KB
    int16_t *dtile = reinterpret_cast<int16_t *>(smem + ST_HDR_BYTES + 2 * ST_PIECE * ring::D * 8);  // 32 x L x 48 B <= 12 KB
    u64 *pbar = reinterpret_cast<u64 *>(smem);                                                   // [2]
    const u32 L = sh.L;
    if (threadIdx.x == 0) {
        mbar_init(&pbar[0], 1);
        mbar_init(&pbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (fw.ready_flag)  // a ticketed step: the upload runs on a copy engine, its ticket lands behind the data
            spin_until_equals(fw.ready_flag, fw.ready_value, fw.guard, SPIN_UPLOAD_TICKET, fw.ready_value);
    }
    __syncthreads();
    // the passes of this CTA, in order: (chunk c, elements [p0, p1)) with p1 - p0 <= ST_PIECE
    u32 c = 0;
    u64 lo = 0, hi = 0, p0 = 0;
    sh.tr_share(0, x, lo, hi);
    p0 = lo;
    auto advance = [&](u32 &cc, u64 &l, u64 &h, u64 &p) {  // to the next non-empty pass; cc == C when there is none
        p = min(h, p + ST_PIECE);
        while (p >= h && cc < sh.C) {
            ++cc;
            if (cc < sh.C) {
                sh.tr_share(cc, x, l, h);
                p = l;
            }
        }
    };
    auto issue = [&](u64 p, u64 h, u32 slot) {  // one thread: fetch elements [p, min(h, p + ST_PIECE))
        const u32 bytes = (u32)(min(h, p + ST_PIECE) - p) * ring::D * 8;
        mbar_arrive_expect_tx(&pbar[slot], bytes);
        tma_bulk_g2s(piece + slot * ST_PIECE * ring::D, fw.w + p * ring::D, bytes, &pbar[slot]);
    };
    while (lo >= hi && c < sh.C) {  // skip leading empty chunks
        ++c;
        if (c < sh.C) {
            sh.tr_share(c, x, lo, hi);
            p0 = lo;
        }
    }
    // cursor of the NEXT pass (for the prefetch)
    u32 nc = c;
    u64 nlo = lo, nhi = hi, np = p0;
    if (c < sh.C) {
        if (threadIdx.x == 0) issue(p0, hi, 0);
        advance(nc, nlo, nhi, np);
        if (threadIdx.x == 0 && nc < sh.C) issue(np, nhi, 1);
    }
    u32 pass = 0, done_c = 0;  // chunks [0, done_c) have been counted
    while (c < sh.C) {
        const u32 slot = pass & 1;
        mbar_wait(&pbar[slot], (pass >> 1) & 1);
        const u32 ne = (u32)(min(hi, p0 + ST_PIECE) - p0);
        transform_piece<MONT>(piece + slot * ST_PIECE * ring::D, p0, ne, L, (u32)fw.log2b, dtile, fw.f16, fw.fx, fw.flag);
        // (transform_piece ends with a barrier: the slot is free)  next-next pass into this slot
        const u32 pc = c;
        c = nc; lo = nlo; hi = nhi; p0 = np;
        if (nc < sh.C) advance(nc, nlo, nhi, np);
        if (threadIdx.x == 0 && c < sh.C && nc < sh.C) issue(np, nhi, slot);
        ++pass;
        // chunks this CTA has finished (its share of pc is complete when the cursor left it; empty shares count as well)
        if (c != pc) {
            __threadfence();  // the rows written by every thread of the block (barrier above) before the count
            if (threadIdx.x == 0)
                for (u32 k = done_c; k < min(c, sh.C); ++k) atomicAdd(&chunk_done[k], 1u);
            done_c = min(c, sh.C);
        }
    }
    if (threadIdx.x == 0)
        for (u32 k = done_c; k < sh.C; ++k) atomicAdd(&chunk_done[k], 1u);  // (only when this CTA had nothing at all)
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------------
template <bool MONT>
__global__ void __launch_bounds__(ST_THREADS, 2)
step_kernel(const u64 *__restrict__ A_dev, MatLayout lay, u64 *__restrict__ ws, u64 *__restrict__ cms, u32 *__restrict__ sync,
            u32 chunks, MacReport report, FusedWitness fw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ u32 s_role, s_idx, s_last;
    u32 *sm_count = sync, *role_count = sync + 256, *chunk_done = sync + 258;
    if (threadIdx.x == 0) {
        // one MAC CTA and one transform CTA per SM: the first arrival on an SM takes the matrix, the second the transform;
        // should the pairing come out uneven, the surplus switches sides so that both sides have exactly half the grid
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        u32 role = atomicAdd(&sm_count[smid & 255], 1u) & 1u;
        u32 idx = atomicAdd(&role_count[role], 1u);
        if (idx >= gridDim.x / 2) {
            role ^= 1u;
            idx = atomicAdd(&role_count[role], 1u);
        }
        s_role = role;
        s_idx = idx;
    }
    __syncthreads();
    StepShape sh;
    sh.ntiles = lay.ntiles; sh.n = lay.n; sh.w_len = fw.w_len; sh.L = (u32)fw.L; sh.C = chunks;
    sh.M = gridDim.x / 2; sh.X = gridDim.x - sh.M;
    const u32 lane = threadIdx.x & 31;
    gl::Fq3Acc acc;
    acc.clear();

    if (s_role == 1) {
        transform_role<MONT>(smem_raw, sh, s_idx, fw, chunk_done);
    } else {
        u64 *bars = reinterpret_cast<u64 *>(smem_raw);             // [tile full x2]
        u32 *released = reinterpret_cast<u32 *>(bars + ST_STAGES);  // [x2]
        StepHdr *hd = reinterpret_cast<StepHdr *>(smem_raw + 48);
        unsigned char *stages = smem_raw + ST_HDR_BYTES;
        // Request the next tile of this CTA (cursor in shared memory) into stage st: the matrix half at once, the witness
        // half once the tile's chunk is complete.  One lane at a time gets here (the refills of a CTA are ordered).
        auto issue_next = [&](u32 st) {
            u32 c = hd->iss_c;
            u64 t = hd->iss_t, lo, hi;
            hd->sh.mac_share(c, hd->m, lo, hi);
            while (t >= hi) {  // next non-empty share (the caller knows that a tile is left)
                ++c;
                hd->sh.mac_share(c, hd->m, lo, hi);
                t = lo;
            }
            const u32 fb = (u32)min((u64)ST_TJ, hd->sh.n - t * ST_TJ) * ST_FX * 8;
            mbar_arrive_expect_tx(&bars[st], ST_TILE_BYTES + fb);
#ifndef LAT_NO_L2_HINT
            tma_bulk_g2s_hint(stages + (size_t)st * ST_STAGE_BYTES, hd->A + t * ST_TILE_ELEMS, ST_TILE_BYTES, &bars[st], L2_EVICT_FIRST);
#else
            tma_bulk_g2s(stages + (size_t)st * ST_STAGE_BYTES, hd->A + t * ST_TILE_ELEMS, ST_TILE_BYTES, &bars[st]);
#endif
            if (c >= hd->ready_chunks) {  // acquire the chunk: every transform CTA has counted itself done for it
                spin_until_equals_u32(hd->chunk_done + c, hd->sh.X, hd->guard, SPIN_CHUNK, c);
                hd->ready_chunks = c + 1;
            }
            // rows written by other CTAs through the generic proxy, fetched by this bulk copy (async proxy)
            asm volatile("fence.proxy.async;" ::: "memory");
            tma_bulk_g2s(stages + (size_t)st * ST_STAGE_BYTES + ST_TILE_BYTES, hd->fx + t * ST_TJ * ST_FX, fb, &bars[st]);
            hd->iss_c = c;
            hd->iss_t = t + 1;
        };
        if (threadIdx.x == 0) {
            hd->sh = sh;
            hd->A = A_dev; hd->fx = fw.fx; hd->chunk_done = chunk_done; hd->guard = fw.guard;
            hd->m = s_idx;
            u32 total = 0;
            for (u32 c = 0; c < sh.C; ++c) {
                u64 lo, hi;
                sh.mac_share(c, s_idx, lo, hi);
                total += (u32)(hi - lo);
            }
            hd->my_tiles = total;
            u64 lo, hi;
            sh.mac_share(0, s_idx, lo, hi);
            hd->iss_c = 0;
            hd->iss_t = lo;
            hd->ready_chunks = 0;
            for (u32 st = 0; st < ST_STAGES; ++st) {
                mbar_init(&bars[st], 1);
                released[st] = 0;
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (u32 t = 0; t < min((u32)ST_STAGES, total); ++t) issue_next(t);
        }
        __syncthreads();
        const u32 my_tiles = hd->my_tiles;
        u32 st = 0, ph = 0;
        bool ready = false;
        for (u32 t = 0; t < my_tiles; ++t) {
            if (!ready) mbar_wait(&bars[st], ph);
            const u64 *sa = reinterpret_cast<const u64 *>(stages + (size_t)st * ST_STAGE_BYTES) + threadIdx.x;
            const ulonglong2 *sf = reinterpret_cast<const ulonglong2 *>(stages + (size_t)st * ST_STAGE_BYTES + ST_TILE_BYTES) + (lane & 7) * 3;
            u32 st_n = st + 1, ph_n = ph;
            if (st_n == ST_STAGES) {
                st_n = 0;
                ph_n ^= 1;
            }
            ready = (t + 1 < my_tiles) && mbar_test(&bars[st_n], ph_n);
#pragma unroll
            for (int jj = 0; jj < ST_TJ; ++jj) {
                const u64 *pa = sa + jj * (3 * ST_RB * 8);
                const u64 a0 = pa[0], a1 = pa[ST_RB * 8], a2 = pa[2 * ST_RB * 8];
                const ulonglong2 x = sf[jj * (ST_FX / 2)], y = sf[jj * (ST_FX / 2) + 1], z = sf[jj * (ST_FX / 2) + 2];
                acc.mac(a0, a1, a2, x.x, x.y, y.x, y.y, z.x, z.y);
            }
            __syncwarp();
            if (lane == 0) {
                if (atomicAdd(&released[st], 1u) == ST_THREADS / 32 - 1) {
                    released[st] = 0;
                    if (t + ST_STAGES < my_tiles) issue_next(st);
                }
            }
            st = st_n;
            ph = ph_n;
        }
    }

    // ===== epilogue: as mac_kernel; transform CTAs contribute nothing but are counted ============================================
    const u32 il = threadIdx.x >> 3, s = lane & 7;
    if (s_role == 0 && il < lay.kappa) {
        u64 c[3];
        acc.finish(c[0], c[1], c[2]);
        u64 *dst = ws + 2 * ((u64)il * ring::D + s * 3);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            atomicAdd(reinterpret_cast<unsigned long long *>(dst + 2 * q), c[q] & 0xFFFFFFFFull);
            atomicAdd(reinterpret_cast<unsigned long long *>(dst + 2 * q + 1), c[q] >> 32);
        }
    }
    const u64 nout = (u64)lay.kappa * ring::D;
    u64 *counter = ws + 2 * nout;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = (atomicAdd(reinterpret_cast<unsigned long long *>(counter), 1ull) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (u64 i = threadIdx.x; i < nout; i += blockDim.x) {
            const u64 lo = __ldcg(ws + 2 * i), hi = __ldcg(ws + 2 * i + 1);
            const u64 v_lo = lo + (hi << 32);
            const u64 v_hi = (hi >> 32) + (v_lo < lo ? 1ull : 0ull);
            const u64 v = gl::reduce128(v_lo, v_hi);
            cms[i] = v;
            if (report.cm_host) report.cm_host[i] = v;
            ws[2 * i] = 0;
            ws[2 * i + 1] = 0;
        }
        for (u32 i = threadIdx.x; i < ST_SYNC_WORDS; i += blockDim.x) sync[i] = 0;  // every other CTA is past its last use
        if (threadIdx.x == 0) *counter = 0;
        if (threadIdx.x == 0 && report.flag_dev) {
            *report.flag_host = *reinterpret_cast<volatile int *>(report.flag_dev);
            *report.flag_dev = 0;
        }
        if (report.done_host) {
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(report.done_host), "l"(report.done_value) : "memory");
        }
    }
}

// 0 = launched; otherwise the cudaError_t (as int) of a launch that did not happen -- in particular when the grid cannot be
// co-resident -- and the caller takes the two-kernel chain.
int launch_step(const u64 *A_dev, const MatLayout &lay, int sm_count, u64 *workspace, u64 *cms, uint32_t *sync, cudaStream_t stream, bool mont,
                const FusedWitness &fw, cudaEvent_t ev_begin, cudaEvent_t ev_end, const MacReport &report) {
    static int fits_on[64] = {};  // 0 unknown, 1 yes, -1 no
    int dev = 0;
    cudaGetDevice(&dev);
    int &fits = fits_on[dev & 63];
    if (!fits) {
        cudaFuncSetAttribute(step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM);
        cudaFuncSetAttribute(step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM);
        int occ = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_kernel<true>, ST_THREADS, ST_SMEM);
        fits = (e == cudaSuccess && occ >= 2) ? 1 : -1;
        cudaGetLastError();
    }
    if (fits < 0) return (int)cudaErrorCooperativeLaunchTooLarge;
    uint32_t chunks = 8;
    if (const char *e = getenv("LAT_STEP_CHUNKS")) chunks = (uint32_t)atoi(e);
    if (chunks < 1) chunks = 1;
    if (chunks > ST_MAX_CHUNKS) chunks = ST_MAX_CHUNKS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * sm_count);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = ST_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ev_begin) cudaEventRecord(ev_begin, stream);
    cudaError_t e;
    if (mont) e = cudaLaunchKernelEx(&cfg, step_kernel<true>, A_dev, lay, workspace, cms, sync, chunks, report, fw);
    else e = cudaLaunchKernelEx(&cfg, step_kernel<false>, A_dev, lay, workspace, cms, sync, chunks, report, fw);
    if (ev_end) cudaEventRecord(ev_end, stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fits = -1;
    }
    return (int)e;
}

}  // namespace lat

// Internal launcher interface between engine.cu (the C ABI) and the kernel translation units.
// Everything here is asynchronous on `stream`; pointers are device pointers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace lat {

typedef unsigned long long u64;

// Every in-kernel wait for something another engine or another GPU delivers (the upload ticket of a pipelined step,
// the peers' flags of an exchange) is bounded by %globaltimer: after `timeout_ns` the waiter stores
// code | detail << 8 into *status (a page-locked word mapped into the device, read by the host in lat_ajtai_wait /
// lat_ajtai_synchronize / lat_device_wait_status) and carries on, so that a protocol slip ends in LAT_E_CUDA with a
// message instead of a hung GPU.  timeout_ns = 0 waits for ever; status may be null.
struct SpinGuard {
    unsigned long long *status = nullptr;
    unsigned long long timeout_ns = 0;
};
constexpr unsigned long long SPIN_UPLOAD_TICKET = 1, SPIN_PEER_FLAG = 2;

// ---- ring_kernels.cu ------------------------------------------------------------------------------------
// Batched CRT / iCRT of `count` ring elements of 24 u64 (in-place allowed).
void launch_crt(const u64 *in, u64 *out, u64 count, cudaStream_t stream);
void launch_icrt(const u64 *in, u64 *out, u64 count, cudaStream_t stream);

// Witness::from_w_ccs fused: w (CRT form unless in_coeff; `mont` says which representation) -> iCRT -> balanced
// digits base 2^log2b, L limbs -> CRT of every limb.
//   f16      : w_len*L x 24 int16 digits, element-major / limb-minor (always written)
//   f_coeff  : the digits as u64 field elements in the caller's representation, or nullptr
//   f_plain  : CRT form, w_len*L x 24, or nullptr
//   fx       : CRT form in the MAC kernel's extended layout, w_len*L x 48, or nullptr
//   flag     : device int, OR-ed with 1 when a coefficient does not fit in L digits
//   stage_input: w lives in mapped page-locked host memory: fetch each block's elements with one bulk copy (w 16-B aligned)
void launch_witness(const u64 *w, u64 w_len, int log2b, int L, bool mont, bool in_coeff, int16_t *f16, u64 *f_coeff,
                    u64 *f_plain, u64 *fx, int *flag, cudaStream_t stream, bool overlap_previous = false,
                    const unsigned long long *ready_flag = nullptr, unsigned long long ready_value = 0,
                    const SpinGuard &guard = SpinGuard(), bool stage_input = false);

// int16 coefficients -> K base-2 digit planes: plane k of element j = sign * bit_k(|c|).
//   planes_f     : K x n x 24 CRT form, or nullptr
//   planes_fx    : CRT form in the extended layout, Toom-3 form (see FX_WORDS), or nullptr: planes 1 .. K-1 as
//                  (K-1) x n x 48 starting here, plane 0 (never committed: y_0 is derived) as n x 48 at planes_fx0 --
//                  so that the committed planes of several decompositions can sit back to back for ONE launch
//   planes_coeff : K x n x 24 coefficient form, or nullptr
//   lut          : the 3 x 256 x 8 subset-sum table of launch_planes_lut for the same representation
void launch_planes(const int16_t *f16, u64 n, int K, bool mont, const u64 *lut, u64 *planes_f, u64 *planes_fx,
                   u64 *planes_fx0, u64 *planes_coeff, cudaStream_t stream);
// lut[(c * 256 + pat) * 8 + s] = sum over the bits i of pat of the word of slot s that CRT(X^(3i+c)) is nonzero in
// (PLANES_LUT_WORDS u64, canonical values in the representation `mont` names); see planes_kernel.
constexpr int PLANES_LUT_WORDS = 3 * 256 * 8;
void launch_planes_lut(bool mont, u64 *lut, cudaStream_t stream);

// int16 digits (n x 24) -> Witness::get_fhat tables: fhat[j][i][3 s] = digit 8 j + s of element i as a field element in
// the caller's representation, the other two components of every slot zero; fhat: 3 x n x 24.
void launch_fhat(const int16_t *f16, u64 n, bool mont, u64 *fhat, cudaStream_t stream);

// u64 coefficient-form elements (caller's representation) -> int16, flag |= 1 unless |c| < 2^bits for all c.
void launch_pack_coeff(const u64 *f_coeff, u64 count, bool mont, int bits, int16_t *f16, int *flag,
                       cudaStream_t stream);

// ---- mac_kernels.cu -------------------------------------------------------------------------------------
// Geometry of the device-resident matrix, fixed at creation.
struct MatLayout {
    uint32_t kappa;      // logical rows
    uint32_t kappa_pad;  // rows incl. zero padding
    uint32_t rb;         // rows per row block (multiple of 4, <= 32)
    uint32_t nrb;        // row blocks
    uint32_t rg;         // row groups (warps along rows) per CTA = rb / 4
    uint32_t cg;         // column groups (warps along columns) per CTA
    uint32_t tj;         // columns per tile
    uint32_t nc;         // u64 per Fq3 entry: 3 (components; the single-witness kernel) or 5 (Toom-3 evaluations; several witnesses)
    u64 n;               // logical columns
    u64 n_pad;           // columns incl. zero padding (multiple of tj)
    u64 ntiles;          // n_pad / tj
    __host__ __device__ u64 tile_elems() const { return (u64)tj * nc * rb * 8; }          // u64 per tile
    __host__ __device__ u64 total_elems() const { return tile_elems() * ntiles * nrb; }   // u64 in the whole matrix
};
MatLayout make_layout(uint32_t kappa, u64 n, bool toom = false);

// The 5-word matrix of the several-witness kernels, derived on the device from the 3-word one (both canonical): entry
// (a0, a1, a2) -> its values at 0, infinity, 1, -1, 2 = (a0, a2, a0+a1+a2, a0-a1+a2, a0+2 a1+4 a2) mod q.  A5_dev must be
// zero where it is padding (columns >= n); only columns < n are written.
void launch_derive_toom(const u64 *A_dev, const MatLayout &lay, u64 *A5_dev, const MatLayout &lay5, cudaStream_t stream);

// rows: nrows x row_stride x 24 (CRT form, caller's representation) -> tiles of A_dev (canonical form).
void launch_relayout(const u64 *rows, uint32_t row0, uint32_t nrows, u64 row_stride, bool mont, const MatLayout &lay,
                     u64 *A_dev, cudaStream_t stream);

// Number of partial-sum slots mac needs for `planes` witnesses, and the workspace size in u64.
struct MacPlan {
    uint32_t pt;        // planes per thread
    uint32_t grid_x;    // column-chunk CTAs
    uint32_t nslots;    // partial sums per output (= CTAs x column groups contributing to it)
    size_t ws_elems;    // u64 of workspace (must be zeroed once; launches leave it zeroed)
    size_t smem_bytes;
    uint32_t stages;
};
MacPlan plan_mac(const MatLayout &lay, uint32_t planes, int sm_count);

// Extended witness layout consumed by the MAC kernel: per element 8 slots x 6 u64, in one of two forms --
//   Karatsuba (one witness per launch):  (f0, f1, f2, f0+f1, f0+f2, f1+f2)
//   Toom-3 (several witnesses, planes):  (f0, f1, f2, f(1), f(-1), f(2)) = (.., f0+f1+f2, f0-f1+f2, f0+2 f1+4 f2)
// The producer and the launch must agree: witness_kernel writes the first form, planes_kernel the second, fext either.
constexpr int FX_WORDS = 48;
// f (count x 24, CRT form, any representation) -> fx (count x 48)
void launch_fext(const u64 *f, u64 count, u64 *fx, cudaStream_t stream, bool toom = false);

// cms[p][i][24] = sum_j A[i][j] * F[p][j]   for p < planes;  Fx: planes x f_stride x 48 in the extended layout
// (f_stride >= n in elements; no padding needed).  lay.nc selects the form: 3 = A_dev in components and Fx in the
// Karatsuba form, 5 = A_dev in Toom-3 evaluations (launch_derive_toom) and Fx in the Toom-3 form.
// Optional completion report straight into page-locked host memory (pipelined host-buffer steps): the last CTA also
// stores the commitment to cm_host, moves the step's overflow flag to flag_host (clearing the device copy) and then
// publishes done_value in *done_host with system scope -- no copy, no event, nothing between two kernels in the stream.
struct MacReport {
    u64 *cm_host = nullptr;
    int *flag_dev = nullptr;
    int *flag_host = nullptr;
    unsigned long long *done_host = nullptr;
    unsigned long long done_value = 0;
};
void launch_mac(const u64 *A_dev, const MatLayout &lay, const u64 *Fx, u64 f_stride, uint32_t planes, const MacPlan &plan,
                u64 *workspace, u64 *cms, cudaStream_t stream, cudaEvent_t ev_begin = nullptr,
                cudaEvent_t ev_end = nullptr, const MacReport &report = MacReport());

// cms[0] = cm - sum_{k=1..K-1} 2^k cms[k]      (LF/nifs/decomposition.rs:189-197)
void launch_y0(const u64 *cm, u64 *cms, uint32_t K, uint32_t kappa, cudaStream_t stream);

// f0[j] = sum_{i < nplanes} rho[i] (*) planes[i][j]  (slot-wise Fq3), planes given per side as launch_planes writes them
// (Toom-3 form): sides_fx[s] = planes 1 .. planes_per_side-1 of side s, (planes_per_side-1) x n x 48; sides_fx0[s] = its
// plane 0, n x 48.  rho: nplanes x 24 in the caller's representation, side-major.  f0: n x 24.
// elem0 / count: fold only that range of elements (f0 is still the full n x 24 array; count = ~0 means "to the end").
void launch_fold(const u64 *const *sides_fx, const u64 *const *sides_fx0, int nsides, int planes_per_side, u64 n, const u64 *rho,
                 bool mont, u64 *f0, cudaStream_t stream, u64 elem0 = 0, u64 count = ~0ull);

// out[i] = sum_{p < 2K} rho[p] (*) cms[p][i]  over the K commitments of side 0 followed by the K of side 1 (each
// K x kappa x 24); the folded commitment cm_0 (LF/nifs/folding/utils.rs:466-472).  rho in the caller's representation.
void launch_lincomb(const u64 *rho, const u64 *cms0, const u64 *cms1, int K, uint32_t kappa, bool mont, u64 *out,
                    cudaStream_t stream);

// out[i] = sum_l (2^log2b)^l * f[i*L + l], CRT form (scalar multiples, so any representation)
void launch_recompose(const u64 *f, u64 count, int log2b, int L, u64 *out, cudaStream_t stream);

// out[i] = sum_{p < count} parts[p * words + i] mod q
void launch_commitment_sum(const u64 *parts, uint32_t count, u64 words, u64 *out, cudaStream_t stream);

// Peer-memory exchange + fold (see lat_commitment_exchange_dev).  recv/flags: up to 16 peer-mapped addresses.
constexpr int MAX_PEERS = 16;
struct PeerPtrs {
    u64 *recv[MAX_PEERS];
    u64 *flags[MAX_PEERS];
};
void launch_exchange(const u64 *partial, u64 words, int rank, int world, const PeerPtrs &peers, u64 epoch, u64 *out,
                     cudaStream_t stream, u64 *report_cm = nullptr,
                     unsigned long long *report_done = nullptr, unsigned long long done_value = 0,
                     const SpinGuard &guard = SpinGuard());

// Standalone negacyclic NTT over Z_q[X]/(X^d + 1), d = 2^logd (ntt_pow2.cu; not on the drop-in path).  Returns 0 or a
// cudaError_t as int.
int launch_ntt_pow2(const u64 *in, u64 *out, u64 batch, uint32_t logd, bool inverse, cudaStream_t stream);

}  // namespace lat

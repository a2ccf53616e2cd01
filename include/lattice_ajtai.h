/*
 * lattice_ajtai.h -- C ABI of the B200-native Ajtai commitment engine for Latticeum's LatticeFold prover.
 *
 * This is the drop-in boundary for ONE path of Nesquiko/Latticeum: the per-fold-step Ajtai commitment
 * pipeline (iCRT -> balanced gadget decomposition -> CRT -> kappa x n matrix-vector product in CRT form)
 * over the Goldilocks ring Z_q[X]/(X^24 - X^12 + 1), q = 2^64 - 2^32 + 1.  The reference has no FFI of its
 * own (it is pure Rust); each entry point below names the Rust function it replaces.  A `crates/zkvm-cuda`
 * FFI crate binds exactly these symbols.  Paths are relative to
 * /root/reference/latticeum/ :
 *     LF     = crates/latticefold/src
 *     RING   = crates/stark-rings/crates/ring/src
 *     LINALG = crates/stark-rings/crates/linear_algebra/src
 *     GOLD   = RING/cyclotomic_ring/models/goldilocks
 *     ZKVM   = crates/zkvm/src
 *
 * Data layout at the boundary (RING/cyclotomic_ring/flatten.rs:10-17, GOLD/utils.rs:5-23):
 *   - a ring element is 24 contiguous uint64_t, little-endian limbs;
 *     CRT ("NTT") form: index = slot*3 + component (8 slots x Fq3);  coefficient form: index = degree;
 *   - vectors of ring elements are contiguous; the matrix is uploaded row by row (the host's
 *     Matrix<R> is Vec<Vec<R>>, LINALG/matrix.rs:17-21);
 *   - `repr` fixes what a limb means for EVERY buffer crossing this ABI through that handle:
 *       LAT_REPR_CANONICAL  : the integer x in [0, q);
 *       LAT_REPR_MONTGOMERY : x * 2^64 mod q, the in-memory form of ark-ff's Fp64<MontBackend>
 *                             (GOLD/mod.rs:20-24) -- what a Rust caller passes without conversion.
 *
 * Pointers are HOST memory unless the function name ends in `_dev` (then they are device pointers on the
 * handle's GPU and the call is asynchronous on the handle's stream).  A handle is used by one thread at a
 * time (the reference calls this path from its single main thread, ZKVM/main.rs:121-219).
 *
 * Every function returns a lat_status.  There is no CPU fallback: without a usable CUDA device
 * lat_ajtai_create fails with LAT_E_CUDA.
 */
#ifndef LATTICE_AJTAI_H
#define LATTICE_AJTAI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAT_RING_DEGREE 24 /* GOLD/ntt.rs:9  */
#define LAT_RING_SLOTS 8   /* GOLD/ntt.rs:12 */
#define LAT_ABI_VERSION 1

typedef enum lat_status {
    LAT_OK = 0,
    /* CommitmentError::WrongWitnessLength(got, expected)            LF/commitment.rs:15-17 */
    LAT_E_WRONG_WITNESS_LENGTH = 1,
    /* CommitmentError::WrongCommitmentLength                        LF/commitment.rs:18-20 */
    LAT_E_WRONG_COMMITMENT_LENGTH = 2,
    /* CommitmentError::WrongAjtaiMatrixDimensions                   LF/commitment.rs:21-26 */
    LAT_E_WRONG_MATRIX_DIMENSIONS = 3,
    /* A coefficient needs more digits than the padding: the reference panics (index out of bounds at
     * RING/balanced_decomposition/mod.rs:80,85,87); this engine reports it and never truncates.        */
    LAT_E_DIGIT_OVERFLOW = 4,
    LAT_E_INVALID_ARGUMENT = 5,
    LAT_E_CUDA = 6,            /* CUDA runtime failure or no device; see lat_last_error()              */
    LAT_E_MATRIX_INCOMPLETE = 7 /* a commit was requested before all kappa rows were uploaded           */
} lat_status;

typedef enum lat_repr { LAT_REPR_CANONICAL = 0, LAT_REPR_MONTGOMERY = 1 } lat_repr;

/* Opaque engine: owns the device-resident Ajtai matrix, work buffers and a stream on one GPU.
 * Replaces AjtaiCommitmentScheme<GoldilocksRingNTT> (LF/commitment/commitment_scheme.rs:38-58). */
typedef struct lat_ajtai lat_ajtai;

const char *lat_strerror(int status);
/* Message of the last failure on the calling thread (CUDA error string etc.); never NULL. */
const char *lat_last_error(void);
int lat_abi_version(void);

/* ---- construction: AjtaiCommitmentScheme::new / ::rand   LF/commitment/commitment_scheme.rs:43-58 ---------
 * kappa x n matrix over the d = 24 Goldilocks ring.  (log2_B, L) are DecompositionParams::B, ::L and K is
 * ::K with B_SMALL fixed to 2 (LF/decomposition_parameters.rs:11-20; zkVM: 15, 5, 15, ZKVM/ccs.rs:26-34).
 * Requires 1 <= log2_B <= 15, 1 <= L <= 8, 1 <= K <= 15 (digits are held as int16 on the device).
 * `device` is the CUDA ordinal.  The matrix starts empty: upload all rows before committing.             */
int lat_ajtai_create(lat_ajtai **out, uint32_t kappa, uint64_t n, uint32_t log2_B, uint32_t L, uint32_t K,
                     int repr, int device);
void lat_ajtai_destroy(lat_ajtai *h);

/* Upload rows [row0, row0 + nrows) of the matrix.  Row r starts at rows + r * row_stride * 24 and holds n
 * ring elements in CRT form (row_stride >= n, in ring elements; pass n for a dense block, or the full
 * width when uploading a column shard of a wider host matrix).  The engine re-lays the matrix out for
 * streaming and, for LAT_REPR_MONTGOMERY, converts it to canonical form once here.                       */
int lat_ajtai_upload_rows(lat_ajtai *h, uint32_t row0, uint32_t nrows, const uint64_t *rows, uint64_t row_stride);
int lat_ajtai_upload_rows_dev(lat_ajtai *h, uint32_t row0, uint32_t nrows, const uint64_t *rows_dev,
                              uint64_t row_stride);

/* AjtaiCommitmentScheme::kappa / ::width                    LF/commitment/commitment_scheme.rs:85-94 */
uint32_t lat_ajtai_kappa(const lat_ajtai *h);
uint64_t lat_ajtai_width(const lat_ajtai *h);

/* Use an existing CUDA stream (cudaStream_t) for all work of this handle, e.g. the caller's current
 * stream so that its own events bracket the kernels.  NULL restores the handle's own stream; to name
 * the legacy default stream pass cudaStreamLegacy ((void *)1).                                          */
int lat_ajtai_set_stream(lat_ajtai *h, void *cuda_stream);
/* Block until all work queued on the handle's stream has finished; returns the sticky asynchronous status
 * (LAT_E_DIGIT_OVERFLOW raised by a *_dev call, or LAT_OK) and clears it.                                 */
int lat_ajtai_synchronize(lat_ajtai *h);

/* ---- commitments ------------------------------------------------------------------------------------------
 * commit / commit_ntt: cm = A * f.        LF/commitment/commitment_scheme.rs:63-80,101-103 -> LINALG/matrix.rs:168-178
 * f: f_len ring elements in CRT form; f_len != n -> LAT_E_WRONG_WITNESS_LENGTH.  cm: kappa ring elements.  */
int lat_ajtai_commit_ntt(lat_ajtai *h, const uint64_t *f, uint64_t f_len, uint64_t *cm);
int lat_ajtai_commit_ntt_dev(lat_ajtai *h, const uint64_t *f_dev, uint64_t f_len, uint64_t *cm_dev);
/* Batched: `count` witnesses of n elements each, contiguous; cms: count x kappa x 24.  One launch streams the
 * matrix once for the whole batch (the reference loops, LF/nifs/decomposition.rs:185-187).                  */
int lat_ajtai_commit_ntt_batch(lat_ajtai *h, const uint64_t *fs, uint32_t count, uint64_t f_len, uint64_t *cms);
int lat_ajtai_commit_ntt_batch_dev(lat_ajtai *h, const uint64_t *fs_dev, uint32_t count, uint64_t f_len,
                                   uint64_t *cms_dev);
/* commit_coeff: CRT every element, then commit.              LF/commitment/commitment_scheme.rs:107-112 */
int lat_ajtai_commit_coeff(lat_ajtai *h, const uint64_t *f_coeff, uint64_t f_len, uint64_t *cm);
/* decompose_and_commit_coeff: base-B, L-limb balanced decomposition (element-major, limb-minor), CRT,
 * commit.  w_coeff: w_len elements, w_len * L must equal n.  LF/commitment/commitment_scheme.rs:116-127   */
int lat_ajtai_decompose_and_commit_coeff(lat_ajtai *h, const uint64_t *w_coeff, uint64_t w_len, uint64_t *cm);
/* decompose_and_commit_ntt: iCRT first.                      LF/commitment/commitment_scheme.rs:132-139 */
int lat_ajtai_decompose_and_commit_ntt(lat_ajtai *h, const uint64_t *w, uint64_t w_len, uint64_t *cm);

/* ---- Witness::from_w_ccs (+ Witness::commit)                LF/arith.rs:230-248, 357-362; ZKVM/main.rs:348-367
 * w_ccs: w_len elements in CRT form, w_len * L == n.  Computes w_coeff = iCRT(w_ccs),
 * f_coeff = gadget_decompose(w_coeff, B, L) (out[i*L + l] = limb l of element i,
 * RING/balanced_decomposition/mod.rs:163-175), f = CRT(f_coeff) and cm = A * f in one pass on the device.
 * Any of f_coeff (n x 24), f (n x 24), cm (kappa x 24) may be NULL to skip that output (and its copy).
 * The decomposed witness stays resident on the device as the handle's "current witness" for
 * lat_ajtai_decompose_commit_resident.                                                                    */
int lat_ajtai_witness_from_w_ccs(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, uint64_t *f_coeff,
                                 uint64_t *f, uint64_t *cm);
int lat_ajtai_witness_from_w_ccs_dev(lat_ajtai *h, const uint64_t *w_ccs_dev, uint64_t w_len,
                                     uint64_t *f_coeff_dev, uint64_t *f_dev, uint64_t *cm_dev);

/* The same call returning Witness::f_coeff as the int16 digits the device holds (n x 24 int16, 4.7 MB instead of
 * 19 MB at the zkVM's size): base-B limbs satisfy |digit| <= B/2 <= 2^14, so int16 is lossless, and the host widens
 * them into Fq (negative digits are q - |d|) where its MLE code wants field elements -- f_hat (LF/arith.rs:273-297) is
 * a re-layout of exactly these digits.  f_coeff16 and f may be NULL.                                            */
int lat_ajtai_witness_from_w_ccs_compact(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, int16_t *f_coeff16,
                                         uint64_t *f, uint64_t *cm);

/* Non-blocking form of the same call: lat_ajtai_submit_w_ccs queues upload -> iCRT/decompose/CRT -> A * f -> report
 * of cm and returns at once with a ticket; lat_ajtai_wait blocks until that ticket's cm has been written (and reports
 * its LAT_E_DIGIT_OVERFLOW, if any).  What it is for: work that is INDEPENDENT of the commitment -- the host's own
 * MLE/sumcheck work of the same step, or independent provers sharing one matrix.  It does NOT pipeline the steps of
 * one zkVM run against each other: step i+1's z is built from ivc_output.{acc, w_acc, folding_proof}
 * (ZKVM/main.rs:140-156), which come out of fold(cm_i, w_i) (ZKVM/main.rs:174-182), so consecutive IVC steps are
 * strictly dependent and a drop-in caller sees the latency of one ticket (bench.py reports that figure as e2e).
 * At most LAT_PIPELINE_DEPTH tickets may be outstanding; w_ccs and cm must stay valid until the ticket has been
 * waited for.  The upload runs on a copy engine; with page-locked w_ccs (lat_host_alloc / cudaHostRegister) it
 * overlaps the kernels of other tickets, pageable memory works but serialises the upload.  The handle's "current
 * witness" is the one submitted last.  All per-slot state is allocated at lat_ajtai_create; nothing is allocated,
 * memset or synchronised on the submit path.                                                                   */
#define LAT_PIPELINE_DEPTH 4
int lat_ajtai_submit_w_ccs(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, uint64_t *cm, uint64_t *ticket);
int lat_ajtai_wait(lat_ajtai *h, uint64_t ticket);

/* ---- LFDecompositionProver::{decompose_witness, commit_witnesses}   LF/nifs/decomposition.rs:162-201
 * f_coeff: n elements in coefficient form whose signed representatives satisfy |c| < 2^K, else
 * LAT_E_DIGIT_OVERFLOW (the reference panics).  Plane k, element j = digit k (base 2, so sign * bit_k(|c|))
 * of f_coeff[j] (LF/nifs/decomposition/utils.rs:45-49).  cm: the commitment of the undecomposed witness
 * (cm_i.cm), kappa x 24.  Outputs, each may be NULL:
 *   planes_coeff : K x n x 24, coefficient form (Witness::f_coeff of each wit_s[k]);
 *   planes_f     : K x n x 24, CRT form        (Witness::f of each wit_s[k], LF/arith.rs:327);
 *   cms          : K x kappa x 24; cms[k] = A * planes_f[k] for k >= 1 and
 *                  cms[0] = cm - fold_rev((acc + cms[k]) * 2) by homomorphism (decomposition.rs:189-197).
 * All K - 1 matrix commits run in one batched launch that streams the matrix once.                         */
int lat_ajtai_decompose_commit(lat_ajtai *h, const uint64_t *f_coeff, uint64_t n, const uint64_t *cm,
                               uint64_t *planes_coeff, uint64_t *planes_f, uint64_t *cms);
int lat_ajtai_decompose_commit_dev(lat_ajtai *h, const uint64_t *f_coeff_dev, uint64_t n, const uint64_t *cm_dev,
                                   uint64_t *planes_coeff_dev, uint64_t *planes_f_dev, uint64_t *cms_dev);
/* In the _dev form cm_dev may be NULL: cms[0] is then left untouched.  Column-sharded callers do this -- their
 * cms[1..K-1] are partial sums over one column block, and y_0 needs the exchanged totals -- and finish with
 * lat_commitment_y0_dev: cms[0] = cm - sum_{k>=1} 2^k cms[k]  (LF/nifs/decomposition.rs:189-197), cms: K x kappa x 24. */
int lat_commitment_y0_dev(const uint64_t *cm_dev, uint64_t *cms_dev, uint32_t K, uint32_t kappa, void *cuda_stream);
/* Same, on the witness left resident by the last lat_ajtai_witness_from_w_ccs* call (no re-upload).        */
int lat_ajtai_decompose_commit_resident(lat_ajtai *h, const uint64_t *cm, uint64_t *planes_coeff,
                                        uint64_t *planes_f, uint64_t *cms);

/* ---- LFFoldingProver::compute_f_0 + Witness::from_f      LF/nifs/folding.rs:110,121,258-268; LF/arith.rs:299-313
 * (SURVEY 8 f1).  A fold step decomposes two witnesses (accumulator side, step side) into K planes each; the
 * folded witness is f_0[j] = sum_{i < 2K} rho_i * f_i[j] over those 2K planes in CRT form, followed by
 * f_0_coeff = iCRT(f_0).  The engine keeps the planes of the last decomposition of EACH side resident:
 * lat_ajtai_select_side chooses which side (0 = first / accumulator, 1 = second / step witness) the following
 * lat_ajtai_decompose_commit* calls fill.  rho: 2K ring elements in CRT form (rho_0..rho_{K-1} for side 0, then
 * side 1).  Outputs f0 (n x 24, CRT form) and f0_coeff (n x 24, coefficient form); either may be NULL.
 * Fails with LAT_E_INVALID_ARGUMENT until both sides have been decomposed.                                   */
int lat_ajtai_select_side(lat_ajtai *h, int side);
int lat_ajtai_fold_witness(lat_ajtai *h, const uint64_t *rho, uint64_t *f0, uint64_t *f0_coeff);
int lat_ajtai_fold_witness_dev(lat_ajtai *h, const uint64_t *rho_dev, uint64_t *f0_dev, uint64_t *f0_coeff_dev);

/* ---- the fold step of one IVC step as two blocking calls            ZKVM/zk_latticefold.rs:37-102, ZKVM/main.rs:174-182
 * The engine keeps the running accumulator witness w_acc resident (its coefficients as int16 digits) together with its
 * commitment, so that per step only w_ccs goes up and commitments + int16 digits come down:
 *
 *   lat_ajtai_set_accumulator : upload the initial accumulator witness (ZKVM/main.rs:306-344: the zero witness),
 *       f_coeff: n x 24 coefficient form with |c| < 2^K else LAT_E_DIGIT_OVERFLOW; cm_acc (kappa x 24) may be NULL if
 *       it is passed to the first begin instead.
 *   lat_ajtai_fold_step_begin : Witness::from_w_ccs + Witness::commit of the step witness (ZKVM/main.rs:348-367), then
 *       decompose_witness + commit_witnesses (LF/nifs/decomposition.rs:162-201) of the accumulator (side 0, against
 *       cm_acc; NULL = the folded commitment left resident by the previous finish) and of the step witness (side 1,
 *       against the commitment just computed) -- all chained on the device, one synchronisation.  Outputs: cm
 *       (kappa x 24), cms (2 x K x kappa x 24: side 0 then side 1, each [y_0 .. y_{K-1}]) and, optionally, the step
 *       witness's f_coeff as int16 digits (n x 24).  The 2K planes stay resident.
 *   -- the host runs its linearization / decomposition / folding sumchecks and derives the challenges rho --
 *   lat_ajtai_fold_step_finish: f_0 = sum_i rho_i * f_i over the 2K planes (LF/nifs/folding.rs:258-268), Witness::from_f's
 *       iCRT (LF/arith.rs:299-313) and cm_0 = sum_i rho_i * cm_i (LF/nifs/folding/utils.rs:466-472).  rho: 2K x 24 CRT
 *       form.  Outputs (each may be NULL): f0_coeff16 (n x 24 int16), f0 (n x 24 CRT form), cm0 (kappa x 24).  f_0 becomes
 *       the resident accumulator of the next step; if one of its coefficients reaches 2^K (the protocol's norm bound,
 *       checked nowhere in the reference, which would panic in the next decomposition) -> LAT_E_DIGIT_OVERFLOW.   */
int lat_ajtai_set_accumulator(lat_ajtai *h, const uint64_t *f_coeff, uint64_t n, const uint64_t *cm_acc);
int lat_ajtai_fold_step_begin(lat_ajtai *h, const uint64_t *w_ccs, uint64_t w_len, const uint64_t *cm_acc,
                              int16_t *f_coeff16, uint64_t *cm, uint64_t *cms);
int lat_ajtai_fold_step_finish(lat_ajtai *h, const uint64_t *rho, int16_t *f0_coeff16, uint64_t *f0, uint64_t *cm0,
                               uint64_t *w_ccs0);
/* (w_ccs0, may be NULL: gadget_recompose(f_0), n / L elements in CRT form -- the w_ccs of Witness::from_f, LF/arith.rs:305.)
 *
 * Witness::get_fhat (LF/arith.rs:273-297) of a resident witness on the device, for MLE code that runs there:
 * fhat_dev: tau = 3 tables x n x 24; table j, element i, slot s = (coefficient 8 j + s of f_coeff[i], 0, 0).
 * which: 0 = the current witness (last from_w_ccs / begin), 1 = the accumulator.  Hosts re-lay out the int16 digits
 * themselves (12x less PCIe than fetching the tables).                                                          */
int lat_ajtai_get_fhat_dev(lat_ajtai *h, int which, uint64_t *fhat_dev);

/* GadgetRecompose for &[R] in CRT form: out[i] = sum_l B^l * f[i*L + l]   (Witness::from_f / from_f_coeff rebuild
 * w_ccs this way, LF/arith.rs:305,330; RING/balanced_decomposition/mod.rs:105-117,177-190).
 * f: count*L elements, out: count elements.                                                                  */
int lat_ring_gadget_recompose(const uint64_t *f, uint64_t count, uint32_t log2_b, uint32_t L, uint64_t *out, int repr,
                              int device);

/* ---- standalone batched ring transforms (no handle; run on `device`, synchronous) ----------------------------
 * CRT::elementwise_crt / ICRT::elementwise_icrt      RING/cyclotomic_ring/crt.rs:10-49 -> GOLD/ntt.rs:135-319
 * `count` ring elements; in-place allowed (in == out).  Both maps are linear, so `repr` does not matter.   */
int lat_ring_crt(const uint64_t *coeff, uint64_t count, uint64_t *ntt, int device);
int lat_ring_icrt(const uint64_t *ntt, uint64_t count, uint64_t *coeff, int device);
int lat_ring_crt_dev(const uint64_t *coeff_dev, uint64_t count, uint64_t *ntt_dev, void *cuda_stream);
int lat_ring_icrt_dev(const uint64_t *ntt_dev, uint64_t count, uint64_t *coeff_dev, void *cuda_stream);
/* GadgetDecompose for &[R]: out[i*L + l] = limb l (base 2^log2_b) of in[i], coefficient form.
 * RING/balanced_decomposition/mod.rs:163-175, coeff_form.rs:588-606.  out: count*L x 24.                   */
int lat_ring_gadget_decompose(const uint64_t *in, uint64_t count, uint32_t log2_b, uint32_t L, uint64_t *out,
                              int repr, int device);

/* ---- column-sharded commitments (multi-GPU) --------------------------------------------------------
 * y = sum_j A_j f_j is a sum over columns: each GPU owns a contiguous column block (its own lat_ajtai handle of
 * width n_g, uploaded with row_stride = full width) and produces a full kappa x 24 partial commitment; the
 * partials are exchanged (NCCL all-gather over NVLink, done by the host layer) and summed mod q here.
 * parts: count x words u64 (words = batch * kappa * 24), out: words u64.  Addition is representation-agnostic.   */
int lat_commitment_sum_dev(const uint64_t *parts_dev, uint32_t count, uint64_t words, uint64_t *out_dev,
                           void *cuda_stream);
int lat_commitment_sum(const uint64_t *parts, uint32_t count, uint64_t words, uint64_t *out, int device);

/* Fused exchange + fold over NVLink peer memory (one kernel per rank, no NCCL call on the data path).
 * Every rank owns a receive buffer of 2 x world x words u64 and a flag array of 2 x world u64 (zero-initialised),
 * both mapped into every peer (e.g. torch symmetric memory / cudaIpc).  recv_ptrs[r] and flag_ptrs[r] are the
 * addresses, valid on THIS GPU, of rank r's buffer and flags (r = rank: the local ones).  The kernel stores this
 * rank's partial into slot (epoch & 1) of every rank's buffer, fences, raises flag[rank] = epoch on every rank,
 * waits until all `world` local flags show `epoch`, and folds the received partials mod q into out_dev.
 * `epoch` must start at 1 and grow by 1 per call on every rank; two slots suffice because a rank can only reach
 * call e+2 after every peer finished reading call e (it needs their call e+1 data first).                      */
int lat_commitment_exchange_dev(const uint64_t *partial_dev, uint64_t words, int rank, int world,
                                const uint64_t *recv_ptrs, const uint64_t *flag_ptrs, uint64_t epoch,
                                uint64_t *out_dev, void *cuda_stream);

/* Column-sharded form of lat_ajtai_submit_w_ccs: once peers are set, every submitted step ends with the fused exchange
 * (as lat_commitment_exchange_dev, with the given mailbox and flag addresses of all ranks) and lat_ajtai_wait returns
 * the FULL commitment, summed over the ranks.  next_epoch is the epoch of the next exchange (the engine counts on from
 * there; pass it again whenever other exchanges used the same mailboxes in between).  world <= 1 or NULL addresses
 * switch back to the single-GPU behaviour.  Every rank must submit the same sequence of steps.                  */
int lat_ajtai_set_peers(lat_ajtai *h, int rank, int world, const uint64_t *recv_ptrs, const uint64_t *flag_ptrs,
                        uint64_t next_epoch);

/* ---- building blocks of a host-buffer pipeline that enqueues nothing but kernels (sharded callers; the single-GPU
 * form is lat_ajtai_submit_w_ccs).  Both keep event waits and copies out of the compute stream, so the kernels of
 * consecutive steps keep overlapping (lat_ajtai_set_step_overlap):
 *  - the gated witness call starts iCRT/decompose/CRT + A * f as usual, but its kernel first polls *ready_flag_dev
 *    until it equals ready_value -- the caller uploads w_ccs_dev on another stream and then copies ready_value there.
 *    Enqueue those copies BEFORE this call: a kernel must never wait for work submitted after it (streams can share
 *    a hardware queue, and the copy would then sit behind the waiting kernel);
 *  - the reporting exchange additionally stores the folded commitment to cm_host and then publishes done_value in
 *    *done_host (both page-locked host memory, e.g. lat_host_alloc), which the host polls instead of synchronising. */
int lat_ajtai_witness_from_w_ccs_gated_dev(lat_ajtai *h, const uint64_t *w_ccs_dev, uint64_t w_len, uint64_t *cm_dev,
                                           const uint64_t *ready_flag_dev, uint64_t ready_value);
int lat_commitment_exchange_report_dev(const uint64_t *partial_dev, uint64_t words, int rank, int world,
                                       const uint64_t *recv_ptrs, const uint64_t *flag_ptrs, uint64_t epoch,
                                       uint64_t *out_dev, uint64_t *cm_host, uint64_t *done_host, uint64_t done_value,
                                       void *cuda_stream);

/* ---- standalone negacyclic NTT over Z_q[X]/(X^d + 1), d = 2^log2_d (BASELINE configs[3]) -------------------
 * NOT part of the drop-in path and ABSENT from the reference, whose ring is Z_q[X]/(X^24 - X^12 + 1): nothing in the
 * reference fixes these values, so this header is the specification and parity is pinned only against the O(d^2)
 * restatement in oracle/ and the transform's algebraic properties.  With psi = 7^((q - 1) / 2d):
 *   forward (inverse = 0): out[i] = sum_j in[j] * psi^((2 i + 1) j)        natural order in and out
 *   inverse (inverse = 1): out[j] = d^-1 * sum_i in[i] * psi^(-(2 i + 1) j)
 * hence forward(a * b mod X^d + 1) = forward(a) (.) forward(b).  `batch` polynomials of d coefficients each, contiguous;
 * in-place allowed.  Linear with canonical constants: Montgomery-form input gives Montgomery-form output.       */
#define LAT_NTT_MAX_LOG2_D 14
int lat_ntt_negacyclic(const uint64_t *in, uint64_t batch, uint32_t log2_d, int inverse, uint64_t *out, int device);
int lat_ntt_negacyclic_dev(const uint64_t *in_dev, uint64_t batch, uint32_t log2_d, int inverse, uint64_t *out_dev,
                           void *cuda_stream);

/* ---- overlap of consecutive steps on the device ---------------------------------------------------------------------
 * With the option enabled, the iCRT/decompose/CRT kernel of a lat_ajtai_witness_from_w_ccs_dev call that directly
 * follows a commitment on the same handle starts while that commitment's matrix-vector kernel is still draining
 * (programmatic dependent launch; the witness goes to a second buffer, and the kernels order themselves so that every
 * result is still produced as if the calls ran back to back).  CONTRACT for a caller-supplied stream: between the
 * previous call on this handle and this one, nothing else enqueued on the stream produces w_ccs_dev or consumes
 * the previous call's outputs -- such work would no longer be ordered before the early-starting kernel.  Off by
 * default; host-buffer entry points are unaffected (they synchronise).                                          */
int lat_ajtai_set_step_overlap(lat_ajtai *h, int enabled);

/* ---- diagnostics: CUDA-event timing of the dominant kernel (bench.py's roofline leg) ---------------------------
 * While enabled, every mac_kernel launch (the matrix-vector kernel alone, not its tiny reduce) is bracketed by a
 * pair of events on the handle's stream, taken from a pool so that nothing synchronises inside a timed loop.
 * lat_ajtai_mac_profile waits for outstanding brackets and returns the summed duration and the launch count
 * since profiling was enabled (or since the last call), then resets both.                                     */
int lat_ajtai_set_profiling(lat_ajtai *h, int enabled);
int lat_ajtai_mac_profile(lat_ajtai *h, double *sum_ms, uint64_t *launches);

/* ---- bounded device-side waits ----------------------------------------------------------------------------------------
 * Two kernels wait inside the GPU for something that another engine or GPU delivers: the witness kernel of a
 * submitted step polls its upload ticket, the exchange kernel polls its peers' flags.  Both waits are bounded by
 * %globaltimer (default 5000 ms; LAT_SPIN_TIMEOUT_MS in the environment or lat_set_spin_timeout_ms; 0 = unbounded):
 * on expiry the kernel records what it was waiting for in a per-device status word and carries on, and the next
 * lat_ajtai_wait / lat_ajtai_synchronize / host-buffer call on that device -- or lat_device_wait_status for callers
 * of the handle-less exchange -- returns LAT_E_CUDA with a message naming the wait (lat_last_error) and clears the
 * word.  *code (may be NULL) receives the raw word: low byte 1 = upload ticket, 2 = peer flag; 0 = none.          */
int lat_set_spin_timeout_ms(uint64_t ms);
int lat_device_wait_status(int device, uint64_t *code);

/* ---- pinned host memory for callers that want the fast copy path (optional) -------------------------------- */
int lat_host_alloc(void **ptr, size_t bytes);
void lat_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* LATTICE_AJTAI_H */

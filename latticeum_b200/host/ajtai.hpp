// Header-only C++ mirror of the reference's commitment API over the C ABI (include/lattice_ajtai.h), for C++ hosts.
// Names follow latticefold/src/commitment/commitment_scheme.rs:38-140 and homomorphic_commitment.rs:12-80.
// Ring elements are 24 contiguous uint64_t (CRT form: slot*3 + component).  No CPU fallback: construction throws
// without a CUDA device.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lattice_ajtai.h"

namespace lat {

using RingElem = uint64_t[LAT_RING_DEGREE];

// latticefold/src/commitment.rs:13-26
struct CommitmentError : std::runtime_error {
    int status;
    size_t got, expected;
    CommitmentError(int st, size_t g, size_t e, const std::string &msg) : std::runtime_error(msg), status(st), got(g), expected(e) {}
};
inline void check(int st, size_t got = 0, size_t expected = 0) {
    if (st == LAT_OK) return;
    if (st == LAT_E_WRONG_WITNESS_LENGTH)
        throw CommitmentError(st, got, expected,
                              "Wrong length of the witness: " + std::to_string(got) + ", expected: " + std::to_string(expected));
    throw CommitmentError(st, got, expected, std::string(lat_strerror(st)) + ": " + lat_last_error());
}

// Commitment<R>: kappa ring elements (homomorphic_commitment.rs:12-14)
struct Commitment {
    std::vector<uint64_t> val;  // kappa * 24
    size_t len() const { return val.size() / LAT_RING_DEGREE; }
    bool operator==(const Commitment &o) const { return val == o.val; }
};

struct DecompositionParams {  // decomposition_parameters.rs:11-20; zkvm/src/ccs.rs:26-34
    uint32_t log2_B = 15, L = 5, K = 15;
};

class AjtaiCommitmentScheme {
   public:
    // AjtaiCommitmentScheme::new: rows[i] points at n ring elements (the host matrix is a vector of rows)
    AjtaiCommitmentScheme(const std::vector<const uint64_t *> &rows, uint64_t n, DecompositionParams p = {},
                          lat_repr repr = LAT_REPR_CANONICAL, int device = 0)
        : kappa_((uint32_t)rows.size()), n_(n), p_(p) {
        check(lat_ajtai_create(&h_, kappa_, n, p.log2_B, p.L, p.K, repr, device));
        for (uint32_t i = 0; i < kappa_; ++i) check(lat_ajtai_upload_rows(h_, i, 1, rows[i], n));
    }
    ~AjtaiCommitmentScheme() { lat_ajtai_destroy(h_); }
    AjtaiCommitmentScheme(const AjtaiCommitmentScheme &) = delete;
    AjtaiCommitmentScheme &operator=(const AjtaiCommitmentScheme &) = delete;

    size_t kappa() const { return kappa_; }  // :85
    size_t width() const { return n_; }      // :92

    // commit / commit_ntt (:63-80, :101-103)
    Commitment commit_ntt(const uint64_t *f, size_t f_len) const {
        Commitment cm{std::vector<uint64_t>((size_t)kappa_ * LAT_RING_DEGREE)};
        check(lat_ajtai_commit_ntt(h_, f, f_len, cm.val.data()), f_len, n_);
        return cm;
    }
    Commitment commit(const uint64_t *f, size_t f_len) const { return commit_ntt(f, f_len); }
    // commit_coeff (:107-112)
    Commitment commit_coeff(const uint64_t *f_coeff, size_t f_len) const {
        Commitment cm{std::vector<uint64_t>((size_t)kappa_ * LAT_RING_DEGREE)};
        check(lat_ajtai_commit_coeff(h_, f_coeff, f_len, cm.val.data()), f_len, n_);
        return cm;
    }
    // decompose_and_commit_ntt (:132-139)
    Commitment decompose_and_commit_ntt(const uint64_t *w, size_t w_len) const {
        Commitment cm{std::vector<uint64_t>((size_t)kappa_ * LAT_RING_DEGREE)};
        check(lat_ajtai_decompose_and_commit_ntt(h_, w, w_len, cm.val.data()), w_len * p_.L, n_);
        return cm;
    }
    // Witness::from_w_ccs + Witness::commit (arith.rs:230-248, 357-362); f_coeff / f may be null
    Commitment witness_from_w_ccs(const uint64_t *w_ccs, size_t w_len, uint64_t *f_coeff, uint64_t *f) const {
        Commitment cm{std::vector<uint64_t>((size_t)kappa_ * LAT_RING_DEGREE)};
        check(lat_ajtai_witness_from_w_ccs(h_, w_ccs, w_len, f_coeff, f, cm.val.data()), w_len * p_.L, n_);
        return cm;
    }
    // decompose_witness + commit_witnesses (nifs/decomposition.rs:162-201): K commitments, y_0 by homomorphism
    std::vector<Commitment> decompose_commit(const uint64_t *f_coeff, size_t n, const Commitment &cm, uint64_t *planes_coeff = nullptr,
                                             uint64_t *planes_f = nullptr) const {
        std::vector<uint64_t> cms((size_t)p_.K * kappa_ * LAT_RING_DEGREE);
        check(lat_ajtai_decompose_commit(h_, f_coeff, n, cm.val.data(), planes_coeff, planes_f, cms.data()), n, n_);
        std::vector<Commitment> out(p_.K);
        for (uint32_t k = 0; k < p_.K; ++k)
            out[k].val.assign(cms.begin() + (size_t)k * kappa_ * LAT_RING_DEGREE, cms.begin() + (size_t)(k + 1) * kappa_ * LAT_RING_DEGREE);
        return out;
    }
    // Non-blocking Witness::from_w_ccs + commit for INDEPENDENT work (consecutive IVC steps are dependent,
    // zkvm/src/main.rs:140-182, and use the blocking calls): submit returns a ticket at once, wait blocks for that
    // commitment (written to `cm`, kappa x 24, which like w_ccs must stay valid until then).
    uint64_t submit_w_ccs(const uint64_t *w_ccs, size_t w_len, uint64_t *cm) const {
        uint64_t ticket = 0;
        check(lat_ajtai_submit_w_ccs(h_, w_ccs, w_len, cm, &ticket), w_len * p_.L, n_);
        return ticket;
    }
    void wait(uint64_t ticket) const { check(lat_ajtai_wait(h_, ticket)); }
    // Which side (0 = accumulator, 1 = step witness) the following decompose_commit calls fill (kept resident).
    void select_side(int side) const { check(lat_ajtai_select_side(h_, side)); }
    // LFFoldingProver::compute_f_0 (nifs/folding.rs:258-268) + Witness::from_f's iCRT (arith.rs:299-313) over
    // the 2K resident planes; rho: 2K ring elements (CRT form); f0 / f0_coeff: n x 24, either may be null.
    void fold_witness(const uint64_t *rho, uint64_t *f0, uint64_t *f0_coeff) const { check(lat_ajtai_fold_witness(h_, rho, f0, f0_coeff)); }
    // Witness::from_w_ccs + commit fetching f_coeff as the device's int16 digits (n x 24); see digits_to_fq / get_fhat below
    Commitment witness_from_w_ccs_compact(const uint64_t *w_ccs, size_t w_len, int16_t *f_coeff16, uint64_t *f = nullptr) const {
        Commitment cm{std::vector<uint64_t>((size_t)kappa_ * LAT_RING_DEGREE)};
        check(lat_ajtai_witness_from_w_ccs_compact(h_, w_ccs, w_len, f_coeff16, f, cm.val.data()), w_len * p_.L, n_);
        return cm;
    }
    // The fold step of one IVC step as two blocking calls (zkvm/src/zk_latticefold.rs:37-102), accumulator resident.
    void set_accumulator(const uint64_t *f_coeff, size_t n, const Commitment &cm_acc) const {
        check(lat_ajtai_set_accumulator(h_, f_coeff, n, cm_acc.val.data()), n, n_);
    }
    struct FoldBegin {
        Commitment cm;                               // of the step witness
        std::vector<Commitment> ys_acc, ys_step;     // K each: y_0 .. y_{K-1} of the accumulator / the step witness
    };
    FoldBegin fold_step_begin(const uint64_t *w_ccs, size_t w_len, const Commitment *cm_acc = nullptr, int16_t *f_coeff16 = nullptr) const {
        const size_t cw = (size_t)kappa_ * LAT_RING_DEGREE;
        FoldBegin r;
        r.cm.val.resize(cw);
        std::vector<uint64_t> cms(2 * p_.K * cw);
        check(lat_ajtai_fold_step_begin(h_, w_ccs, w_len, cm_acc ? cm_acc->val.data() : nullptr, f_coeff16, r.cm.val.data(), cms.data()),
              w_len * p_.L, n_);
        r.ys_acc.resize(p_.K);
        r.ys_step.resize(p_.K);
        for (uint32_t k = 0; k < p_.K; ++k) {
            r.ys_acc[k].val.assign(cms.begin() + k * cw, cms.begin() + (k + 1) * cw);
            r.ys_step[k].val.assign(cms.begin() + (p_.K + k) * cw, cms.begin() + (p_.K + k + 1) * cw);
        }
        return r;
    }
    // returns cm_0; f0_coeff16 (n x 24), f0 (n x 24) and w_ccs0 (n / L x 24) may be null
    Commitment fold_step_finish(const uint64_t *rho, int16_t *f0_coeff16, uint64_t *f0 = nullptr, uint64_t *w_ccs0 = nullptr) const {
        Commitment cm0{std::vector<uint64_t>((size_t)kappa_ * LAT_RING_DEGREE)};
        check(lat_ajtai_fold_step_finish(h_, rho, f0_coeff16, f0, cm0.val.data(), w_ccs0));
        return cm0;
    }
    lat_ajtai *handle() const { return h_; }

   private:
    lat_ajtai *h_ = nullptr;
    uint32_t kappa_;
    uint64_t n_;
    DecompositionParams p_;
};

// GadgetRecompose for a CRT-form vector (arith.rs:305,330): out[i] = sum_l B^l f[i*L + l]
inline void gadget_recompose(const uint64_t *f, size_t count, DecompositionParams p, uint64_t *out, lat_repr repr = LAT_REPR_CANONICAL,
                             int device = 0) {
    check(lat_ring_gadget_recompose(f, count, p.log2_B, p.L, out, repr, device));
}

// The engine's int16 digit -> Fq limb in the caller's representation (negative digits are q - |d|; Montgomery: x 2^64).
inline uint64_t digit_to_fq(int16_t d, lat_repr repr = LAT_REPR_CANONICAL) {
    const uint64_t q = 0xFFFFFFFF00000001ull;
    uint64_t m = (uint64_t)(d < 0 ? -(int)d : (int)d);
    if (repr == LAT_REPR_MONTGOMERY) m = (m << 32) - m;  // m (2^32 - 1) = m 2^64 mod q, exact for |d| < 2^31
    return (d < 0 && m) ? q - m : m;
}
// Witness::get_fhat (arith.rs:273-297) straight from the digits: fhat[(j * n + i) * 24 + 3 s] = digit 8 j + s of element i,
// the other two components of each slot zero; fhat: 3 x n x 24 (zero-initialised by the caller or here).
inline void get_fhat_from_digits(const int16_t *digits, size_t n, uint64_t *fhat, lat_repr repr = LAT_REPR_CANONICAL) {
    for (size_t j = 0; j < 3; ++j)
        for (size_t i = 0; i < n; ++i)
            for (size_t s = 0; s < 8; ++s) {
                uint64_t *o = fhat + ((j * n + i) * LAT_RING_DEGREE + 3 * s);
                o[0] = digit_to_fq(digits[i * LAT_RING_DEGREE + 8 * j + s], repr);
                o[1] = 0;
                o[2] = 0;
            }
}

}  // namespace lat

// The folding loop of the reference's zkVM on the C++ mirror: zkvm/src/main.rs:140-182 (one IVC step per trace, each
// depending on the previous accumulator) with the commitment work of zkvm/src/zk_latticefold.rs:37-102 as the two
// blocking calls fold_step_begin / fold_step_finish.  Synthetic witnesses; the checks are the scheme's own algebra:
//   cm = sum_k 2^k y_k on both sides (nifs/decomposition.rs:162-201), cm_0 = A f_0 (the commitment is homomorphic),
//   w_ccs of the folded witness = gadget_recompose(f_0) (arith.rs:305,330), and f_0 of step i is the accumulator of i+1.
// Needs a GPU to run.   g++ -std=c++17 example_ivc.cpp -L../lib -llattice_ajtai -Wl,-rpath,'$ORIGIN/../lib' -o example_ivc
#include <cstdio>

#include "ajtai.hpp"

namespace {
const uint64_t Q = 0xFFFFFFFF00000001ull;
uint64_t state = 0x9E3779B97F4A7C15ull;
uint64_t next_u64() {  // splitmix64
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t mulmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % Q); }
uint64_t addmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a + b) % Q); }

// sum_k 2^k y_k: a scalar multiplies every CRT component
lat::Commitment recombine(const std::vector<lat::Commitment> &ys) {
    lat::Commitment out{std::vector<uint64_t>(ys[0].val.size(), 0)};
    uint64_t s = 1;
    for (const auto &y : ys) {
        for (size_t i = 0; i < out.val.size(); ++i) out.val[i] = addmod(out.val[i], mulmod(s, y.val[i]));
        s = addmod(s, s);
    }
    return out;
}
}  // namespace

int main() {
    const uint32_t kappa = 4;
    const lat::DecompositionParams dp;
    const uint64_t wl = 64, n = wl * dp.L;
    std::vector<std::vector<uint64_t>> rows(kappa, std::vector<uint64_t>(n * 24));
    std::vector<const uint64_t *> ptrs;
    for (auto &r : rows) {
        for (auto &v : r) v = next_u64() % Q;
        ptrs.push_back(r.data());
    }
    try {
        lat::AjtaiCommitmentScheme scheme(ptrs, n), checker(ptrs, n);
        std::vector<uint64_t> w(wl * 24), f_coeff(n * 24), f0(n * 24), w0(wl * 24), w0_check(wl * 24);
        std::vector<int16_t> d16(n * 24), f0_16(n * 24);
        // initialize_accumulator (main.rs:348-367): the first witness, committed, becomes the running instance
        for (auto &v : w) v = next_u64() % Q;
        lat::Commitment cm_acc = scheme.witness_from_w_ccs(w.data(), wl, f_coeff.data(), nullptr);
        scheme.set_accumulator(f_coeff.data(), n, cm_acc);
        for (int step = 0; step < 3; ++step) {
            for (auto &v : w) v = next_u64() % Q;  // this step's CCS witness
            auto b = scheme.fold_step_begin(w.data(), wl, nullptr, d16.data());
            if (!(recombine(b.ys_step) == b.cm) || !(recombine(b.ys_acc) == cm_acc)) {
                std::printf("MISMATCH: cm != sum 2^k y_k at step %d\n", step);
                return 1;
            }
            if (!(b.cm == checker.witness_from_w_ccs(w.data(), wl, nullptr, nullptr))) return 1;
            // the verifier's challenges (transcript, host side): 2K short ring elements, CRT form
            std::vector<uint64_t> rho_coeff(2 * dp.K * 24), rho(2 * dp.K * 24);
            for (auto &v : rho_coeff) {
                const int c = (int)(next_u64() % 64) - 32;
                v = c < 0 ? Q - (uint64_t)(-c) : (uint64_t)c;
            }
            lat::check(lat_ring_crt(rho_coeff.data(), 2 * dp.K, rho.data(), 0));
            lat::Commitment cm0 = scheme.fold_step_finish(rho.data(), f0_16.data(), f0.data(), w0.data());
            if (!(cm0 == checker.commit_ntt(f0.data(), n))) {
                std::printf("MISMATCH: cm_0 != A f_0 at step %d\n", step);
                return 1;
            }
            lat::gadget_recompose(f0.data(), wl, dp, w0_check.data());
            if (w0 != w0_check) return 1;
            cm_acc = cm0;  // the folded instance is the next step's accumulator (main.rs:174-182)
        }
        std::printf("ivc example ok: 3 dependent fold steps, cm = sum 2^k y_k, cm_0 = A f_0, w_ccs(f_0) = recompose(f_0)\n");
        return 0;
    } catch (const lat::CommitmentError &e) {
        std::printf("engine error: %s\n", e.what());
        return 2;
    }
}

// Minimal C++ host using the mirror API: commits the reference's closed-form test matrix
// (latticefold/src/commitment/commitment_scheme.rs:150-185) and checks the result.  Needs a GPU to run.
// g++ -std=c++17 example.cpp -L../lib -llattice_ajtai -Wl,-rpath,'$ORIGIN/../lib' -o example
#include <cstdio>

#include "ajtai.hpp"

int main() {
    const uint32_t kappa = 9;
    const uint64_t n = 1 << 15;
    std::vector<std::vector<uint64_t>> rows(kappa, std::vector<uint64_t>(n * 24, 0));
    std::vector<const uint64_t *> ptrs;
    for (uint32_t i = 0; i < kappa; ++i) {
        for (uint64_t j = 0; j < n; ++j)
            for (int s = 0; s < 8; ++s) rows[i][(j * 24) + 3 * s] = i * n + j;  // scalar in every slot
        ptrs.push_back(rows[i].data());
    }
    try {
        lat::AjtaiCommitmentScheme scheme(ptrs, n);
        std::vector<uint64_t> w(n * 24, 0);
        for (uint64_t j = 0; j < n; ++j)
            for (int s = 0; s < 8; ++s) w[j * 24 + 3 * s] = 2;
        lat::Commitment cm = scheme.commit_ntt(w.data(), n);
        for (uint32_t i = 0; i < kappa; ++i) {
            uint64_t expected = n * (2 * i * n + (n - 1));
            for (int s = 0; s < 8; ++s)
                if (cm.val[i * 24 + 3 * s] != expected || cm.val[i * 24 + 3 * s + 1] || cm.val[i * 24 + 3 * s + 2]) {
                    std::printf("MISMATCH row %u\n", i);
                    return 1;
                }
        }
        try {
            scheme.commit_ntt(w.data(), n - 1);
            return 1;
        } catch (const lat::CommitmentError &e) {
            if (e.status != LAT_E_WRONG_WITNESS_LENGTH || e.got != n - 1 || e.expected != n) return 1;
        }
        std::printf("example ok: closed-form commitment matches, WrongWitnessLength raised\n");
        return 0;
    } catch (const lat::CommitmentError &e) {
        std::printf("engine error: %s\n", e.what());
        return 2;
    }
}

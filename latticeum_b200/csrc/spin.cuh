// Bounded device-side waits (see lat::SpinGuard in kernels.h).
#pragma once
#include "kernels.h"

namespace lat {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until *word == want (acquire, system scope: the word is written by a copy engine or by another GPU).
// Returns false if the guard's deadline passed first; the code goes to the host's status word.
__device__ __forceinline__ bool spin_until_equals(const unsigned long long *word, unsigned long long want, const SpinGuard &g,
                                                  unsigned long long code, unsigned long long detail) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(word) : "memory");
    if (v == want) return true;
    const unsigned long long t0 = global_timer_ns();
    for (unsigned spins = 0;; ++spins) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(word) : "memory");
        if (v == want) return true;
        if ((spins & 63) == 63 && g.timeout_ns && global_timer_ns() - t0 > g.timeout_ns) {
            if (g.status) {
                // plain store: the word lives in host memory (PCIe atomics are not a given); any one report is enough
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(g.status), "l"(code | (detail << 8)) : "memory");
                __threadfence_system();
            }
            return false;
        }
    }
}

}  // namespace lat

// Host build of latticeum_b200/csrc/ring96.cuh (plain C++: the carry chains use their portable fallbacks) so that the
// butterfly network in Z/(2^96 + 1) can be checked against the oracle without a GPU (tests/test_ring96_cpu.py).
#include "../../latticeum_b200/csrc/ring96.cuh"

extern "C" {
void r96_crt(unsigned long long *x, unsigned long long count) {
    for (unsigned long long e = 0; e < count; ++e) r96::crt24(*reinterpret_cast<r96::u64(*)[24]>(x + e * 24));
}
void r96_icrt(unsigned long long *x, unsigned long long count) {
    for (unsigned long long e = 0; e < count; ++e) r96::icrt24(*reinterpret_cast<r96::u64(*)[24]>(x + e * 24));
}
void r96_crt_small(const int *d, unsigned long long count, int mont, unsigned long long *out) {
    for (unsigned long long e = 0; e < count; ++e) {
        const int(&de)[24] = *reinterpret_cast<const int(*)[24]>(d + e * 24);
        r96::u64(&oe)[24] = *reinterpret_cast<r96::u64(*)[24]>(out + e * 24);
        if (mont) r96::crt24_small<true>(de, oe);
        else r96::crt24_small<false>(de, oe);
    }
}
}

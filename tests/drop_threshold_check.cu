// Host-side check of the integer keep threshold of sat::DropGen (sat_linear.cuh) against the float32 formula
// floor(keep + u), u = k * 2^-24, for EVERY 24-bit count k.  Built and run by tests/test_host_logic.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../show-attend-and-tell_b200/csrc/sat_linear.cuh"

int main(int argc, char** argv) {
    long bad = 0;
    for (int a = 1; a < argc; ++a) {
        const float keep = (float)atof(argv[a]);
        const sat::DropGen g = sat::drop_gen(1, 2, keep);
        long mism = 0;
        for (uint32_t k = 0; k < (1u << 24); ++k) {
            const float u = (float)k * 5.9604644775390625e-08f;
            const bool kept_float = floorf(keep + u) >= 1.0f;   // (== 1 except keep == 1 with the largest u, where the sum rounds to 2)
            mism += kept_float != (k >= g.kt);
        }
        printf("keep %.9g threshold %u mismatches %ld\n", keep, g.kt, mism);
        bad += mism;
    }
    // the generator itself, for the numpy copy to be compared with
    printf("u %.9g %.9g %.9g\n", sat::rng_u24(1234, 5, 0), sat::rng_u24(1234, 5, 1), sat::rng_u24(77, 16 * 19 + 7, (1ull << 33) + 5));
    return bad ? 1 : 0;
}

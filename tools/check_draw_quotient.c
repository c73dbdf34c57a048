// Exhaustive check of the dither generator's draw / RAND_MAX (dsf2flac_b200/csrc/dsd_rand.cuh, rand_draw_quotient):
// for every one of the 2^31 values rand() can return, fma(x, 2^-31 + 2^-62, x) with x = a * 2^-31 equals the correctly
// rounded quotient a / 2147483647.0 the reference computes (reference src/dsd_decimator.cpp:203-204).
//   gcc -O2 -ffp-contract=off -o check_draw_quotient tools/check_draw_quotient.c -lm && ./check_draw_quotient   (about 30 s)
// Last run: "one-FMA mismatches: 0" (a plain multiplication by the reciprocal is wrong for 9 437 184 of the values).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void) {
    const double M = 2147483647.0;
    // c = 2^-31 + 2^-62 exactly representable
    const double c = ldexp(1.0, -31) + ldexp(1.0, -62);
    uint64_t bad1 = 0, bad2 = 0;
    for (uint64_t a = 0; a < (1ull << 31); ++a) {
        const double want = (double)a / M;
        const double x = ldexp((double)a, -31);          // exact
        const double q1 = fma(x, c, x);                  // one FMA
        if (q1 != want) { if (bad1 < 5) printf("1fma a=%llu q=%.17g want=%.17g\n", (unsigned long long)a, q1, want); ++bad1; }
        // variant: q = fma(a, rcpM_lo?, ...) -- also test plain a * (1/M) for reference
        const double q2 = (double)a * (1.0 / M);
        if (q2 != want) ++bad2;
    }
    printf("one-FMA mismatches: %llu ; plain mul by reciprocal mismatches: %llu\n", (unsigned long long)bad1, (unsigned long long)bad2);
    return 0;
}

/* Exhaustive check behind uw_div162() (gr-uwspr_b200/csrc/fine.cu): for EVERY fp32 value x (all 2^32 bit patterns),
 *     q = RN(a * y),  r = RN(a - 162 q) (exact, one fma),  q' = RN(q + r * y),   a = (double)x, y = RN(1/162)
 * equals the correctly rounded double quotient a / 162.0 that the reference computes (sync_and_demodulate_impl.cc
 * :243-244 divide a float promoted to double by 162.0).  Non-finite inputs are excluded: the kernel sends them through
 * the IEEE division.  The sign of a zero result may differ (-0/162 = -0, the sequence gives +0): the quotient is only
 * ever added to an accumulator that starts at +0, where +-0 are indistinguishable; the count is reported.
 *   gcc -O2 -march=native -fopenmp tools/div162_exhaustive.c -o /tmp/div162 -lm && /tmp/div162
 * prints the number of mismatches (0); seconds with a hardware fma. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    /* optional argument: visit every stride-th bit pattern only (machines without a hardware fma) */
    const long long stride = argc > 1 ? atoll(argv[1]) : 1;
    const double y = 1.0 / 162.0;
    long long bad = 0, zero_sign = 0, checked = 0;
#pragma omp parallel for reduction(+ : bad, zero_sign, checked) schedule(static)
    for (long long b = 0; b < (1LL << 32); b += stride) {
        const uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        if (!isfinite(x)) continue;
        const double a = (double)x;
        const volatile double want = a / 162.0;
        const double q = a * y;
        const double r = fma(-162.0, q, a);
        const double got = fma(r, y, q);
        uint64_t w, g;
        double wv = want;
        memcpy(&w, &wv, 8);
        memcpy(&g, &got, 8);
        checked++;
        if (w != g) {
            if (wv == 0.0 && got == 0.0) zero_sign++;
            else bad++;
        }
    }
    printf("checked %lld finite fp32 values: %lld mismatches, %lld zero-sign differences\n", checked, bad, zero_sign);
    return bad != 0;
}

// Exhaustive check that a constant division can be replaced by the 3-operation FMA sequence (used by the
// discriminator layer-0 input normalisation): gcc -O2 -mfma -fopenmp -ffp-contract=off verify_const_div.c -lm && ./a.out 127.5
// prints the largest |a| for which fma(fma(-a*rc, c, a), rc, a*rc) != a / c  (127.5: 2^-119, denormal quotients only).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    float c = strtof(argv[1], 0);
    float rc = 1.0f / c;
    float maxbad = 0;
    #pragma omp parallel for reduction(max:maxbad)
    for (int64_t u = 0; u < (1LL << 32); u++) {
        uint32_t b = (uint32_t)u; float a; memcpy(&a, &b, 4);
        if (!isfinite(a)) continue;
        float q0 = a * rc;
        float r = fmaf(-q0, c, a);
        float q = fmaf(r, rc, q0);
        float t = a / c;
        if (memcmp(&q, &t, 4) != 0) { if (fabsf(a) > maxbad) maxbad = fabsf(a); }
    }
    printf("c=%.9g largest |a| with a mismatch = %g (2^%d)\n", c, maxbad, (int)floor(log2(maxbad)));
    return 0;
}

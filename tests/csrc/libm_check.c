/* Compares csrc/ogn_libm.h (the replay the CUDA kernels use) with this box's libm on every `stride`-th float bit
 * pattern. Prints "<expf mismatches> <tanhf mismatches> <count>". stride 1 = all 2^32 inputs (about 12 core-seconds). */
#include "../../ogmaneo2_b200/csrc/ogn_libm.h"
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    long long stride = argc > 1 ? atoll(argv[1]) : 16, off = argc > 2 ? atoll(argv[2]) : 0;
    unsigned long long be = 0, bt = 0, n = 0;
#pragma omp parallel for reduction(+ : be, bt, n) schedule(static)
    for (long long i = off; i < (1LL << 32); i += stride) {
        float x = ogn_u2f((uint32_t)i);
        float a = expf(x), b = ogn_expf(x), c = tanhf(x), d = ogn_tanhf(x);
        if (ogn_f2u(a) != ogn_f2u(b) && !(a != a && b != b)) be++;
        if (ogn_f2u(c) != ogn_f2u(d) && !(c != c && d != d)) bt++;
        n++;
    }
    printf("%llu %llu %llu\n", be, bt, n);
    return 0;
}

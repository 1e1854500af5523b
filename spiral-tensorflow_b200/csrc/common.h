// common.h -- shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

int spiral_set_error(int code, const char *fmt, ...);
int spiral_cuda_error(cudaError_t e, const char *what);

#define SPIRAL_REQUIRE(p)                                                                          \
    do {                                                                                           \
        if (!(p)) return spiral_set_error(-1, "%s: null or invalid argument `%s`", __func__, #p);  \
    } while (0)

// common.h -- shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

int spiral_set_error(int code, const char *fmt, ...);
int spiral_cuda_error(cudaError_t e, const char *what);

#define SPIRAL_REQUIRE(p)                                                                          \
    do {                                                                                           \
        if (!(p)) return spiral_set_error(-1, "%s: null or invalid argument `%s`", __func__, #p);  \
    } while (0)

// Checked build (csrc/build.sh also produces libspiral_b200_checked.so with -DSPIRAL_BOUNDS_CHECK): compute-sanitizer is
// closed on the B200 pool, so the kernels carry their own index checks -- every shared-memory / tile / list index that
// is computed from data is tested against its bound and violations are counted in a device-side counter that
// spiral_debug_bounds_violations() reads (tests/test_gpu_raster.py runs the parity cases through the checked library and
// requires zero).  In the product build the macro expands to nothing.
#if defined(SPIRAL_BOUNDS_CHECK) && defined(__CUDACC__)
static __device__ unsigned long long g_spiral_bounds_violations = 0ULL;      // one counter per translation unit (no RDC)
static inline long long spiral_tu_bounds_violations()
{
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_spiral_bounds_violations, sizeof(v)) != cudaSuccess) return -2;
    return (long long)v;
}
#define SP_CHECK(cond)                                                        \
    do {                                                                      \
        if (!(cond)) atomicAdd(&g_spiral_bounds_violations, 1ULL);            \
    } while (0)
#else
#define SP_CHECK(cond) do { } while (0)
#endif

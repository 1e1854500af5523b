#!/bin/bash
# Builds libspiral_b200.so for sm_100a in-tree (the .so is git-ignored but travels with gpurun).
#   --fmad=false : the rasteriser's float/double arithmetic must not be contracted (bit-exact
#                  against the oracle); fused operations are written as explicit fma().
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
cd "$HERE"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math $ARCH"
mkdir -p build
pids=()
compile() { # src extra-flags...
  local src=$1; shift
  if [ ! -f build/${src%.cu}.o ] || [ $src -nt build/${src%.cu}.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer build/${src%.cu}.o)" ] || [ ../../include/spiral_b200.h -nt build/${src%.cu}.o ]; then
    $NVCC $COMMON "$@" -c $src -o build/${src%.cu}.o &
    pids+=($!)
  fi
}
compile raster.cu --fmad=false ${RASTER_DEFS:-} ${PTXAS_V:+-Xptxas -v}
compile api_env.cu --fmad=false
compile a2c.cu ${A2C_DEFS:-} ${PTXAS_V:+-Xptxas -v}
for f in sample.cu optim.cu policy_conv.cu policy_learn.cu disc.cu disc_tc.cu disc_conv0_mma.cu conv_ops.cu conv_tc.cu; do [ -f $f ] && compile $f ${PTXAS_V:+-Xptxas -v}; done
# checked twins of the kernels whose indices are computed from data (common.h SP_CHECK): compute-sanitizer is closed on
# the B200 pool, so tests run the parity cases through libspiral_b200_checked.so and require zero violations
mkdir -p build/checked
cpids=()
for src in raster.cu policy_learn.cu api_env.cu; do
  obj=build/checked/${src%.cu}.o
  if [ ! -f $obj ] || [ $src -nt $obj ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer $obj)" ] || [ ../../include/spiral_b200.h -nt $obj ]; then
    extra=""; [ $src != policy_learn.cu ] && extra="--fmad=false"
    $NVCC $COMMON $extra -DSPIRAL_BOUNDS_CHECK -c $src -o $obj &
    cpids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
for p in "${cpids[@]:-}"; do [ -n "$p" ] && wait $p; done
OBJS=$(ls build/*.o)
OUT=${SPIRAL_LIB_NAME:-libspiral_b200.so}
$NVCC $ARCH -shared -o $OUT $OBJS -lcudart
COBJS=$(ls build/*.o | grep -v -E 'build/(raster|policy_learn|api_env)\.o') 
$NVCC $ARCH -shared -o libspiral_b200_checked.so $COBJS build/checked/raster.o build/checked/policy_learn.o build/checked/api_env.o -lcudart
echo "built $HERE/$OUT (+ libspiral_b200_checked.so)"

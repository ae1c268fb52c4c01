#!/bin/bash
# dev/build_variants.sh -- builds libxyz_b200.so variants that differ in compile-time knobs of the splat kernels
# (csrc/splat_kernels.cuh) into xyz-autodiff-cuda_b200/lib_variants/<name>/, for dev/variant_time.py.
set -e
cd "$(dirname "$0")/../xyz-autodiff-cuda_b200/csrc"
make -j8 >/dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
FAST="-use_fast_math -ftz=true -prec-div=false -prec-sqrt=false"
OTHERS=$(ls ../build/*.o | grep -v splat_fast.o)
build() {  # name, defines...
  name=$1; shift
  mkdir -p ../lib_variants/$name
  nvcc $NVFLAGS $FAST "$@" -c splat_fast.cu -o ../lib_variants/$name/splat_fast.o
  nvcc $ARCH -shared -o ../lib_variants/$name/libxyz_b200.so $OTHERS ../lib_variants/$name/splat_fast.o -ldl
  echo built $name
}
build base &
build bwd_min4 -DXYZ_BWD_MINBLOCKS=4 &
build bwd_min6 -DXYZ_BWD_MINBLOCKS=6 &
build bwd_unroll2 -DXYZ_BWD_ROW_UNROLL=2 &
wait
build bwd_nodx2_min6 -DXYZ_BWD_DX2=0 -DXYZ_BWD_MINBLOCKS=6 &
build bwd_unroll2_min4 -DXYZ_BWD_ROW_UNROLL=2 -DXYZ_BWD_MINBLOCKS=4 &
build fwd_unroll8 -DXYZ_FWD_UNROLL=8 &
build fwd_unroll2 -DXYZ_FWD_UNROLL=2 &
wait

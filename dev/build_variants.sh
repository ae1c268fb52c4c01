#!/bin/bash
# dev/build_variants.sh -- builds libxyz_b200.so variants that differ in compile-time knobs into
# xyz-autodiff-cuda_b200/lib_variants/<name>/, for dev/variant_time.py and the dev/*_time.py scripts (XYZ_B200_LIB selects
# the library).  TUS = the translation units the knobs live in (default: the three splat units, which share
# splat_common.cuh's knobs); every other object is taken from the regular build.
#   usage: [TUS="lsq_kernels"] dev/build_variants.sh name1 "-DX=1 -DY=2" name2 "-DZ=3" ...
set -e
cd "$(dirname "$0")/../xyz-autodiff-cuda_b200/csrc"
make -j8 >/dev/null
TUS=${TUS:-"splat_fast splat_precise splat_host"}
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
FAST="-use_fast_math -ftz=true -prec-div=false -prec-sqrt=false"
PATTERN=$(echo $TUS | sed 's/ /.o\\|/g').o
OTHERS=$(ls ../build/*.o | grep -v "$PATTERN")
build() {  # name, defines...
  name=$1; shift
  mkdir -p ../lib_variants/$name
  objs=""
  for tu in $TUS; do
    extra=""; [ "$tu" = "splat_fast" ] && extra="$FAST"
    nvcc $NVFLAGS $extra $@ -c $tu.cu -o ../lib_variants/$name/$tu.o 2>/dev/null
    objs="$objs ../lib_variants/$name/$tu.o"
  done
  nvcc $ARCH -shared -o ../lib_variants/$name/libxyz_b200.so $OTHERS $objs -ldl
  echo built $name
}
while [ $# -ge 2 ]; do build "$1" $2 & shift 2; done
wait

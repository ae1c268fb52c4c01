#!/bin/bash
# dev/build_variants.sh -- builds libxyz_b200.so variants that differ in compile-time knobs of the splat kernels
# (csrc/splat_kernels.cuh) into xyz-autodiff-cuda_b200/lib_variants/<name>/, for dev/variant_time.py.
#   usage: dev/build_variants.sh name1 "-DX=1 -DY=2" name2 "-DZ=3" ...
set -e
cd "$(dirname "$0")/../xyz-autodiff-cuda_b200/csrc"
make -j8 >/dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="$ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
FAST="-use_fast_math -ftz=true -prec-div=false -prec-sqrt=false"
OTHERS=$(ls ../build/*.o | grep -v "splat_fast.o\|splat_host.o\|splat_precise.o")
build() {  # name, defines...  (the three splat translation units share splat_common.cuh's knobs)
  name=$1; shift
  mkdir -p ../lib_variants/$name
  nvcc $NVFLAGS $FAST $@ -c splat_fast.cu -o ../lib_variants/$name/splat_fast.o
  nvcc $NVFLAGS $@ -c splat_precise.cu -o ../lib_variants/$name/splat_precise.o
  nvcc $NVFLAGS $@ -c splat_host.cu -o ../lib_variants/$name/splat_host.o 2>/dev/null
  nvcc $ARCH -shared -o ../lib_variants/$name/libxyz_b200.so $OTHERS ../lib_variants/$name/splat_fast.o \
       ../lib_variants/$name/splat_precise.o ../lib_variants/$name/splat_host.o -ldl
  echo built $name
}
while [ $# -ge 2 ]; do build "$1" $2 & shift 2; done
wait

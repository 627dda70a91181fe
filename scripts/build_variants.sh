#!/bin/bash
# A/B builds of libb200accel.so with experiment macros -> lucille_b200/variants/lib_<name>.so (git-ignored, travels with gpurun).
#   scripts/build_variants.sh name1:"-DFLAG1 -DFLAG2" name2:"" ...
set -e
cd "$(dirname "$0")/../lucille_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false --prec-div=true --prec-sqrt=true \
       -Xcompiler -fPIC,-ffp-contract=off,-O2,-Wno-unused-function $flags -shared accel.cu bvh_build.cpp -o ../variants/lib_$name.so -lpthread &
done
wait
ls -la ../variants

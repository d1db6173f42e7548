#!/bin/bash
# Build library variants with different compile-time knobs into ab/lib_<tag>.so (they travel with gpurun):
#   tools/ab_variants.sh tag1 "-DLSQ_WG_UMUL=1 -DLSQ_WG_MINB_NUM=1 -DLSQ_WG_MINB_DEN=1" tag2 "..." ...
set -e
cd "$(dirname "$0")/../lsqfakequantize-pytorch_b200/csrc"
mkdir -p ../../ab
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  rm -rf build_ab; make -j8 BUILD=build_ab OUT=../../ab/lib_$tag.so EXTRA_NVCCFLAGS="$flags" > /dev/null
  echo "built ab/lib_$tag.so ($flags)"
done
rm -rf build_ab

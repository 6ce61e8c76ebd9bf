#!/bin/bash
# Build libjds variants that differ only in jds_ssim.cu (A/B runs on the GPU box):
#   tools/ssim_variants.sh            -> build/variants/libjds_<name>.so
# then on the box:  for f in build/variants/*.so; do JDS_LIB=$f python tools/ssim_time.py; done
# `v1` is the round-1 kernel (kept under build/ssim_v1 when present).
set -e
cd "$(dirname "$0")/.."
CS=jpeg_dsp_studio_b200/csrc
OUT=build/variants
mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
for f in jds_api jds_kernels jds_fused jds_fused_exact jds_preview jds_ops jds_alias jds_entropy; do
  [ $OUT/obj/$f.o -nt $CS/$f.cu ] || nvcc $FLAGS -c $CS/$f.cu -o $OUT/obj/$f.o &
done
wait
build() { # name, source, defines...
  name=$1; src=$2; shift; shift
  nvcc $FLAGS -I$CS "$@" -c $src -o $OUT/obj/ssim_$name.o
  nvcc -shared -o $OUT/libjds_$name.so $OUT/obj/jds_*.o $OUT/obj/ssim_$name.o
}
rm -f $OUT/libjds_*.so
build base $CS/jds_ssim.cu &
for v in "$@"; do   # extra variants: name:-DFLAG=1,-DOTHER=2
  name=${v%%:*}; defs=${v#*:}
  build $name $CS/jds_ssim.cu ${defs//,/ } &
done
wait
ls -la $OUT/*.so

#!/bin/bash
# Build a full libjds variant with extra -D flags (A/B runs on the GPU box):
#   tools/lib_variant.sh name -DJDS_XL_MIN_CTAS=3 ...   -> build/variants/libjds_<name>.so
# Only the translation units named in FILES are recompiled with the flags; the rest come from
# the object cache of tools/ssim_variants.sh (run that first).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
CS=jpeg_dsp_studio_b200/csrc
OUT=build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
mkdir -p $OUT/obj_$name
for f in jds_api jds_kernels jds_fused jds_fused_exact jds_preview jds_ops jds_alias jds_entropy jds_ssim; do
  nvcc $FLAGS "$@" -c $CS/$f.cu -o $OUT/obj_$name/$f.o &
done
wait
nvcc -shared -o $OUT/libjds_$name.so $OUT/obj_$name/*.o
ls -la $OUT/libjds_$name.so

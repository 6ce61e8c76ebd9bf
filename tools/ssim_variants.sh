#!/bin/bash
# Build libjds variants that differ only in jds_ssim.cu's tuning knobs (A/B runs on the GPU box):
#   tools/ssim_variants.sh            -> build/variants/libjds_<name>.so
# then on the box:  for f in build/variants/*.so; do JDS_LIB=$f python tools/ssim_time.py; done
set -e
cd "$(dirname "$0")/.."
CS=jpeg_dsp_studio_b200/csrc
OUT=build/variants
mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
for f in jds_api jds_kernels jds_fused jds_preview jds_ops jds_alias jds_entropy; do
  [ $OUT/obj/$f.o -nt $CS/$f.cu ] || nvcc $FLAGS -c $CS/$f.cu -o $OUT/obj/$f.o &
done
wait
build() { # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -c $CS/jds_ssim.cu -o $OUT/obj/ssim_$name.o
  nvcc -shared -o $OUT/libjds_$name.so $OUT/obj/jds_*.o $OUT/obj/ssim_$name.o
}
build base &
build ctas24 -DJDS_SSIM_CTAS_PER_SM=24 &
build ctas32 -DJDS_SSIM_CTAS_PER_SM=32 &
build ctas40 -DJDS_SSIM_CTAS_PER_SM=40 &
build naive24 -DJDS_SSIM_COMPENSATED=0 -DJDS_SSIM_CTAS_PER_SM=24 &
wait
ls -la $OUT/*.so

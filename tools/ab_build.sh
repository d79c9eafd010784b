#!/bin/bash
# A/B tuning builds of the C-ABI library: tools/ab_build.sh <tag> [-DNAME=VALUE ...]
#   -> gpurun_ab/libvosd_<tag>.so ; select it at run time with VOSD_B200_LIB=gpurun_ab/libvosd_<tag>.so
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p gpurun_ab/obj_$tag
for f in api roialign proposals collect paste flow_align mask_nms overlaps; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
       -c vosdetectron_b200/csrc/$f.cu -o gpurun_ab/obj_$tag/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_ab/libvosd_$tag.so gpurun_ab/obj_$tag/*.o
rm -rf gpurun_ab/obj_$tag
echo gpurun_ab/libvosd_$tag.so

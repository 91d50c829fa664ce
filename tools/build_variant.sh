#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG=VALUE ...] -- tuning build of the C-ABI library with extra macros, written to
# build/variants/libzj_NAME.so (select it with ZJ_LIB_PATH; build/ is git-ignored but travels with gpurun).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/zune-jpeg_b200/csrc
out=$root/build/variants
mkdir -p $out/obj_$name
for f in zj_kernels.cu zj_capi.cu zj_entropy.cu zj_consumer.cu zj_host_decoder.cpp; do
  o=$out/obj_$name/${f%.*}.o
  if [ $f = zj_kernels.cu ] || [ $f = zj_entropy.cu ] || [ ! -f $o ] || [ $src/$f -nt $o ]; then
    rm -f $o
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O3 "$@" -x cu -c $src/$f -o $o &
  fi
done
wait
for f in zj_kernels zj_capi zj_entropy zj_consumer zj_host_decoder; do [ -f $out/obj_$name/$f.o ] || { echo "FAILED: $f"; exit 1; }; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libzj_$name.so $out/obj_$name/*.o -cudart shared -lpthread
echo built $out/libzj_$name.so

#!/bin/bash
# tools/ab_build.sh <name> [extra nvcc flags...] : variant build of librtw_cuda.so for tools/ab.py
set -e
N=$1; shift
cd "$(dirname "$0")/../raytracer-weekend_b200/csrc"
mkdir -p build_$N ../lib/ab_$N
for f in rtw_api rtw_bvh rtw_trace rtw_render rtw_multi rtw_mem; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC,-ffp-contract=off "$@" -c $f.cu -o build_$N/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../lib/ab_$N/librtw_cuda.so build_$N/*.o
echo built ab_$N

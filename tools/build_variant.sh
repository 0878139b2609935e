#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>  ->  cocodr_b200/libcocodr_b200_<name>.so (A/B experiments;
# select with COCODR_B200_LIB=...)
name=$1; shift
out=cocodr_b200/libcocodr_b200_$name.so
mkdir -p build/var_$name
pids=()
for f in cocodr_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $f -o build/var_$name/$(basename $f .cu).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p || exit 1; done
nvcc -shared -o $out build/var_$name/*.o -lcudart && echo built $out

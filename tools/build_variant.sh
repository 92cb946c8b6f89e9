#!/bin/bash
# Build a variant of the CUDA library with extra nvcc defines into tools/_bin/
#   tools/build_variant.sh NAME -DINF_MIN_BLOCKS=3 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/tools/_bin/var_$name
mkdir -p $out
cd $root/infimum_b200/csrc
for f in poseidon_t2 poseidon_t3 poseidon_t4 poseidon_t5 poseidon_t6 poseidon_t7 poseidon_t8 dense_generic leaves tree_paths imad_peak multi capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wno-unknown-pragmas -I ../../include -I . "$@" -c $f.cu -o $out/$f.o &
done
g++ -O2 -std=c++17 -fPIC -Wno-unknown-pragmas -I ../../include -I . "$@" -c host_params.cpp -o $out/host_params.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/tools/_bin/libinfimum_b200_$name.so $out/*.o -lcudart_static -lpthread -ldl -lrt
echo built $root/tools/_bin/libinfimum_b200_$name.so

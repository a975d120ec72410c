#!/bin/bash
# Build a tuning variant of libocean_b200.so: tools/build_variant.sh <name> <file.cu> "<-D flags>"  ->  gpurun_variants/lib_<name>.so
# (only <file.cu> is recompiled with the extra flags; select the result with OB200_LIB=gpurun_variants/lib_<name>.so)
set -e
cd "$(dirname "$0")/../oceanscalingtests.jl_b200/csrc"
name=$1; src=$2; flags=$3
mkdir -p ../../gpurun_variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -fmad=false -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c $src -o /tmp/variant_$name.o
objs=""
for o in ocean_b200 kernels_column kernels_baro kernels_halo kernels_tend kernels_leith; do
  if [ "$o.cu" == "$src" ]; then objs="$objs /tmp/variant_$name.o"; else objs="$objs $o.o"; fi
done
$NVCC -shared $ARCH -Xlinker -Bsymbolic -o ../../gpurun_variants/lib_$name.so $objs -L/usr/lib/x86_64-linux-gnu -l:libnccl.so.2 -lcudart
echo built gpurun_variants/lib_$name.so

#!/bin/bash
# Build a measurement variant of libsnb.so: scripts/build_variant.sh <name> <-D flags...>
# Recompiles snb_tc.cu with the extra defines, links it with the objects of the regular build into
# switch_nerf_b200/variants/libsnb_<name>.so (git-ignored, travels with gpurun).  Regular build must be current.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p switch_nerf_b200/variants switch_nerf_b200/csrc/build/var_$name
obj=switch_nerf_b200/csrc/build/var_$name/snb_tc.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cudart static \
  --expt-relaxed-constexpr "$@" -c switch_nerf_b200/csrc/snb_tc.cu -o $obj
objs=$(ls switch_nerf_b200/csrc/build/*.o | grep -v snb_tc.o)
/usr/local/cuda/bin/nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o switch_nerf_b200/variants/libsnb_$name.so $obj $objs
echo "VARIANT OK switch_nerf_b200/variants/libsnb_$name.so"

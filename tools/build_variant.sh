#!/bin/bash
# builds moby_b200/libb200moby_$1.so with extra nvcc flags $2.. (experiments; the product library is built by __graft_entry__.build)
set -e
name=$1; shift
mkdir -p build/$name
for f in capi_common lcp_kernels sim_kernels k_fused k_advance k_impact_warp k_impact_thread k_impact_block64 k_impact_block128 k_impact_block256 k_rc k_stabilize; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -O2 -w "$@" -c moby_b200/csrc/$f.cu -o build/$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o moby_b200/libb200moby_$name.so build/$name/*.o -lcudart
echo built moby_b200/libb200moby_$name.so

#!/bin/bash
# build container: three builds of the library that differ only in the tile-GEMM epilogue arithmetic (A/B on the GPU box
# through HDPO_LIB_PATH): all scalar (default) | scalar sums + packed ELU | both packed | packed sums + scalar ELU
set -e
cd /root/repo
PKG=neural_inventory_control_b200
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I $PKG/csrc"
mkdir -p tools/_build
for v in "s0e1:-DHDPO_EPI_PACK_SUM=0 -DHDPO_EPI_PACK_ELU=1" "s1e1:-DHDPO_EPI_PACK_SUM=1 -DHDPO_EPI_PACK_ELU=1" "s1e0:-DHDPO_EPI_PACK_SUM=1 -DHDPO_EPI_PACK_ELU=0"; do
  tag=${v%%:*}; flags=${v#*:}
  $NV $flags -c $PKG/csrc/gemm_tc.cu -o tools/_build/gemm_tc_$tag.o
  objs=$(ls $PKG/build/*.o | grep -v gemm_tc.cu.o)
  nvcc -shared -o tools/_build/libhdpo_$tag.so $objs tools/_build/gemm_tc_$tag.o -gencode arch=compute_100a,code=sm_100a -cudart static
done
ls -la tools/_build/*.so

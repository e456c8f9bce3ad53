#!/bin/bash
# build container: alternative builds of the library that differ in ONE compile-time switch of one source file, for A/B
# timing on the GPU box through HDPO_LIB_PATH (tools/wg_ab.sh runs them interleaved). Spec: tag:source:flags
set -e
cd /root/repo
PKG=neural_inventory_control_b200
python -m neural_inventory_control_b200.build > /dev/null
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I $PKG/csrc"
mkdir -p tools/_build
rm -f tools/_build/libhdpo_*.so tools/_build/ab_*.o
SPECS=("$@")
if [ ${#SPECS[@]} -eq 0 ]; then
  SPECS=("pdl1:gemm_tc.cu:-DHDPO_PDL_TRIGGER=1" "pdl2:gemm_tc.cu:-DHDPO_PDL_TRIGGER=2" "pdl3:gemm_tc.cu:-DHDPO_PDL_TRIGGER=3"
         "hw2:rollout_wide.cu:-DHDPO_HEAD_WARPS=2" "hw8:rollout_wide.cu:-DHDPO_HEAD_WARPS=8")
fi
for v in "${SPECS[@]}"; do
  tag=${v%%:*}; rest=${v#*:}; src=${rest%%:*}; flags=${rest#*:}
  $NV $flags -c $PKG/csrc/$src -o tools/_build/ab_$tag.o &
done
wait
for v in "${SPECS[@]}"; do
  tag=${v%%:*}; rest=${v#*:}; src=${rest%%:*}
  objs=$(ls $PKG/build/*.o | grep -v "/$src.o")
  nvcc -shared -o tools/_build/libhdpo_$tag.so $objs tools/_build/ab_$tag.o -gencode arch=compute_100a,code=sm_100a -cudart static
done
ls tools/_build/*.so

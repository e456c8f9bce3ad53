# GPU box: A/B of compile-time variants of the tile GEMM built by tools/epi_ab.sh (interleaved runs; first a parity check)
out=gpurun_out/r2n_epi_ab.log; rm -f $out
VARIANTS="${VARIANTS:-default pipe pipe_l2 l2}"
for v in $VARIANTS; do
  if [ $v != default ]; then
    HDPO_LIB_PATH=/root/repo/tools/_build/libhdpo_$v.so timeout 300 python -m pytest tests/test_kernels_abi.py tests/test_gemm_tc.py -m gpu -x -q -k "wide or gemm" 2>&1 | tail -1 | sed "s/^/$v parity: /" >> $out
  fi
done
for rep in 1 2 3; do
for v in $VARIANTS; do
  if [ $v = default ]; then unset HDPO_LIB_PATH; else export HDPO_LIB_PATH=/root/repo/tools/_build/libhdpo_$v.so; fi
  echo -n "$v: " >> $out
  timeout 200 python tools/wide_ab.py one_warehouse_lost_demand 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out
done; done
cat $out

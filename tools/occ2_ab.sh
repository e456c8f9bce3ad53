export HDPO_TC_OCC2=1
timeout 600 python -m pytest tests/test_kernels_abi.py tests/test_full_size_properties.py -m gpu -x -q -k "wide or warehouse" > gpurun_out/r2c_occ2_tests.log 2>&1
tail -3 gpurun_out/r2c_occ2_tests.log
for occ in 0 1 2; do for ch in 2 4 8; do
  echo "OCC2=$occ CHUNKS=$ch" >> gpurun_out/r2c_occ2_time.log
  HDPO_TC_OCC2=$occ HDPO_WIDE_CHUNKS=$ch timeout 120 python tools/graph_replay_time.py 2>&1 | head -1 >> gpurun_out/r2c_occ2_time.log
done; done
cat gpurun_out/r2c_occ2_time.log

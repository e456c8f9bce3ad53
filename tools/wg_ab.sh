rm -f gpurun_out/r2m_epi_ab.log
for rep in 1 2 3; do
for v in default s0e1 s0e0 s1e0; do
  if [ $v = default ]; then unset HDPO_LIB_PATH; else export HDPO_LIB_PATH=/root/repo/tools/_build/libhdpo_$v.so; fi
  echo -n "$v: " >> gpurun_out/r2m_epi_ab.log
  timeout 200 python tools/wide_ab.py one_warehouse_lost_demand 2>&1 | tail -1 >> gpurun_out/r2m_epi_ab.log
done; done
cat gpurun_out/r2m_epi_ab.log

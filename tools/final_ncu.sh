# GPU box: ncu launch list of one step of the default workload + full-set captures (final code): forward tile GEMMs,
# adjoint tile GEMMs, CTA-pair weight-gradient tiles
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2c_launches_wide.csv python tools/profile_one_step.py one_warehouse_lost_demand > gpurun_out/r2c_ncu1.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c_launches_wide.csv > gpurun_out/r2c_launches_wide.summary.txt 2>&1
head -14 gpurun_out/r2c_launches_wide.summary.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 6 \
    -o gpurun_out/r2c_gemm_fwd -f python tools/profile_one_step.py one_warehouse_lost_demand > gpurun_out/r2c_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c_gemm_fwd.ncu-rep > gpurun_out/r2c_ncu_gemm_fwd.txt 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tc_kernel -s 1100 -c 6 \
    -o gpurun_out/r2c_gemm_bwd -f python tools/profile_one_step.py one_warehouse_lost_demand > gpurun_out/r2c_ncu3.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c_gemm_bwd.ncu-rep > gpurun_out/r2c_ncu_gemm_bwd.txt 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:"gemm_tc_kernel<128, 4" -c 3 \
    -o gpurun_out/r2c_gemm_wgrad -f python tools/profile_one_step.py one_warehouse_lost_demand > gpurun_out/r2c_ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c_gemm_wgrad.ncu-rep > gpurun_out/r2c_ncu_gemm_wgrad.txt 2>&1
grep -h "^##" -A1 gpurun_out/r2c_ncu_gemm_wgrad.txt | head -8
grep -h "^##" gpurun_out/r2c_ncu_gemm_bwd.txt | head -8

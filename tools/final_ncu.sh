# GPU box: ncu launch list of one step of the default workload + full-set capture of the tile GEMM launches (final code)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2b_launches_wide.csv python tools/profile_one_step.py one_warehouse_lost_demand > gpurun_out/r2b_ncu1.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches_wide.csv > gpurun_out/r2b_launches_wide.summary.txt 2>&1
head -12 gpurun_out/r2b_launches_wide.summary.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 10 \
    -o gpurun_out/r2b_gemm_wide -f python tools/profile_one_step.py one_warehouse_lost_demand > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out/r2b_gemm_wide.ncu-rep
python tools/ncu_summary.py gpurun_out/r2b_gemm_wide.ncu-rep > gpurun_out/r2b_ncu_gemm_wide.txt 2>&1
head -8 gpurun_out/r2b_ncu_gemm_wide.txt
ncu --set full --clock-control none -k regex:step_ -s 20 -c 4 -o gpurun_out/r2b_step_k3 -f python tools/step_time.py > gpurun_out/r2b_ncu3.log 2>&1
python tools/ncu_summary.py gpurun_out/r2b_step_k3.ncu-rep > gpurun_out/r2b_ncu_step_k3.txt 2>&1
head -8 gpurun_out/r2b_ncu_step_k3.txt

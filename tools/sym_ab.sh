out=gpurun_out/r2s_sym_wg.log; rm -f $out
timeout 600 python -m pytest tests/test_symmetry_aware_fused.py -m gpu -x -q 2>&1 | tail -1 >> $out
HDPO_WIDE_WG_SYM=1 timeout 600 python -m pytest tests/test_symmetry_aware_fused.py -m gpu -x -q 2>&1 | tail -1 >> $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py one_warehouse_lost_demand_symmetry_aware 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
for rep in 1 2; do
run HDPO_X=default
run HDPO_WIDE_WG_SYM=1
run HDPO_WIDE_WG_SYM=1 HDPO_WIDE_WG_GROUP=10
run HDPO_WIDE_WG_SYM=1 HDPO_WIDE_WG_GROUP=2
done
cat $out

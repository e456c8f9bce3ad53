out=gpurun_out/r3f_chunks_small.log; rm -f $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
for WL in one_warehouse_lost_demand many_warehouses_lost_demand; do
for b in 2048 4096; do
run HDPO_AB_BATCH=$b
run HDPO_AB_BATCH=$b HDPO_WIDE_CHUNKS=2
run HDPO_AB_BATCH=$b HDPO_WIDE_CHUNKS=2 HDPO_WIDE_KSPLIT=1 HDPO_WIDE_WG_OVERLAP=1
run HDPO_AB_BATCH=$b HDPO_WIDE_CHUNKS=4
run HDPO_AB_BATCH=$b HDPO_WIDE_CHUNKS=4 HDPO_WIDE_KSPLIT=1
done; done
cat $out

out=gpurun_out/r2y_chunks.log; rm -f $out
run() { echo -n "$*: " >> $out; env "$@" timeout 200 python tools/wide_ab.py $WL 2>&1 | tail -1 | sed 's/\[.*\]//' >> $out; }
WL=one_warehouse_lost_demand
for rep in 1 2; do
for ch in 3 4 5 6; do run HDPO_WIDE_CHUNKS=$ch; done
done
cat $out

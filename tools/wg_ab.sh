out=gpurun_out/r3a_hichunks.log; rm -f $out
for h in 1 2; do
echo "HDPO_TC_HI_CHUNKS=$h parity:" >> $out
HDPO_TC_HI_CHUNKS=$h timeout 900 python -m pytest tests/test_kernels_abi.py tests/test_full_size_properties.py -m gpu -q -k "wide or warehouse" 2>&1 | tail -6 >> $out
done
for h in 3 2 1; do HDPO_TC_HI_CHUNKS=$h python tools/wg_accuracy.py hc$h | tail -1 >> $out; done
python tools/wg_accuracy.py fp32 | tail -1 >> $out
python tools/wg_accuracy.py compare hc3 hc2 hc1 fp32 >> $out
python tools/wg_accuracy.py compare fp32 hc3 hc2 hc1 >> $out
rm -f gpurun_out/wg_grad_*.npy
cat $out

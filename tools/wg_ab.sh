for k in 128 512 1024 2048 4096; do HDPO_WG_KPS=$k python tools/wg_accuracy.py $k | tail -1; done
python tools/wg_accuracy.py fp32 | tail -1
python tools/wg_accuracy.py compare 128 512 1024 2048 4096 fp32
python tools/wg_accuracy.py compare fp32 128 512 1024 2048 4096
rm -f gpurun_out/wg_grad_*.npy

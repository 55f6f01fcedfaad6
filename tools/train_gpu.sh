# training step: stage-by-stage check against the fp32 PyTorch restatement, steady-state timing, ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout -k 10 600 python tools/train_check.py 2 64 > gpurun_out/train_check.log 2>&1; echo "rc=$?" >> gpurun_out/train_check.log
timeout -k 10 300 python tools/train_timing.py 512 100 > gpurun_out/train_timing.log 2>&1
AZ_TRAIN_PDL=0 timeout -k 10 300 python tools/train_timing.py 512 100 >> gpurun_out/train_timing.log 2>&1
cat gpurun_out/train_timing.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/train_launches.csv python tools/train_timing.py 512 2 > gpurun_out/train_ncu.log 2>&1
timeout -k 10 900 ncu --set full --import-source on --clock-control none -k regex:'k_wgrad$|k_conv|k_bn_apply|k_bn_bwd_apply|k_wgrad_reduce|k_heads' -s 400 -c 14 -o gpurun_out/train_full -f python tools/train_timing.py 512 2 > gpurun_out/train_ncu_full.log 2>&1
ls -la gpurun_out/train_full.ncu-rep

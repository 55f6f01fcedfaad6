set -x
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_net_pair -s 300 -c 1 -f -o gpurun_out/r02_pair python tools/profile_pool.py 2048 800 400 > gpurun_out/r02_ncu_pair.log 2>&1
tail -3 gpurun_out/r02_ncu_pair.log

set -x
mkdir -p gpurun_out
python tools/profile_pool.py 2048 800 96 > gpurun_out/pool_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 120 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_pool.py 2048 800 96 > gpurun_out/ncu_launch.log 2>&1
python tools/profile_net.py 2048 4 > gpurun_out/net_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 1 -c 2 -o gpurun_out/net_tc_r01 python tools/profile_net.py 2048 4 > gpurun_out/ncu_net.log 2>&1
tail -3 gpurun_out/*.log
ls -la gpurun_out

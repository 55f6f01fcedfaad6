set -x
mkdir -p gpurun_out
python tools/profile_pool.py 2048 800 64 > gpurun_out/pool_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_pool.py 2048 800 64 > gpurun_out/ncu_launch.log 2>&1
python tools/profile_pool.py 512 800 40 > gpurun_out/pool_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tree_tick -s 36 -c 2 -o gpurun_out/tree_tick_r01 python tools/profile_pool.py 512 800 40 > gpurun_out/ncu_tree.log 2>&1
ls -la gpurun_out

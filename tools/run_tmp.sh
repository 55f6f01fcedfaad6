python tools/profile_pool.py 2048 800 64 > gpurun_out/pool_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_tree_tick" -s 60 -c 2 -o gpurun_out/tree_r01b python tools/profile_pool.py 2048 800 64 > gpurun_out/ncu_tree_b.log 2>&1
tail -2 gpurun_out/ncu_tree_b.log

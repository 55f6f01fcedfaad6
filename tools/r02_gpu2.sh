# r02 GPU call 2: full parity suite, net error report, finish-time distribution of the tree kernel, ncu of the tree kernel, full bench
set -x
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -25 gpurun_out/r02_pytest2.log
timeout -k 10 600 python tools/net_error.py > gpurun_out/r02_net_error.txt 2>&1; tail -60 gpurun_out/r02_net_error.txt
AZ_POOL_PROFILE=1 timeout -k 10 300 python tools/profile_pool.py 2048 800 1200 > gpurun_out/r02_phase2.log 2>&1
grep -a "profile\|^ok" gpurun_out/r02_phase2.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_tree_tick -s 900 -c 1 -f -o gpurun_out/r02_tree python tools/profile_pool.py 2048 800 1000 > gpurun_out/r02_ncu_tree.log 2>&1
tail -3 gpurun_out/r02_ncu_tree.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench2.log 2> gpurun_out/r02_bench2.err; echo "bench rc=$?" >> gpurun_out/r02_bench2.err
cat gpurun_out/r02_bench2.log; tail -5 gpurun_out/r02_bench2.err

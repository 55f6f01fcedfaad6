# r02 GPU call 4: speculative single-tree search, favourite-child prefetch, pipelined softmax helper
set -x
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -40 gpurun_out/r02_pytest4.log
{
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_TREE_FAVOURITE=0 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_TREE_SOFTMAX=1 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_TREE_SOFTMAX=1 AZ_TREE_FAVOURITE=0 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
timeout -k 10 300 python tools/tick_timing.py 256 400 2048
timeout -k 10 300 python tools/net_timing.py
} > gpurun_out/r02_ticks4.log 2>&1
grep -v "^+" gpurun_out/r02_ticks4.log
AZ_POOL_PROFILE=1 timeout -k 10 300 python tools/profile_pool.py 2048 800 1200 > gpurun_out/r02_phase4.log 2>&1
grep -a "profile\|^ok" gpurun_out/r02_phase4.log
timeout -k 10 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench4.log 2> gpurun_out/r02_bench4.err; echo "bench rc=$?" >> gpurun_out/r02_bench4.err
cat gpurun_out/r02_bench4.log; tail -5 gpurun_out/r02_bench4.err

# r02 GPU call 3: softmax front half in the net kernel, fused sym8, f16 mode: parity, A/B timing, full bench
set -x
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -25 gpurun_out/r02_pytest3.log
{
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_TREE_SOFTMAX=1 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_LEVELS_PER_TICK=32 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_LEVELS_PER_TICK=40 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
timeout -k 10 300 python tools/tick_timing.py 256 400 2048
timeout -k 10 300 python tools/net_timing.py
} > gpurun_out/r02_ticks3.log 2>&1
grep -v "^+" gpurun_out/r02_ticks3.log
AZ_POOL_PROFILE=1 timeout -k 10 300 python tools/profile_pool.py 2048 800 1200 > gpurun_out/r02_phase3.log 2>&1
grep -a "profile\|^ok" gpurun_out/r02_phase3.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench3.log 2> gpurun_out/r02_bench3.err; echo "bench rc=$?" >> gpurun_out/r02_bench3.err
cat gpurun_out/r02_bench3.log; tail -5 gpurun_out/r02_bench3.err

# r02 GPU call 5: early (programmatic) launch of the net kernel behind the tree kernel
set -x
mkdir -p gpurun_out
{
# smallest possible exposure first: a hang here must not cost more than a minute
timeout -k 5 90 python tools/tick_timing.py 256 400 256 || echo "EARLY-MODE SMOKE FAILED rc=$?"
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_POOL_EARLY=0 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_LEVELS_PER_TICK=64 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_LEVELS_PER_TICK=96 AZ_REQ_CAP=2048 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
timeout -k 10 300 python tools/tick_timing.py 256 400 2048
AZ_POOL_EARLY=0 timeout -k 10 300 python tools/tick_timing.py 256 400 2048
} > gpurun_out/r02_ticks5.log 2>&1
grep -v "^+" gpurun_out/r02_ticks5.log | tail -20
timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
tail -30 gpurun_out/r02_pytest5.log
timeout -k 10 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench5.log 2> gpurun_out/r02_bench5.err; echo "bench rc=$?" >> gpurun_out/r02_bench5.err
cat gpurun_out/r02_bench5.log | cut -c1-3000; tail -5 gpurun_out/r02_bench5.err

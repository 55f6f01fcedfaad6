set -x
mkdir -p gpurun_out
{
timeout -k 5 90 python tools/tick_timing.py 256 400 256 || echo "EARLY-MODE SMOKE FAILED rc=$?"
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_POOL_EARLY=0 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
timeout -k 10 300 python tools/tick_timing.py 256 400 2048
AZ_POOL_EARLY=0 timeout -k 10 300 python tools/tick_timing.py 256 400 2048
AZ_POOL_TRACE=1 timeout -k 10 300 python tools/tick_timing.py 2048 800 256
} > gpurun_out/r02_ticks6.log 2>&1
grep -v "^+" gpurun_out/r02_ticks6.log | tail -20

# r02 GPU call 1: parity of the compact tree kernel, tick timing against the round-1 build (_v1/), budget sweeps, phase profile
set -x
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -25 gpurun_out/r02_pytest1.log
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke1.log
tail -3 gpurun_out/r02_smoke1.log
{
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
(cd _v1 && timeout -k 10 300 python tools/tick_timing.py 2048 800 1024)
for L in 16 24 32 64; do AZ_LEVELS_PER_TICK=$L timeout -k 10 300 python tools/tick_timing.py 2048 800 1024; done
for C in 50000 70000 90000 120000; do AZ_TICK_CYCLES=$C AZ_LEVELS_PER_TICK=200 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024; done
AZ_REQ_CAP=1480 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
AZ_REQ_CAP=2048 timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
timeout -k 10 300 python tools/tick_timing.py 256 400 2048
(cd _v1 && timeout -k 10 300 python tools/tick_timing.py 256 400 2048)
} > gpurun_out/r02_ticks1.log 2>&1
cat gpurun_out/r02_ticks1.log | grep -v "^+"
AZ_POOL_PROFILE=1 timeout -k 10 300 python tools/profile_pool.py 2048 800 512 > gpurun_out/r02_phase1.log 2>&1
grep profile gpurun_out/r02_phase1.log

# quick GPU check: full parity suite + steady-state tick timing
set -x
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -25 gpurun_out/pytest_quick.log
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024 2>&1 | tail -1

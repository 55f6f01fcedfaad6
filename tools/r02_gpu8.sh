set -x
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
tail -12 gpurun_out/r02_pytest8.log
{
timeout -k 10 300 python tools/tick_timing.py 2048 800 1024
timeout -k 10 300 python tools/tick_timing.py 256 400 2048
} > gpurun_out/r02_ticks8.log 2>&1
grep -v "^+" gpurun_out/r02_ticks8.log
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke8.log 2>&1; tail -2 gpurun_out/r02_smoke8.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench8.log 2> gpurun_out/r02_bench8.err; echo "bench rc=$?" >> gpurun_out/r02_bench8.err
cat gpurun_out/r02_bench8.log | cut -c1-1500; tail -3 gpurun_out/r02_bench8.err

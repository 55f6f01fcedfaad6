# One GPU call: tests, smoke, bench (both arms), ncu launch list of the bench command, phase profile of the tree kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.log
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
BENCH_SMALL="python bench.py --steps 2 --warmup 3 --ticks-per-step 24 --no-cpu-baseline --no-single-tree"
$BENCH_SMALL > gpurun_out/bench_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $BENCH_SMALL > gpurun_out/ncu_launch.log 2>&1
AZ_POOL_PROFILE=1 python tools/profile_pool.py 2048 800 1200 > gpurun_out/phase.log 2>&1
grep -a "profile" gpurun_out/phase.log
ls -la gpurun_out | head -40

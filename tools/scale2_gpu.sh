# 2-GPU sanity of the driver's scaling launch: torchrun, one rank per GPU over NCCL (counters only)
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
cut -c1-1200 gpurun_out/bench_n2.log; tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "rc=$?" >> gpurun_out/bench_ref_n2.err
cut -c1-400 gpurun_out/bench_ref_n2.log; tail -2 gpurun_out/bench_ref_n2.err

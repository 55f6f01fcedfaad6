python -m pytest tests/test_cli_gpu.py tests/test_gpu_server.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo rc=$?; cat gpurun_out/bench_n8.log | cut -c1-900; tail -3 gpurun_out/bench_n8.err

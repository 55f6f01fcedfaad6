set -x
mkdir -p gpurun_out
{
AZ_NET_PAIR=1 timeout -k 5 100 python tools/net_error.py 2>&1 | head -40 || echo "PAIR net_error FAILED/HUNG rc=$?"
AZ_NET_PAIR=1 timeout -k 5 200 python -m pytest tests/test_net_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15
AZ_NET_PAIR=1 timeout -k 5 100 python tools/net_timing.py
timeout -k 5 100 python tools/net_timing.py
AZ_NET_PAIR=1 timeout -k 10 200 python tools/tick_timing.py 2048 800 1024
timeout -k 10 200 python tools/tick_timing.py 2048 800 1024
} > gpurun_out/r02_pair9.log 2>&1
grep -v "^+" gpurun_out/r02_pair9.log | tail -70

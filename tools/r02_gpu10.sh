set -x
mkdir -p gpurun_out
{
for R in 2 1 0; do AZ_NET_PAIR=1 AZ_PAIR_RELAY=$R timeout -k 5 100 python tools/net_timing.py 2>&1 | grep "B=2048\|B=256 "; done
AZ_NET_PAIR=1 AZ_PAIR_MAX_CTAS=148 timeout -k 5 100 python tools/net_timing.py 2>&1 | grep "B=2048\|B=256 "
AZ_NET_MAX_CTAS=148 timeout -k 5 100 python tools/net_timing.py 2>&1 | grep "B=2048\|B=256 "
AZ_NET_PAIR=1 timeout -k 5 200 python -m pytest tests/test_net_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
} > gpurun_out/r02_pair10.log 2>&1
grep -v "^+" gpurun_out/r02_pair10.log | tail -30

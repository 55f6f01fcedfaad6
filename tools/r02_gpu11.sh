set -x
mkdir -p gpurun_out
{
AZ_NET_PAIR=1 AZ_PAIR_RELAY=2 timeout -k 10 200 python tools/tick_timing.py 2048 800 1024
AZ_NET_PAIR=1 AZ_PAIR_RELAY=1 timeout -k 10 200 python tools/tick_timing.py 2048 800 1024
timeout -k 10 200 python tools/tick_timing.py 2048 800 1024
AZ_NET_PAIR=1 timeout -k 10 200 python tools/tick_timing.py 256 400 2048
timeout -k 10 200 python tools/tick_timing.py 256 400 2048
AZ_NET_PAIR=1 timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5
} > gpurun_out/r02_pair11.log 2>&1
grep -v "^+" gpurun_out/r02_pair11.log | tail -30

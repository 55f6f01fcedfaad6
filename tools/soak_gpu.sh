set -x
mkdir -p gpurun_out
{
echo "# python tools/soak.py 120 800   (1 x B200, 2048 games x 800 visits from the standard opening, bf16 net on CTA pairs, noise on, JSON records written)"
timeout -k 10 600 python tools/soak.py 120 800
echo "# python tools/soak.py 60 100    (same pool, 100 visits per move: 180 finished games/s, exercises recycling + the one-copy drain)"
timeout -k 10 600 python tools/soak.py 60 100
} > gpurun_out/r02_soak.txt 2>&1
grep -v "^+" gpurun_out/r02_soak.txt
timeout -k 10 300 python -m pytest tests/test_net_gpu.py -q -m gpu -p no:cacheprovider -k "bitwise" 2>&1 | tail -3

set -x
mkdir -p gpurun_out
{
AZ_POOL_PROFILE=1 timeout -k 10 200 python tools/profile_pool.py 2048 800 1023
AZ_POOL_PROFILE=1 timeout -k 10 200 python tools/profile_pool.py 256 400 1023
} > gpurun_out/r02_early7.log 2>&1
grep -a "profile" gpurun_out/r02_early7.log

export AZ_POOL_TRACE=1 AZ_NET_CLUSTER=2
AZ_POOL_GROUPS=2 AZ_POOL_NET_TILES=2 timeout 200 python tools/tick_timing.py 2048 800 512 2>&1 | grep -v "^$"
AZ_POOL_GROUPS=2 AZ_POOL_NET_TILES=1 timeout 200 python tools/tick_timing.py 2048 800 512 2>&1 | grep -v "^$"
AZ_POOL_GROUPS=2 AZ_POOL_NET_TILES=2 AZ_NET_EXPERIMENT=2 timeout 200 python tools/tick_timing.py 2048 800 512 2>&1 | grep -v "^$"

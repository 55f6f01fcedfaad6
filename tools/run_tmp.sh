python -m pytest tests/test_mcts_gpu.py tests/test_selfplay_gpu.py -x -q 2>&1 | tail -3
for lv in 48 64 80; do AZ_POOL_PROFILE=1 AZ_LEVELS_PER_TICK=$lv timeout 200 python tools/tick_timing.py 2048 800 1024 2>&1 | grep -v "^$" | grep -v "populate\|backup\|expand\|make_move"; done

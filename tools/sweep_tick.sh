for cfg in "48 0" "48 1776" "96 1776" "100000 1776" "64 1184" "100000 2048"; do set -- $cfg
AZ_LEVELS_PER_TICK=$1 AZ_REQ_CAP=$2 timeout 200 python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import bench, ataxxzero_b200 as az
from ataxxzero_b200 import model, net, search
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
pool = search.Pool(ctx, 2048, 800, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=1)
pool.set_roots(bench.synthetic_roots(ctx, 2048, 0))
pool.selfplay_ticks(768)
s0 = pool.stats(); import time; t0 = time.perf_counter()
pool.selfplay_ticks(1024)
dt = time.perf_counter() - t0; s1 = pool.stats()
d = {k: s1[k] - s0[k] for k in s1}
print("levels=%s cap=%s: %.3f ms/tick tree %.3f net %.3f  pos/s %.0f evals/tick %.0f evals/s %.0f" % (os.environ["AZ_LEVELS_PER_TICK"], os.environ["AZ_REQ_CAP"], dt / 1024 * 1e3, d["tree_seconds"] / 1024 * 1e3, d["net_seconds"] / 1024 * 1e3, d["positions"] / dt, d["evals"] / 1024, d["evals"] / dt))
PY
done

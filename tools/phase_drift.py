"""How the positions/s metric drifts while the pool's mix of game phases settles (2048 games x 800 visits)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ataxxzero_b200 as az
from ataxxzero_b200 import model, net, search
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
pool = search.Pool(ctx, 2048, 800, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=1000)
pool.set_roots(bench.synthetic_roots(ctx, 2048, 0))
prev = pool.stats(); t_prev = time.perf_counter()
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    pool.selfplay_ticks(1024)
    st = pool.stats(); t = time.perf_counter()
    d = {k: st[k] - prev[k] for k in st}
    print("step %2d: %.2fs  pos/s %6.0f  evals/s %8.0f  evals/pos %5.0f  term/steps %.3f finished %d tree %.3f net %.3f" % (
        step, t - t_prev, d["positions"] / (t - t_prev), d["evals"] / (t - t_prev), d["evals"] / max(d["positions"], 1),
        d["terminal_steps"] / max(d["steps"], 1), st["games_finished"], d["tree_seconds"], d["net_seconds"]), flush=True)
    prev, t_prev = st, t

"""Steady-state tick timing of the self-play pool (tree ms, net ms, evals/tick, positions/s)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ataxxzero_b200 as az
from ataxxzero_b200 import model, net, search
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
V = int(sys.argv[2]) if len(sys.argv) > 2 else 800
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
pool = search.Pool(ctx, G, V, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=1)
pool.set_roots(bench.synthetic_roots(ctx, G, 0))
pool.selfplay_ticks(768)
s0 = pool.stats(); t0 = time.perf_counter()
pool.selfplay_ticks(T)
dt = time.perf_counter() - t0; s1 = pool.stats()
d = {k: s1[k] - s0[k] for k in s1}
print("G=%d V=%d %s: %.3f ms/tick tree %.3f net %.3f  pos/s %.0f evals/tick %.0f evals/s %.0f steps/tick %.0f" % (
    G, V, " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("AZ_")),
    dt / T * 1e3, d["tree_seconds"] / max(d["timed_ticks"], 1) * 1e3, d["net_seconds"] / max(d["timed_ticks"], 1) * 1e3, d["positions"] / dt, d["evals"] / T, d["evals"] / dt, d["steps"] / T), flush=True)

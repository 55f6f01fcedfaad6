import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, search
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
V = int(sys.argv[2]) if len(sys.argv) > 2 else 800
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
for mode in (search.EVAL_BF16,):
    pool = search.Pool(ctx, G, V, eval_mode=mode, noise=True, auto_play=True, seed=1)
    pool.selfplay_ticks(64)
    s0 = pool.stats(); t0 = time.perf_counter()
    pool.selfplay_ticks(T)
    dt = time.perf_counter() - t0; s1 = pool.stats()
    d = {k: s1[k] - s0[k] for k in s1}
    print("G=%d V=%d ticks=%d: %.3f s, %.3f ms/tick, evals/s %.0f, steps/s %.0f, positions %d -> %.1f pos/s; net %.3f s tree %.3f s; max_depth %d; finished %d"
          % (G, V, T, dt, dt / T * 1e3, d["evals"] / dt, d["steps"] / dt, d["positions"], d["positions"] / dt, d["net_seconds"], d["tree_seconds"], s1["max_depth"], s1["games_finished"]), flush=True)
    pool.close()

"""A few self-play ticks for ncu: G games x V visits from synthetic mid-game roots."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, rules, search
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
V = int(sys.argv[2]) if len(sys.argv) > 2 else 800
T = int(sys.argv[3]) if len(sys.argv) > 3 else 96
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
pool = search.Pool(ctx, G, V, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=1)
pool.set_roots(bench.synthetic_roots(ctx, G, 0))
pool.selfplay_ticks(T)
print("ok", pool.stats())

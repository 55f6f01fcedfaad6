"""Single-tree search (BASELINE configs[4]) under AZ_POOL_PROFILE=1: per-phase cycles of the one tree warp."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, rules, search
V = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
spec = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
fen = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mcts_golden.json")))["midgame_fen"]
with search.Pool(ctx, 1, 64, eval_mode=search.EVAL_BF16, node_capacity=V + 64, steps_per_tick=64, speculate=spec) as tree:
    tree.set_root(0, rules.set_board(fen))
    tree.run()
    tree.set_visits(V)
    s0 = tree.stats(); t0 = time.perf_counter()
    tree.run()
    dt = time.perf_counter() - t0; s1 = tree.stats()
    print("visits/s %.0f ticks %d steps/tick %.2f us/tick %.1f" % ((V - 64) / dt, s1["ticks"] - s0["ticks"], (V - 64) / max(s1["ticks"] - s0["ticks"], 1),
                                                                 dt * 1e6 / max(s1["ticks"] - s0["ticks"], 1)))

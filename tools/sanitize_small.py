"""A small pass over every kernel family for compute-sanitizer --tool memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, rules, search, train_data
ctx = az.Context(0)
p = rules.set_board(rules.START_FEN)
print("perft4", rules.perft(ctx, p, 4), "movegen", len(rules.movegen_batch(ctx, [p])[0]))
print("playouts", rules.random_playouts(ctx, p, 64, 400, 1)[1].sum())
net.load_weights(ctx, model.Network.random_init(seed=0))
f = np.zeros((5, 7, 7, 4), np.float32); f[..., 0] = 1
for mode in (net.FP32, net.BF16):
    print("net", mode, float(net.forward(ctx, f, mode)[1].sum()), float(net.evaluate_symmetric(ctx, f[:2], mode)[1].sum()))
with search.Pool(ctx, 24, 20, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=3) as pool:
    st = pool.selfplay_ticks(300, "/tmp/san.json")
    print("selfplay", st["positions"], st["games_finished"])
with search.Pool(ctx, 3, 60, eval_mode=search.EVAL_BF16) as pool:
    pool.set_root(0, p); pool.run(); r = pool.root(0); pool.play(0, r["moves"][int(np.argmax(r["visits"]))]); pool.run()
    print("search", sum(pool.root(0)["visits"]), pool.principal_variation(0)[:3])
if os.path.exists("/tmp/san.json") and os.path.getsize("/tmp/san.json"):
    packed = train_data.pack_entries(train_data.load_entries(["/tmp/san.json"]))
    print("samples", train_data.minibatch(ctx, packed, 64)[1].sum())
print("done")

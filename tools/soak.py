"""Soak: BASELINE configs[3] shard (2048 games x 800 visits) from the standard opening for N seconds, every finished game
replayed through the CPU oracle (legality, boards, result, k/N distributions)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, search
from oracle.cpu import Oracle, START_FEN
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
visits = int(sys.argv[2]) if len(sys.argv) > 2 else 800
ctx = az.Context(0, seed=99)
net.load_weights(ctx, model.Network.random_init(seed=0))
out = "/tmp/soak.json"
if os.path.exists(out):
    os.unlink(out)
pool = search.Pool(ctx, 2048, visits, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=99)
t0 = time.time()
stats = pool.selfplay(out, max_seconds=seconds)
dt = time.time() - t0
print("ran %.0f s: %s" % (dt, {k: stats[k] for k in ("ticks", "steps", "evals", "positions", "games_finished", "games_skipped", "max_depth")}))
print("positions/s %.0f evals/s %.0f evals/position %.0f" % (stats["positions"] / dt, stats["evals"] / dt, stats["evals"] / max(stats["positions"], 1)))
oracle = Oracle()
games = plies = 0
lengths = []
for ln in open(out):
    g = json.loads(ln)
    p = oracle.set_board(START_FEN)
    for board, move, dist in zip(g["boards"], g["moves"], g["dists"]):
        assert board == oracle.board_json(p) and oracle.result(p) == 0
        legal = {oracle.move_string(m): m for m in oracle.movegen(p)}
        assert move in dist and set(dist) <= set(legal) and abs(sum(dist.values()) - 1) < 1e-9
        p = oracle.makemove(p, legal[move])
        plies += 1
    assert oracle.result(p) == g["result"]
    games += 1
    lengths.append(len(g["moves"]))
print("replayed %d finished games, %d plies, all legal; mean length %.1f (min %d max %d)" % (games, plies, sum(lengths) / max(games, 1), min(lengths or [0]), max(lengths or [0])))

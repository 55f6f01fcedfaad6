"""Speculative leaf evaluation in search pools (az_pool_config::speculate; engine.py:387-392 queues likely children the same
way): a cached evaluation changes WHEN a leaf is linked, never what the search computes -- visit counts, edge scores and
priors must equal the one-leaf-per-tick search bit for bit, in fewer ticks."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _search(ctx, fens, visits, speculate, mode=None, plays=()):
    from ataxxzero_b200 import rules, search
    mode = search.EVAL_BF16 if mode is None else mode
    with search.Pool(ctx, len(fens), visits, eval_mode=mode, speculate=speculate, steps_per_tick=64) as pool:
        for i, f in enumerate(fens):
            pool.set_root(i, rules.set_board(f))
        assert pool.run()
        for mv in plays:                       # re-rooting keeps the cache: transpositions of the old tree stay evaluated
            for i in range(len(fens)):
                r = pool.root(i)
                if r["moves"]:
                    pool.play(i, r["moves"][int(np.argmax(r["visits"]))])
            assert pool.run()
        roots = [pool.root(i) for i in range(len(fens))]
        return roots, pool.stats()


def _same(a, b):
    assert a["visits"] == b["visits"]
    assert [float(x).hex() for x in a["total_score"]] == [float(x).hex() for x in b["total_score"]]
    assert [float(x).hex() for x in a["prior"]] == [float(x).hex() for x in b["prior"]]
    assert a["root_visits"] == b["root_visits"] and float(a["value"]).hex() == float(b["value"]).hex()


@pytest.mark.parametrize("weights", ["init", "trained"])
def test_speculative_search_is_bit_identical(ctx, weights):
    from ataxxzero_b200 import model, net, rules
    from oracle import net_numpy
    if weights == "init":
        net.load_weights(ctx, model.Network.random_init(seed=0))
    else:
        net.load_weights(ctx, model.Network(*net_numpy.trained_scale_weights(seed=0)))
    fen = load_golden("mcts_golden.json")["midgame_fen"]
    fens = [rules.START_FEN, rules.OPEN_FEN, fen]
    base, st0 = _search(ctx, fens, 5000, 0)
    for k in (1, 4, 16):
        got, st = _search(ctx, fens, 5000, k)
        for a, b in zip(got, base):
            _same(a, b)
        assert st["steps"] == st0["steps"]
        assert st["ticks"] < st0["ticks"], (k, st["ticks"], st0["ticks"])


def test_speculative_search_with_reroot_and_f16(ctx):
    from ataxxzero_b200 import model, net, rules, search
    net.load_weights(ctx, model.Network.random_init(seed=5))
    fen = load_golden("mcts_golden.json")["midgame_fen"]
    for mode in (search.EVAL_BF16, search.EVAL_F16):
        base, _ = _search(ctx, [fen, rules.START_FEN], 800, 0, mode, plays=(0, 1, 2))
        got, _ = _search(ctx, [fen, rules.START_FEN], 800, 4, mode, plays=(0, 1, 2))
        for a, b in zip(got, base):
            _same(a, b)


def test_speculative_100k_visits(ctx):
    """BASELINE configs[4]: 100 000 visits from the midgame position, speculative against one leaf per tick"""
    from ataxxzero_b200 import model, net
    net.load_weights(ctx, model.Network.random_init(seed=0))
    fen = load_golden("mcts_golden.json")["midgame_fen"]
    base, st0 = _search(ctx, [fen], 100000, 0)
    got, st = _search(ctx, [fen], 100000, 4)
    _same(got[0], base[0])
    assert st["ticks"] * 2 < st0["ticks"], (st["ticks"], st0["ticks"])


def test_many_small_trees_share_the_batch(ctx):
    """64 trees speculating at once: the request batch is bounded, blocked leaves always get their slot"""
    from ataxxzero_b200 import model, net, rules
    net.load_weights(ctx, model.Network.random_init(seed=0))
    fens = [rules.START_FEN] * 32 + [rules.OPEN_FEN] * 32
    base, _ = _search(ctx, fens, 300, 0)
    got, _ = _search(ctx, fens, 300, 8)
    for a, b in zip(got, base):
        _same(a, b)

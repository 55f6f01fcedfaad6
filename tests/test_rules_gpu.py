"""GPU parity of the batched rule kernels (through the C ABI) against the oracle and the
reference-generated goldens: bit-exact move lists (order included), makemove, adjudication,
features and dilations."""
import random

import numpy as np
import pytest

from conftest import golden_position, load_golden

pytestmark = pytest.mark.gpu


def _positions_from_golden(rules_mod):
    from ataxxzero_b200 import Position
    g = load_golden("rules_golden.json")
    return g, [golden_position(Position, e) for e in g["positions"]]


def test_movegen_bit_exact_vs_golden(ctx):
    from ataxxzero_b200 import rules
    g, pos = _positions_from_golden(rules)
    lists = rules.movegen_batch(ctx, pos)
    for e, mv in zip(g["positions"], lists):
        assert [rules.pack_move(m) for m in mv] == e["moves"]


def test_result_features_boards_vs_golden(ctx, oracle):
    from ataxxzero_b200 import rules
    from oracle.cpu import Position as OPos
    g, pos = _positions_from_golden(rules)
    res = rules.result_batch(ctx, pos)
    assert res.tolist() == [e["result"] for e in g["positions"]]
    feats = rules.features_batch(ctx, pos)
    for e, f in zip(g["positions"][::5], feats[::5]):
        want = oracle.features(golden_position(OPos, e))
        assert np.array_equal(f, want)


def test_makemove_vs_golden(ctx):
    from ataxxzero_b200 import rules
    g, pos = _positions_from_golden(rules)
    sel = [(p, e) for p, e in zip(pos, g["positions"]) if "after" in e]
    out = rules.makemove_batch(ctx, [p for p, _ in sel], [rules.unpack_move(e["after"]["move"]) for _, e in sel])
    for rec, (_, e) in zip(out, sel):
        a = e["after"]["pos"]
        assert (int(rec["turn"]), int(rec["ply"]), int(rec["blockers"]), int(rec["pieces"][0]), int(rec["pieces"][1])) == \
            (a["turn"], a["ply"], a["blockers"], a["x"], a["o"])


def test_dilations_random(ctx, oracle):
    from ataxxzero_b200 import rules
    rng = random.Random(5)
    bbs = [0, (1 << 49) - 1, 1, 1 << 48] + [rng.getrandbits(49) & rng.getrandbits(49) for _ in range(4000)]
    s, d = rules.jump_bb_batch(ctx, bbs)
    for bb, a, b in zip(bbs, s.tolist(), d.tolist()):
        assert a == oracle.single_jump_bb(bb) and b == oracle.double_jump_bb(bb)


def test_random_playouts_vs_oracle_large(ctx, oracle):
    """>= 10^4 positions incl. blockers and both sides to move, played on the GPU:
    the device's makemove drives the games, the oracle checks every move list."""
    from ataxxzero_b200 import rules
    from oracle.cpu import Position as OPos
    rng = random.Random(11)
    fens = [rules.START_FEN, rules.OPEN_FEN, "x5o/7/2-1-2/7/2-1-2/7/o5x x", "x5o/7/7/7/7/7/o5x o"]
    games = [rules.set_board(fens[i % 4]) for i in range(128)]
    arr = rules.positions_array(games)
    total = 0
    for ply in range(120):
        lists = rules.movegen_batch(ctx, arr)
        res = rules.result_batch(ctx, arr)
        live, moves = [], []
        for i, (rec, mv) in enumerate(zip(arr, lists)):
            op = OPos()
            op.ply, op.turn, op.blockers = int(rec["ply"]), int(rec["turn"]), int(rec["blockers"])
            op.pieces[0], op.pieces[1] = int(rec["pieces"][0]), int(rec["pieces"][1])
            assert mv == oracle.movegen(op)
            assert int(res[i]) == oracle.result(op)
            total += 1
            if mv and res[i] == 0:
                live.append(i)
                moves.append(rng.choice(mv))
        if not live:
            break
        sub = rules.makemove_batch(ctx, arr[live], moves)
        for k, i in enumerate(live):
            op = OPos()
            rec = arr[i]
            op.ply, op.turn, op.blockers = int(rec["ply"]), int(rec["turn"]), int(rec["blockers"])
            op.pieces[0], op.pieces[1] = int(rec["pieces"][0]), int(rec["pieces"][1])
            want = oracle.makemove(op, moves[k])
            assert (int(sub[k]["turn"]), int(sub[k]["pieces"][0]), int(sub[k]["pieces"][1]), int(sub[k]["ply"])) == \
                (want.turn, want.pieces[0], want.pieces[1], want.ply)
        arr = np.concatenate([sub, np.delete(arr, live)]) if len(live) < len(arr) else sub
        arr = arr[[r == 0 for r in rules.result_batch(ctx, arr)]] if len(arr) else arr
        if len(arr) == 0:
            break
    assert total >= 10000


def test_empty_batches(ctx):
    from ataxxzero_b200 import rules
    assert rules.movegen_batch(ctx, []) == []
    assert len(rules.result_batch(ctx, [])) == 0
    assert len(rules.perft_batch(ctx, [], 3)) == 0


def test_random_playouts_on_device(ctx, oracle):
    """az_random_playouts: every game replays legally through the oracle (boards, moves, adjudication), the games differ,
    the same seed reproduces them, and the length statistics look like the reference's random play (SURVEY App. C-4:
    mean ~185 plies)."""
    import numpy as np
    from ataxxzero_b200 import rules
    from oracle.cpu import OPEN_FEN, START_FEN
    for fen in (OPEN_FEN, START_FEN):
        start = rules.set_board(fen)
        plies, n_plies, result = rules.random_playouts(ctx, start, 600, 400, seed=5)
        again = rules.random_playouts(ctx, start, 600, 400, seed=5)
        assert np.array_equal(plies["move"], again[0]["move"]) and np.array_equal(n_plies, again[1])
        other = rules.random_playouts(ctx, start, 600, 400, seed=6)
        assert not np.array_equal(n_plies, other[1])
        for g in range(0, 600, 3):
            p = oracle.set_board(fen)
            for k in range(int(n_plies[g])):
                assert oracle.result(p) == 0
                assert (int(plies["x"][g, k]), int(plies["o"][g, k])) == (p.pieces[0], p.pieces[1])
                m = int(plies["move"][g, k])
                mv = (m & 0xff, m >> 8)
                assert mv in oracle.movegen(p)
                p = oracle.makemove(p, mv)
            assert oracle.result(p) == int(result[g])
            assert result[g] != 0 or n_plies[g] == 400
        assert 140 < n_plies[result != 0].mean() < 240
        first = np.bincount(plies["move"][:, 0], minlength=1)            # 16 legal first moves, each drawn ~1/16 of the time
        assert (first > 0).sum() == 16 and first[first > 0].min() > 600 / 16 * 0.4
    assert len(rules.random_playouts(ctx, rules.set_board(OPEN_FEN), 0)[1]) == 0

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


EDGE_FENS = [
    "7/7/7/3x3/7/7/7 x",                    # o has no pieces at all: x wins (x to move, 24 moves available)
    "7/7/7/3x3/7/7/7 o",                    # ... also when the piece-less side is to move
    "7/7/7/3o3/7/7/7 x",                    # the side to move has no pieces
    "xxxxxxx/xxxxxxx/xxxxxxx/xxxxxxx/xxxxxxx/xxxxxxx/xxxxxx1 o",      # o has no pieces, one empty cell
    "xxxxxxx/xxxxxxx/xxxxxxx/ooooooo/ooooooo/ooooooo/oooooo1 x",      # x cannot reach the last empty cell: stuck, o is credited
    "xxxxxxx/xxxxxxx/xxxxxxx/xxxoooo/ooooooo/ooooooo/ooooooo x",      # full board: piece count decides (24 : 25)
    "xxxxxxx/xxxxxxx/xxxxxxx/xxxxooo/ooooooo/ooooooo/ooooooo o",      # full board: 25 : 24
    "xoo4/ooo4/ooo4/7/7/7/7 x",             # x walled in by enemy pieces: stuck, the opponent is credited every empty cell
    "x-1o3/--5/7/7/7/7/7 x",                # blockers next to x: only jumps remain
    "x1x1x1x/7/x1x1x1x/7/x1x1x1x/7/x1x1x1o x",   # very many moves for x (jumps + clones)
    "7/7/2ooo2/2oxo2/2ooo2/7/7 x",          # only jumps available to x
    "x5o/7/3-3/2-1-2/3-3/7/o5x o",          # the C++ start position with o to move
    "-------/-------/-------/--xx---/-------/-------/------o x",     # blockers everywhere: nobody can move, counts decide
]


def test_edge_positions_vs_oracle_and_reference(ctx, oracle):
    """Adjudication and move generation on hand-made corner cases (zero pieces, stuck sides, full boards, walled-in
    pieces, maximum fan-out, jump-only positions) against the oracle and, when present, the compiled reference."""
    from ataxxzero_b200 import rules
    from oracle.cpu import Reference
    ref = Reference() if Reference.available() else None
    positions = [rules.set_board(f) for f in EDGE_FENS]
    got_moves = rules.movegen_batch(ctx, positions)
    got_res = rules.result_batch(ctx, positions)
    feats = rules.features_batch(ctx, positions)
    fan = 0
    for fen, p, mv, res, f in zip(EDGE_FENS, positions, got_moves, got_res, feats):
        op = oracle.set_board(fen)
        assert mv == oracle.movegen(op), fen
        assert int(res) == oracle.result(op), fen
        assert np.array_equal(f.reshape(-1), np.asarray(oracle.features(op), dtype=np.float32).reshape(-1)), fen
        if ref is not None:
            assert mv == ref.movegen(op) and int(res) == ref.result(op), fen
        fan = max(fan, len(mv))
        if mv and int(res) == 0:            # every move applies identically
            nxt = rules.makemove_batch(ctx, [p] * len(mv), mv)
            for q, m in zip(rules.array_to_positions(nxt), mv):
                assert q.key() == oracle.makemove(op, m).key(), (fen, m)
    assert fan >= 100
    assert [int(r) for r in got_res[:8]] == [1, 1, 2, 1, 2, 2, 1, 2]

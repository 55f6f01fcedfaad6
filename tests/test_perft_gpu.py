"""GPU perft through the C ABI: bit-exact against SURVEY App. C / perft_ref goldens, the
batched workload against the oracle, plus size-independent properties at full depth."""
import random

import numpy as np
import pytest

from conftest import golden_position, load_golden

pytestmark = pytest.mark.gpu


def test_perft_tables_depth_1_to_7(ctx):
    from ataxxzero_b200 import rules
    g = load_golden("perft_golden.json")
    for fen, table in g["perft_ref"].items():
        p = rules.set_board(fen)
        for depth, want in enumerate(table, start=1):
            assert rules.perft(ctx, p, depth) == want, (fen, depth)
    assert rules.perft(ctx, rules.set_board(rules.START_FEN), 0) == 1


def test_perft_depth_8_survey_goldens(ctx):
    from ataxxzero_b200 import rules
    g = load_golden("perft_golden.json")
    for fen, table in g["survey_app_c"].items():
        assert rules.perft(ctx, rules.set_board(fen), 8) == table[7], fen
        stats = rules.perft_last_stats(ctx)
        assert stats["count_nodes"] == table[6]          # parents counted at the last ply = perft(7)


def test_perft_batch_vs_golden_and_oracle(ctx, oracle):
    from ataxxzero_b200 import Position, rules
    from oracle.cpu import Position as OPos
    g = load_golden("perft_golden.json")
    pos = [golden_position(Position, e["pos"]) for e in g["batch"]]
    assert rules.perft_batch(ctx, pos, 2).tolist() == [e["depth2"] for e in g["batch"]]
    assert rules.perft_batch(ctx, pos, 3).tolist() == [e["depth3"] for e in g["batch"]]
    # ragged: positions with no moves / terminal positions mixed in
    r = load_golden("rules_golden.json")
    mixed = [golden_position(Position, e) for e in r["positions"][::3]]
    want = [oracle.perft(golden_position(OPos, e), 3) for e in r["positions"][::3]]
    assert rules.perft_batch(ctx, mixed, 3).tolist() == want
    assert 0 in want or any(len(e["moves"]) == 0 for e in r["positions"][::3])


def test_perft_additivity_property(ctx):
    """perft(p, d) == sum over children perft(child, d-1), at a depth the CPU could not check quickly."""
    from ataxxzero_b200 import rules
    p = rules.set_board(rules.START_FEN)
    moves = rules.movegen_batch(ctx, [p])[0]
    children = rules.makemove_batch(ctx, [p] * len(moves), moves)
    assert int(rules.perft_batch(ctx, children, 6).sum()) == rules.perft(ctx, p, 7) == 3044225260


def test_perft_16k_batch_consistency(ctx, oracle):
    """The BASELINE batched workload: 16384 positions from seeded random playouts, depth 3.
    Checked exactly on a 256-position sample by the oracle, and as a whole by additivity."""
    from ataxxzero_b200 import rules
    from oracle.cpu import Position as OPos
    rng = random.Random(0)
    start = rules.set_board(rules.START_FEN)
    arr = rules.positions_array([start] * 16384)
    plies = np.array([rng.randrange(0, 120) for _ in range(16384)])
    for ply in range(120):
        res = rules.result_batch(ctx, arr)
        lists = rules.movegen_batch(ctx, arr)
        idx = [i for i in range(16384) if plies[i] > ply and res[i] == 0 and lists[i]]
        if not idx:
            break
        sub = rules.makemove_batch(ctx, arr[idx], [rng.choice(lists[i]) for i in idx])
        arr[idx] = sub
    d3 = rules.perft_batch(ctx, arr, 3)
    d2 = rules.perft_batch(ctx, arr, 2)
    d1 = rules.perft_batch(ctx, arr, 1)
    lists = rules.movegen_batch(ctx, arr)
    assert d1.tolist() == [len(m) for m in lists]
    for i in range(0, 16384, 64):
        rec = arr[i]
        op = OPos()
        op.turn, op.blockers = int(rec["turn"]), int(rec["blockers"])
        op.pieces[0], op.pieces[1] = int(rec["pieces"][0]), int(rec["pieces"][1])
        assert int(d3[i]) == oracle.perft(op, 3)
        assert int(d2[i]) == oracle.perft(op, 2)
    # additivity over the whole batch for a slice of roots
    for i in range(0, 16384, 2048):
        kids = rules.makemove_batch(ctx, arr[[i] * len(lists[i])], lists[i]) if lists[i] else []
        assert int(rules.perft_batch(ctx, kids, 2).sum()) == int(d3[i])

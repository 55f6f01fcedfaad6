"""Host-side twins of the reference's Python interfaces (no GPU): the bitboard AtaxxState against reference-generated
games (tests/golden/python_rules_golden.json, made by importing /root/reference/ataxx_rules.py) and against the C
oracle incl. blockers; the UAI codec vectors (uai_interface.py:34-39); engine.py's feature / policy-plane helpers
against the oracle's encoding (self_play_client.cpp:77-86,174-202)."""
import random

import numpy as np
import pytest

from conftest import load_golden

START_FEN = "x5o/7/3-3/2-1-2/3-3/7/o5x x"


def _tuple_move(move):
    return ("c", tuple(move[1])) if move[0] == "c" else (tuple(move[0]), tuple(move[1]))


def _compact(m):
    return "c%d%d" % m[1] if m[0] == "c" else "%d%d%d%d" % (m[0] + m[1])


def test_state_follows_reference_python_games():
    from ataxxzero_b200 import ataxx_rules as ar
    g = load_golden("python_rules_golden.json")
    for game in g["games"]:
        s = ar.AtaxxState.initial()
        for board, move, legal, fen in zip(game["boards"], game["moves"], game["legal_sorted"], game["fens"]):
            assert list(s.board) == board and s.result() is None
            assert sorted(_compact(m) for m in s.legal_moves()) == legal
            assert s.fen() == fen and ar.AtaxxState.from_fen(fen) == s
            s.move(_tuple_move(move))
        assert list(s.board) == game["final_board"] and s.result() == game["result"]


def test_state_perft_matches_reference():
    from ataxxzero_b200 import ataxx_rules as ar

    def perft(b, d):
        if d == 0:
            return 1
        total = 0
        for m in b.legal_moves():
            if m != "pass":
                c = b.copy()
                c.move(m)
                total += perft(c, d - 1)
        return total
    assert [perft(ar.AtaxxState.initial(), d) for d in (1, 2, 3)] == [16, 256, 6460]
    assert [perft(ar.AtaxxState.from_fen(START_FEN), d) for d in (1, 2, 3)] == [16, 256, 5948]


def test_state_agrees_with_oracle_with_blockers(oracle):
    """Random playouts from the C++ start position (4 blockers): same move lists IN THE SAME ORDER, same boards, results."""
    from ataxxzero_b200 import ataxx_rules as ar, rules
    rng = random.Random(11)
    for _ in range(12):
        s, p = ar.AtaxxState.from_fen(START_FEN), oracle.set_board(START_FEN)
        while True:
            want = oracle.movegen(p)
            got = [m for m in s.legal_moves() if m != "pass"]
            assert [rules.from_reference_move(m) for m in got] == want
            assert oracle.board_json(p) == list(s.board)
            assert (oracle.result(p) or None) == s.result()
            q = s.to_position()
            assert (q.turn, q.blockers, q.pieces[0], q.pieces[1]) == p.key()
            if not want or oracle.result(p):
                break
            k = rng.randrange(len(want))
            s.move(got[k])
            p = oracle.makemove(p, want[k])


def test_state_item_access_and_blocked():
    from ataxxzero_b200 import ataxx_rules as ar
    s = ar.AtaxxState.from_fen(START_FEN)
    assert s[0, 0] == 1 and s[6, 0] == 2 and s[0, 6] == 2 and s[6, 6] == 1 and s.to_move == 1
    assert s.blocked == frozenset({(3, 2), (2, 3), (4, 3), (3, 4)})            # d5, c4, e4, d3 (SURVEY C-2)
    s[1, 1] = 2
    assert s[1, 1] == 2 and s.board[1 + 7 * 1] == 2
    with pytest.raises(AssertionError):
        s.move(("c", (3, 2)))                                                   # a blocker is not an empty cell
    with pytest.raises(ValueError):
        ar.AtaxxState.from_fen("x5o/7/7/7/7/o5x x")


def test_uai_codec_vectors():
    from ataxxzero_b200.cli import uai_interface as u
    for text in ["f2", "c3d5"]:
        assert u.uai_encode_move(u.uai_decode_move(text)) == text
    for m in [("c", (4, 3)), ((4, 3), (2, 5))]:
        assert u.uai_decode_move(u.uai_encode_move(m)) == m
    assert u.uai_decode_move("f2") == ("c", (5, 5)) and u.uai_decode_move("c3d5") == ((2, 4), (3, 2))   # SURVEY C-2
    assert u.uai_encode_move(("c", (4, 3))) == "e4" and u.uai_encode_move(((4, 3), (2, 5))) == "e4c2"
    assert u.uai_encode_move("pass") == "0000" and u.uai_decode_move("0000") == "pass"


def test_engine_feature_and_policy_helpers_match_oracle(oracle):
    from ataxxzero_b200 import ataxx_rules as ar, engine, rules
    rng = random.Random(2)
    p, s = oracle.set_board(START_FEN), ar.AtaxxState.from_fen(START_FEN)
    for _ in range(40):
        f = engine.board_to_features(s)
        assert f.dtype == np.int8 and f.shape == (7, 7, 4)
        assert np.array_equal(f.astype(np.float32).reshape(-1), np.asarray(oracle.features(p), dtype=np.float32).reshape(-1))
        moves = oracle.movegen(p)
        if not moves or oracle.result(p):
            break
        logits = np.asarray([rng.uniform(-2, 2) for _ in range(833)], dtype=np.float32)
        pri = np.asarray(oracle.priors(logits, moves))
        sm = engine.softmax(logits.astype(np.float64)).reshape(7, 7, 17)
        mine = np.array([engine.get_move_score(sm, rules.to_reference_move(m)) for m in moves])
        assert np.allclose(mine / mine.sum(), pri, rtol=1e-9, atol=1e-12)         # same plane for every move (App. A-2)
        heat = engine.encode_move_as_heatmap(rules.to_reference_move(moves[0]))
        assert heat.sum() == 1 and heat.shape == (7, 7, 17)
        k = rng.randrange(len(moves))
        s.move(rules.to_reference_move(moves[k]))
        p = oracle.makemove(p, moves[k])
    assert abs(sum(engine.add_dirichlet_noise_to_posterior({"a": 0.5, "b": 0.5}, 0.15, 0.25).values()) - 1) < 1e-9
    assert engine.sample_by_weight({"x": 1.0}) == "x"


def test_looper_paths_and_counts(tmp_path):
    from ataxxzero_b200.cli import looper
    args = looper.build_parser().parse_args(["--prefix", str(tmp_path), "--gpus", "2"])
    args.processes = args.parallel_games_processes or args.gpus
    assert looper.index_to_model_path(args, 7).endswith("models/model-007.npy")
    paths = looper.index_to_games_paths(args, 1)
    assert [p.split("/")[-1] for p in paths] == ["model-001-0.json", "model-001-1.json"]      # looper.py:72-74
    (tmp_path / "games").mkdir()
    open(paths[0], "w").write("{}\n\n{}\n")
    assert looper.count_games(paths) == 2


def test_generate_games_supervised_with_a_scripted_uai_engine(tmp_path):
    """--supervised CMD (generate_games.py:27-36,98-99): games recorded from an external UAI engine; here a scripted
    engine that answers `go movetime` with a random legal move.  No GPU involved."""
    import json
    import sys
    from conftest import ROOT
    from ataxxzero_b200 import ataxx_rules as ar
    from ataxxzero_b200.cli import generate_games
    script = tmp_path / "fake_uai.py"
    script.write_text(
        "import sys, random\n"
        "sys.path.insert(0, %r)\n"
        "from ataxxzero_b200 import ataxx_rules\n"
        "from ataxxzero_b200.cli.uai_interface import uai_encode_move\n"
        "board, rng = ataxx_rules.AtaxxState.initial(), random.Random(1)\n"
        "for line in sys.stdin:\n"
        "    line = line.strip()\n"
        "    if line == 'quit': break\n"
        "    if line == 'uai': print('uaiok')\n"
        "    elif line == 'isready': print('readyok')\n"
        "    elif line.startswith('position fen '): board = ataxx_rules.AtaxxState.from_fen(line[13:])\n"
        "    elif line.startswith('go '): print('bestmove ' + uai_encode_move(rng.choice(board.legal_moves())))\n"
        "    sys.stdout.flush()\n" % ROOT)
    out = tmp_path / "sup.json"
    n = generate_games.main(["--supervised", "%s %s" % (sys.executable, script), "--supervised-ms", "1", "--game-count", "2",
                             "--output-games", str(out)])
    assert n == 2
    for line in out.read_text().splitlines():
        game = json.loads(line)
        board = ar.AtaxxState.initial()
        for cells, move in zip(game["boards"], game["moves"]):
            assert list(board.board) == cells
            mv = ("c", tuple(move[1])) if move[0] == "c" else (tuple(move[0]), tuple(move[1]))
            assert mv in board.legal_moves()
            board.move(mv)
        assert board.result() == game["result"]


def test_uai_ringmaster_between_scripted_engines(tmp_path):
    """uai_ringmaster.py flags and flow (pairings both ways round, PGN append, win tally) with two scripted engines."""
    import sys
    from conftest import ROOT
    from ataxxzero_b200.cli import uai_ringmaster
    script = tmp_path / "fake_uai.py"
    script.write_text(
        "import sys, random\n"
        "sys.path.insert(0, %r)\n"
        "from ataxxzero_b200 import ataxx_rules\n"
        "from ataxxzero_b200.cli.uai_interface import uai_encode_move\n"
        "board, rng = ataxx_rules.AtaxxState.initial(), random.Random(int(sys.argv[1]))\n"
        "for line in sys.stdin:\n"
        "    line = line.strip()\n"
        "    if line == 'quit': break\n"
        "    if line.startswith('position fen '): board = ataxx_rules.AtaxxState.from_fen(line[13:])\n"
        "    elif line.startswith('go '): print('bestmove ' + uai_encode_move(rng.choice(board.legal_moves())))\n"
        "    sys.stdout.flush()\n" % ROOT)
    pgn = tmp_path / "games.pgn"
    wins = uai_ringmaster.main(["--engine", "%s %s 1" % (sys.executable, script), "--engine", "%s %s 2" % (sys.executable, script),
                                "--tc", "0.001", "--games", "2", "--pgn-out", str(pgn), "--opening", "f2, a2"])
    assert sum(wins.values()) == 2
    text = pgn.read_text()
    assert text.count('[Result "') == 2 and text.count('[Opening "f2, a2"]') == 2 and '[TimeControl "+0.001"]' in text
    first_line = [l for l in text.splitlines() if l and not l.startswith("[")][0]
    assert first_line.startswith("f2 a2 ")

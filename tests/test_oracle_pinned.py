"""Pins the CPU oracle (oracle/ataxx_oracle.c) against (1) the golden vectors the compiled
reference produced (tests/golden/*.json) and (2) the compiled reference
itself when oracle/_ref is present.  CPU only."""
import random

import numpy as np
import pytest

from conftest import golden_position, load_golden
from oracle.cpu import OPEN_FEN, START_FEN, Position


def unpack(m):
    return (m & 0xff, m >> 8)


def test_ring_tables_and_dilations(oracle):
    g = load_golden("rules_golden.json")
    for sq in range(49):
        assert oracle.single_ring(sq) == g["rings"]["single"][sq]
        assert oracle.double_ring(sq) == g["rings"]["double"][sq]
    for e in g["positions"]:
        own = e["x"] if e["turn"] == 0 else e["o"]
        assert oracle.single_jump_bb(own) == e["single_jump_own"]
        assert oracle.double_jump_bb(own) == e["double_jump_own"]


def test_fen_parser_status_codes(oracle):
    g = load_golden("rules_golden.json")
    for fen, want in g["fens"].items():
        assert oracle.set_board_rc(fen) == want["status"], fen
        if want["status"] == 0:
            p = oracle.set_board(fen)
            assert (p.turn, p.blockers, p.pieces[0], p.pieces[1]) == \
                (want["pos"]["turn"], want["pos"]["blockers"], want["pos"]["x"], want["pos"]["o"])


def test_movegen_makemove_result_against_golden(oracle):
    g = load_golden("rules_golden.json")
    assert len(g["positions"]) > 300
    saw_terminal = saw_big = False
    for e in g["positions"]:
        p = golden_position(Position, e)
        moves = oracle.movegen(p)
        assert [m[0] | (m[1] << 8) for m in moves] == e["moves"]          # exact order
        assert oracle.result(p) == e["result"]
        assert oracle.board_json(p) == e["board_json"]
        saw_terminal |= e["result"] != 0
        saw_big |= len(moves) > 100
        if "after" in e:
            m = unpack(e["after"]["move"])
            q = oracle.makemove(p, m)
            a = e["after"]["pos"]
            assert (q.turn, q.ply, q.blockers, q.pieces[0], q.pieces[1]) == (a["turn"], a["ply"], a["blockers"], a["x"], a["o"])
            assert oracle.move_string(m) == e["after"]["string"]
            assert oracle.legal_move(p, m)
    assert saw_terminal and saw_big


def test_encoding_golden_c2(oracle):
    """plane sums and non-zero cells at the C++ start position."""
    p = oracle.set_board(START_FEN)
    f = oracle.features(p)
    assert f.sum(axis=(0, 1)).tolist() == [49, 2, 2, 4]
    assert sorted(zip(*np.nonzero(f[..., 1]))) == [(0, 0), (6, 6)]
    assert sorted(zip(*np.nonzero(f[..., 2]))) == [(0, 6), (6, 0)]
    assert sorted(zip(*np.nonzero(f[..., 3]))) == [(2, 3), (3, 2), (3, 4), (4, 3)]
    assert oracle.board_json(p) == [1, 0, 0, 0, 0, 0, 2] + [0] * 35 + [2, 0, 0, 0, 0, 0, 1]
    order = "g1e1 g1e2 g1e3 g1f3 g1g3 a7a5 a7b5 a7c5 a7c6 a7c7 f1 f2 g2 a6 b6 b7".split()
    assert [oracle.move_string(m) for m in oracle.movegen(p)] == order


def test_populate_features_and_priors_against_golden(oracle):
    g = load_golden("populate_golden.json")
    n_term = 0
    for e in g["entries"]:
        p = golden_position(Position, e["pos"])
        if "terminal_value" in e:
            n_term += 1
            res = oracle.result(p)
            assert res != 0
            v = 1.0 if res == 1 else -1.0
            assert (v if p.turn == 0 else -v) == e["terminal_value"]
            continue
        feats = oracle.features(p)
        assert [int(i) for i in feats.reshape(-1).nonzero()[0]] == e["features_nonzero"]
        logits, values = oracle.probe_eval(feats)
        moves = [unpack(m) for m in e["moves"]]
        pri = oracle.priors(logits[0], moves)
        assert [float(x).hex() for x in pri] == e["priors_hex"]            # bit-exact doubles
        assert float(values[0]) == e["value"]
    assert n_term >= 1


def test_perft_goldens(oracle):
    g = load_golden("perft_golden.json")
    for fen, table in g["survey_app_c"].items():
        p = oracle.set_board(fen)
        for depth in range(1, 6):
            assert oracle.perft(p, depth) == table[depth - 1], (fen, depth)
    for fen, table in g["perft_ref"].items():
        assert table[:len(g["survey_app_c"].get(fen, table)[:7])] == g["survey_app_c"].get(fen, table)[:7]
        p = oracle.set_board(fen)
        assert oracle.perft(p, 6, threads=8) == table[5]
    for e in g["batch"]:
        p = golden_position(Position, e["pos"])
        assert oracle.perft(p, 2) == e["depth2"]
        assert oracle.perft(p, 3) == e["depth3"]


def test_policy_plane_table(oracle):
    """SURVEY A-2: plane index per (dx, dy) and flat index 119*to_x + 17*to_y + plane."""
    planes = {(-2, -2): 0, (-2, -1): 1, (-2, 0): 2, (-2, 1): 3, (-2, 2): 4, (-1, -2): 5, (-1, 2): 6, (0, -2): 7,
              (0, 2): 8, (1, -2): 9, (1, 2): 10, (2, -2): 11, (2, -1): 12, (2, 0): 13, (2, 1): 14, (2, 2): 15}
    for frm in range(49):
        fx, fy = frm % 7, 6 - frm // 7
        for to in range(49):
            tx, ty = to % 7, 6 - to // 7
            d = (tx - fx, ty - fy)
            if frm == to:
                assert oracle.policy_index(frm, to) == 119 * tx + 17 * ty + 16
            elif d in planes:
                assert oracle.policy_index(frm, to) == 119 * tx + 17 * ty + planes[d]


def test_umap_iteration_order_golden(oracle):
    g = load_golden("mcts_golden.json")
    assert len(g["umap"]) > 50
    for e in g["umap"]:
        moves = [unpack(m) for m in e["moves"]]
        fresh, buckets = oracle.umap_order(moves)
        assert fresh == e["fresh"] and buckets == e["buckets"]
        again, _ = oracle.umap_order(moves, buckets)
        assert again == e["reinserted"]


def test_survey_b4_orders(oracle):
    p = oracle.set_board(START_FEN)
    moves = oracle.movegen(p)
    fresh, b = oracle.umap_order(moves)
    names = [oracle.move_string(moves[i]) for i in fresh]
    assert names == "b6 a6 a7b5 b7 a7c5 g1f3 g1e2 a7c7 g1g3 a7a5 g1e1 a7c6 g1e3 f1 f2 g2".split()
    again, _ = oracle.umap_order(moves, b)
    names = [oracle.move_string(moves[i]) for i in again]
    assert names == "b6 a6 g2 f2 b7 a7c5 a7b5 a7a5 g1g3 g1f3 f1 g1e3 a7c7 g1e2 a7c6 g1e1".split()


def test_mcts_search_goldens(oracle):
    g = load_golden("mcts_golden.json")
    for s in g["searches"]:
        if s["visits"] > 1000:
            continue
        tree = oracle.tree(oracle.set_board(s["fen"]), s["evaluator"])
        tree.search(s["visits"])
        d = tree.dist()
        assert [m[0] | (m[1] << 8) for m, v, w, p in d] == s["moves"]
        assert [v for m, v, w, p in d] == s["edge_visits"], (s["evaluator"], s["fen"], s["visits"])
        assert [float(w).hex() for m, v, w, p in d] == s["edge_total_hex"]
        assert tree.evals == s["evals"]
        tree.close()


def test_mcts_survey_c3(oracle):
    dist, evals = oracle.search(START_FEN, 400, "probe")
    got = {oracle.move_string(m): v for m, v in dist if v}
    assert got == {"a7a5": 46, "a7c5": 14, "a7c7": 127, "b7": 3, "f1": 7, "f2": 78, "g1e1": 13, "g1e2": 13,
                   "g1e3": 42, "g1g3": 57}
    assert evals == 401
    dist, evals = oracle.search(START_FEN, 800, "probe")
    got = {oracle.move_string(m): v for m, v in dist if v}
    assert got == {"a7a5": 192, "a7b5": 2, "a7c5": 17, "a7c7": 345, "b7": 3, "f1": 11, "f2": 86, "g1e1": 13,
                   "g1e2": 13, "g1e3": 57, "g1g3": 61}
    assert evals == 801


def test_mcts_game_goldens_with_tree_reuse(oracle):
    g = load_golden("mcts_golden.json")
    for game in g["games"]:
        if game["visits"] > 100:
            continue
        tree = oracle.tree(oracle.set_board(game["fen"]), game["evaluator"])
        for ply in game["plies"]:
            tree.search(game["visits"])
            d = tree.dist()
            assert [v for m, v, w, p in d] == ply["visits"]
            tree.play(unpack(ply["played"]))
        assert oracle.result(tree.root_position()) == game["result"]
        assert tree.evals == game["evals"]
        if game["evaluator"] == "uniform":
            assert tree.ties > 0          # the tie path (libstdc++ order) really was exercised
        tree.close()


def test_python_twin_agrees_without_blockers(oracle):
    """ataxx_rules.py (no blockers) and the C++ rules agree on legal moves,
    move application and results along reference-generated random games."""
    g = load_golden("python_rules_golden.json")
    assert g["perft"] == [16, 256, 6460, 155888]

    def compact(m):
        (fx, fy), (tx, ty) = (m[0] % 7, 6 - m[0] // 7), (m[1] % 7, 6 - m[1] // 7)
        return "c%d%d" % (tx, ty) if m[0] == m[1] else "%d%d%d%d" % (fx, fy, tx, ty)

    for game in g["games"]:
        p = oracle.set_board(OPEN_FEN)
        for board, move, legal in zip(game["boards"], game["moves"], game["legal_sorted"]):
            assert oracle.board_json(p) == board
            assert oracle.result(p) == 0
            assert sorted(compact(m) for m in oracle.movegen(p)) == legal
            if move[0] == "c":
                sq = move[1][0] + 7 * (6 - move[1][1])
                mv = (sq, sq)
            else:
                mv = (move[0][0] + 7 * (6 - move[0][1]), move[1][0] + 7 * (6 - move[1][1]))
            p = oracle.makemove(p, mv)
        assert oracle.board_json(p) == game["final_board"]
        assert oracle.result(p) == game["result"]


# ---------------- live cross-checks against the compiled reference ----------------

def test_live_reference_random_playouts(oracle, reference):
    rng = random.Random(7)
    checked = 0
    for g in range(40):
        p = oracle.set_board([START_FEN, OPEN_FEN][g % 2])
        while True:
            mo, mr = oracle.movegen(p), reference.movegen(p)
            assert mo == mr
            assert oracle.result(p) == reference.result(p)
            checked += 1
            if not mo or oracle.result(p):
                break
            m = rng.choice(mo)
            a, b = oracle.makemove(p, m), reference.makemove(p, m)
            assert a.key() == b.key() and a.ply == b.ply
            p = a
    assert checked > 3000


def test_live_reference_mcts(oracle, reference):
    for ev, ptr in (("probe", oracle.probe_eval_ptr), ("uniform", oracle.uniform_eval_ptr)):
        ours, evals = oracle.search(START_FEN, 300, ev)
        theirs, revals = reference.search(START_FEN, 300, ptr)
        assert [(m, v) for m, v in ours] == [(m, v) for m, v, w in theirs]
        assert evals == revals


def test_live_umap_order(oracle, reference):
    rng = random.Random(3)
    for _ in range(200):
        p = oracle.set_board(START_FEN)
        for _k in range(rng.randrange(0, 120)):
            mo = oracle.movegen(p)
            if not mo or oracle.result(p):
                break
            p = oracle.makemove(p, rng.choice(mo))
        mo = oracle.movegen(p)
        if not mo:
            continue
        a, ab = oracle.umap_order(mo)
        b, bb = reference.umap_order(mo)
        assert a == b and ab == bb
        assert oracle.umap_order(mo, ab)[0] == reference.umap_order(mo, True)[0]


def test_torch_cpu_net_matches_numpy_restatement():
    """The torch-CPU forward used by bench.py's reference arm is the same function as the NumPy oracle."""
    from oracle import net_numpy, net_torch
    conv, bn = net_numpy.init_weights(seed=0)
    bn = net_numpy.randomize_bn(bn)
    feats = net_numpy.random_features(6, seed=2)
    want_p, want_v = net_numpy.forward(feats, conv, bn, dtype=np.float64)
    got_p, got_v = net_torch.TorchNet(conv, bn).forward(feats)
    assert got_p.shape == (6, 7, 7, 17) and got_v.shape == (6, 1)
    assert np.abs(got_p - want_p).max() < 1e-5 and np.abs(got_v - want_v).max() < 1e-5


def test_weight_file_layout_roundtrip(tmp_path):
    """model.py:179-196: np.save(path, [conv_list, bn_list]) / np.load(allow_pickle=True)."""
    from oracle import net_numpy
    from ataxxzero_b200 import model
    conv, bn = net_numpy.init_weights(seed=4)
    path = str(tmp_path / "m.npy")
    net_numpy.save_model(path, conv, bn)
    a = model.Network.load(path)                      # product loader reads the oracle-written file
    assert a.filters == 128 and a.blocks == 12 and a.total_parameters == 3545906
    a.save(str(tmp_path / "n.npy"))
    conv2, bn2 = net_numpy.load_model(str(tmp_path / "n.npy"))      # and the other way round
    assert all(np.array_equal(x, y) for x, y in zip(conv, conv2)) and all(np.array_equal(x, y) for x, y in zip(bn, bn2))
    assert a.packed().size == 3545906 + 50 * 128


def test_live_reference_edge_positions(oracle, reference):
    """Hand-made corner cases (zero pieces, stuck sides, full boards, walled-in pieces, 103 legal moves, jump-only
    positions): the oracle and the compiled reference agree on move lists, adjudication and move application."""
    from test_rules_gpu import EDGE_FENS
    for fen in EDGE_FENS:
        p = oracle.set_board(fen)
        assert reference.set_board(fen).key() == p.key()
        mo = oracle.movegen(p)
        assert mo == reference.movegen(p), fen
        assert oracle.result(p) == reference.result(p), fen
        if oracle.result(p) == 0:
            for m in mo:
                assert oracle.makemove(p, m).key() == reference.makemove(p, m).key(), (fen, m)


def test_net_oracle_has_three_unrelated_witnesses():
    """The net oracle is parity-unpinned (no TensorFlow here, no reference goldens).  Three evaluations that share no
    code path must agree on it: NumPy shifted-window matmuls (the oracle), torch/oneDNN conv2d, and scipy's direct
    cross-correlation on an explicitly padded board with a hand-written value head (oracle/net_scipy.py) -- at random-init
    scale and at trained scale (calibrated batch-norm, large logits, saturating values)."""
    from oracle import net_numpy, net_scipy, net_torch
    cases = [net_numpy.init_weights(seed=3), net_numpy.trained_scale_weights(seed=1)]
    cases[0] = (cases[0][0], net_numpy.randomize_bn(cases[0][1], seed=2))
    for conv, bn in cases:
        feats = net_numpy.random_features(2, seed=17)
        want_p, want_v = net_numpy.forward(feats, conv, bn, dtype=np.float64)
        scale = max(1.0, float(np.abs(want_p).max()))
        sp, sv = net_scipy.forward_one(feats[0], conv, bn)
        assert np.abs(sp - want_p[0]).max() < 1e-12 * scale and abs(sv - want_v[0, 0]) < 1e-12
        tp, tv = net_torch.TorchNet(conv, bn).forward(feats)
        assert np.abs(tp - want_p).max() < 2e-5 * scale and np.abs(tv - want_v).max() < 2e-5

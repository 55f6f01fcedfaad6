"""GPU tests of the reference-facing Python surfaces: engine.py twin (MCTS / MCTSEngine on a device tree), the UAI
front end, and the command-line twins (perft, generate_games --random-play, accelerated_generate_games, looper)."""
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
START_FEN = "x5o/7/3-3/2-1-2/3-3/7/o5x x"


@pytest.fixture(scope="module")
def engine_mod(tmp_path_factory):
    """engine.initialize_model() on a reference-format .npy written by model.Network.save()."""
    from ataxxzero_b200 import engine, model
    path = str(tmp_path_factory.mktemp("models") / "model-001.npy")
    model.Network.random_init(seed=0).save(path)
    if not engine.initialized:
        engine.setup_evaluator(use_rpc=False)
        engine.initialize_model(path, device=0, seed=5)
    engine.model_path = path
    return engine


def test_engine_mcts_matches_pool_and_python_formula(engine_mod):
    """engine.MCTS is the device tree: same root statistics as a raw pool; the snapshot's total_action_score is the
    reference's PUCT formula (engine.py:268-276) on those statistics."""
    from ataxxzero_b200 import ataxx_rules as ar, rules, search
    board = ar.AtaxxState.from_fen(START_FEN)
    m = engine_mod.MCTS(board)
    m.search(300)
    root = m.root_node
    assert root.all_edge_visits == 300 == sum(e.edge_visits for e in root.outgoing_edges.values())
    with search.Pool(engine_mod.context, 1, 300, eval_mode=search.EVAL_BF16) as pool:
        pool.set_root(0, board.to_position())
        assert pool.run()
        r = pool.root(0)
    want = {rules.to_reference_move(mv): n for mv, n in zip(r["moves"], r["visits"]) if n}
    assert {mv: e.edge_visits for mv, e in root.outgoing_edges.items()} == want
    assert abs(sum(root.posterior.values()) - 1.0) < 1e-9 and set(root.posterior) == set(board.legal_moves())
    best = root.select_action()
    scores = {mv: root.total_action_score(mv) for mv in root.posterior}
    assert scores[best] == max(scores.values())
    # step(): one more visit, returns the root edge it went through
    edge = m.step()
    assert m.root_node.all_edge_visits == 301 and edge.move in m.root_node.outgoing_edges
    # principal variation (best=True) starts with the most visited root edge
    _, _, pv = m.select_principal_variation(best=True)
    top = max(m.root_node.outgoing_edges.values(), key=lambda e: e.edge_visits)
    assert pv and pv[0].move == top.move and pv[0].edge_visits == top.edge_visits
    # play(): subtree kept (visits carry over, SURVEY A-5)
    m.play(board.to_move, top.move)
    assert m.root_node.all_edge_visits == top.edge_visits - 1
    m.close()


def test_nn_evaluator_contract(engine_mod, oracle):
    """NNEvaluator.populate: posterior over legal moves (sums to ~1), value in [-1,1], game_over adjudication."""
    from ataxxzero_b200 import ataxx_rules as ar
    ev = engine_mod.NNEvaluator(temperature=0.0)
    b = ar.AtaxxState.initial()
    ev.populate(b)
    assert set(b.evaluations.posterior) == set(b.legal_moves()) and not b.evaluations.game_over
    assert abs(sum(b.evaluations.posterior.values()) - 1.0) < 1e-4 and -1.0 <= b.evaluations.value <= 1.0
    won = ar.AtaxxState.from_fen("xxxxxxx/xxxxxxx/xxxxxxx/xxxxxxx/xxxxxxx/xxxxxxx/xxxxxx1 o")
    ev.populate(won)
    assert won.evaluations.game_over and won.evaluations.value == -1.0          # o to move, x wins


def test_mcts_engine_genmove_and_set_state(engine_mod):
    from ataxxzero_b200 import ataxx_rules as ar
    eng = engine_mod.MCTSEngine()
    eng.MAX_STEPS = 200
    move = eng.genmove(1000000.0, early_out=False)
    assert move in eng.state.legal_moves() and eng.mcts.root_node.all_edge_visits == 200
    # two plies later the subtree is reused (engine.py:467-479)
    b = eng.state.copy()
    b.move(move)
    reply = b.legal_moves()[0]
    b.move(reply)
    eng.set_state(b)
    assert eng.state == b and eng.mcts.board == b
    move2 = eng.genmove(1000000.0, use_weighted_exponent=5.0)
    assert move2 in b.legal_moves()
    eng.set_state(ar.AtaxxState.from_fen(START_FEN))                           # unrelated position: tree rebuilt
    assert eng.mcts.root_node.all_edge_visits == 0
    eng.mcts.close()


def test_uai_session(engine_mod):
    import argparse
    from ataxxzero_b200.cli import uai_interface as u
    out = io.StringIO()
    args = argparse.Namespace(visits=150, safety_ms=0, show_game=False)
    lines = ["uai", "isready", "uainewgame", "moves g1f2 a1a2", "go movetime 100", "position fen " + START_FEN, "go movetime 100",
             "showboard", "quit"]
    eng = u.main(args, engine_mod, lines=lines, out=out)
    text = out.getvalue().splitlines()
    assert "uaiok" in text and "readyok" in text and "boardok" in text
    best = [l.split()[1] for l in text if l.startswith("bestmove ")]
    assert len(best) == 2
    from ataxxzero_b200 import ataxx_rules as ar
    assert u.uai_decode_move(best[1]) in ar.AtaxxState.from_fen(START_FEN).legal_moves()
    eng.mcts.close()


def test_perft_cli_matches_reference_divide(capsys):
    from ataxxzero_b200.cli import perft
    assert perft.main(["--depth", "4"]) == 155888                              # perft.py "Total:" (SURVEY C-1)
    out = capsys.readouterr().out
    sizes = {l.split("Size:")[0].strip(): int(l.split("Size:")[1]) for l in out.splitlines() if l.startswith("Move:")}
    assert sizes["Move: ('c', (0, 1))"] == 9138 and sizes["Move: ('c', (1, 1))"] == 10562
    assert sizes["Move: ('c', (5, 5))"] == 10562 and sizes["Move: ('c', (6, 5))"] == 9138
    assert perft.main(["--depth", "6", "--fen", START_FEN]) == 97538324


def test_generate_games_random_play(tmp_path, oracle):
    """Config 1: --random-play --game-count N; every record replays legally through the oracle rules."""
    from ataxxzero_b200.cli import generate_games
    out = str(tmp_path / "random.json")
    assert generate_games.main(["--random-play", "--game-count", "200", "--output-games", out, "--seed", "3"]) == 200
    lines = open(out).read().splitlines()
    assert len(lines) == 200
    plies = []
    for ln in lines:
        game = json.loads(ln)
        assert set(game) == {"boards", "moves", "result"} and game["result"] in (1, 2)
        p = oracle.set_board("x5o/7/7/7/7/7/o5x x")
        for board, move in zip(game["boards"], game["moves"]):
            assert oracle.board_json(p) == board and oracle.result(p) == 0
            if move[0] == "c":
                sq = move[1][0] + 7 * (6 - move[1][1])
                mv = (sq, sq)
            else:
                mv = (move[0][0] + 7 * (6 - move[0][1]), move[1][0] + 7 * (6 - move[1][1]))
            assert oracle.legal_move(p, mv)
            p = oracle.makemove(p, mv)
        assert oracle.result(p) == game["result"]
        plies.append(len(game["moves"]))
    assert 120 < np.mean(plies) < 260                                           # reference: mean ~184 plies


def test_accelerated_generate_games_and_looper(tmp_path, engine_mod, oracle):
    from test_selfplay_gpu import replay_and_check
    prefix = tmp_path / "run"
    (prefix / "models").mkdir(parents=True)
    (prefix / "games").mkdir()
    os.link(engine_mod.model_path, prefix / "models" / "model-001.npy") if hasattr(os, "link") else None
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "ataxxzero_b200.cli.looper", "--prefix", str(prefix), "--visits", "24", "--game-count", "6",
                        "--buffer-size", "16", "--poll-seconds", "1", "--iterations", "1"], capture_output=True, text=True, env=env,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = open(prefix / "games" / "model-001-0.json").read().splitlines()
    assert len(lines) >= 6
    for ln in lines:
        replay_and_check(oracle, ln, 24)
    assert "Training is out of scope" in r.stdout


def test_perft_sharded_single_rank(ctx):
    """dist.perft_sharded with one rank equals the plain device perft (the N>1 path is the same code plus one all-reduce)."""
    from ataxxzero_b200 import dist, rules
    for fen, depth, want in ((rules.OPEN_FEN, 5, 4752668), (rules.START_FEN, 6, 97538324), (rules.OPEN_FEN, 2, 256)):
        assert dist.perft_sharded(ctx, rules.set_board(fen), depth) == want

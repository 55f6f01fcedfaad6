"""generate_games.py of the reference (same flags and record format), with the rules batched on the GPU.

``--random-play`` (BASELINE config 1, the only mode of the reference that still runs): all
``--game-count`` games advance together -- one ``movegen`` / ``makemove`` / ``get_board_result`` kernel call per ply over
every unfinished game -- and each is written as the reference's Python record
``{"boards": [[49 ints]...], "moves": [[[sx,sy],[ex,ey]] | ["c",[x,y]] ...], "result": 1|2}`` (generate_games.py:20-68).
Without ``--random-play`` the moves come from ``engine.MCTS`` (device-resident tree) as the reference intended:
search until the most-visited root edge has ``--visit-count`` visits (at most 10x that many steps), pick a move
proportionally to visits, and randomise the opening plies with the reference's schedule.
"""
import argparse
import json
import os
import random

from .. import Context, ataxx_rules, rules

MAX_STEP_RATIO = 10
MAXIMUM_GAME_PLIES = 400
LOGIT_TEMPERATURE = 0.0
OPENING_RANDOMIZATION_SCHEDULE = [0.2 * (0.5 ** (i / 2)) for i in range(10)]


def _json_move(move):
    return list(map(list, move)) if move[0] != "c" else ["c", list(move[1])]


_BOARD_LUT = None


def _board_cells(x, o):
    """49 ints (1 = x, 2 = o), index x + 7*y with y = 0 at the top, from the two piece bitboards (bit sq = x + 7*(6-y))."""
    global _BOARD_LUT
    if _BOARD_LUT is None:
        _BOARD_LUT = [(i % 7) + 7 * (6 - i // 7) for i in range(49)]
    return [1 if x >> sq & 1 else 2 if o >> sq & 1 else 0 for sq in _BOARD_LUT]


def random_play_games(ctx, count, seed=0, start_fen="x5o/7/7/7/7/7/o5x x"):
    """`count` uniformly random games played on the GPU (az_random_playouts: one thread per game); returns the entries in
    the reference's Python record format.  `result` is None for a game that reached 400 plies undecided."""
    start = ataxx_rules.AtaxxState.from_fen(start_fen).to_position()
    plies, n_plies, results = rules.random_playouts(ctx, start, count, MAXIMUM_GAME_PLIES, seed)
    entries = []
    for g in range(count):
        n = int(n_plies[g])
        xs, os_, mvs = plies["x"][g, :n].tolist(), plies["o"][g, :n].tolist(), plies["move"][g, :n].tolist()
        entries.append({"boards": [_board_cells(x, o) for x, o in zip(xs, os_)],
                        "moves": [_json_move(rules.to_reference_move((m & 0xff, m >> 8))) for m in mvs],
                        "result": int(results[g]) or None})
    return entries


class UAIPlayer:
    """A UAI engine in a subprocess (what ``--supervised CMD`` talks to; the reference keeps this class in
    uai_ringmaster.py:9-60): ``set_state(board)`` sends the FEN, ``genmove(ms)`` asks for a move."""

    def __init__(self, cmd):
        import shlex
        import subprocess
        self.cmd = shlex.split(cmd) if isinstance(cmd, str) else list(cmd)
        self.proc = subprocess.Popen(self.cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        for line in ("uai", "isready", "uainewgame"):
            self.send(line)

    def send(self, text):
        self.proc.stdin.write((text + "\n").encode("utf8"))
        self.proc.stdin.flush()

    def set_state(self, board):
        self.send("position fen %s" % board.fen())

    def genmove(self, ms=1000):
        from .uai_interface import uai_decode_move
        self.send("go movetime %i" % ms)
        while True:
            line = self.proc.stdout.readline().decode("utf8")
            if not line:
                raise ValueError("the UAI engine %r closed its output" % (self.cmd,))
            if line.startswith("bestmove "):
                return uai_decode_move(line.split()[1])

    def quit(self):
        try:
            self.send("quit")
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()


def supervised_game(args):
    """Record a game played by an external UAI engine (generate_games.py:27-36): its move is both played and the target."""
    board = ataxx_rules.AtaxxState.initial()
    entry = {"boards": [], "moves": []}
    for ply in range(MAXIMUM_GAME_PLIES):
        args.uai_player.set_state(board)
        move = args.uai_player.genmove(args.supervised_ms)
        if move not in board.legal_moves():
            raise ValueError("illegal move %r from the UAI engine at %s" % (move, board.fen()))
        entry["boards"].append(list(board.board))
        entry["moves"].append(_json_move(move) if move != "pass" else "pass")
        board.move(move)
        if board.result() is not None:
            break
    entry["result"] = board.result()
    print("[%3i] Generated a %i ply game with result %r." % (args.group_index, len(entry["boards"]), entry["result"]))
    return entry


def mcts_game(args, engine):
    """One self-play game driven through the engine.py interface (generate_games.py:16-77)."""
    board = ataxx_rules.AtaxxState.initial()
    m = engine.MCTS(board.copy(), use_dirichlet_noise=True)
    entry = {"boards": [], "moves": []}
    all_steps = 0
    helper = engine.MCTSEngine.__new__(engine.MCTSEngine)       # sample_with_exponential_weight needs only .mcts
    helper.mcts = m
    for ply in range(MAXIMUM_GAME_PLIES):
        while True:
            root = m.root_node
            most = max([e.edge_visits for e in root.outgoing_edges.values()] or [0])
            if most >= args.visit_count or root.all_edge_visits >= args.visit_count * MAX_STEP_RATIO:
                break
            burst = min(64, max(1, args.visit_count - most))
            m.search(root.all_edge_visits + burst)
            all_steps += burst
        training_move = selected_move = helper.sample_with_exponential_weight(1.0)
        if ply < len(OPENING_RANDOMIZATION_SCHEDULE) and random.random() < OPENING_RANDOMIZATION_SCHEDULE[ply]:
            selected_move = random.choice(board.legal_moves())
        entry["boards"].append(list(board.board))
        entry["moves"].append(_json_move(training_move))
        m.play(board.to_move, selected_move, print_variation_count=False)
        board.move(selected_move)
        if board.result() is not None:
            break
        if args.show_game:
            print(board)
        if args.die_if_present and os.path.exists(args.die_if_present):
            print("Exiting due to signal file!")
            raise SystemExit
    m.close()
    entry["result"] = board.result()
    print("[%3i] Generated a %i ply game (%.2f avg steps) with result %r." % (
        args.group_index, len(entry["boards"]), all_steps / float(ply + 1), entry["result"]))
    return entry


def build_parser():
    parser = argparse.ArgumentParser(description="Generates games in the .json format that train.py reads.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--network", metavar="PATH", default="", help="Path of the model to load.")
    parser.add_argument("--output-games", metavar="PATH", type=str, default=None, help="Path to write .json games to.")
    parser.add_argument("--group-index", metavar="N", default=0, type=int, help="Our index in the work group.")
    parser.add_argument("--use-rpc", action="store_true", help="Use RPC for NN evaluation (not supported here).")
    parser.add_argument("--random-play", action="store_true", help="Generate games by totally random play.")
    parser.add_argument("--visit-count", metavar="N", default=200, type=int,
                        help="Perform MCTS steps until the PV move has at least N visits.")
    parser.add_argument("--die-if-present", metavar="PATH", default=None, type=str, help="Die once a file is present at the target path.")
    parser.add_argument("--show-game", action="store_true", help="Show the game while it's generating.")
    parser.add_argument("--game-count", metavar="N", default=None, type=int, help="Maximum number of games to generate.")
    parser.add_argument("--no-write", action="store_true", help="Don't write out generated games at all.")
    parser.add_argument("--supervised", metavar="CMD", default=None, type=str, help="Command for a UAI engine.")
    parser.add_argument("--supervised-ms", metavar="N", default=100, type=int,
                        help="Number of milliseconds per move for supervised generation.")
    parser.add_argument("--device", metavar="N", default=0, type=int, help="GPU index.")
    parser.add_argument("--seed", metavar="N", default=None, type=int, help="Seed for Python's RNG (random play).")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.use_rpc:
        raise SystemExit("--use-rpc: the gevent/mprpc transport is out of scope for this package")
    if args.seed is not None:
        random.seed(args.seed)
    if args.output_games is None:
        name = os.path.splitext(os.path.basename(args.network))[0] or "random"
        directory = os.path.join("games", name)
        os.makedirs(directory, exist_ok=True)
        output_path = os.path.join(directory, os.urandom(8).hex() + ".json")
    else:
        output_path = args.output_games
    if args.no_write:
        output_path = "/dev/null"
    print("[%3i] Writing to: %s" % (args.group_index, output_path))
    written = 0
    with open(output_path, "w") as f:          # "w", as the reference does despite its help text (generate_games.py:127)
        def emit(entry):
            nonlocal written
            if entry["result"] is None:
                print("[%3i] Skipping game with null result." % (args.group_index,))
                return
            json.dump(entry, f)
            f.write("\n")
            f.flush()
            written += 1
        if args.random_play:
            print("Doing random play! Loading no model, and not using RPC.")
            with Context(device=args.device) as ctx:
                target = args.game_count if args.game_count is not None else 1 << 62
                batch = 0
                while written < target:
                    seed = (args.seed if args.seed is not None else random.getrandbits(62)) + batch
                    for entry in random_play_games(ctx, min(4096, target - written), seed=seed):
                        emit(entry)
                    batch += 1
        elif args.supervised is not None:
            args.uai_player = UAIPlayer(args.supervised)
            try:
                while args.game_count is None or written < args.game_count:
                    emit(supervised_game(args))
            finally:
                args.uai_player.quit()
        else:
            from .. import engine
            engine.setup_evaluator(use_rpc=False, temperature=LOGIT_TEMPERATURE)
            engine.initialize_model(args.network, device=args.device)
            while args.game_count is None or written < args.game_count:
                emit(mcts_game(args, engine))
    print("Done generating games.")
    return written


if __name__ == "__main__":
    main()

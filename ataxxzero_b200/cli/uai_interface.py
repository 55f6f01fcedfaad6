"""UAI front end with the reference's protocol and flags (uai_interface.py:6-107), searching on the GPU.

``uai`` / ``uainewgame`` / ``isready`` / ``moves ...`` / ``position fen ...`` / ``go movetime N`` / ``showboard`` /
``quit``; squares are ``a1``..``g7`` with rank 7 at the top, clones are the destination only, pass is ``0000``.
Unlike the reference's ``from_fen`` the position parser accepts ``-`` blockers.
"""
import string
import sys

from .. import ataxx_rules


def uai_encode_square(xy):
    return "%s%i" % (string.ascii_lowercase[xy[0]], 7 - xy[1])


def uai_encode_move(move):
    if move == "pass":
        return "0000"
    start, end = move
    if start == "c":
        return uai_encode_square(end)
    return uai_encode_square(start) + uai_encode_square(end)


def uai_decode_square(s):
    return string.ascii_lowercase.index(s[0].lower()), 7 - int(s[1])


def uai_decode_move(s):
    if s in ("pass", "none", "0000"):
        return "pass"
    if len(s) == 2:
        return "c", uai_decode_square(s)
    if len(s) == 4:
        return uai_decode_square(s[:2]), uai_decode_square(s[2:])
    raise Exception("Bad UAI move string: %r" % s)


def main(args, engine, lines=None, out=sys.stdout):
    board = ataxx_rules.AtaxxState.initial()
    eng = engine.MCTSEngine()
    if args.visits is not None:
        eng.MAX_STEPS = args.visits
    for line in (lines if lines is not None else iter(input, None)):
        line = line.strip()
        if line == "quit":
            break
        elif line == "uai":
            print("id name AtaxxZero-B200", file=out)
            print("id author ataxxzero_b200 (protocol of Peter Schmidt-Nielsen's AtaxxZero)", file=out)
            print("uaiok", file=out)
        elif line == "uainewgame":
            board = ataxx_rules.AtaxxState.initial()
            eng.set_state(board)
        elif line == "isready":
            print("readyok", file=out)
        elif line.startswith("moves "):
            for text in line[6:].split():
                board.move(uai_decode_move(text))
            eng.set_state(board.copy())
        elif line.startswith("position fen "):
            board = ataxx_rules.AtaxxState.from_fen(line[13:])
            eng.set_state(board)
            if args.show_game:
                print("===\n%s" % (board,), file=sys.stderr)
        elif line.startswith("go movetime "):
            ms = int(line[12:]) - args.safety_ms
            move = eng.genmove(ms * 1e-3 if args.visits is None else 1000000.0, use_weighted_exponent=5.0)
            print("bestmove %s" % (uai_encode_move(move),), file=out)
        elif line == "showboard":
            print(board, file=out)
            print("boardok", file=out)
        out.flush()
    return eng


def parse_args(argv=None):
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("--network-path", metavar="NETWORK", type=str, help="Name of the model to load.")
    parser.add_argument("--visits", metavar="VISITS", default=None, type=int, help="Number of visits during MCTS.")
    parser.add_argument("--safety-ms", metavar="MS", default=0, type=int, help="Number of milliseconds to shave off of each movetime for safety.")
    parser.add_argument("--show-game", action="store_true", help="Show the game on stderr.")
    parser.add_argument("--device", metavar="N", default=0, type=int, help="GPU to search on.")
    return parser.parse_args(argv)


if __name__ == "__main__":
    from .. import engine as _engine
    _args = parse_args()
    print(_args, file=sys.stderr)
    _engine.setup_evaluator(use_rpc=False)
    _engine.initialize_model(_args.network_path, device=_args.device)
    main(_args, _engine)

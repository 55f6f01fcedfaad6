"""UAI front end speaking the reference's protocol (uai_interface.py:41-88), searching on the GPU.

Commands: ``uai``, ``uainewgame``, ``isready``, ``moves m1 m2 ...``, ``position fen FEN``, ``go movetime MS``,
``showboard``, ``quit``.  Squares are ``a1``..``g7`` with rank 7 at the top (y = 0), a clone move is written as its
destination only, a jump as source + destination, a pass as ``0000`` (uai_interface.py:6-32).  Unlike the reference's
FEN reader, ``position fen`` accepts ``-`` blockers.
"""
import sys

from .. import ataxx_rules

FILES = "abcdefg"
PASS_WORDS = ("pass", "none", "0000")


def uai_encode_square(xy):
    x, y = xy
    return FILES[x] + str(7 - y)


def uai_decode_square(text):
    return FILES.index(text[0].lower()), 7 - int(text[1])


def uai_encode_move(move):
    if move == "pass":
        return "0000"
    squares = [sq for sq in move if sq != "c"]
    return "".join(uai_encode_square(sq) for sq in squares)


def uai_decode_move(text):
    if text in PASS_WORDS:
        return "pass"
    if len(text) not in (2, 4):
        raise Exception("Bad UAI move string: %r" % text)
    squares = [uai_decode_square(text[i:i + 2]) for i in range(0, len(text), 2)]
    return ("c", squares[0]) if len(squares) == 1 else (squares[0], squares[1])


class Session:
    """One UAI conversation: the current board, an ``engine.MCTSEngine`` searching it, and a line dispatcher."""

    def __init__(self, args, engine, out=sys.stdout):
        self.args, self.engine, self.out = args, engine, out
        self.board = ataxx_rules.AtaxxState.initial()
        self.eng = engine.MCTSEngine()
        if args.visits is not None:
            self.eng.MAX_STEPS = args.visits
        self.commands = [("moves ", self.cmd_moves), ("position fen ", self.cmd_position), ("go movetime ", self.cmd_go)]
        self.words = {"uai": self.cmd_uai, "uainewgame": self.cmd_newgame, "isready": self.cmd_isready, "showboard": self.cmd_showboard}

    def say(self, text):
        print(text, file=self.out)

    def cmd_uai(self):
        self.say("id name AtaxxZero-B200")
        self.say("id author ataxxzero_b200 (UAI protocol as in petersn/AtaxxZero)")
        self.say("uaiok")

    def cmd_newgame(self):
        self.board = ataxx_rules.AtaxxState.initial()
        self.eng.set_state(self.board)

    def cmd_isready(self):
        self.say("readyok")

    def cmd_showboard(self):
        self.say(str(self.board))
        self.say("boardok")

    def cmd_moves(self, rest):
        for text in rest.split():
            self.board.move(uai_decode_move(text))
        self.eng.set_state(self.board.copy())

    def cmd_position(self, rest):
        self.board = ataxx_rules.AtaxxState.from_fen(rest)
        self.eng.set_state(self.board)
        if self.args.show_game:
            print("===\n%s" % (self.board,), file=sys.stderr)

    def cmd_go(self, rest):
        budget = (int(rest) - self.args.safety_ms) * 1e-3
        if self.args.visits is not None:
            budget = 1000000.0                     # the visit limit (MAX_STEPS) ends the search, as in the reference
        move = self.eng.genmove(budget, use_weighted_exponent=5.0)
        self.say("bestmove %s" % (uai_encode_move(move),))
        if self.args.show_game and move != "pass":
            after = self.board.copy()
            after.move(move)
            print(after, file=sys.stderr)

    def feed(self, line):
        """Handle one input line; returns False on ``quit``."""
        line = line.strip()
        if line == "quit":
            return False
        if line in self.words:
            self.words[line]()
        else:
            for prefix, fn in self.commands:
                if line.startswith(prefix):
                    fn(line[len(prefix):])
                    break
        self.out.flush()
        return True


def main(args, engine, lines=None, out=sys.stdout):
    session = Session(args, engine, out)
    source = lines if lines is not None else iter(sys.stdin.readline, "")
    for line in source:
        if not session.feed(line):
            break
    return session.eng


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

"""Interface twin of the reference's ``ataxx_rules.py`` (ataxx_rules.py:37-179): the board object
that ``engine.py`` / ``uai_interface.py`` / ``generate_games.py`` pass around.

Same surface -- ``AtaxxState.initial/from_fen/copy/move/legal_moves/result/fen``, ``board`` as an
``array('b')`` indexed ``x + 7*y`` with y = 0 at the top, moves as ``((sx,sy),(ex,ey))`` /
``("c",(x,y))`` / ``"pass"`` -- but the state is ALSO kept as the bitboards the GPU library uses
(``to_position()``), and unlike the reference twin it understands blockers (``-`` in FENs), so the
C++ start position ``x5o/7/3-3/2-1-2/3-3/7/o5x x`` can be passed through the same interface.
Host-side bookkeeping for the CLIs only: searches and batched rules run in libataxxzero.so.
"""
import array

from ._native import Position

SIZE = 7
OTHER_PLAYER = {1: 2, 2: 1}
NEAR_NEIGHBOR_OFFSETS = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]
FAR_NEIGHBOR_OFFSETS = [(a, b) for a in (-2, -1, 0, 1, 2) for b in (-2, -1, 0, 1, 2)
                        if (a, b) != (0, 0) and (a, b) not in NEAR_NEIGHBOR_OFFSETS]
BLOCKED_CELLS = frozenset()      # the reference twin's default (ataxx_rules.py:10); per-state blockers below


def Linf_distance(a, b):
    return max(abs(a[0] - b[0]), abs(a[1] - b[1]))


class AtaxxState:
    def __init__(self, board, to_move=1, legal_moves_cache=None, blocked=frozenset()):
        self.board = board
        self.to_move = to_move
        self.legal_moves_cache = legal_moves_cache
        self.blocked = frozenset(blocked)

    @staticmethod
    def initial():
        s = AtaxxState(array.array("b", [0] * 49))
        s[0, 0] = 1
        s[6, 6] = 1
        s[6, 0] = 2
        s[0, 6] = 2
        return s

    @staticmethod
    def from_fen(fen):
        mapping = {"x": 1, "o": 2}
        parts = fen.lower().split()
        rows, to_move = parts[0], (parts[1] if len(parts) > 1 else "x")
        s = AtaxxState(array.array("b", [0] * 49), to_move=mapping[to_move])
        blocked = set()
        for y, chunk in enumerate(rows.split("/")):
            x = 0
            for c in chunk:
                if c in "1234567":
                    x += int(c)
                    continue
                if c == "-":
                    blocked.add((x, y))
                else:
                    s[x, y] = mapping[c]
                x += 1
        s.blocked = frozenset(blocked)
        return s

    @staticmethod
    def from_position(pos):
        s = AtaxxState(array.array("b", [0] * 49), to_move=pos.turn + 1)
        blocked = set()
        for sq in range(49):
            xy = (sq % 7, 6 - sq // 7)
            if pos.pieces[0] >> sq & 1:
                s[xy] = 1
            elif pos.pieces[1] >> sq & 1:
                s[xy] = 2
            elif pos.blockers >> sq & 1:
                blocked.add(xy)
        s.blocked = frozenset(blocked)
        return s

    def to_position(self, ply=0):
        p = Position()
        p.ply, p.turn = ply, self.to_move - 1
        for y in range(7):
            for x in range(7):
                bit = 1 << (x + 7 * (6 - y))
                if self[x, y] == 1:
                    p.pieces[0] |= bit
                elif self[x, y] == 2:
                    p.pieces[1] |= bit
                elif (x, y) in self.blocked:
                    p.blockers |= bit
        return p

    def copy(self):
        return AtaxxState(self.board[:], self.to_move, self.legal_moves_cache, self.blocked)

    def __setitem__(self, index, value):
        self.board[index[0] + index[1] * SIZE] = value

    def __getitem__(self, index):
        return self.board[index[0] + index[1] * SIZE]

    def __eq__(self, other):
        return self.to_move == other.to_move and self.board == other.board and self.blocked == other.blocked

    def _legal_spot(self, xy):
        return xy not in self.blocked and 0 <= xy[0] < SIZE and 0 <= xy[1] < SIZE

    def __str__(self):
        return "\n".join(" ".join("#" if (x, y) in self.blocked else {0: ".", 1: "X", 2: "O"}[self[x, y]] for x in range(SIZE))
                         for y in range(SIZE))

    def fen(self):
        s = "/".join("".join("-" if (x, y) in self.blocked else {0: ".", 1: "x", 2: "o"}[self[x, y]] for x in range(SIZE))
                     for y in range(SIZE)) + " " + {1: "x", 2: "o"}[self.to_move]
        for i in range(SIZE, 0, -1):
            s = s.replace("." * i, str(i))
        return s

    def move(self, desc):
        self.legal_moves_cache = None
        if desc == "pass":
            self.to_move = OTHER_PLAYER[self.to_move]
            return
        start, end = desc
        if start != "c":
            assert self[start] == self.to_move
        assert end not in self.blocked
        assert self[end] == 0
        self[end] = self.to_move
        if start != "c":
            distance = Linf_distance(start, end)
            assert distance in (1, 2)
            if distance == 2:
                self[start] = 0
        for i, j in NEAR_NEIGHBOR_OFFSETS:
            n = (end[0] + i, end[1] + j)
            if self._legal_spot(n) and self[n] != 0:
                self[n] = self.to_move
        self.to_move = OTHER_PLAYER[self.to_move]

    def legal_moves(self):
        if self.legal_moves_cache is None:
            moves, clones = [], []
            seen = set()
            for x in range(SIZE):
                for y in range(SIZE):
                    if self[x, y] != self.to_move:
                        continue
                    for i, j in FAR_NEIGHBOR_OFFSETS:
                        d = (x + i, y + j)
                        if self._legal_spot(d) and self[d] == 0:
                            moves.append(((x, y), d))
                    for i, j in NEAR_NEIGHBOR_OFFSETS:
                        d = (x + i, y + j)
                        if self._legal_spot(d) and self[d] == 0 and d not in seen:
                            seen.add(d)
                            clones.append(("c", d))
            self.legal_moves_cache = (moves + clones) or ["pass"]
        return self.legal_moves_cache

    def result(self):
        nb = len(self.blocked)
        counts = {i: self.board.count(i) for i in (0, 1, 2)}
        assert counts[1] != 0 or counts[2] != 0
        if self.legal_moves() == ["pass"]:
            counts[OTHER_PLAYER[self.to_move]] += counts[0] - nb
            counts[0] = nb
            return max(counts, key=counts.__getitem__)
        if counts[1] == 0:
            return 2
        if counts[2] == 0:
            return 1
        if counts[0] != nb:
            return None
        return max((1, 2), key=lambda i: self.board.count(i))

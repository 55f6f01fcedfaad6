"""Board object with the interface of the reference's ``ataxx_rules.AtaxxState`` (ataxx_rules.py:37-179),
implemented on the bitboards the GPU library uses.

``engine.py`` / ``uai_interface.py`` / ``generate_games.py`` pass this object around: ``initial()``,
``from_fen()``, ``copy()``, ``move()``, ``legal_moves()``, ``result()``, ``fen()``, ``board`` (an
``array('b')`` indexed ``x + 7*y`` with y = 0 at the top), ``state[x, y]``, ``to_move`` in {1, 2}; moves
are ``((sx, sy), (ex, ey))`` jumps, ``("c", (x, y))`` clones or ``"pass"``.  Here the state *is* an
``az_position`` (two piece bitboards + blockers, bit ``sq = x + 7*(6-y)``, cpp/ataxx.hpp:13-34), so it can
be handed to the library without conversion, and -- unlike the reference twin, whose ``BLOCKED_CELLS`` is
empty and whose ``from_fen`` cannot read ``-`` -- it understands blockers, so the C++
start position ``x5o/7/3-3/2-1-2/3-3/7/o5x x`` goes through the same interface.  ``legal_moves()``
returns moves in the C++ movegen order (cpp/movegen.cpp:16-66: jumps by source/destination ascending,
then one clone per destination ascending), i.e. the order the GPU trees use.
Host-side bookkeeping for the CLIs only; batched rules and searches run in libataxxzero.so.
"""
import array

from ._native import Position

SIZE = 7
OTHER_PLAYER = {1: 2, 2: 1}
BOARD_MASK = (1 << 49) - 1
NEAR_NEIGHBOR_OFFSETS = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]
FAR_NEIGHBOR_OFFSETS = [(a, b) for a in (-2, -1, 0, 1, 2) for b in (-2, -1, 0, 1, 2) if max(abs(a), abs(b)) == 2]
BLOCKED_CELLS = frozenset()      # the reference module's default (ataxx_rules.py:10); blockers live per state here


def _sq(xy):
    return xy[0] + 7 * (6 - xy[1])


def _xy(sq):
    return (sq % 7, 6 - sq // 7)


def _ring(sq, radius):
    """Squares at Chebyshev distance exactly `radius` from sq, as a bitboard."""
    f, r = sq % 7, sq // 7
    out = 0
    for dr in range(-radius, radius + 1):
        for df in range(-radius, radius + 1):
            if max(abs(dr), abs(df)) == radius and 0 <= f + df < 7 and 0 <= r + dr < 7:
                out |= 1 << ((r + dr) * 7 + f + df)
    return out


RING1 = [_ring(s, 1) for s in range(49)]
RING2 = [_ring(s, 2) for s in range(49)]


def _bits(bb):
    while bb:
        low = bb & -bb
        yield low.bit_length() - 1
        bb ^= low


def Linf_distance(a, b):
    return max(abs(a[0] - b[0]), abs(a[1] - b[1]))


class AtaxxState:
    __slots__ = ("pieces", "blockers", "to_move", "legal_moves_cache", "evaluations")

    def __init__(self, board=None, to_move=1, legal_moves_cache=None, blocked=()):
        self.pieces = [0, 0]
        self.blockers = 0
        for xy in blocked:
            self.blockers |= 1 << _sq(xy)
        self.to_move = to_move
        self.legal_moves_cache = legal_moves_cache
        if board is not None:
            for i, v in enumerate(board):
                if v:
                    self.pieces[v - 1] |= 1 << _sq((i % 7, i // 7))

    # ---- constructors ----
    @staticmethod
    def initial():
        return AtaxxState.from_fen("x5o/7/7/7/7/7/o5x x")

    @staticmethod
    def from_fen(fen):
        parts = fen.split()
        s = AtaxxState(to_move={"x": 1, "o": 2}[(parts[1] if len(parts) > 1 else "x").lower()])
        rows = parts[0].split("/")
        if len(rows) != 7:
            raise ValueError("bad FEN %r" % (fen,))
        for y, row in enumerate(rows):
            x = 0
            for ch in row:
                if ch.isdigit():
                    x += int(ch)
                    continue
                bit = 1 << _sq((x, y))
                if ch in "xX":
                    s.pieces[0] |= bit
                elif ch in "oO":
                    s.pieces[1] |= bit
                elif ch == "-":
                    s.blockers |= bit
                else:
                    raise ValueError("bad FEN %r" % (fen,))
                x += 1
            if x != 7:
                raise ValueError("bad FEN %r" % (fen,))
        return s

    @staticmethod
    def from_position(pos):
        s = AtaxxState(to_move=pos.turn + 1)
        s.pieces = [int(pos.pieces[0]), int(pos.pieces[1])]
        s.blockers = int(pos.blockers)
        return s

    def to_position(self, ply=0):
        p = Position()
        p.ply, p.turn, p.blockers = ply, self.to_move - 1, self.blockers
        p.pieces[0], p.pieces[1] = self.pieces
        return p

    def copy(self):
        s = AtaxxState(to_move=self.to_move, legal_moves_cache=self.legal_moves_cache)
        s.pieces = list(self.pieces)
        s.blockers = self.blockers
        return s

    # ---- the reference's array view ----
    @property
    def board(self):
        cells = array.array("b", [0] * 49)
        for player in (0, 1):
            for sq in _bits(self.pieces[player]):
                x, y = _xy(sq)
                cells[x + 7 * y] = player + 1
        return cells

    @property
    def blocked(self):
        return frozenset(_xy(sq) for sq in _bits(self.blockers))

    def __getitem__(self, xy):
        bit = 1 << _sq(xy)
        return 1 if self.pieces[0] & bit else 2 if self.pieces[1] & bit else 0

    def __setitem__(self, xy, value):
        bit = 1 << _sq(xy)
        self.pieces[0] &= ~bit
        self.pieces[1] &= ~bit
        if value:
            self.pieces[value - 1] |= bit
        self.legal_moves_cache = None

    def __eq__(self, other):
        return (self.to_move, self.pieces, self.blockers) == (other.to_move, other.pieces, other.blockers)

    def __hash__(self):
        return hash((self.to_move, self.pieces[0], self.pieces[1], self.blockers))

    def __str__(self):
        glyph = {0: ".", 1: "X", 2: "O"}
        return "\n".join(" ".join("#" if self.blockers >> _sq((x, y)) & 1 else glyph[self[x, y]] for x in range(7)) for y in range(7))

    def fen(self):
        rows = []
        for y in range(7):
            row, gap = "", 0
            for x in range(7):
                ch = "-" if self.blockers >> _sq((x, y)) & 1 else {0: "", 1: "x", 2: "o"}[self[x, y]]
                if not ch:
                    gap += 1
                    continue
                row += (str(gap) if gap else "") + ch
                gap = 0
            rows.append(row + (str(gap) if gap else ""))
        return "/".join(rows) + " " + "xo"[self.to_move - 1]

    # ---- rules ----
    def _empty(self):
        return BOARD_MASK & ~(self.pieces[0] | self.pieces[1] | self.blockers)

    def legal_moves(self):
        if self.legal_moves_cache is None:
            own, empty = self.pieces[self.to_move - 1], self._empty()
            moves, reach = [], 0
            for src in _bits(own):
                reach |= RING1[src]
                moves += [(_xy(src), _xy(dst)) for dst in _bits(RING2[src] & empty)]
            moves += [("c", _xy(dst)) for dst in _bits(reach & empty)]
            self.legal_moves_cache = moves or ["pass"]
        return self.legal_moves_cache

    def move(self, desc):
        self.legal_moves_cache = None
        me = self.to_move - 1
        self.to_move = OTHER_PLAYER[self.to_move]
        if desc == "pass":
            return
        start, end = desc
        dst = 1 << _sq(end)
        assert dst & self._empty(), "destination %r is not an empty cell" % (end,)
        if start == "c":
            assert RING1[_sq(end)] & self.pieces[me], "no piece next to %r" % (end,)
        else:
            src = 1 << _sq(start)
            assert src & self.pieces[me], "no piece of the side to move on %r" % (start,)
            distance = Linf_distance(start, end)
            assert distance in (1, 2)
            if distance == 2:
                self.pieces[me] ^= src
        flipped = RING1[_sq(end)] & self.pieces[me ^ 1]
        self.pieces[me] |= dst | flipped
        self.pieces[me ^ 1] ^= flipped

    def result(self):
        """None while the game goes on, else the winner 1 / 2 (ataxx_rules.py:159-179, cpp get_board_result)."""
        count = [bin(self.pieces[0]).count("1"), bin(self.pieces[1]).count("1")]
        assert count[0] or count[1]
        empties = bin(self._empty()).count("1")
        if self.legal_moves() == ["pass"]:
            count[OTHER_PLAYER[self.to_move] - 1] += empties     # the side that can still move fills the board
            empties = 0
        elif count[0] == 0:
            return 2
        elif count[1] == 0:
            return 1
        if empties:
            return None
        return 1 if count[0] >= count[1] else 2

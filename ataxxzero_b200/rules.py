"""Host-side mirror of the reference rules interface over the C ABI.

Reference interfaces mirrored: cpp/movegen.hpp:10 ``movegen``, cpp/makemove.hpp:12-16
``makemove``, cpp/ataxx.hpp:36 ``set_board``, cpp/move.hpp:35 ``move_string``,
cpp/self_play_client.cpp:109 ``get_board_result``, :174-202 features; plus the coordinate
conventions of ataxx_rules.py / uai_interface.py:6-32 (``(x, y)`` with y = 0 at the top).
Moves are ``(from_sq, to_sq)`` tuples with ``from == to`` for clones, squares 0..48
(``sq = rank*7 + file``).  Batched calls run on the GPU; text helpers are host-only.
"""
import ctypes as C

import numpy as np

from ._native import AZ_FEATURES, AZ_MAX_MOVES, AzError, Position, check, lib

START_FEN = "x5o/7/3-3/2-1-2/3-3/7/o5x x"     # cpp/self_play_client.cpp:23
OPEN_FEN = "x5o/7/7/7/7/7/o5x x"               # ataxx_rules.AtaxxState.initial() / perft.py:18
NO_MOVE = (50, 50)                             # cpp/move.hpp:33

POSITION_DTYPE = np.dtype([("ply", "<i4"), ("turn", "<i4"), ("blockers", "<u8"), ("pieces", "<u8", (2,))])
assert POSITION_DTYPE.itemsize == C.sizeof(Position) == 32


def set_board(fen):
    """cpp/ataxx.cpp:14 set_board; raises ValueError with the reference's status code."""
    p = Position()
    rc = lib().az_set_board(C.byref(p), fen.encode())
    if rc != 0:
        raise ValueError("set_board(%r) failed with status %d" % (fen, rc))
    return p


def set_board_status(fen):
    p = Position()
    return lib().az_set_board(C.byref(p), fen.encode())


def fen(pos):
    buf = C.create_string_buffer(64)
    lib().az_fen(C.byref(pos), buf, 64)
    return buf.value.decode()


def pack_move(move):
    return int(move[0]) | (int(move[1]) << 8)


def unpack_move(m):
    return (int(m) & 0xff, int(m) >> 8)


def move_string(move):
    """cpp/move.cpp:11 move_string: ``"b6"`` / ``"a7b5"``."""
    buf = C.create_string_buffer(8)
    lib().az_move_string(pack_move(move), buf)
    return buf.value.decode()


def parse_move(text):
    m = lib().az_parse_move(text.encode())
    mv = unpack_move(m)
    if mv == NO_MOVE:
        raise ValueError("bad move string %r" % (text,))
    return mv


def sq_to_xy(sq):
    """square -> the Python reference's (x, y) with y = 0 at the top rank (SURVEY A-1)."""
    return (sq % 7, 6 - sq // 7)


def xy_to_sq(xy):
    return xy[0] + 7 * (6 - xy[1])


def to_reference_move(move):
    """(from, to) -> ataxx_rules move: ``("c", (x, y))`` or ``((sx, sy), (ex, ey))``."""
    if move[0] == move[1]:
        return ("c", sq_to_xy(move[1]))
    return (sq_to_xy(move[0]), sq_to_xy(move[1]))


def from_reference_move(desc):
    start, end = desc
    if start == "c":
        s = xy_to_sq(end)
        return (s, s)
    return (xy_to_sq(start), xy_to_sq(end))


def positions_array(positions):
    """list of Position / structured array -> contiguous structured ndarray (az_position[n])."""
    if isinstance(positions, np.ndarray) and positions.dtype == POSITION_DTYPE:
        return np.ascontiguousarray(positions)
    arr = np.zeros(len(positions), dtype=POSITION_DTYPE)
    for i, p in enumerate(positions):
        arr[i] = (p.ply, p.turn, p.blockers, (p.pieces[0], p.pieces[1]))
    return arr


def array_to_positions(arr):
    out = []
    for rec in arr:
        p = Position()
        p.ply, p.turn, p.blockers = int(rec["ply"]), int(rec["turn"]), int(rec["blockers"])
        p.pieces[0], p.pieces[1] = int(rec["pieces"][0]), int(rec["pieces"][1])
        out.append(p)
    return out


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def movegen_batch(ctx, positions):
    """movegen over a batch on the GPU -> list of move lists in reference order."""
    arr = positions_array(positions)
    n = len(arr)
    moves = np.zeros((max(n, 1), AZ_MAX_MOVES), dtype=np.uint16)
    counts = np.zeros(max(n, 1), dtype=np.int32)
    check(lib().az_movegen_batch(ctx.handle, _ptr(arr), n, _ptr(moves), _ptr(counts)))
    return [[unpack_move(m) for m in moves[i, :counts[i]]] for i in range(n)]


def makemove_batch(ctx, positions, moves):
    arr = positions_array(positions).copy()
    mv = np.array([pack_move(m) for m in moves], dtype=np.uint16)
    if len(mv) != len(arr):
        raise AzError(-1, "makemove_batch: %d positions but %d moves" % (len(arr), len(mv)))
    check(lib().az_makemove_batch(ctx.handle, _ptr(arr), _ptr(mv), len(arr)))
    return arr


def result_batch(ctx, positions):
    arr = positions_array(positions)
    out = np.zeros(max(len(arr), 1), dtype=np.int32)
    check(lib().az_result_batch(ctx.handle, _ptr(arr), len(arr), _ptr(out)))
    return out[:len(arr)]


def features_batch(ctx, positions):
    arr = positions_array(positions)
    out = np.zeros((max(len(arr), 1), 7, 7, 4), dtype=np.float32)
    assert out[0].size == AZ_FEATURES
    check(lib().az_features_batch(ctx.handle, _ptr(arr), len(arr), _ptr(out)))
    return out[:len(arr)]


def jump_bb_batch(ctx, bitboards):
    bb = np.ascontiguousarray(bitboards, dtype=np.uint64)
    s = np.zeros(max(len(bb), 1), dtype=np.uint64)
    d = np.zeros(max(len(bb), 1), dtype=np.uint64)
    check(lib().az_jump_bb_batch(ctx.handle, _ptr(bb), len(bb), _ptr(s), _ptr(d)))
    return s[:len(bb)], d[:len(bb)]


def perft_batch(ctx, positions, depth):
    arr = positions_array(positions)
    out = np.zeros(max(len(arr), 1), dtype=np.uint64)
    check(lib().az_perft_batch(ctx.handle, _ptr(arr), len(arr), int(depth), _ptr(out)))
    return out[:len(arr)]


def perft(ctx, position, depth):
    """Leaf count at ``depth`` (perft.py:5-16 semantics; a stuck side has no children)."""
    return int(perft_batch(ctx, [position], depth)[0])


def perft_last_stats(ctx):
    cn = C.c_uint64()
    la = C.c_int32()
    fr = C.c_int32()
    check(lib().az_perft_last_stats(ctx.handle, C.byref(cn), C.byref(la), C.byref(fr)))
    return {"count_nodes": cn.value, "launches": la.value, "frontier_items": fr.value}


PLAYOUT_DTYPE = np.dtype([("x", "<u8"), ("o", "<u8"), ("move", "<u4"), ("pad", "<u4")])


def random_playouts(ctx, start, n_games, max_plies=400, seed=0):
    """``n_games`` uniformly random games from ``start`` played on the GPU (generate_games.py --random-play).
    Returns (plies [n_games, max_plies] structured array of PLAYOUT_DTYPE, n_plies [n_games], result [n_games])."""
    plies = np.zeros((max(n_games, 1), max_plies), dtype=PLAYOUT_DTYPE)
    n_plies = np.zeros(max(n_games, 1), dtype=np.int32)
    result = np.zeros(max(n_games, 1), dtype=np.int32)
    check(lib().az_random_playouts(ctx.handle, C.byref(start), int(n_games), int(max_plies), int(seed) & (2**64 - 1), _ptr(plies),
                                   _ptr(n_plies), _ptr(result)))
    return plies[:n_games], n_plies[:n_games], result[:n_games]

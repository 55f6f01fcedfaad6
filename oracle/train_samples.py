"""CPU restatement of the reference's training-sample extraction (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/train.py:11-77 -- ``apply_symmetry`` (:11-23), ``apply_symmetry_to_move`` (:25-40) and the body
of ``get_sample_from_entries`` (:43-77) for one explicit pick (entry, ply, symmetry) instead of ``random`` draws -- with
the pieces it calls: ``engine.board_to_features`` (engine.py:53-73; BLOCKED_CELLS is empty, so plane 3 stays 0, SURVEY
App. B-1), ``engine.add_move_to_heatmap`` (engine.py:79-88), ``uai_interface.uai_decode_move`` (uai_interface.py:20-32).

Pinned by tests/golden/train_samples_golden.json, which tests/golden/make_train_golden.py produced by calling the
reference's own ``get_sample_from_entries`` (TensorFlow mocked out, ``random`` scripted).
"""
import numpy as np

FAR = [(a, b) for a in (-2, -1, 0, 1, 2) for b in (-2, -1, 0, 1, 2) if max(abs(a), abs(b)) == 2]
PLANE = {d: i for i, d in enumerate(FAR)}           # engine.py:75 over ataxx_rules.py:17-20


def decode_move(s):
    """uai_interface.py:20-32"""
    def sq(t):
        return "abcdefg".index(t[0].lower()), 6 - (int(t[1]) - 1)
    if s in ("pass", "none", "0000"):
        return "pass"
    if len(s) == 2:
        return "c", sq(s)
    return sq(s[:2]), sq(s[2:])


def encode_move(move):
    """uai_interface.py:6-18"""
    def sq(xy):
        return "%s%i" % ("abcdefg"[xy[0]], 7 - xy[1])
    if move == "pass":
        return "0000"
    if move[0] == "c":
        return sq(move[1])
    return sq(move[0]) + sq(move[1])


def symmetry_coord(index, xy):
    x, y = xy
    if index & 1:
        x = 6 - x
    if index & 2:
        y = 6 - y
    if index & 4:
        x, y = y, x
    return x, y


def apply_symmetry(index, arr):
    arr = np.array(arr).copy()
    if index & 1:
        arr = arr[::-1, :, :].copy()
    if index & 2:
        arr = arr[:, ::-1, :].copy()
    if index & 4:
        arr = np.swapaxes(arr, 0, 1).copy()
    return arr


def sample(entry, ply, symmetry):
    """(features int8 [7,7,4], policy float32 [7,7,17], value [1]) or None when the recorded move is a pass."""
    to_move = 1 if ply % 2 == 0 else 2
    board = entry["boards"][ply]
    move = entry["moves"][ply]
    if move == "pass":
        return None
    features = np.zeros((7, 7, 4), dtype=np.int8)
    for y in range(7):
        for x in range(7):
            features[x, y, 0] = 1
            piece = board[x + 7 * y]
            if piece:
                features[x, y, 1 if piece == to_move else 2] = 1
    value = [1 if entry["result"] == to_move else -1]
    features = apply_symmetry(symmetry, features)
    policy = np.zeros((7, 7, 17), dtype=np.float32)
    if "dists" not in entry:
        if not isinstance(move, str):
            move = ("c", tuple(move[1])) if move[0] == "c" else (tuple(move[0]), tuple(move[1]))
            move = encode_move(move)
        dist = {move: 1}
    else:
        dist = entry["dists"][ply]
    for mv, p in dist.items():
        mv = decode_move(mv)
        start, end = mv
        if start == "c":
            ex, ey = symmetry_coord(symmetry, end)
            policy[ex, ey, 16] += p
        else:
            (sx, sy), (ex, ey) = symmetry_coord(symmetry, start), symmetry_coord(symmetry, end)
            policy[ex, ey, PLANE[(ex - sx, ey - sy)]] += p
    return features, policy, value

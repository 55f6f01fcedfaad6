"""Training-sample extraction on the GPU: the step right after self-play (train.py:43-77 of the reference).

``load_entries`` reads JSON-lines game files (both record formats: the C++ client's, with ``"dists"``, and
generate_games.py's), ``pack_entries`` turns them into the binary ply table the kernel reads, ``draw`` picks
``(entry, ply, symmetry)`` triples the way ``get_sample_from_entries`` does, and ``extract`` returns the minibatch
``(features int8 [n,7,7,4], policy float32 [n,7,7,17], value float32 [n,1])`` computed by ``az_samples_extract``.
"""
import ctypes as C
import json
import random

import numpy as np

from . import _native
from ._native import AZ_FEATURES, AZ_LOGITS, AzError, check, lib
from .rules import from_reference_move, pack_move, parse_move

_vp = C.c_void_p
_native.register("az_samples_extract", C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_int, _vp, _vp, _vp])
_native.register("az_samples_extract_dev", C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp])


def load_entries(paths, shuffle=True, rng=random):
    """train.py:79-90."""
    entries = []
    for path in paths:
        with open(path) as f:
            entries += [json.loads(line) for line in f if line.strip()]
    if shuffle:
        rng.shuffle(entries)
    return entries


def _move_code(move):
    """"a7b5" / "b6" / [[sx,sy],[ex,ey]] / ["c",[x,y]] -> from | to << 8; None for a pass."""
    if move in ("pass", "0000", "none"):
        return None
    if isinstance(move, str):
        return pack_move(parse_move(move))
    start, end = move
    return pack_move(from_reference_move(("c", tuple(end)) if start == "c" else (tuple(start), tuple(end))))


class PackedGames:
    """Binary ply table: ``words`` (uint32), ``offsets[g][ply]`` word offsets, per-entry results / flags."""

    def __init__(self, entries):
        words, self.offsets, self.results, self.has_dists, self.random_ply, self.is_pass = [], [], [], [], [], []
        for entry in entries:
            offs, passes = [], []
            dists = entry.get("dists")
            for ply, (board, move) in enumerate(zip(entry["boards"], entry["moves"])):
                offs.append(len(words))
                x = o = 0
                for i, v in enumerate(board):
                    if v:
                        bit = 1 << ((i % 7) + 7 * (6 - i // 7))
                        if v == 1:
                            x |= bit
                        else:
                            o |= bit
                code = _move_code(move)
                passes.append(code is None)
                pairs = []
                if dists is not None:
                    for mv, p in dists[ply].items():
                        pairs.append((_move_code(mv), int(np.float32(p).view(np.uint32))))
                words += [x & 0xffffffff, x >> 32, o & 0xffffffff, o >> 32, (code or 0) | (len(pairs) << 16), 0]
                for mv, bits in pairs:
                    words += [mv, bits]
            self.offsets.append(offs)
            self.results.append(entry["result"])
            self.has_dists.append(dists is not None)
            self.random_ply.append(entry.get("random_ply"))
            self.is_pass.append(passes)
        self.words = np.asarray(words, dtype=np.uint32)

    def __len__(self):
        return len(self.offsets)


def pack_entries(entries):
    return PackedGames(entries)


def draw(packed, n, rng=random):
    """n picks ``(entry, ply, symmetry)`` as get_sample_from_entries makes them (train.py:44-52,59): a uniformly random
    game, a uniformly random ply of it (or ``random_ply + 1``), passes redrawn, a uniformly random symmetry."""
    picks = []
    while len(picks) < n:
        g = rng.randrange(len(packed))
        ply = rng.randrange(len(packed.offsets[g]))
        if packed.random_ply[g] is not None:
            ply = packed.random_ply[g] + 1
        if packed.is_pass[g][ply]:
            continue
        picks.append((g, ply, rng.randrange(8)))
    return picks


def _flat(packed):
    """Per-game arrays for the vectorised draw (built once per PackedGames)."""
    if getattr(packed, "_flat_cache", None) is None:
        n_plies = np.array([len(o) for o in packed.offsets], dtype=np.int64)
        start = np.concatenate([[0], np.cumsum(n_plies)[:-1]]).astype(np.int64)
        packed._flat_cache = {
            "n_plies": n_plies, "start": start,
            "offsets": np.array([w for o in packed.offsets for w in o], dtype=np.uint64),
            "is_pass": np.array([p for ps in packed.is_pass for p in ps], dtype=bool),
            "results": np.array(packed.results, dtype=np.uint32),
            "has_dists": np.array(packed.has_dists, dtype=np.uint32),
            "random_ply": np.array([-1 if r is None else int(r) for r in packed.random_ply], dtype=np.int64),
        }
    return packed._flat_cache


def draw_arrays(packed, n, rng):
    """``draw`` for a whole minibatch at once with a NumPy ``Generator``: the same distribution (uniform game, uniform ply of
    it or ``random_ply + 1``, passes redrawn, uniform symmetry; train.py:44-52,59), returned as the ``(offsets uint64 [n],
    meta uint32 [n])`` sample description ``az_samples_extract`` / ``az_trainer_step_picks`` take."""
    f = _flat(packed)
    offsets = np.empty(n, dtype=np.uint64)
    meta = np.empty(n, dtype=np.uint32)
    have = 0
    while have < n:
        m = (n - have) + (n - have) // 8 + 16
        g = rng.integers(0, len(packed), size=m)
        ply = rng.integers(0, f["n_plies"][g])
        rp = f["random_ply"][g]
        ply = np.where(rp >= 0, rp + 1, ply)
        sym = rng.integers(0, 8, size=m).astype(np.uint32)
        idx = f["start"][g] + ply
        keep = np.nonzero(~f["is_pass"][idx])[0][:n - have]
        k = len(keep)
        g, ply, sym, idx = g[keep], ply[keep], sym[keep], idx[keep]
        offsets[have:have + k] = f["offsets"][idx]
        meta[have:have + k] = (ply & 1).astype(np.uint32) | (f["results"][g] << 1) | (sym << 3) | (f["has_dists"][g] << 6)
        have += k
    return offsets, meta


def extract(ctx, packed, picks):
    n = len(picks)
    offsets = np.zeros(max(n, 1), dtype=np.uint64)
    meta = np.zeros(max(n, 1), dtype=np.uint32)
    for i, (g, ply, sym) in enumerate(picks):
        if not 0 <= sym < 8:
            raise AzError(-1, "symmetry index %r out of range" % (sym,))
        if packed.is_pass[g][ply]:
            raise AzError(-1, "pick %d is a pass move (train.py:53-54 skips those)" % i)
        to_move = ply % 2                      # train.py:50: player 1 moves on even plies
        offsets[i] = packed.offsets[g][ply]
        meta[i] = to_move | (int(packed.results[g]) << 1) | (sym << 3) | (int(packed.has_dists[g]) << 6)
    features = np.zeros((max(n, 1), 7, 7, 4), dtype=np.int8)
    policy = np.zeros((max(n, 1), 7, 7, 17), dtype=np.float32)
    value = np.zeros((max(n, 1), 1), dtype=np.float32)
    assert features[0].size == AZ_FEATURES and policy[0].size == AZ_LOGITS
    check(lib().az_samples_extract(ctx.handle, _vp(packed.words.ctypes.data), packed.words.size, _vp(offsets.ctypes.data),
                                   _vp(meta.ctypes.data), n, _vp(features.ctypes.data), _vp(policy.ctypes.data), _vp(value.ctypes.data)))
    return features[:n], policy[:n], value[:n]


def minibatch(ctx, packed, n, rng=random):
    """``n`` fresh samples: what one training step of train.py consumes."""
    return extract(ctx, packed, draw(packed, n, rng))

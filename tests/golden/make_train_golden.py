"""Generates tests/golden/train_samples_golden.json by running the REFERENCE's own sample extraction
(/root/reference/train.py:43-77 get_sample_from_entries) on a few recorded games, with TensorFlow mocked out (only
model.py touches it at import time) and ``random`` scripted so that every (entry, ply, symmetry) pick is known.

    python tests/golden/make_train_golden.py      # needs /root/reference; run in the build container only

Game records: two C++-format self-play records (with "dists") produced by the compiled reference client through
oracle/_ref/libref.so's greedy self-play, and two Python-format random-play records from the reference's ataxx_rules.
"""
import json
import os
import random
import sys
import unittest.mock

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.modules["tensorflow"] = unittest.mock.MagicMock()
import train as ref_train            # noqa: E402  the reference module
import ataxx_rules as ref_rules      # noqa: E402
import uai_interface as ref_uai      # noqa: E402


def cpp_style_games():
    """Self-play records in the C++ client's format, built with the reference's own rules + a synthetic visit table."""
    rng = random.Random(5)
    games = []
    for g in range(2):
        board = ref_rules.AtaxxState.initial()
        entry = {"boards": [], "dists": [], "moves": []}
        while board.result() is None and len(entry["moves"]) < 60 + 40 * g:
            moves = board.legal_moves()
            picks = rng.sample(moves, min(len(moves), rng.randrange(1, 7)))
            counts = [rng.randrange(1, 300) for _ in picks]
            total = sum(counts)
            dist = {ref_uai.uai_encode_move(m): c / total for m, c in zip(picks, counts)}
            entry["boards"].append(list(board.board))
            entry["dists"].append(dict(sorted(dist.items())))
            entry["moves"].append(ref_uai.uai_encode_move(picks[0]))
            board.move(picks[0])
        entry["result"] = board.result() or 1
        games.append(entry)
    return games


def python_style_games():
    rng = random.Random(9)
    games = []
    for g in range(2):
        board = ref_rules.AtaxxState.initial()
        entry = {"boards": [], "moves": []}
        while board.result() is None and len(entry["moves"]) < 400:
            m = rng.choice(board.legal_moves())
            entry["boards"].append(list(board.board))
            entry["moves"].append(m)
            board.move(m)
        entry["result"] = board.result()
        games.append(json.loads(json.dumps(entry)))        # tuples -> lists, as train.py sees them after json.loads
    return games


class Scripted:
    """Stands in for the `random` module inside train.py."""
    def __init__(self, entry_index, ply, symmetry):
        self.entry_index, self.queue = entry_index, [ply, symmetry]

    def choice(self, seq):
        return seq[self.entry_index]

    def randrange(self, n):
        v = self.queue.pop(0)
        assert 0 <= v < n
        return v


def main():
    entries = cpp_style_games() + python_style_games()
    rng = random.Random(1)
    samples = []
    for e_idx, entry in enumerate(entries):
        for _ in range(24):
            ply, sym = rng.randrange(len(entry["boards"])), rng.randrange(8)
            ref_train.random = Scripted(e_idx, ply, sym)
            feats, policy, value = ref_train.get_sample_from_entries(entries)
            nz = [[int(i), int(j), int(k), float(policy[i, j, k]).hex()] for i, j, k in zip(*policy.nonzero())]
            samples.append({"entry": e_idx, "ply": ply, "symmetry": sym, "features": feats.astype(int).reshape(-1).tolist(),
                            "policy_nonzero": nz, "value": int(value[0])})
    ref_train.random = random
    out = {"entries": entries, "samples": samples,
           "note": "produced by /root/reference/train.py get_sample_from_entries with scripted random; policy values are float32 hex"}
    path = os.path.join(HERE, "train_samples_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %s: %d entries, %d samples, %d bytes" % (path, len(entries), len(samples), os.path.getsize(path)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the COMPILED REFERENCE (oracle/_ref, built by
oracle/Makefile from /root/reference/cpp) and, for the Python twin, by importing
/root/reference/ataxx_rules.py.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The fixtures are what travels to the GPU box; /root/reference does not.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.cpu import OPEN_FEN, START_FEN, Oracle, Reference, ref_perft  # noqa: E402

ALT_FEN = "x5o/7/2-1-2/7/2-1-2/7/o5x x"   # the commented-out start position in cpp/ataxx.cpp:11


def pos_dict(p):
    return {"turn": p.turn, "ply": p.ply, "blockers": p.blockers, "x": p.pieces[0], "o": p.pieces[1]}


def fen_of(p):
    rows = []
    for rank in range(6, -1, -1):
        row, run = "", 0
        for file in range(7):
            bit = 1 << (rank * 7 + file)
            c = "x" if p.pieces[0] & bit else "o" if p.pieces[1] & bit else "-" if p.blockers & bit else ""
            if not c:
                run += 1
                continue
            row += (str(run) if run else "") + c
            run = 0
        rows.append(row + (str(run) if run else ""))
    return "/".join(rows) + " " + "xo"[p.turn]


def main():
    ref = Reference()
    orc = Oracle()       # only used for its fixture evaluators (function pointers) and feature hashing
    rng = random.Random(20240607)

    # ---- rules: positions from seeded random playouts, everything computed by the reference ----
    rules = []
    for g in range(12):
        fen = [START_FEN, OPEN_FEN, ALT_FEN][g % 3]
        p = ref.set_board(fen)
        if g % 2:
            p.turn = 1 - p.turn if g % 4 == 1 else p.turn
        ply = 0
        while True:
            moves = ref.movegen(p)
            res = ref.result(p)
            if ply % 4 == g % 4 or res != 0 or len(moves) > 100:
                entry = pos_dict(p)
                entry["moves"] = [m[0] | (m[1] << 8) for m in moves]
                entry["result"] = res
                entry["board_json"] = ref.board_json(p)
                entry["single_jump_own"] = ref.single_jump_bb(p.pieces[p.turn])
                entry["double_jump_own"] = ref.double_jump_bb(p.pieces[p.turn])
                if moves:
                    m = moves[rng.randrange(len(moves))]
                    q = ref.makemove(p, m)
                    entry["after"] = {"move": m[0] | (m[1] << 8), "string": ref.move_string(m), "pos": pos_dict(q)}
                rules.append(entry)
            if res != 0 or not moves:
                break
            p = ref.makemove(p, rng.choice(moves))
            ply += 1
    rings = {"single": [ref.single_ring(s) for s in range(49)], "double": [ref.double_ring(s) for s in range(49)]}
    fens = {}
    for fen in [START_FEN, OPEN_FEN, ALT_FEN, "startpos", "x5o/7/7/7/7/7/o5x o", "7/7/7/7/7/7/7", "x5o/7/7/7/7/7/o5x z",
                "x5o/7/7", "x5o/7/7/7/7/7/o5x x extra", "xxxxxxx/ooooooo/7/7/7/7/7 o", "x5o/7/7/7/7/7/o5y x", "",
                "x5o/7/7/7/7/7/o5x  x", "x5o/7/7/7/7/7/o5x ", "X5O/7/3-3/2-1-2/3-3/7/O5X O", "x6o/7/7/7/7/7/o5x x"]:
        rc = ref.set_board_rc(fen)
        fens[fen] = {"status": rc, "pos": pos_dict(ref.set_board(fen)) if rc == 0 else None}
    json.dump({"source": "oracle/_ref/libref.so (reference cpp/ compiled in place)", "positions": rules,
               "rings": rings, "fens": fens}, open(os.path.join(HERE, "rules_golden.json"), "w"))
    print("rules:", len(rules), "positions; max moves", max(len(e["moves"]) for e in rules))

    # ---- populate(): features + priors for a handful of positions (probe evaluator) ----
    pop = []
    for e in rules[::7] + [t for t in rules if t["result"] != 0][:6]:
        p = ref.set_board(START_FEN)
        p.turn, p.ply, p.blockers = e["turn"], e["ply"], e["blockers"]
        p.pieces[0], p.pieces[1] = e["x"], e["o"]
        feats, moves, priors, value = ref.populate(p, orc.probe_eval_ptr)
        if feats is None:
            pop.append({"pos": pos_dict(p), "terminal_value": value})
            continue
        pop.append({"pos": pos_dict(p), "features_nonzero": [int(i) for i in feats.reshape(-1).nonzero()[0]],
                    "moves": [m[0] | (m[1] << 8) for m in moves], "priors_hex": [float(x).hex() for x in priors],
                    "value": value})
    json.dump({"evaluator": "ao_probe_eval (oracle/ataxx_oracle.c; SURVEY App. D)", "entries": pop},
              open(os.path.join(HERE, "populate_golden.json"), "w"))
    print("populate:", len(pop))

    # ---- perft ----
    perft = {"survey_app_c": {OPEN_FEN: [16, 256, 6460, 155888, 4752668, 141865520, 5023479496, 176821532236],
                              START_FEN: [16, 256, 5948, 133264, 3639856, 97538324, 3044225260, 94167145692]},
             "perft_ref": {}, "batch": []}
    for fen in (OPEN_FEN, START_FEN, ALT_FEN):
        perft["perft_ref"][fen] = [ref_perft(fen, d, 8)[0] for d in range(1, 8)]
        print("perft", fen, perft["perft_ref"][fen])
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    for e in rules[::5][:48]:
        p = ref.set_board(START_FEN)
        p.turn, p.ply, p.blockers = e["turn"], e["ply"], e["blockers"]
        p.pieces[0], p.pieces[1] = e["x"], e["o"]
        # perft over the reference's movegen/makemove through the shim (recursion written here)
        def walk(q, d):
            mv = ref.movegen(q)
            if d == 1:
                return len(mv)
            return sum(walk(ref.makemove(q, m), d - 1) for m in mv)
        perft["batch"].append({"pos": pos_dict(p), "depth2": walk(p, 2), "depth3": walk(p, 3)})
    json.dump(perft, open(os.path.join(HERE, "perft_golden.json"), "w"))

    # ---- MCTS (reference lines 1..582 + injected evaluator, noise off) ----
    mid = None
    mcts = {"searches": [], "games": [], "umap": []}
    while True:        # ply 60 of a seeded random playout that is still ongoing
        p = ref.set_board(START_FEN)
        for _ in range(60):
            if ref.result(p) != 0:
                break
            p = ref.makemove(p, rng.choice(ref.movegen(p)))
        if p.ply == 60 and ref.result(p) == 0:
            break
    mid = fen_of(p)
    for ev, ptr in (("probe", orc.probe_eval_ptr), ("uniform", orc.uniform_eval_ptr)):
        for fen in (START_FEN, OPEN_FEN, mid):
            for visits in (50, 400, 800) + ((5000,) if ev == "probe" and fen == START_FEN else ()):
                dist, evals = ref.search(fen, visits, ptr)
                mcts["searches"].append({"evaluator": ev, "fen": fen, "visits": visits, "evals": evals,
                                         "moves": [m[0] | (m[1] << 8) for m, v, w in dist],
                                         "edge_visits": [v for m, v, w in dist],
                                         "edge_total_hex": [float(w).hex() for m, v, w in dist]})
        for visits in (60, 200):
            plies, result, evals = ref.selfplay_greedy(START_FEN, visits, 400, ptr)
            mcts["games"].append({"evaluator": ev, "fen": START_FEN, "visits": visits, "result": result, "evals": evals,
                                  "plies": [{"visits": pl["visits"], "played": pl["played"][0] | (pl["played"][1] << 8)}
                                            for pl in plies]})
            print("game", ev, visits, "plies", len(plies), "result", result, "evals", evals)
    for e in rules[::9]:
        mv = [(m & 0xff, m >> 8) for m in e["moves"]]
        if not mv:
            continue
        fresh, b = ref.umap_order(mv, False)
        again, _ = ref.umap_order(mv, True)
        mcts["umap"].append({"moves": e["moves"], "fresh": fresh, "reinserted": again, "buckets": b})
    mcts["midgame_fen"] = mid
    json.dump(mcts, open(os.path.join(HERE, "mcts_golden.json"), "w"))
    print("mcts:", len(mcts["searches"]), "searches,", len(mcts["games"]), "games,", len(mcts["umap"]), "umap orders")

    # ---- Python twin: perft.py numbers + a random game through ataxx_rules (no blockers) ----
    sys.path.insert(0, "/root/reference")
    import ataxx_rules
    py = {"perft": [], "games": []}
    state = ataxx_rules.AtaxxState.initial()

    def py_perft(s, d):
        ens = [s.copy()]
        for _ in range(d):
            nxt = []
            for b in ens:
                for m in b.legal_moves():
                    c = b.copy()
                    c.move(m)
                    nxt.append(c)
            ens = nxt
        return len(ens)
    py["perft"] = [py_perft(state, d) for d in range(1, 5)]
    def compact(m):
        return "c%d%d" % m[1] if m[0] == "c" else "%d%d%d%d" % (m[0] + m[1])
    for g in range(2):
        s = ataxx_rules.AtaxxState.initial()
        game = {"boards": [], "moves": [], "legal_sorted": [], "fens": []}
        while s.result() is None:
            legal = s.legal_moves()
            game["boards"].append(list(s.board))
            game["fens"].append(s.fen())
            game["legal_sorted"].append(sorted(compact(m) for m in legal))
            m = rng.choice(sorted(legal, key=json.dumps))
            game["moves"].append(m)
            s.move(m)
        game["result"] = s.result()
        game["final_board"] = list(s.board)
        py["games"].append(game)
    json.dump(py, open(os.path.join(HERE, "python_rules_golden.json"), "w"))
    print("python twin: perft", py["perft"], "games", [len(g["moves"]) for g in py["games"]])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".json"):
            print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()

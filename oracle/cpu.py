"""ctypes bindings for the CPU checker (TEST INFRASTRUCTURE ONLY).

``Oracle``   -> oracle/libataxx_oracle.so : our C restatement (oracle/ataxx_oracle.c)
``Reference`` -> oracle/_ref/libref.so     : the reference's own compiled code behind
                                              our shim (oracle/ref_harness/ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libataxx_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")
REF_PERFT = os.path.join(HERE, "_ref", "perft_ref")
REF_CLIENT_SO = os.path.join(HERE, "_ref", "self_play_client.so")

START_FEN = "x5o/7/3-3/2-1-2/3-3/7/o5x x"      # cpp/self_play_client.cpp:23
OPEN_FEN = "x5o/7/7/7/7/7/o5x x"                # the position perft.py:18 uses


class Position(C.Structure):
    """cpp/ataxx.hpp:28-34"""
    _fields_ = [("ply", C.c_int32), ("turn", C.c_int32), ("blockers", C.c_uint64),
                ("pieces", C.c_uint64 * 2)]

    def key(self):
        return (self.turn, self.blockers, self.pieces[0], self.pieces[1])

    def clone(self):
        q = Position()
        C.pointer(q)[0] = self
        return q


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float))
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)


def build(force=False):
    """(Re)build the checker libraries with oracle/Makefile."""
    sources = [os.path.join(HERE, f) for f in ("ataxx_oracle.c", "tree_model.c", "ataxx_oracle.h")]
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(f) for f in sources):
        subprocess.check_call(["make", "-s", "-C", HERE, "libataxx_oracle.so"])
    if os.path.isdir("/root/reference/cpp") and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class _Base:
    prefix = ""

    def _fn(self, name, restype, *argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    # --- shared surface (same semantics on both libraries) ---
    def set_board(self, fen):
        p = Position()
        rc = self._set_board(C.byref(p), fen.encode())
        if rc != 0:
            raise ValueError("set_board(%r) -> %d" % (fen, rc))
        return p

    def set_board_rc(self, fen):
        p = Position()
        return self._set_board(C.byref(p), fen.encode())

    def movegen(self, pos):
        f = (C.c_int32 * 256)()
        t = (C.c_int32 * 256)()
        n = self._movegen(C.byref(pos), f, t)
        return [(f[i], t[i]) for i in range(n)]

    def makemove(self, pos, move):
        q = pos.clone()
        self._makemove(C.byref(q), move[0], move[1])
        return q

    def result(self, pos):
        return self._result(C.byref(pos))

    def legal_move(self, pos, move):
        return bool(self._legal_move(C.byref(pos), move[0], move[1]))

    def move_string(self, move):
        buf = C.create_string_buffer(8)
        self._move_string(move[0], move[1], buf)
        return buf.value.decode()

    def board_json(self, pos):
        out = (C.c_int32 * 49)()
        self._board_json(C.byref(pos), out)
        return list(out)


class Oracle(_Base):
    prefix = "ao_"

    def __init__(self):
        build()
        self.lib = C.CDLL(ORACLE_SO)
        fn = self._fn
        self._set_board = fn("set_board", C.c_int, C.POINTER(Position), C.c_char_p)
        self._movegen = fn("movegen", C.c_int, C.POINTER(Position), _i32p, _i32p)
        self._makemove = fn("makemove", None, C.POINTER(Position), C.c_int, C.c_int)
        self._result = fn("result", C.c_int, C.POINTER(Position))
        self._legal_move = fn("legal_move", C.c_int, C.POINTER(Position), C.c_int, C.c_int)
        self._move_string = fn("move_string", C.c_int, C.c_int, C.c_int, C.c_char_p)
        self._board_json = fn("board_json", None, C.POINTER(Position), _i32p)
        self.single_ring = fn("single_ring", C.c_uint64, C.c_int)
        self.double_ring = fn("double_ring", C.c_uint64, C.c_int)
        self.single_jump_bb = fn("single_jump_bb", C.c_uint64, C.c_uint64)
        self.double_jump_bb = fn("double_jump_bb", C.c_uint64, C.c_uint64)
        self.policy_index = fn("policy_index", C.c_int, C.c_int, C.c_int)
        self._features = fn("features", None, C.POINTER(Position), _f32p)
        self._perft = fn("perft", C.c_uint64, C.POINTER(Position), C.c_int)
        self._perft_mt = fn("perft_mt", C.c_uint64, C.POINTER(Position), C.c_int, C.c_int)
        self._priors = fn("priors", None, _f32p, _i32p, _i32p, C.c_int, _f64p)
        self._umap_order = fn("umap_order", C.c_int, _i32p, _i32p, C.c_int, C.c_int, _i32p)
        self._search_fen = fn("search_fen", C.c_int, C.c_char_p, C.c_int, C.c_int, _i32p, _i32p, _i32p,
                              C.POINTER(C.c_long))
        self._mcts_new = fn("mcts_new", C.c_void_p, C.POINTER(Position), C.c_void_p, C.c_void_p)
        self._mcts_free = fn("mcts_free", None, C.c_void_p)
        self._mcts_step = fn("mcts_step", None, C.c_void_p)
        self._mcts_root_visits = fn("mcts_root_visits", C.c_int, C.c_void_p)
        self._mcts_eval_count = fn("mcts_eval_count", C.c_long, C.c_void_p)
        self._mcts_tie_count = fn("mcts_tie_count", C.c_long, C.c_void_p)
        self._mcts_root_dist = fn("mcts_root_dist", C.c_int, C.c_void_p, _i32p, _i32p, _i32p, _f64p, _f64p)
        self._mcts_play = fn("mcts_play", C.c_int, C.c_void_p, C.c_int, C.c_int)
        self._mcts_root_position = fn("mcts_root_position", None, C.c_void_p, C.POINTER(Position))
        self.probe_eval_ptr = C.cast(self.lib.ao_probe_eval, C.c_void_p)
        self.uniform_eval_ptr = C.cast(self.lib.ao_uniform_eval, C.c_void_p)

    def features(self, pos):
        out = np.zeros(196, dtype=np.float32)
        self._features(C.byref(pos), out.ctypes.data_as(_f32p))
        return out.reshape(7, 7, 4)

    def perft(self, pos, depth, threads=1):
        if threads > 1:
            return int(self._perft_mt(C.byref(pos), depth, threads))
        return int(self._perft(C.byref(pos), depth))

    def priors(self, logits, moves):
        logits = np.ascontiguousarray(logits, dtype=np.float32).reshape(833)
        n = len(moves)
        f = (C.c_int32 * max(n, 1))(*[m[0] for m in moves])
        t = (C.c_int32 * max(n, 1))(*[m[1] for m in moves])
        out = np.zeros(max(n, 1), dtype=np.float64)
        self._priors(logits.ctypes.data_as(_f32p), f, t, n, out.ctypes.data_as(_f64p))
        return out[:n]

    def umap_order(self, moves, start_buckets=0):
        n = len(moves)
        f = (C.c_int32 * max(n, 1))(*[m[0] for m in moves])
        t = (C.c_int32 * max(n, 1))(*[m[1] for m in moves])
        out = (C.c_int32 * max(n, 1))()
        b = self._umap_order(f, t, n, start_buckets, out)
        return list(out[:n]), b

    def probe_eval(self, feats):
        """feats [B,196]/[B,7,7,4] f32 -> (logits [B,833] f32, values [B] f32) via ao_probe_eval."""
        feats = np.ascontiguousarray(feats, dtype=np.float32).reshape(-1, 196)
        logits = np.zeros((feats.shape[0], 833), dtype=np.float32)
        values = np.zeros(feats.shape[0], dtype=np.float32)
        fn = self.lib.ao_probe_eval_batch
        fn.restype = None
        fn.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        fn(feats.ctypes.data_as(_f32p), feats.shape[0], logits.ctypes.data_as(_f32p), values.ctypes.data_as(_f32p))
        return logits, values

    def search(self, fen, visits, evaluator="probe"):
        f = (C.c_int32 * 256)()
        t = (C.c_int32 * 256)()
        v = (C.c_int32 * 256)()
        evals = C.c_long()
        n = self._search_fen(fen.encode(), visits, 1 if evaluator == "uniform" else 0, f, t, v, C.byref(evals))
        if n < 0:
            raise ValueError(fen)
        return [((f[i], t[i]), v[i]) for i in range(n)], evals.value

    class Tree:
        def __init__(self, oracle, pos, evaluator="probe", py_eval=None):
            self.o = oracle
            self._cb = None
            if py_eval is not None:
                def _cb(ctx, feats, logits, value):
                    fa = np.ctypeslib.as_array(feats, shape=(196,))
                    lo, va = py_eval(fa)
                    np.ctypeslib.as_array(logits, shape=(833,))[:] = lo
                    value[0] = va
                self._cb = EVAL_FN(_cb)
                fn = C.cast(self._cb, C.c_void_p)
            else:
                fn = oracle.uniform_eval_ptr if evaluator == "uniform" else oracle.probe_eval_ptr
            self.h = oracle._mcts_new(C.byref(pos), fn, None)

        def step(self):
            self.o._mcts_step(self.h)

        def search(self, visits):
            while self.o._mcts_root_visits(self.h) < visits:
                self.o._mcts_step(self.h)

        @property
        def root_visits(self):
            return self.o._mcts_root_visits(self.h)

        @property
        def evals(self):
            return self.o._mcts_eval_count(self.h)

        @property
        def ties(self):
            return self.o._mcts_tie_count(self.h)

        def root_position(self):
            p = Position()
            self.o._mcts_root_position(self.h, C.byref(p))
            return p

        def dist(self):
            f = (C.c_int32 * 256)()
            t = (C.c_int32 * 256)()
            v = (C.c_int32 * 256)()
            w = (C.c_double * 256)()
            p = (C.c_double * 256)()
            n = self.o._mcts_root_dist(self.h, f, t, v, w, p)
            return [((f[i], t[i]), v[i], w[i], p[i]) for i in range(n)]

        def play(self, move):
            return self.o._mcts_play(self.h, move[0], move[1])

        def close(self):
            if self.h:
                self.o._mcts_free(self.h)
                self.h = None

        def __del__(self):
            try:
                self.close()
            except Exception:
                pass

    def tree(self, pos, evaluator="probe", py_eval=None):
        return Oracle.Tree(self, pos, evaluator, py_eval)


class Reference(_Base):
    """The reference's compiled code (only where /root/reference was available at build time)."""
    prefix = "ref_"

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        build()
        self.lib = C.CDLL(REF_SO)
        fn = self._fn
        self._set_board = fn("set_board", C.c_int, C.POINTER(Position), C.c_char_p)
        self._movegen = fn("movegen", C.c_int, C.POINTER(Position), _i32p, _i32p)
        self._makemove = fn("makemove", None, C.POINTER(Position), C.c_int, C.c_int)
        self._result = fn("result", C.c_int, C.POINTER(Position))
        self._legal_move = fn("legal_move", C.c_int, C.POINTER(Position), C.c_int, C.c_int)
        self._move_string = fn("move_string", C.c_int, C.c_int, C.c_int, C.c_char_p)
        self._board_json = fn("board_json", None, C.POINTER(Position), _i32p)
        self.single_ring = fn("single_ring", C.c_uint64, C.c_int)
        self.double_ring = fn("double_ring", C.c_uint64, C.c_int)
        self.single_jump_bb = fn("single_jump_bb", C.c_uint64, C.c_uint64)
        self.double_jump_bb = fn("double_jump_bb", C.c_uint64, C.c_uint64)
        self._populate = fn("populate", C.c_int, C.POINTER(Position), C.c_void_p, C.c_void_p, _f32p, _i32p, _i32p,
                            _f64p, _f64p)
        self._umap_order = fn("umap_order", C.c_int, _i32p, _i32p, C.c_int, C.c_int, _i32p)
        self._search = fn("search", C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, _i32p, _i32p, _i32p,
                          _f64p, C.POINTER(C.c_long))
        self._selfplay_greedy = fn("selfplay_greedy", C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   _i32p, _i32p, _i32p, _i32p, C.POINTER(C.c_long))

    def populate(self, pos, eval_ptr):
        feats = np.zeros(196, dtype=np.float32)
        f = (C.c_int32 * 256)()
        t = (C.c_int32 * 256)()
        pr = (C.c_double * 256)()
        val = C.c_double()
        n = self._populate(C.byref(pos), eval_ptr, None, feats.ctypes.data_as(_f32p), f, t, pr, C.byref(val))
        if n < 0:
            return None, None, None, val.value
        return feats.reshape(7, 7, 4), [(f[i], t[i]) for i in range(n)], np.array(pr[:n]), val.value

    def umap_order(self, moves, reinsert=False):
        n = len(moves)
        f = (C.c_int32 * max(n, 1))(*[m[0] for m in moves])
        t = (C.c_int32 * max(n, 1))(*[m[1] for m in moves])
        out = (C.c_int32 * max(n, 1))()
        b = self._umap_order(f, t, n, int(reinsert), out)
        return list(out[:n]), b

    def search(self, fen, visits, eval_ptr):
        f = (C.c_int32 * 256)()
        t = (C.c_int32 * 256)()
        v = (C.c_int32 * 256)()
        w = (C.c_double * 256)()
        evals = C.c_long()
        n = self._search(fen.encode(), visits, eval_ptr, None, f, t, v, w, C.byref(evals))
        if n < 0:
            raise ValueError(fen)
        return [((f[i], t[i]), v[i], w[i]) for i in range(n)], evals.value

    def selfplay_greedy(self, fen, visits, max_plies, eval_ptr):
        nm = (C.c_int32 * max_plies)()
        dist = (C.c_int32 * (max_plies * 256))()
        played = (C.c_int32 * max_plies)()
        result = C.c_int32()
        evals = C.c_long()
        plies = self._selfplay_greedy(fen.encode(), visits, max_plies, eval_ptr, None, nm, dist, played,
                                      C.byref(result), C.byref(evals))
        out = []
        for p in range(plies):
            out.append({"n_moves": nm[p], "visits": [dist[p * 256 + i] for i in range(nm[p])],
                        "played": (played[p] & 0xff, played[p] >> 8)})
        return out, result.value, evals.value


def ref_perft(fen, depth, threads=1):
    """Run oracle/_ref/perft_ref; returns (nodes, seconds)."""
    out = subprocess.check_output([REF_PERFT, fen, str(depth), str(threads)]).decode().split()
    return int(out[1]), float(out[3])

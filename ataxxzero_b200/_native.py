"""ctypes binding of libataxxzero.so (the C ABI declared in include/ataxxzero.h).

This is the reference-side binding style (link.py:6-32 uses ctypes too).  The library is
the only compute path: if it is missing, or no B200 is visible, calls raise -- there is no
Python/NumPy fallback anywhere in this package.
"""
import ctypes as C
import os
import weakref

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AZ_LIB_PATH") or os.path.join(HERE, "libataxxzero.so")

AZ_MAX_MOVES = 256
AZ_FEATURES = 196
AZ_LOGITS = 833


class AzError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libataxxzero error %d: %s" % (code, message))
        self.code = code


class Position(C.Structure):
    """az_position == cpp/ataxx.hpp:28-34 Position (32 bytes)."""
    _fields_ = [("ply", C.c_int32), ("turn", C.c_int32), ("blockers", C.c_uint64), ("pieces", C.c_uint64 * 2)]

    def key(self):
        return (self.turn, self.blockers, self.pieces[0], self.pieces[1])

    def clone(self):
        q = Position()
        C.pointer(q)[0] = self
        return q

    def __repr__(self):
        return "Position(turn=%d, x=%#x, o=%#x, blockers=%#x, ply=%d)" % (
            self.turn, self.pieces[0], self.pieces[1], self.blockers, self.ply)


_lib = None

_vp = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "az_last_error": (C.c_char_p, []),
    "az_version": (C.c_char_p, []),
    "az_create": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "az_destroy": (None, [_vp]),
    "az_device_count": (C.c_int, []),
    "az_sync": (C.c_int, [_vp]),
    "az_stream": (_vp, [_vp]),
    "az_set_board": (C.c_int, [C.POINTER(Position), C.c_char_p]),
    "az_move_string": (C.c_int, [C.c_uint16, C.c_char_p]),
    "az_parse_move": (C.c_uint16, [C.c_char_p]),
    "az_fen": (C.c_int, [C.POINTER(Position), C.c_char_p, C.c_size_t]),
    "az_movegen_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "az_makemove_batch": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "az_result_batch": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "az_features_batch": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "az_jump_bb_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "az_perft_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "az_perft": (C.c_int, [_vp, C.POINTER(Position), C.c_int, C.POINTER(C.c_uint64)]),
    "az_perft_batch_dev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "az_perft_last_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "az_random_playouts": (C.c_int, [_vp, C.POINTER(Position), C.c_int, C.c_int, C.c_uint64, _vp, _vp, _vp]),
}


def lib():
    """Load libataxxzero.so (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or python ataxxzero_b200/build.py). There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the ABI and the header disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def register(name, restype, argtypes):
    """Used by sibling modules to declare the entry points they bind."""
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(code):
    if code != 0:
        raise AzError(code, lib().az_last_error().decode(errors="replace"))
    return code


class Context:
    """Owns an az_context (one per GPU / process rank)."""

    def __init__(self, device=0, seed=0):
        self._h = _vp()
        check(lib().az_create(int(device), int(seed) & (2**64 - 1), C.byref(self._h)))
        self.device = device
        self._dependents = weakref.WeakSet()         # pools / trainers created on this context: they hold pointers into it

    def adopt(self, obj):
        """Objects of the library that live on this context register here: ``close()`` closes them first, so a pool or trainer that
        outlives its context (an exception unwinding, interpreter shutdown order) never touches freed memory."""
        self._dependents.add(obj)

    @property
    def handle(self):
        if not self._h:
            raise AzError(-3, "context already destroyed")
        return self._h

    def sync(self):
        check(lib().az_sync(self.handle))

    @property
    def stream(self):
        return lib().az_stream(self.handle)

    def close(self):
        if self._h:
            for obj in list(self._dependents):
                try:
                    obj.close()
                except Exception:
                    pass
            lib().az_destroy(self._h)
            self._h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

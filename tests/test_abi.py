"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, exports every
symbol include/ataxxzero.h declares, refuses to compute without a GPU, and its host-only
helpers (FEN / move codec) agree with the oracle and the reference's codec vectors."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_golden


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ataxxzero.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(az_[a-z0-9_]+|launch_threads|get_workload|complete_workload|shutdown)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(native):
    names = header_symbols()
    assert len(names) >= 20
    handle = C.CDLL(os.path.join(ROOT, "ataxxzero_b200", "libataxxzero.so"))
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, "declared in include/ataxxzero.h but not exported: %s" % missing


def test_binding_signatures_cover_header(native):
    from ataxxzero_b200 import _native
    import ataxxzero_b200.rules  # noqa: F401  (registers nothing extra, but must import cleanly)
    for opt in ("net", "search", "selfplay", "link", "train_data", "trainer"):
        try:
            __import__("ataxxzero_b200." + opt)
        except ImportError:
            pass
    unbound = [n for n in header_symbols() if n not in _native._SIGNATURES]
    assert not unbound, "no ctypes signature for: %s" % unbound


def test_no_cpu_fallback(native):
    """Without a GPU the product must fail loudly, never compute on the host."""
    import ataxxzero_b200
    if native.az_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(ataxxzero_b200.AzError) as e:
        ataxxzero_b200.Context()
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is the checker, never the product path."""
    pkg = os.path.join(ROOT, "ataxxzero_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(dirpath, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
            assert "ataxx_oracle" not in text and "libref.so" not in text and "net_numpy" not in text, f


def test_fen_parser_matches_reference_codes(native):
    from ataxxzero_b200 import rules
    g = load_golden("rules_golden.json")
    for fen, want in g["fens"].items():
        assert rules.set_board_status(fen) == want["status"], fen
        if want["status"] == 0:
            p = rules.set_board(fen)
            assert (p.turn, p.blockers, p.pieces[0], p.pieces[1]) == \
                (want["pos"]["turn"], want["pos"]["blockers"], want["pos"]["x"], want["pos"]["o"])
            q = rules.set_board(rules.fen(p))           # fen round trip (ataxx_rules.py:193)
            assert q.key() == p.key()


def test_move_codec_vectors(native):
    """uai_interface.py:34-39 self-test vectors + cpp/move.cpp:11-21 strings."""
    from ataxxzero_b200 import rules
    assert rules.to_reference_move(rules.parse_move("f2")) == ("c", (5, 5))
    assert rules.to_reference_move(rules.parse_move("c3d5")) == ((2, 4), (3, 2))
    assert rules.move_string(rules.from_reference_move(("c", (4, 3)))) == "e4"
    assert rules.move_string(rules.from_reference_move(((4, 3), (2, 5)))) == "e4c2"
    for bad in ("", "h1", "a8", "a1b", "a1b2c"):
        with pytest.raises(ValueError):
            rules.parse_move(bad)
    g = load_golden("rules_golden.json")
    for e in g["positions"][:200]:
        if "after" in e:
            m = (e["after"]["move"] & 0xff, e["after"]["move"] >> 8)
            assert rules.move_string(m) == e["after"]["string"]
            assert rules.parse_move(e["after"]["string"]) == m

"""The device tree algorithm (visited list + one candidate per node, csrc/az_tree.cu) restated on the CPU
(oracle/tree_model.c) must reproduce the reference-order search (oracle ao_mcts_*, pinned to the compiled
reference in test_oracle_pinned.py) bit for bit: visit counts, edge_total_score and priors at the root, across
re-rooting, with exact ties (uniform evaluator) and with adversarially perturbed priors."""
import ctypes as C
import json
import os

import pytest

from oracle import cpu as ocpu

MIDGAME = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mcts_golden.json"))).get("midgame_fen", ocpu.START_FEN)
FENS = [ocpu.START_FEN, ocpu.OPEN_FEN, MIDGAME, "7/7/7/7/ooooooo/ooooooo/xxxxxxx x", "x5o/7/7/7/7/7/o5x o"]


@pytest.fixture(scope="module")
def lib():
    o = ocpu.Oracle()
    o.lib.tm_selftest.restype = C.c_long
    o.lib.tm_selftest.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
    o.lib.ao_set_prior_perturbation.argtypes = [C.c_int]
    yield o.lib
    o.lib.ao_set_prior_perturbation(0)


def run(lib, fen, visits, evaluator, plays, force_slow=0, perturb=0):
    counters = (C.c_long * 4)()
    lib.ao_set_prior_perturbation(perturb)
    try:
        bad = lib.tm_selftest(fen.encode(), visits, evaluator, plays, force_slow, counters)
    finally:
        lib.ao_set_prior_perturbation(0)
    return bad, list(counters)


@pytest.mark.parametrize("fen", FENS)
@pytest.mark.parametrize("evaluator", [0, 1])
def test_model_matches_reference_order_search(lib, fen, evaluator):
    bad, counters = run(lib, fen, 800, evaluator, plays=6)
    assert bad == 0, (fen, evaluator, counters)


def test_forced_slow_path_is_equivalent(lib):
    for fen in FENS[:3]:
        for evaluator in (0, 1):
            bad, counters = run(lib, fen, 400, evaluator, plays=3, force_slow=1)
            assert bad == 0 and counters[0] > 0, (fen, evaluator, counters)


@pytest.mark.parametrize("perturb", [1, 2, 3])
def test_perturbed_priors(lib, perturb):
    """quantised priors (exact ties between different moves) and adjacent doubles (different priors whose products
    with sqrt(1+N) collide): the candidate shortcut must fall back to the full scan exactly when it has to"""
    total_near = 0
    for fen in FENS[:3]:
        bad, counters = run(lib, fen, 1500, 0, plays=4, perturb=perturb)
        assert bad == 0, (fen, perturb, counters)
        total_near += counters[2]
    if perturb >= 2:
        assert total_near > 0, "the adjacent-double perturbation never produced a product collision"


def test_deep_search(lib):
    bad, counters = run(lib, MIDGAME, 20000, 0, plays=1)
    assert bad == 0, counters


def test_speculative_evaluation_tick_model(lib):
    """design evidence for the speculative single-tree search (csrc/az_tree.cu, cache mode): the number of net round trips
    with the top-k children of every consumed node evaluated in the same batch, against one round trip per visit"""
    lib.tm_spec_sim.restype = C.c_long
    lib.tm_spec_sim.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
    evals = C.c_long()
    base = lib.tm_spec_sim(MIDGAME.encode(), 3000, 2, 0, C.byref(evals))
    assert base in (3000, 3001) and evals.value == base
    for k, least in ((1, 2.0), (16, 3.0)):
        ticks = lib.tm_spec_sim(MIDGAME.encode(), 3000, 2, k, C.byref(evals))
        assert 3000 / ticks > least, (k, ticks)
        assert evals.value <= 3001 * (k + 1)

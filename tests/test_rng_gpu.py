"""Statistical parity of the two random rows of the path (a11 Dirichlet root noise,
self_play_client.cpp:250-271; a18 move sampling proportional to visits, :495-506) plus the host-visible
consequences of the compact node layout (forced full-scan selection, legacy thread counts, row-count checks).

The reference draws from a racy global std::mt19937; its streams cannot be reproduced, so the check is on the
DISTRIBUTIONS its code defines: g_k ~ Gamma(0.15, 1), noise = g / sum(g) ~ Dirichlet(0.15), P <- 0.25 noise + 0.75 P,
and P(move i) = n_i / N."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALPHA, WEIGHT = 0.15, 0.25     # self_play_client.cpp:36-37


def _lib():
    from ataxxzero_b200 import _native
    lib = _native.lib()
    lib.az_debug_gamma.restype = C.c_int
    lib.az_debug_gamma.argtypes = [C.c_void_p, C.c_double, C.c_uint64, C.c_int, C.c_void_p]
    lib.az_debug_sample_moves.restype = C.c_int
    lib.az_debug_sample_moves.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_void_p]
    lib.az_debug_exp.restype = C.c_int
    lib.az_debug_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.az_debug_div.restype = C.c_int
    lib.az_debug_div.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def test_paired_division_is_the_library_division(ctx):
    """select_action's two divisions per edge (sqrt(1+N)/(1+n) and W/n, self_play_client.cpp:316-323) run as one
    interleaved block (az_tree.cu div_pair): bit-identical to __ddiv_rn on 2 M operand pairs -- PUCT-shaped operands
    (square roots of counts over counts, score sums over counts), zero and tiny numerators, huge and subnormal quotients
    (those take the fallback) -- and equal to the host's IEEE division."""
    from ataxxzero_b200 import _native
    rng = np.random.default_rng(11)
    n = 1 << 20
    counts = rng.integers(1, 1 << 23, n).astype(np.float64)
    small = rng.integers(1, 900, n).astype(np.float64)
    a1 = np.sqrt(1.0 + rng.integers(0, 1 << 23, n))
    b1 = np.where(rng.random(n) < 0.7, 1.0 + small, 1.0 + counts)
    b2 = np.where(rng.random(n) < 0.7, small, counts)
    a2 = b2 * rng.random(n)                                    # a total score between 0 and n
    a2[rng.random(n) < 0.05] = 0.0                             # an edge that only ever lost
    # the variant the kernel runs on exactly these operand shapes (only the score sum is range-checked there)
    puct_in = np.ascontiguousarray(np.stack([a1, b1, a2, b2], axis=1))
    tiny = rng.random(n) < 0.02                                # one visit whose value was (v+1)/2 of a float v just above -1
    puct_in[tiny, 2] = (np.float32(-1.0) + np.float32(2.0) ** rng.integers(-24, -1, tiny.sum()).astype(np.float32) + 1.0).astype(np.float64) / 2.0
    puct_out = np.empty_like(puct_in)
    _native.check(_lib().az_debug_div(ctx.handle, C.c_void_p(puct_in.ctypes.data), len(puct_in), 1, C.c_void_p(puct_out.ctypes.data)))
    pb = puct_out.view(np.uint64)
    assert np.array_equal(pb[:, 0], pb[:, 2]) and np.array_equal(pb[:, 1], pb[:, 3])
    assert np.array_equal((puct_in[:, 0] / puct_in[:, 1]).view(np.uint64), pb[:, 0])
    assert np.array_equal((puct_in[:, 2] / puct_in[:, 3]).view(np.uint64), pb[:, 1])
    m = 1 << 16                                                # the corners of the fast path
    ea = np.concatenate([2.0 ** rng.uniform(-1074, -960, m), 2.0 ** rng.uniform(900, 1023, m), rng.uniform(0, 1, m), np.zeros(m)])
    eb = np.concatenate([rng.integers(1, 1 << 23, 2 * m).astype(np.float64), 2.0 ** rng.uniform(-900, 900, 2 * m)])
    a1 = np.concatenate([a1, ea]); b1 = np.concatenate([b1, eb])
    a2 = np.concatenate([a2, ea[::-1]]); b2 = np.concatenate([b2, eb[::-1]])
    inp = np.ascontiguousarray(np.stack([a1, b1, a2, b2], axis=1))
    out = np.empty_like(inp)
    _native.check(_lib().az_debug_div(ctx.handle, C.c_void_p(inp.ctypes.data), len(inp), 0, C.c_void_p(out.ctypes.data)))
    bits = out.view(np.uint64)
    assert np.array_equal(bits[:, 0], bits[:, 2]) and np.array_equal(bits[:, 1], bits[:, 3])
    with np.errstate(all="ignore"):
        assert np.array_equal((a1 / b1).view(np.uint64), bits[:, 2]) and np.array_equal((a2 / b2).view(np.uint64), bits[:, 3])


def test_straight_line_exp_is_the_library_exp(ctx):
    """The softmax numerators (self_play_client.cpp:210-214) come from a written-out copy of CUDA's exp(double) fast path
    (az_tree.cu exp_inline, interleaved with the sequential sums): bit-identical to exp() on 4 M floats -- the logit
    range, every binade up to the range limit, the limit itself, and values beyond it (which take the library call) --
    and within 1 ulp of the host's exp(), the function the reference calls."""
    from ataxxzero_b200 import _native
    rng = np.random.default_rng(5)
    x = np.concatenate([
        rng.normal(0, 3, 1 << 21), rng.uniform(-708, 708, 1 << 20), rng.uniform(-30, 30, 1 << 19),
        (rng.uniform(1, 2, 1 << 18) * 2.0 ** rng.integers(-40, 10, 1 << 18) * rng.choice([-1, 1], 1 << 18)),
        np.array([0.0, -0.0, 707.99994, -707.99994, 708.0, -708.0, 709.5, -745.0, 800.0, -800.0, 1e-30, -1e-30, 88.7, -103.9,
                  np.inf, -np.inf, np.nan]),
    ]).astype(np.float32)
    out = np.empty((len(x), 2), np.float64)
    _native.check(_lib().az_debug_exp(ctx.handle, C.c_void_p(x.ctypes.data), len(x), C.c_void_p(out.ctypes.data)))
    a, b = out[:, 0].view(np.uint64), out[:, 1].view(np.uint64)
    assert np.array_equal(a, b), "first mismatch at x = %r" % x[np.nonzero(a != b)[0][:4]]
    finite = np.isfinite(x) & (np.abs(x) < 700)
    host = np.exp(x[finite].astype(np.float64))
    ulp = np.abs(out[finite, 1].view(np.int64) - host.view(np.int64))
    assert ulp.max() <= 1


def test_gamma_sampler_distribution(ctx):
    """std::gamma_distribution<double>(0.15, 1.0) (:252): mean = variance = 0.15, and the whole CDF (KS test)."""
    from scipy import stats
    from ataxxzero_b200 import _native
    n = 200000
    out = np.zeros(n, dtype=np.float64)
    _native.check(_lib().az_debug_gamma(ctx.handle, ALPHA, 12345, n, C.c_void_p(out.ctypes.data)))
    assert (out > 0).all() and np.isfinite(out).all()
    se_mean = (ALPHA / n) ** 0.5
    assert abs(out.mean() - ALPHA) < 4 * se_mean, (out.mean(), se_mean)
    # Var of the sample variance of Gamma(a): (mu4 - sigma^4)/n with mu4 = 3a^2 + 6a
    se_var = ((3 * ALPHA ** 2 + 6 * ALPHA - ALPHA ** 2) / n) ** 0.5
    assert abs(out.var() - ALPHA) < 4 * se_var, (out.var(), se_var)
    ks = stats.kstest(out, "gamma", args=(ALPHA,))
    assert ks.pvalue > 1e-4, ks
    # a different seed gives a different stream, the same seed the same one
    again, other = np.zeros(1000), np.zeros(1000)
    _native.check(_lib().az_debug_gamma(ctx.handle, ALPHA, 12345, 1000, C.c_void_p(again.ctypes.data)))
    _native.check(_lib().az_debug_gamma(ctx.handle, ALPHA, 54321, 1000, C.c_void_p(other.ctypes.data)))
    assert (again == out[:1000]).all() and (other != out[:1000]).any()


def test_dirichlet_root_noise_marginals(ctx, oracle):
    """20 000 fresh roots of the opening position, uniform evaluation: (P_noisy - 0.75 P) / 0.25 must be a
    Dirichlet(0.15 x 16) draw -- components sum to 1, each marginal is Beta(0.15, 2.25) (mean 1/16, known
    variance, KS test), different games get different draws."""
    from scipy import stats
    from ataxxzero_b200 import search
    games = 20000
    with search.Pool(ctx, games, 1, eval_mode=search.EVAL_EXTERNAL, noise=True, auto_play=False, node_capacity=6, seed=99) as pool:
        feats = pool.collect()
        assert len(feats) == games
        pool.provide(np.zeros((games, 833), dtype=np.float32), np.zeros(games, dtype=np.float32))
        pool.collect()                                    # consumes the evaluations: priors + noise are in the roots now
        roots = [pool.root(g) for g in range(0, games, 1)]
    L = len(roots[0]["moves"])
    assert L == 16
    noisy = np.array([r["prior"] for r in roots])
    clean = 1.0 / L                                        # uniform logits: every legal move has prior 1/16
    d = (noisy - (1 - WEIGHT) * clean) / WEIGHT
    assert np.abs(d.sum(axis=1) - 1.0).max() < 1e-9
    assert d.min() > -1e-12
    a0 = ALPHA * L
    var = ALPHA * (a0 - ALPHA) / (a0 * a0 * (a0 + 1))
    se = (var / games) ** 0.5
    for k in range(L):
        col = np.clip(d[:, k], 0, 1)
        assert abs(col.mean() - 1.0 / L) < 4.5 * se, (k, col.mean(), se)
        assert abs(col.var() - var) < 0.12 * var, (k, col.var(), var)
        ks = stats.kstest(col, "beta", args=(ALPHA, a0 - ALPHA))
        assert ks.pvalue > 1e-5, (k, ks)
    assert len({tuple(np.round(row, 12)) for row in d[:500]}) == 500       # independent streams per game


def test_move_sampling_proportional_to_visits(ctx):
    """fixed root visit table, 50 000 independent draws of the kernel's sampler: chi-square against n_i / N; moves
    without an edge are never played"""
    from scipy import stats
    from ataxxzero_b200 import _native
    visits = np.array([0, 5, 0, 120, 3, 472, 0, 1, 60, 0, 0, 139, 1, 0], dtype=np.int32)
    n = 50000
    out = np.zeros(n, dtype=np.int32)
    _native.check(_lib().az_debug_sample_moves(ctx.handle, C.c_void_p(visits.ctypes.data), len(visits), 777, n, C.c_void_p(out.ctypes.data)))
    counts = np.bincount(out, minlength=len(visits))
    assert counts[visits == 0].sum() == 0
    live = visits > 0
    expected = visits[live] / visits.sum() * n
    chi = stats.chisquare(counts[live], expected)
    assert chi.pvalue > 1e-4, (chi, counts, expected)
    # single-edge table: always that edge
    one = np.array([0, 0, 800, 0], dtype=np.int32)
    _native.check(_lib().az_debug_sample_moves(ctx.handle, C.c_void_p(one.ctypes.data), 4, 1, 1000, C.c_void_p(out.ctypes.data)))
    assert (out[:1000] == 2).all()


def test_selfplay_first_moves_follow_visit_table(ctx):
    """end to end through generate_game: with noise off every game of a pool searches the opening identically, so the
    first moves of many games are draws from ONE visit distribution (the one a single search tree reports)"""
    from scipy import stats
    from ataxxzero_b200 import model, net, rules, search
    from oracle import net_numpy
    # a net with trained-scale statistics: its sharper priors and varied values spread the 200 visits of the opening over
    # several moves (the random-init net puts them all on one, FPU = 0)
    net.load_weights(ctx, model.Network(*net_numpy.trained_scale_weights(seed=0)))
    games, visits = 2048, 200
    os.environ["AZ_REQ_CAP"] = "0"          # serve every request every tick: the games stay in lockstep up to their first move
    try:
        pool = search.Pool(ctx, games, visits, eval_mode=search.EVAL_BF16, noise=False, auto_play=True, seed=6)
    finally:
        del os.environ["AZ_REQ_CAP"]
    with pool:
        for _ in range(4 * visits):
            pool.selfplay_ticks(1)
            if pool.root(0)["position"].ply >= 1:
                break
        after_first = [pool.root(g)["position"] for g in range(games)]
    assert all(p.ply == 1 for p in after_first)
    with search.Pool(ctx, 1, visits, eval_mode=search.EVAL_BF16, noise=False, auto_play=False) as tree:
        tree.run()
        root = tree.root(0)
    weights = np.array(root["visits"], dtype=np.float64)
    assert weights.sum() == root["root_visits"] >= visits
    start = rules.set_board(rules.START_FEN)
    after = rules.makemove_batch(ctx, rules.positions_array([start] * len(root["moves"])), root["moves"])
    index = {(int(a["pieces"][0]), int(a["pieces"][1])): i for i, a in enumerate(after)}
    assert len(index) == len(root["moves"])               # distinct moves lead to distinct positions at the opening
    counts = np.zeros(len(root["moves"]))
    for p in after_first:
        counts[index[(p.pieces[0], p.pieces[1])]] += 1
    live = weights > 0
    assert counts[~live].sum() == 0
    chi = stats.chisquare(counts[live], weights[live] / weights.sum() * games)
    assert live.sum() >= 3, weights
    assert chi.pvalue > 1e-4, (chi, counts, weights)


def test_forced_full_scan_selection_matches_goldens(ctx, oracle):
    """AZ_TREE_FORCE_SLOW=1 resolves every candidate by the full scan over all moves without an edge (the path taken when
    two different priors round to the same product): visit counts and total scores must not change."""
    from ataxxzero_b200 import rules, search
    from oracle.cpu import START_FEN
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mcts_golden.json")))
    fens = [START_FEN, golden.get("midgame_fen", START_FEN)]
    os.environ["AZ_TREE_FORCE_SLOW"] = "1"
    try:
        for fen in fens:
            uniform = lambda f: (np.zeros((len(f), 833), dtype=np.float32), np.zeros(len(f), dtype=np.float32))
            for name, evalfn in (("probe", oracle.probe_eval), ("uniform", uniform)):
                with search.Pool(ctx, 1, 600, eval_mode=search.EVAL_EXTERNAL) as pool:
                    pool.set_root(0, rules.set_board(fen))
                    pool.run_external(evalfn)
                    got = pool.root(0)
                tree = oracle.tree(oracle.set_board(fen), name)
                tree.search(600)
                want = tree.dist()
                tree.close()
                assert got["visits"] == [w[1] for w in want], (fen, name)
                assert [float(x).hex() for x in got["total_score"]] == [float(w[2]).hex() for w in want], (fen, name)
    finally:
        del os.environ["AZ_TREE_FORCE_SLOW"]


def test_provide_rejects_wrong_row_count(ctx):
    from ataxxzero_b200 import AzError, search
    with search.Pool(ctx, 8, 10, eval_mode=search.EVAL_EXTERNAL) as pool:
        feats = pool.collect()
        assert len(feats) == 8
        with pytest.raises(AzError):
            pool.provide(np.zeros((5, 833), dtype=np.float32), np.zeros(5, dtype=np.float32))
        pool.provide(np.zeros((8, 833), dtype=np.float32), np.zeros(8, dtype=np.float32))   # still answerable afterwards


def test_legacy_abi_accepts_fewer_threads(tmp_path, ctx, oracle):
    """self_play_client.cpp:683-706 takes any thread_count <= 2*buffer_entries; with buffer_entries <= threads < 2x only one
    workload is outstanding at a time and every workload is a full buffer"""
    import ctypes
    from ataxxzero_b200 import link, model, net
    from test_selfplay_gpu import replay_and_check
    net.load_weights(ctx, model.Network.random_init(seed=0))
    out = str(tmp_path / "legacy_few.json")
    entries, threads = 16, 23
    bufs = [np.zeros((entries, 7, 7, 4), dtype=np.float32) for _ in (0, 1)]
    link.launch_threads(out.encode(), 16, ctypes.c_void_p(bufs[0].ctypes.data), ctypes.c_void_p(bufs[1].ctypes.data), entries, threads)
    try:
        seen = []
        for _ in range(8000):
            i = link.get_workload()
            seen.append(i)
            assert bufs[i][..., 0].min() == 1.0
            p, v = net.forward(ctx, bufs[i], net.BF16)
            link.complete_workload(i, ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(v.ctypes.data))
            if os.path.exists(out) and open(out).read().count("\n") >= 3:
                break
        assert seen[:4] == [0, 1, 0, 1]
    finally:
        link.shutdown()
    lines = open(out).read().splitlines()
    assert len(lines) >= 3
    for ln in lines:
        replay_and_check(oracle, ln, 16)

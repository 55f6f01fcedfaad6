"""Training-sample extraction (train.py:43-77).  CPU: the oracle restatement against goldens produced by
the reference's own get_sample_from_entries.  GPU: az_samples_extract against the goldens and the oracle, bit-exact."""
import random

import numpy as np
import pytest

from conftest import load_golden


def _dense(sample):
    policy = np.zeros((7, 7, 17), dtype=np.float32)
    for i, j, k, hexval in sample["policy_nonzero"]:
        policy[i, j, k] = np.float32(float.fromhex(hexval))
    return np.asarray(sample["features"], dtype=np.int8).reshape(7, 7, 4), policy, sample["value"]


def test_oracle_pinned_to_reference_goldens():
    from oracle import train_samples
    g = load_golden("train_samples_golden.json")
    assert len(g["samples"]) == 96
    for s in g["samples"]:
        want_f, want_p, want_v = _dense(s)
        got_f, got_p, got_v = train_samples.sample(g["entries"][s["entry"]], s["ply"], s["symmetry"])
        assert np.array_equal(got_f, want_f) and got_f.dtype == np.int8
        assert np.array_equal(got_p.view(np.uint32), want_p.view(np.uint32))        # float32 bit patterns
        assert got_v == [want_v]
        assert abs(float(got_p.sum()) - 1.0) < 1e-3                                 # train.py:70


def test_packing_and_draw_follow_the_reference_rules():
    from ataxxzero_b200 import train_data
    g = load_golden("train_samples_golden.json")
    packed = train_data.pack_entries(g["entries"])
    assert len(packed) == 4 and packed.has_dists == [True, True, False, False]
    picks = train_data.draw(packed, 500, random.Random(3))
    assert all(0 <= ply < len(g["entries"][e]["boards"]) and 0 <= s < 8 for e, ply, s in picks)
    assert {e for e, _, _ in picks} == {0, 1, 2, 3} and {s for _, _, s in picks} == set(range(8))
    entry = dict(g["entries"][2], random_ply=4)
    packed = train_data.pack_entries([entry])
    assert {ply for _, ply, _ in train_data.draw(packed, 20, random.Random(1))} == {5}      # train.py:47-49


@pytest.mark.gpu
def test_gpu_extraction_matches_reference_goldens(ctx):
    from ataxxzero_b200 import train_data
    g = load_golden("train_samples_golden.json")
    packed = train_data.pack_entries(g["entries"])
    picks = [(s["entry"], s["ply"], s["symmetry"]) for s in g["samples"]]
    feats, policy, value = train_data.extract(ctx, packed, picks)
    assert feats.shape == (96, 7, 7, 4) and feats.dtype == np.int8 and policy.shape == (96, 7, 7, 17) and value.shape == (96, 1)
    for i, s in enumerate(g["samples"]):
        want_f, want_p, want_v = _dense(s)
        assert np.array_equal(feats[i], want_f), i
        assert np.array_equal(policy[i].view(np.uint32), want_p.view(np.uint32)), i
        assert value[i, 0] == want_v


@pytest.mark.gpu
def test_gpu_extraction_from_our_selfplay_records(tmp_path, ctx):
    """Records written by the self-play kernels (n/N distributions) through the JSON files train.py would read:
    every ply x every symmetry of a few games, against the oracle, bit-exact."""
    from ataxxzero_b200 import model, net, search, train_data
    from oracle import train_samples
    net.load_weights(ctx, model.Network.random_init(seed=0))
    out = str(tmp_path / "games.json")
    with search.Pool(ctx, 32, 30, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=11) as pool:
        pool.selfplay(out, target_games=6, max_seconds=240)
    entries = train_data.load_entries([out], shuffle=False)[:6]
    packed = train_data.pack_entries(entries)
    picks = [(e, ply, sym) for e in range(len(entries)) for ply in range(len(entries[e]["boards"])) for sym in range(8)]
    feats, policy, value = train_data.extract(ctx, packed, picks)
    assert len(picks) > 200
    for i in range(0, len(picks), 3):
        e, ply, sym = picks[i]
        f, p, v = train_samples.sample(entries[e], ply, sym)
        assert np.array_equal(feats[i], f) and np.array_equal(policy[i].view(np.uint32), p.view(np.uint32)) and value[i, 0] == v[0]
    assert np.allclose(policy.reshape(len(picks), -1).sum(axis=1), 1.0, atol=1e-3)
    # edge cases: empty batch, bad symmetry
    f0, p0, v0 = train_data.extract(ctx, packed, [])
    assert len(f0) == len(p0) == len(v0) == 0
    import ataxxzero_b200 as az
    with pytest.raises(az.AzError):
        train_data.extract(ctx, packed, [(0, 0, 8)])

"""Training side (last row).  CPU part: the checker itself is pinned -- the fp32 PyTorch restatement of model.py's
network (tests/torch_train_reference.py) equals the NumPy oracle in inference form and its loss follows model.py:81-96; the
sampling / encoding rules of train.py:43-77 reproduce the reference goldens; cli/train.py refuses to run without the GPU
library.  The hand-written training step itself is checked on the GPU (tests/test_train_gpu.py)."""
import json
import random

import pytest

import numpy as np

from conftest import load_golden


def test_torch_network_matches_numpy_restatement():
    import torch
    from ataxxzero_b200 import model
    import torch_train_reference as train
    from oracle import net_numpy
    network = model.Network.random_init(seed=3, filters=16, blocks=2)
    network.bn = net_numpy.randomize_bn(network.bn, seed=4)
    net = train.build_torch_network(16, 2)
    train.load_into(net, network)
    net.eval()
    feats = net_numpy.random_features(5, seed=6)
    want_p, want_v = net_numpy.forward(feats, network.conv, network.bn, dtype=np.float64)
    with torch.no_grad():
        logits, value = net(torch.from_numpy(np.ascontiguousarray(feats.transpose(0, 3, 1, 2))))
    assert np.abs(logits.numpy().reshape(5, 7, 7, 17) - want_p).max() < 1e-5
    assert np.abs(value.numpy() - want_v).max() < 1e-5
    back = train.export(net)                                   # round trip through the .npy layout
    assert all(np.array_equal(a, b) for a, b in zip(back.conv, network.conv))
    assert all(np.allclose(a, b) for a, b in zip(back.bn, network.bn))


def test_loss_definition_and_sampling_goldens(tmp_path):
    import torch
    from ataxxzero_b200 import model
    import torch_train_reference as train
    g = load_golden("train_samples_golden.json")
    games = tmp_path / "games.json"
    games.write_text("\n".join(json.dumps(e) for e in g["entries"] * 4) + "\n")
    # host minibatches (no GPU here) agree with the reference goldens: same picks -> same tensors
    fn = train.host_minibatch_fn(g["entries"])
    class Scripted(random.Random):
        pass
    for s in g["samples"][:20]:
        seq = iter([s["entry"], s["ply"], s["symmetry"]])
        rng = type("R", (), {"randrange": lambda self, n: next(seq)})()
        f, p, v = fn(1, rng)
        assert f[0].reshape(-1).tolist() == s["features"] and v[0, 0] == s["value"]
        want = np.zeros((7, 7, 17), np.float32)
        for i, j, k, h in s["policy_nonzero"]:
            want[i, j, k] = np.float32(float.fromhex(h))
        assert np.array_equal(p[0], want)
    # loss terms on a tiny net against a NumPy evaluation of model.py:81-96
    net = train.build_torch_network(8, 1)
    train.load_into(net, model.Network.random_init(seed=1, filters=8, blocks=1))
    net.eval()
    batch = train.to_torch_batch(fn(16, random.Random(2)), torch.device("cpu"))
    with torch.no_grad():
        pl, vl, reg = train.loss_terms(net, *batch)
        logits, out = net(batch[0])
    lg = logits.numpy().astype(np.float64)
    lse = np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1)) + lg.max(1)
    want_pl = float(np.mean(-(batch[1].numpy().reshape(16, -1) * (lg - lse[:, None])).sum(1)))
    want_vl = float(np.mean((batch[2].numpy() - out.numpy()) ** 2))
    want_reg = 0.0001 * sum(0.5 * float((p.detach().numpy().astype(np.float64) ** 2).sum()) for p in net.parameters())
    assert abs(float(pl) - want_pl) < 1e-5 and abs(float(vl) - want_vl) < 1e-6 and abs(float(reg) - want_reg) < 1e-7


def test_vectorised_draw_follows_the_sampling_rules():
    """train_data.draw_arrays (one NumPy call per minibatch) against train.py:44-52: only plies of the table, never a pass,
    ``random_ply + 1`` when the record names one, every symmetry, and the same meta word ``extract`` builds for that pick."""
    from ataxxzero_b200 import train_data
    g = load_golden("train_samples_golden.json")
    entries = [dict(e) for e in g["entries"]]
    entries[1] = dict(entries[1], random_ply=3)
    packed = train_data.pack_entries(entries)
    offsets, meta = train_data.draw_arrays(packed, 6000, np.random.default_rng(5))
    where = {int(packed.offsets[gi][ply]): (gi, ply) for gi in range(len(packed)) for ply in range(len(packed.offsets[gi]))}
    seen_games, seen_sym = set(), set()
    for off, m in zip(offsets.tolist(), meta.tolist()):
        gi, ply = where[off]
        assert not packed.is_pass[gi][ply]
        if gi == 1:
            assert ply == 4
        sym = (m >> 3) & 7
        assert m == (ply % 2) | (int(packed.results[gi]) << 1) | (sym << 3) | (int(packed.has_dists[gi]) << 6)
        seen_games.add(gi)
        seen_sym.add(sym)
    assert seen_games == set(range(len(packed))) and seen_sym == set(range(8))
    counts = np.bincount([where[o][0] for o in offsets.tolist()], minlength=len(packed))
    assert counts.min() > 0.8 * 6000 / len(packed)            # uniform over games (train.py:45), not over plies


def test_train_cli_needs_the_gpu_library(tmp_path):
    """No CPU training path: without a visible B200 the CLI stops at az_create instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from ataxxzero_b200 import AzError, model
    from ataxxzero_b200.cli import train
    g = load_golden("train_samples_golden.json")
    games = tmp_path / "games.json"
    games.write_text("\n".join(json.dumps(e) for e in g["entries"] * 2) + "\n")
    old = tmp_path / "model-001.npy"
    model.Network.random_init(seed=5, blocks=1).save(str(old))
    with pytest.raises((AzError, ImportError)):
        train.main(["--games", str(games), "--old-path", str(old), "--new-path", str(tmp_path / "model-002.npy"), "--steps", "2"])
    assert not (tmp_path / "model-002.npy").exists()


@pytest.mark.gpu
def test_training_on_gpu_minibatches_and_looper_round(tmp_path, ctx, oracle):
    """One looper iteration end to end on the GPU: self-play with model-001 (128 filters x 12 blocks, device trees + tensor-core
    net), minibatches from az_samples_extract, a few optimiser steps, model-002.npy written and loadable by the net kernel."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from ataxxzero_b200 import model, net
    prefix = tmp_path / "run"
    (prefix / "models").mkdir(parents=True)
    (prefix / "games").mkdir()
    model.Network.random_init(seed=0).save(str(prefix / "models" / "model-001.npy"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "ataxxzero_b200.cli.looper", "--prefix", str(prefix), "--visits", "24", "--game-count", "24",
           "--buffer-size", "32", "--poll-seconds", "1", "--iterations", "1", "--training-steps-const", "10", "--training-steps-linear", "0",
           "--train-command", "%s -m ataxxzero_b200.cli.train --minibatch-size 64" % sys.executable]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BEGINNING TRAINING" in r.stdout and "Saved model to" in r.stdout
    new = model.Network.load(str(prefix / "models" / "model-002.npy"))
    assert new.filters == 128 and new.blocks == 12
    net.load_weights(ctx, new)                                  # the trained file feeds straight back into the hot path
    logits, values = net.forward(ctx, np.zeros((2, 7, 7, 4), dtype=np.float32) + np.array([1, 0, 0, 0], dtype=np.float32))
    assert np.isfinite(logits).all() and np.isfinite(values).all()

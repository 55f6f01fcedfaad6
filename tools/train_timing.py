"""Steady-state timing of the hand-written training step (reference shape: 128 filters x 12 blocks, minibatch 512)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import ataxxzero_b200 as az
from ataxxzero_b200 import model, trainer
from test_train_gpu import synthetic_batch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = az.Context(0)
tr = trainer.Trainer(ctx, model.Network.random_init(seed=1), max_batch=N)
batch = synthetic_batch(N, 2)
for _ in range(3):
    tr.train(*batch, learning_rate=1e-3)
t0 = time.perf_counter(); l0 = tr.launches
dev_ms = 0.0
for _ in range(STEPS):
    tr.train(*batch, learning_rate=1e-3)
    dev_ms += tr.last_step_ms
dt = (time.perf_counter() - t0) / STEPS
print("native step %d boards x 12 blocks: %.3f ms/step end to end (host batch in, losses out), %.3f ms/step on the device, %.0f samples/s, %d launches/step" % (
    N, dt * 1e3, dev_ms / STEPS, N / dt, (tr.launches - l0) // STEPS))

# the same step on samples of device-resident games (az_trainer_step_picks): what cli/train.py runs per step
import json
from ataxxzero_b200 import train_data
entries = json.load(open(os.path.join(ROOT, "tests", "golden", "train_samples_golden.json")))["entries"] * 64
packed = train_data.pack_entries(entries)
tr.set_games(packed)
rng = np.random.default_rng(3)
for _ in range(3):
    tr.train_picks(*train_data.draw_arrays(packed, N, rng), learning_rate=1e-3)
t0 = time.perf_counter(); draw_s = 0.0
for _ in range(STEPS):
    d0 = time.perf_counter()
    picks = train_data.draw_arrays(packed, N, rng)
    draw_s += time.perf_counter() - d0
    tr.train_picks(*picks, learning_rate=1e-3)
dt = (time.perf_counter() - t0) / STEPS
print("resident games (%d games, %.1f MB ply table): %.3f ms/step incl. %.3f ms NumPy draw, %.0f samples/s" % (
    len(entries), packed.words.nbytes / 1e6, dt * 1e3, draw_s / STEPS * 1e3, N / dt))

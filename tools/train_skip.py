"""In-situ cost of each kernel class of the training step: device ms/step with that class skipped (AZ_TRAIN_SKIP; results are
wrong in those runs, only the timing is used)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import ataxxzero_b200 as az
from ataxxzero_b200 import model, trainer
from test_train_gpu import synthetic_batch
ctx = az.Context(0)
tr = trainer.Trainer(ctx, model.Network.random_init(seed=1), max_batch=512)
b = synthetic_batch(512, 2)
for _ in range(5): tr.train(*b, learning_rate=0.0)
ms = 0.0
for _ in range(50):
    tr.train(*b, learning_rate=0.0); ms += tr.last_step_ms
print("%%.3f" %% (ms / 50))
''' % (ROOT, ROOT)
names = {0: "full step", 1: "forward conv", 2: "data gradient conv", 4: "weight gradient + reduce", 8: "k_bn_apply", 16: "k_bn_bwd_apply", 32: "k_heads",
         63: "everything above (k_sgd, k_images, k_stage_input, last k_bn_bwd_stats, memsets remain)"}
base = None
for bits, name in names.items():
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, AZ_TRAIN_SKIP=str(bits)), capture_output=True, text=True)
    try:
        ms = float(out.stdout.strip().splitlines()[-1])
    except Exception:
        print(name, "failed:", out.stderr[-300:]); continue
    if bits == 0:
        base = ms
        print("full step: %.3f ms on the device" % ms)
    else:
        print("without %-28s %.3f ms  -> in-situ cost %.3f ms (%.0f %%)" % (name + ":", ms, base - ms, 100 * (base - ms) / base))

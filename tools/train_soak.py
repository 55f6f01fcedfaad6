"""Training soak on real self-play data: the hand-written step (csrc/az_train.cu) and the fp32 PyTorch restatement are fed the SAME
minibatches, drawn from games the self-play pool has just generated, for several hundred momentum steps; the held-out losses of
both (inference-mode batch-norm) are printed side by side.  Then the trained network is exported and played by the self-play
kernels.  GPU box only."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net as aznet, search, train_data, trainer
import torch_train_reference as ref

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 600
LR = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ctx = az.Context(0, seed=5)
network = model.Network.random_init(seed=0)
aznet.load_weights(ctx, network)
path = "/tmp/train_soak_games.json"
if os.path.exists(path):
    os.unlink(path)
t0 = time.time()
with search.Pool(ctx, 512, 100, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=5) as pool:
    st = pool.selfplay(path, target_games=400)
entries = train_data.load_entries([path], shuffle=False)
print("self-play: %d games, %d plies in %.1f s (512 games x 100 visits, random-init net)" % (len(entries), sum(len(e["moves"]) for e in entries), time.time() - t0))
held_out, train_entries = entries[:40], entries[40:]
packed, packed_val = train_data.pack_entries(train_entries), train_data.pack_entries(held_out)
rng = np.random.default_rng(1)

def batch(p, n):
    o, m = train_data.draw_arrays(p, n, rng)
    f = np.zeros((n, 7, 7, 4), np.int8); pol = np.zeros((n, 7, 7, 17), np.float32); val = np.zeros((n, 1), np.float32)
    train_data.check(train_data.lib().az_samples_extract(ctx.handle, train_data._vp(p.words.ctypes.data), p.words.size, train_data._vp(o.ctypes.data),
                                                         train_data._vp(m.ctypes.data), n, train_data._vp(f.ctypes.data), train_data._vp(pol.ctypes.data),
                                                         train_data._vp(val.ctypes.data)))
    return f, pol, val

val_batch = batch(packed_val, 2048)
dev = torch.device("cuda", 0)
tnet = ref.build_torch_network(128, 12).to(dev)
ref.load_into(tnet, network)
opt = torch.optim.SGD(tnet.parameters(), lr=LR, momentum=0.9)
tval = ref.to_torch_batch(val_batch, dev)
tr = trainer.Trainer(ctx, network, max_batch=512)

def report(step):
    p1, v1 = tr.losses(*val_batch)
    tnet.eval()
    with torch.no_grad():
        p2, v2, _ = ref.loss_terms(tnet, *tval)
    tnet.train()
    print("step %4d   held-out loss   native: policy %.4f value %.4f   |   PyTorch fp32: policy %.4f value %.4f" % (step, p1, v1, float(p2), float(v2)), flush=True)
    return p1 + v1, float(p2) + float(v2)

t_native = t_torch = 0.0
for step in range(STEPS + 1):
    if step % 100 == 0:
        a, b = report(step)
    b_ = batch(packed, 512)
    t0 = time.perf_counter(); tr.train(*b_, learning_rate=LR); t_native += time.perf_counter() - t0
    xb = ref.to_torch_batch(b_, dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p2, v2, r2 = ref.loss_terms(tnet, *xb); opt.zero_grad(); (p2 + v2 + r2).backward(); opt.step()
    torch.cuda.synchronize(); t_torch += time.perf_counter() - t0
print("per step: native %.2f ms, PyTorch fp32 eager %.2f ms" % (t_native / (STEPS + 1) * 1e3, t_torch / (STEPS + 1) * 1e3))
assert abs(a - b) < 0.05 * max(abs(b), 1e-6), "held-out losses diverged: native %.4f PyTorch %.4f" % (a, b)
trained = tr.network()
aznet.load_weights(ctx, trained)
with search.Pool(ctx, 256, 100, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=6) as pool:
    st = pool.selfplay_ticks(2000)
print("the trained network in the self-play kernels: %d positions, %d games finished in 2000 ticks" % (st["positions"], st["games_finished"]))

"""Device-timed throughput of the downstream kernels against the measured HBM peak (profiles/r01_side_kernels.txt)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ataxxzero_b200 as az
from ataxxzero_b200 import _native, rules, train_data, model, net, search
ctx = az.Context(0)
lib = _native.lib()
stream = torch.cuda.ExternalStream(ctx.stream)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0

# ---- training-sample extraction: records of real self-play games, 1 M samples per launch ----
net.load_weights(ctx, model.Network.random_init(seed=0))
out = "/tmp/side_games.json"
if os.path.exists(out): os.unlink(out)
with search.Pool(ctx, 256, 50, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=5) as pool:
    pool.selfplay(out, target_games=200, max_seconds=120)
entries = train_data.load_entries([out], shuffle=False)
packed = train_data.pack_entries(entries)
import random
picks = train_data.draw(packed, 1 << 20, random.Random(1))
offsets = np.array([packed.offsets[g][p] for g, p, s in picks], dtype=np.uint64)
meta = np.array([(p % 2) | (int(packed.results[g]) << 1) | (s << 3) | (int(packed.has_dists[g]) << 6) for g, p, s in picks], dtype=np.uint32)
n = len(picks)
d_plies = torch.from_numpy(packed.words.view(np.int32)).cuda(); d_off = torch.from_numpy(offsets.view(np.int64)).cuda(); d_meta = torch.from_numpy(meta.view(np.int32)).cuda()
d_f = torch.empty(n * 196, dtype=torch.int8, device="cuda"); d_p = torch.empty(n * 833, dtype=torch.float32, device="cuda"); d_v = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
def launch():
    _native.check(lib.az_samples_extract_dev(ctx.handle, C.c_void_p(d_plies.data_ptr()), C.c_void_p(d_off.data_ptr()), C.c_void_p(d_meta.data_ptr()), n,
                                             C.c_void_p(d_f.data_ptr()), C.c_void_p(d_p.data_ptr()), C.c_void_p(d_v.data_ptr())))
for _ in range(3): launch()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10): launch()
e1.record(stream); ctx.sync()
ms = e0.elapsed_time(e1) / 10
entries_avg = float(np.mean([(packed.words[o + 4] >> 16) for o in offsets[:20000]]))
bytes_per = 196 + 3332 + 4 + 8 + 4 + 24 + 8 * entries_avg
print("k_extract_samples: %d samples in %.3f ms = %.1f M samples/s; algorithmic %.0f B/sample -> %.0f GB/s = %.2f of the measured HBM peak (%.0f GB/s)" % (
    n, ms, n / ms / 1e3, bytes_per, n * bytes_per / ms / 1e6, n * bytes_per / ms / 1e6 / peak, peak))
f, p, v = train_data.extract(ctx, packed, picks[:64])
assert np.array_equal(d_f[:64 * 196].cpu().numpy().reshape(64, 7, 7, 4), f) and np.array_equal(d_p[:64 * 833].cpu().numpy().reshape(64, 7, 7, 17), p)

# ---- random play ----
start = rules.set_board(rules.OPEN_FEN)
import time
for games in (2000, 200000):
    rules.random_playouts(ctx, start, games, 400, 1)
    t0 = time.perf_counter(); _, n_plies, _ = rules.random_playouts(ctx, start, games, 400, 2); dt = time.perf_counter() - t0
    print("k_random_playouts: %d games, %d plies, %.1f ms incl. D2H of %.0f MB of records = %.1f M positions/s" % (games, n_plies.sum(), dt * 1e3, games * 400 * 24 / 1e6, n_plies.sum() / dt / 1e6))

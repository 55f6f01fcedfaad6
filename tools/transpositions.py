"""How many nodes of an 800-visit search tree are transpositions of an earlier node of the same tree?"""
import ctypes as C, os, sys, json, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ataxxzero_b200 as az, bench
from ataxxzero_b200 import model, net, rules, search, _native
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
lib = _native.lib()
lib.az_pool_debug_nodes.restype = C.c_int
G, V = 64, 800
roots = bench.synthetic_roots(ctx, G, 3)
pool = search.Pool(ctx, G, V, eval_mode=search.EVAL_BF16, noise=True, auto_play=False)
pool.set_roots(roots)
pool.run()
tot = dup = 0
for g in range(G):
    cap = V + 64
    own = np.zeros(cap, np.uint64); opp = np.zeros(cap, np.uint64); turn = np.zeros(cap, np.int32); vis = np.zeros(cap, np.int32)
    n = C.c_int32()
    _native.check(lib.az_pool_debug_nodes(pool._h, g, C.c_void_p(own.ctypes.data), C.c_void_p(opp.ctypes.data), C.c_void_p(turn.ctypes.data),
                                          C.c_void_p(vis.ctypes.data), cap, C.byref(n)))
    keys = list(zip(own[:n.value].tolist(), opp[:n.value].tolist(), turn[:n.value].tolist()))
    tot += len(keys); dup += len(keys) - len(set(keys))
print("nodes %d, transposition duplicates %d (%.1f %%)" % (tot, dup, 100.0 * dup / tot))

"""Actual bf16 / fp32 error of the GPU net against the fp64 restatement (tolerances: 2e-2 / 1e-5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net
from oracle import net_numpy
ctx = az.Context(0)
for seed, rand_bn in ((0, False), (0, True), (7, True)):
    network = model.Network.random_init(seed=seed)
    if rand_bn:
        network.bn = net_numpy.randomize_bn(network.bn, seed=seed + 1)
    net.load_weights(ctx, network)
    feats = net_numpy.random_features(64, seed=seed + 2)
    want_p, want_v = net_numpy.forward(feats, network.conv, network.bn, dtype=np.float64)
    for mode, name in ((net.FP32, "fp32"), (net.BF16, "bf16")):
        p, v = net.forward(ctx, feats, mode)
        print("seed %d rand_bn %d %s: max |dlogit| %.3e (logit scale %.2f), max |dvalue| %.3e" % (
            seed, rand_bn, name, np.abs(p - want_p).max(), np.abs(want_p).max(), np.abs(v - want_v).max()))

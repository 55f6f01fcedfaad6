"""Error of the GPU net (fp32 CUDA-core mode, bf16 tcgen05 mode) against the fp64 restatement of model.py, at the
random-init scale the benchmark runs on and at TRAINED scale (calibrated batch-norm statistics, logits of standard
deviation 2, saturating values; oracle/net_numpy.trained_scale_weights), per conv layer and for the heads.
Also against the fp64 evaluation with bf16-rounded operands ("emulated"): that isolates the kernel from the format.

    python tools/net_error.py > profiles/rNN_net_error.txt
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ataxxzero_b200 as az
from ataxxzero_b200 import _native, model, net
from oracle import net_numpy


def bf16(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


def towers(feats, conv, bn):
    """per conv layer: (exact fp64 post-activation, bf16-operand emulation of the kernel's arithmetic)"""
    layers = len(bn) // 2
    xe = np.asarray(feats, dtype=np.float64)
    xq = bf16(feats)
    res_e = res_q = None
    out = []
    for l in range(layers):
        w = np.asarray(conv[l], dtype=np.float64)
        sc = 1.0 / np.sqrt(np.asarray(bn[2 * l + 1], dtype=np.float64) + 1e-3)
        sh = -np.asarray(bn[2 * l], dtype=np.float64) * sc
        ye = net_numpy._conv_same(xe, w) * sc + sh
        yq = net_numpy._conv_same(xq, bf16((w * sc).astype(np.float32))) + sh.astype(np.float32).astype(np.float64)
        second = l > 0 and l % 2 == 0
        if second:
            ye, yq = ye + res_e, yq + res_q
        ye, yq = np.maximum(ye, 0), np.maximum(yq, 0)
        if l == 0 or second:
            res_e, res_q = ye, yq.astype(np.float32).astype(np.float64)
        out.append((ye, yq))
        xe, xq = ye, bf16(yq.astype(np.float32))
    return out


def report(ctx, name, conv, bn, n=32, seed=5):
    lib = _native.lib()
    lib.az_net_debug_tower.restype = C.c_int
    lib.az_net_debug_tower.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    network = model.Network(conv, bn)
    net.load_weights(ctx, network)
    feats = net_numpy.random_features(n, seed=seed)
    want_p, want_v = net_numpy.forward(feats, conv, bn, dtype=np.float64)
    ref = towers(feats, conv, bn)
    print("== %s: logits std %.3f max|.| %.3f; |value| mean %.3f max %.3f" % (
        name, want_p.std(), np.abs(want_p).max(), np.abs(want_v).mean(), np.abs(want_v).max()))
    print("   layer   act rms    bf16 vs fp64: max abs / rms abs / rms rel      bf16 vs emulated: max abs")
    for layers in (1, 2, 3, 5, 9, 13, 17, 21, 25):
        out = np.zeros((n, 49, 128), dtype=np.float32)
        _native.check(lib.az_net_debug_tower(ctx.handle, C.c_void_p(feats.ctypes.data), n, layers, C.c_void_p(out.ctypes.data)))
        got = out.reshape(n, 7, 7, 128).astype(np.float64)
        exact, emu = ref[layers - 1]
        rms = np.sqrt((exact ** 2).mean())
        d = got - exact
        print("   %5d   %7.3f    %.3e / %.3e / %.3e                 %.3e" % (
            layers, rms, np.abs(d).max(), np.sqrt((d ** 2).mean()), np.sqrt((d ** 2).mean()) / rms, np.abs(got - emu).max()))
    for mode, label in ((net.FP32, "fp32"), (net.BF16, "bf16")):
        p, v = net.forward(ctx, feats, mode)
        dp, dv = p - want_p, v - want_v
        print("   heads %s: logits max abs %.3e rms abs %.3e rms rel %.3e (max abs / max|logit| %.3e); value max abs %.3e" % (
            label, np.abs(dp).max(), np.sqrt((dp ** 2).mean()), np.sqrt((dp ** 2).mean()) / want_p.std(),
            np.abs(dp).max() / np.abs(want_p).max(), np.abs(dv).max()))


def main():
    ctx = az.Context(0)
    conv, bn = net_numpy.init_weights(seed=0)
    report(ctx, "random init (model-001, the benchmark's weights)", conv, bn)
    report(ctx, "random init, randomised batch-norm statistics", conv, net_numpy.randomize_bn(bn, seed=1))
    for std in (0.5, 2.0):
        conv, bn = net_numpy.trained_scale_weights(seed=0, logit_std=std)
        report(ctx, "trained scale (calibrated BN), logit std %.1f" % std, conv, bn)


if __name__ == "__main__":
    main()

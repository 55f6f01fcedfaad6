"""Third, structurally unrelated evaluation of the reference network (TEST INFRASTRUCTURE ONLY).

oracle/net_numpy.py is the net oracle and is PARITY UNPINNED (no TensorFlow, no reference goldens).  To give it a
witness that shares none of its code paths, this file evaluates model.py:38-79 with

  * `scipy.signal.correlate(..., mode="valid", method="direct")` on an explicitly zero-padded board, one call per output
    channel, in float64 -- the textbook definition of tf.nn.conv2d(padding="SAME") (cross-correlation, no kernel flip;
    model.py:68,74,118), no im2col, no matmul;
  * batch-norm written out as `(v - mean) / sqrt(var + 1e-3)` (model.py:120-124, gamma = 1, beta = 0);
  * the value head as an explicit double loop over the 49 cells in x-major order (model.py:76-79).

Slow (seconds per position); used on two or three positions in tests/test_oracle_pinned.py."""
import math

import numpy as np
from scipy import signal

BN_EPS = 1e-3


def _conv_same(x, w):
    """x [7,7,Cin], w [kh,kw,Cin,Cout] -> [7,7,Cout]"""
    kh, kw = w.shape[0], w.shape[1]
    xp = np.zeros((x.shape[0] + kh - 1, x.shape[1] + kw - 1, x.shape[2]))
    xp[kh // 2:kh // 2 + x.shape[0], kw // 2:kw // 2 + x.shape[1], :] = x
    out = np.zeros((x.shape[0], x.shape[1], w.shape[3]))
    for o in range(w.shape[3]):
        out[:, :, o] = signal.correlate(xp, w[:, :, :, o], mode="valid", method="direct")[:, :, 0]
    return out


def forward_one(features, conv, bn):
    """features [7,7,4] -> (logits [7,7,17], value float); float64 throughout"""
    conv = [np.asarray(a, dtype=np.float64) for a in conv]
    bn = [np.asarray(a, dtype=np.float64) for a in bn]
    blocks = (len(conv) - 5) // 2

    def conv_bn(v, k):
        return (_conv_same(v, conv[k]) - bn[2 * k]) / np.sqrt(bn[2 * k + 1] + BN_EPS)

    x = np.maximum(conv_bn(np.asarray(features, dtype=np.float64), 0), 0.0)
    for b in range(blocks):
        skip = x
        x = np.maximum(conv_bn(x, 1 + 2 * b), 0.0)
        x = np.maximum(conv_bn(x, 2 + 2 * b) + skip, 0.0)
    logits = _conv_same(x, conv[-4])
    plane = _conv_same(x, conv[-3])[:, :, 0]
    acc = float(conv[-1][0])
    for xx in range(7):
        for yy in range(7):
            acc += plane[xx, yy] * float(conv[-2][7 * xx + yy, 0])
    return logits, math.tanh(acc)

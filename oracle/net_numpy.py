"""NumPy restatement of the reference network's forward pass (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/model.py:38-79 (tower, policy head, value head), :116-142
(conv + batch-norm + residual block), :103-114 (initialisers) and :179-196 (the
``.npy`` layout).  The arithmetic itself lives in TensorFlow 1.x (un-pinned; not
installable here), so this restates TF's published semantics:

  * ``tf.nn.conv2d(..., padding="SAME")`` NHWC with filter ``[kh, kw, Cin, Cout]``:
    ``out[b,i,j,o] = sum_{di,dj,c} in[b, i+di-1, j+dj-1, c] * W[di,dj,c,o]`` (zero pad);
    the first spatial axis is the reference's ``x`` (features are indexed ``[x][y][c]``).
  * ``tf.layers.batch_normalization`` in inference: ``(v - mean) / sqrt(var + 1e-3)``
    with gamma = 1, beta = 0 (model.py never saves gamma/beta).

PARITY UNPINNED: the reference ships no weights, no golden logits and no test that
touches the net, and TensorFlow cannot run in this image.  This restatement *is* the
oracle for the net; tolerance 1e-5 (fp32 mode) / 2e-2 abs (bf16) per north_star.
"""
import numpy as np

BOARD = 7
MOVE_TYPES = 17
FILTERS = 128
BLOCKS = 12
IN_FEATURES = 4
BN_EPS = 1e-3


def conv_shapes(filters=FILTERS, blocks=BLOCKS):
    """Shapes of the 29 arrays of ``conv_list`` in model.py parameter order."""
    shapes = [(3, 3, IN_FEATURES, filters)]
    shapes += [(3, 3, filters, filters)] * (2 * blocks)
    shapes += [(1, 1, filters, MOVE_TYPES), (1, 1, filters, 1), (BOARD * BOARD, 1), (1,)]
    return shapes


def init_weights(seed=0, filters=FILTERS, blocks=BLOCKS):
    """Random init with the reference's distributions (model.py:103-114): truncated normal
    (|z| <= 2) with stddev 0.2*sqrt(2/prod(shape[:-1])), bias 0.01, BN mean 0 / var 1."""
    rng = np.random.default_rng(seed)
    conv = []
    for shape in conv_shapes(filters, blocks):
        if shape == (1,):
            conv.append(np.full(shape, 0.01, dtype=np.float32))
            continue
        std = 0.2 * (2.0 / float(np.prod(shape[:-1]))) ** 0.5
        z = rng.standard_normal(shape)
        bad = np.abs(z) > 2.0
        while bad.any():                       # resample the tails, as tf.truncated_normal does
            z[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(z) > 2.0
        conv.append((z * std).astype(np.float32))
    bn = []
    for _ in range(1 + 2 * blocks):
        bn.append(np.zeros(filters, dtype=np.float32))
        bn.append(np.ones(filters, dtype=np.float32))
    return conv, bn


def randomize_bn(bn, seed=1):
    """Non-trivial moving statistics, so BN folding is actually exercised by the tests."""
    rng = np.random.default_rng(seed)
    out = []
    for i, a in enumerate(bn):
        if i % 2 == 0:
            out.append((0.1 * rng.standard_normal(a.shape)).astype(np.float32))
        else:
            out.append((0.5 + rng.random(a.shape)).astype(np.float32))
    return out


def trained_scale_weights(seed=0, logit_std=2.0, value_std=1.5, batch=64):
    """Random weights with the STATISTICS of a trained network (the reference ships none): every batch-norm layer's moving
    mean / variance are set to the actual per-channel moments of its input on a calibration batch, so each layer emits
    unit-variance channels the way a trained tower does (instead of the 0.2-gain init, whose activations stay tiny), the
    policy head is scaled until the logits have standard deviation `logit_std` (|logit| up to ~4 sigma) and the value head
    until tanh saturates for part of the batch.  Returns (conv, bn) as float32 lists like init_weights()."""
    conv, bn = init_weights(seed)
    conv = [np.asarray(a, dtype=np.float64) for a in conv]
    bn = [np.asarray(a, dtype=np.float64) for a in bn]
    blocks = (len(conv) - 5) // 2
    x = random_features(batch, seed=seed + 100).astype(np.float64)

    def calibrate(v, k):
        y = _conv_same(v, conv[k])
        bn[2 * k] = y.mean(axis=(0, 1, 2))
        bn[2 * k + 1] = y.var(axis=(0, 1, 2)) + 1e-6
        return (y - bn[2 * k]) / np.sqrt(bn[2 * k + 1] + BN_EPS)

    x = np.maximum(calibrate(x, 0), 0)
    for b in range(blocks):
        skip = x
        x = np.maximum(calibrate(x, 1 + 2 * b), 0)
        x = np.maximum(calibrate(x, 2 + 2 * b) + skip, 0)
    conv[-4] = conv[-4] * (logit_std / _conv_same(x, conv[-4]).std())
    v = _conv_same(x, conv[-3]).reshape(batch, BOARD * BOARD) @ conv[-2]
    conv[-3] = conv[-3] * (value_std / v.std())
    return [a.astype(np.float32) for a in conv], [a.astype(np.float32) for a in bn]


def save_model(path, conv, bn):
    """model.py:179-183 -- ``np.save(path, [conv_weights, bn_params])`` (2-element object array)."""
    box = np.empty(2, dtype=object)
    box[0] = list(conv)
    box[1] = list(bn)
    with open(path, "wb") as f:       # np.save would append ".npy" to a bare path
        np.save(f, box, allow_pickle=True)


def load_model(path):
    conv, bn = np.load(path, allow_pickle=True)
    return list(conv), list(bn)


def _conv_same(x, w):
    kh, kw = w.shape[0], w.shape[1]
    ph, pw = kh // 2, kw // 2
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0)))
    out = np.zeros(x.shape[:3] + (w.shape[3],), dtype=x.dtype)
    for di in range(kh):
        for dj in range(kw):
            out += xp[:, di:di + BOARD, dj:dj + BOARD, :] @ w[di, dj]
    return out


def forward(features, conv, bn, dtype=np.float64):
    """features [B,7,7,4] -> (policy logits [B,7,7,17], value [B,1]); model.py:38-79."""
    x = np.asarray(features).astype(dtype)
    conv = [np.asarray(a).astype(dtype) for a in conv]
    bn = [np.asarray(a).astype(dtype) for a in bn]
    blocks = (len(conv) - 5) // 2
    eps = dtype(BN_EPS)

    def conv_bn(v, k):
        v = _conv_same(v, conv[k])
        return (v - bn[2 * k]) / np.sqrt(bn[2 * k + 1] + eps)

    x = np.maximum(conv_bn(x, 0), 0)
    for b in range(blocks):
        skip = x
        x = np.maximum(conv_bn(x, 1 + 2 * b), 0)
        x = conv_bn(x, 2 + 2 * b)
        x = np.maximum(x + skip, 0)
    w_policy, w_value, fc_w, fc_b = conv[-4], conv[-3], conv[-2], conv[-1]
    policy = _conv_same(x, w_policy)
    v = _conv_same(x, w_value).reshape(x.shape[0], BOARD * BOARD)
    value = np.tanh(v @ fc_w + fc_b)
    return policy, value


def random_features(n, seed=0, blockers=True):
    """Structurally valid inputs: c0 = 1, c1/c2 disjoint ~0.3 each, c3 = 4 blockers."""
    rng = np.random.default_rng(seed)
    f = np.zeros((n, 7, 7, 4), dtype=np.float32)
    f[..., 0] = 1
    u = rng.random((n, 7, 7))
    f[..., 1] = u < 0.3
    f[..., 2] = (u >= 0.3) & (u < 0.6)
    if blockers:
        for (x, y) in ((2, 3), (3, 2), (3, 4), (4, 3)):
            f[:, x, y, 1:3] = 0
            f[:, x, y, 3] = 1
    return f

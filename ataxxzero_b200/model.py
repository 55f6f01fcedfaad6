"""Weight-file side of the reference's ``model.py``: the pickled ``.npy`` layout
(model.py:179-196), the initialisers (model.py:103-114) and the packing the C ABI expects.

Only NumPy is needed (PyTorch is optional: ``to_torch_state`` is a convenience for users who
train elsewhere).  The forward pass itself is NOT here -- it runs in libataxxzero.so.
"""
import numpy as np

BOARD_SIZE = 7
MOVE_TYPES = 17


class Network:
    """Holds the parameters of one network in the reference's own list layout.

    ``conv`` (29 arrays for 12 blocks): W_in(3,3,4,F), 2*blocks x (3,3,F,F), W_policy(1,1,F,17),
    W_value(1,1,F,1), fc_w(49,1), fc_b(1,).  ``bn`` (2 per BN layer): moving_mean(F), moving_variance(F).
    Class constants mirror model.Network (model.py:14-19)."""
    INPUT_FEATURE_COUNT = 4
    FILTERS = 128
    CONV_SIZE = 3
    BLOCK_COUNT = 12
    VALUE_FILTERS = 1

    def __init__(self, conv, bn):
        self.conv = [np.asarray(a, dtype=np.float32) for a in conv]
        self.bn = [np.asarray(a, dtype=np.float32) for a in bn]
        self.filters = int(self.conv[0].shape[-1])
        self.blocks = (len(self.conv) - 5) // 2
        self._check()

    def _check(self):
        f, b = self.filters, self.blocks
        want = [(3, 3, 4, f)] + [(3, 3, f, f)] * (2 * b) + [(1, 1, f, MOVE_TYPES), (1, 1, f, 1), (49, 1), (1,)]
        got = [tuple(a.shape) for a in self.conv]
        if got != want:
            raise ValueError("Parameter count mismatch! expected shapes %r, got %r" % (want, got))
        if len(self.bn) != 2 * (1 + 2 * b) or any(a.shape != (f,) for a in self.bn):
            raise ValueError("Bad batch normalization parameter count!")

    @property
    def total_parameters(self):
        return int(sum(a.size for a in self.conv))

    # ---- constructors -------------------------------------------------------------------
    @classmethod
    def random_init(cls, seed=0, filters=None, blocks=None):
        """model.py:103-114: truncated normal (|z|<=2), stddev 0.2*sqrt(2/prod(shape[:-1])); bias 0.01;
        batch-norm moving mean 0 / variance 1 (a fresh tf.layers.batch_normalization)."""
        f = filters or cls.FILTERS
        b = blocks if blocks is not None else cls.BLOCK_COUNT
        rng = np.random.default_rng(seed)
        shapes = [(3, 3, 4, f)] + [(3, 3, f, f)] * (2 * b) + [(1, 1, f, MOVE_TYPES), (1, 1, f, 1), (49, 1)]
        conv = []
        for shape in shapes:
            std = 0.2 * (2.0 / float(np.prod(shape[:-1]))) ** 0.5
            z = rng.standard_normal(shape)
            bad = np.abs(z) > 2.0
            while bad.any():
                z[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(z) > 2.0
            conv.append((z * std).astype(np.float32))
        conv.append(np.full((1,), 0.01, dtype=np.float32))
        bn = []
        for _ in range(1 + 2 * b):
            bn += [np.zeros(f, dtype=np.float32), np.ones(f, dtype=np.float32)]
        return cls(conv, bn)

    @classmethod
    def load(cls, path):
        """model.load_model (model.py:186-196): ``np.load(path, allow_pickle=True)`` -> [conv, bn]."""
        conv, bn = np.load(path, allow_pickle=True)
        return cls(list(conv), list(bn))

    def save(self, path):
        """model.save_model (model.py:179-183): a 2-element object array [conv_list, bn_list]."""
        box = np.empty(2, dtype=object)
        box[0] = [a.copy() for a in self.conv]
        box[1] = [a.copy() for a in self.bn]
        with open(path, "wb") as f:
            np.save(f, box, allow_pickle=True)

    # ---- C-ABI packing ------------------------------------------------------------------
    def packed(self):
        """One float32 vector in the order az_net_load documents (include/ataxxzero.h)."""
        parts = [a.reshape(-1) for a in self.conv] + [a.reshape(-1) for a in self.bn]
        return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)

    def to_torch_state(self):
        """Optional: NCHW torch tensors (conv weights as [Cout, Cin, kh, kw]) for external training code."""
        import torch
        out = {}
        for i, w in enumerate(self.conv[:-2]):
            out["conv%d.weight" % i] = torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1)))
        out["fc.weight"] = torch.from_numpy(self.conv[-2].T.copy())
        out["fc.bias"] = torch.from_numpy(self.conv[-1].copy())
        for i in range(len(self.bn) // 2):
            out["bn%d.running_mean" % i] = torch.from_numpy(self.bn[2 * i].copy())
            out["bn%d.running_var" % i] = torch.from_numpy(self.bn[2 * i + 1].copy())
        return out


def save_model(net, path):
    net.save(path)
    print("\x1b[35mSaved model to:\x1b[0m", path)


def load_model(path):
    return Network.load(path)

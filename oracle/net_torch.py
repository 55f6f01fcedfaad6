"""torch-CPU (oneDNN) forward of the reference network -- the fastest faithful stand-in for the
TensorFlow `sess.run` of accelerated_generate_games.py:57-63 that this image can run on host cores.
TEST / BASELINE INFRASTRUCTURE ONLY (used by bench.py's reference arm).  Same math as
oracle/net_numpy.py (model.py:38-79), checked against it in tests/test_oracle_pinned.py."""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


class TorchNet:
    def __init__(self, conv, bn, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.blocks = (len(conv) - 5) // 2
        # TF filter [kh(x), kw(y), Cin, Cout] -> torch [Cout, Cin, kh, kw]; activations are [B, C, x, y]
        self.w = [torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32).transpose(3, 2, 0, 1)))
                  for a in conv[:-2]]
        self.fc_w = torch.from_numpy(np.asarray(conv[-2], dtype=np.float32).copy())
        self.fc_b = torch.from_numpy(np.asarray(conv[-1], dtype=np.float32).copy())
        self.mean = [torch.from_numpy(np.asarray(bn[2 * i], dtype=np.float32).copy()).view(1, -1, 1, 1) for i in range(len(bn) // 2)]
        self.inv = [torch.rsqrt(torch.from_numpy(np.asarray(bn[2 * i + 1], dtype=np.float32).copy()) + BN_EPS).view(1, -1, 1, 1)
                    for i in range(len(bn) // 2)]

    @torch.no_grad()
    def forward(self, features):
        x = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32)).permute(0, 3, 1, 2).contiguous()

        def conv_bn(v, k):
            return (F.conv2d(v, self.w[k], padding=1) - self.mean[k]) * self.inv[k]
        x = torch.relu(conv_bn(x, 0))
        for b in range(self.blocks):
            skip = x
            x = torch.relu(conv_bn(x, 1 + 2 * b))
            x = torch.relu(conv_bn(x, 2 + 2 * b) + skip)
        policy = F.conv2d(x, self.w[-2]).permute(0, 2, 3, 1).contiguous()          # [B, x, y, 17]
        v = F.conv2d(x, self.w[-1]).reshape(x.shape[0], 49)                          # x-major
        value = torch.tanh(v @ self.fc_w + self.fc_b)
        return policy.numpy(), value.numpy()

"""train.py of the reference (same flags; train.py:92-158 + model.py:81-101,116-142), the last row of SURVEY 8f.

The optimisation itself is far from the self-play hot path, so it is plain PyTorch (autograd + cuDNN): the network of
model.py (conv3x3 + batch-norm + ReLU tower, residual blocks, policy / value heads) in NHWC-equivalent form, the
reference's loss -- mean soft-label cross-entropy over the 833 logits + mean squared value error + 1e-4 * sum of
l2_loss over every trainable variable -- and ``MomentumOptimizer(lr, 0.9)``.  What IS on the device path: minibatches come
from ``az_samples_extract`` (train_data.minibatch) instead of one Python-built sample at a time (train.py:121-128).
Like the reference, only conv / FC weights and the batch-norm MOVING statistics are written to the ``.npy``: the learned
batch-norm gamma / beta are dropped on save (model.py:173-183) -- kept, because the inference side
(reference and ours) assumes gamma = 1, beta = 0.
"""
import argparse
import random

import numpy as np

from .. import model as azmodel


def build_torch_network(filters, blocks):
    import torch
    import torch.nn as nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            def conv(cin, cout, k):
                return nn.Conv2d(cin, cout, k, padding=k // 2, bias=False)
            def bn():
                return nn.BatchNorm2d(filters, eps=1e-3, momentum=0.01)     # tf momentum 0.99
            self.convs = nn.ModuleList([conv(4, filters, 3)] + [conv(filters, filters, 3) for _ in range(2 * blocks)])
            self.bns = nn.ModuleList([bn() for _ in range(1 + 2 * blocks)])
            self.policy = conv(filters, azmodel.MOVE_TYPES, 1)
            self.value = conv(filters, 1, 1)
            self.fc_w = nn.Parameter(torch.zeros(49, 1))
            self.fc_b = nn.Parameter(torch.full((1,), 0.01))

        def forward(self, x):                      # x: [B, 4, 7(x), 7(y)] float
            import torch.nn.functional as F
            h = F.relu(self.bns[0](self.convs[0](x)))
            for b in range((len(self.convs) - 1) // 2):
                y = F.relu(self.bns[1 + 2 * b](self.convs[1 + 2 * b](h)))
                y = self.bns[2 + 2 * b](self.convs[2 + 2 * b](y))
                h = F.relu(y + h)
            logits = self.policy(h).permute(0, 2, 3, 1).reshape(len(x), -1)          # [B, 7*7*17], index 119x + 17y + plane
            v = self.value(h).permute(0, 2, 3, 1).reshape(len(x), 49)                # x-major, like tf.reshape of NHWC
            return logits, torch.tanh(v @ self.fc_w + self.fc_b)

    return Net()


def load_into(net, network):
    """model.Network (TF layout [kh(x), kw(y), Cin, Cout]) -> torch modules (NCHW with H = x, W = y)."""
    import torch
    with torch.no_grad():
        convs = list(net.convs) + [net.policy, net.value]
        for module, w in zip(convs, network.conv[:len(convs)]):
            module.weight.copy_(torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1))))
        net.fc_w.copy_(torch.from_numpy(network.conv[-2]))
        net.fc_b.copy_(torch.from_numpy(network.conv[-1]))
        for i, bn in enumerate(net.bns):
            bn.running_mean.copy_(torch.from_numpy(network.bn[2 * i]))
            bn.running_var.copy_(torch.from_numpy(network.bn[2 * i + 1]))


def export(net):
    """torch modules -> model.Network: conv / FC weights + batch-norm moving statistics (gamma / beta are not saved).
    torch accumulates the UNBIASED batch variance in running_var where TF's moving_variance takes the biased one; with
    minibatch * 49 >= 25 000 samples per channel the factor n / (n - 1) is below 1.00005 and is exported as is."""
    convs = list(net.convs) + [net.policy, net.value]
    conv = [m.weight.detach().cpu().numpy().transpose(2, 3, 1, 0).copy() for m in convs]
    conv += [net.fc_w.detach().cpu().numpy().copy(), net.fc_b.detach().cpu().numpy().copy()]
    bn = []
    for m in net.bns:
        bn += [m.running_mean.detach().cpu().numpy().copy(), m.running_var.detach().cpu().numpy().copy()]
    return azmodel.Network(conv, bn)


def loss_terms(net, features, policies, values):
    """(policy_loss, value_loss, regularization) exactly as model.py:81-96 defines them."""
    import torch
    logits, out = net(features)
    log_sm = torch.log_softmax(logits, dim=1)
    policy_loss = -(policies.reshape(len(features), -1) * log_sm).sum(dim=1).mean()
    value_loss = ((values - out) ** 2).mean()
    reg = 0.0001 * sum(0.5 * (p ** 2).sum() for p in net.parameters())               # l2_regularizer(scale) = scale * l2_loss
    return policy_loss, value_loss, reg


def to_torch_batch(batch, device):
    import torch
    feats, pol, val = batch
    x = torch.from_numpy(np.ascontiguousarray(feats.astype(np.float32).transpose(0, 3, 1, 2))).to(device)    # [B,x,y,c] -> [B,c,x,y]
    return x, torch.from_numpy(pol).to(device), torch.from_numpy(val.astype(np.float32)).to(device)


def make_minibatch_fn(entries, ctx):
    """Minibatches from the GPU extractor when a context is given, else from the same picks through NumPy (CPU tests)."""
    from .. import train_data
    packed = train_data.pack_entries(entries)
    if ctx is not None:
        return lambda size, rng: train_data.minibatch(ctx, packed, size, rng)

    def host(size, rng):                                     # host twin of the kernel, used only without a GPU
        picks = train_data.draw(packed, size, rng)
        feats = np.zeros((size, 7, 7, 4), np.int8)
        pol = np.zeros((size, 7, 7, 17), np.float32)
        val = np.zeros((size, 1), np.float32)
        from ..engine import add_move_to_heatmap
        from ..rules import to_reference_move, unpack_move
        for i, (g, ply, sym) in enumerate(picks):
            entry = entries[g]
            to_move = 1 if ply % 2 == 0 else 2
            def tr(xy):
                x, y = xy
                if sym & 1: x = 6 - x
                if sym & 2: y = 6 - y
                return (y, x) if sym & 4 else (x, y)
            for idx, v in enumerate(entry["boards"][ply]):
                x, y = tr((idx % 7, idx // 7))
                feats[i, x, y, 0] = 1
                if v:
                    feats[i, x, y, 1 if v == to_move else 2] = 1
            off = packed.offsets[g][ply]
            w = packed.words
            n_e = int(w[off + 4] >> 16) if packed.has_dists[g] else 1
            for e in range(n_e):
                if packed.has_dists[g]:
                    mv, p = int(w[off + 6 + 2 * e]) & 0xffff, float(np.uint32(w[off + 7 + 2 * e]).view(np.float32))
                else:
                    mv, p = int(w[off + 4]) & 0xffff, 1.0
                start, end = to_reference_move(unpack_move(mv))
                move = ("c", tr(end)) if start == "c" else (tr(start), tr(end))
                add_move_to_heatmap(pol[i], move, np.float32(p))
            val[i, 0] = 1.0 if entry["result"] == to_move else -1.0
        return feats, pol, val
    return host


def build_parser():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--games", metavar="PATH", required=True, nargs="+", help="Path to .json self-play games files.")
    parser.add_argument("--old-path", metavar="PATH", help="Path for input network.")
    parser.add_argument("--new-path", metavar="PATH", required=True, help="Path for output network.")
    parser.add_argument("--steps", metavar="COUNT", type=int, default=1000, help="Training steps.")
    parser.add_argument("--minibatch-size", metavar="COUNT", type=int, default=512, help="Minibatch size.")
    parser.add_argument("--learning-rate", metavar="LR", type=float, default=0.001, help="Learning rate.")
    parser.add_argument("--device", metavar="N", type=int, default=0, help="GPU index (CPU when no GPU is visible).")
    parser.add_argument("--filters", type=int, default=azmodel.Network.FILTERS)
    parser.add_argument("--blocks", type=int, default=azmodel.Network.BLOCK_COUNT)
    return parser


def main(argv=None):
    import torch
    from .. import train_data
    args = build_parser().parse_args(argv)
    print("Arguments:", args)
    if args.filters != azmodel.Network.FILTERS and not args.old_path:
        # az_net_load only accepts the reference's width (model.py:16): a network of another width could be trained and
        # exported here but not played by the self-play library, and the loop would fail one round later
        raise SystemExit("--filters %d: libataxxzero.so is built for %d filters" % (args.filters, azmodel.Network.FILTERS))
    random.seed(123456789)                                   # train.py:103: shuffle the loaded games deterministically
    entries = train_data.load_entries(args.games, shuffle=True, rng=random)
    print("Found %i games with %i plies." % (len(entries), sum(len(e["moves"]) for e in entries)))
    test_entries, train_entries = entries[:10], entries[10:]
    use_gpu = torch.cuda.is_available()
    device = torch.device("cuda", args.device) if use_gpu else torch.device("cpu")
    ctx = None
    if use_gpu:
        from .. import Context
        ctx = Context(device=args.device)
    if args.old_path:
        print("Loading old model.")
        network = azmodel.Network.load(args.old_path)
    else:
        print("WARNING: Not loading a previous model!")
        network = azmodel.Network.random_init(seed=random.getrandbits(31), filters=args.filters, blocks=args.blocks)
    net = build_torch_network(network.filters, network.blocks).to(device)
    load_into(net, network)
    opt = torch.optim.SGD(net.parameters(), lr=args.learning_rate, momentum=0.9)     # tf.train.MomentumOptimizer(lr, 0.9)
    rng = random.Random(123456789)
    val_batch = to_torch_batch(make_minibatch_fn(test_entries or train_entries, ctx)(min(2048, 64 * max(len(test_entries), 1)), rng), device)
    train_batch = make_minibatch_fn(train_entries or test_entries, ctx)
    print("Model dimensions: %i filters, %i blocks, %i parameters." % (network.filters, network.blocks, network.total_parameters))
    print("=== BEGINNING TRAINING ===")
    history = []
    for step in range(args.steps):
        if step % 100 == 0:
            net.eval()
            with torch.no_grad():
                p_loss, v_loss, _ = loss_terms(net, *val_batch)
            print("Step: %4i -- loss: %.6f  (policy: %.6f  value: %.6f)" % (step, float(p_loss + v_loss), float(p_loss), float(v_loss)))
            history.append((float(p_loss), float(v_loss)))
        net.train()
        p_loss, v_loss, reg = loss_terms(net, *to_torch_batch(train_batch(args.minibatch_size, rng), device))
        opt.zero_grad(set_to_none=True)
        (1.0 * p_loss + v_loss + reg).backward()
        opt.step()
    net.eval()
    with torch.no_grad():
        p_loss, v_loss, _ = loss_terms(net, *val_batch)
    history.append((float(p_loss), float(v_loss)))
    azmodel.save_model(export(net), args.new_path)
    if ctx is not None:
        ctx.close()
    return history


if __name__ == "__main__":
    main()

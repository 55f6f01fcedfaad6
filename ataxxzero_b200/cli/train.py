"""train.py of the reference (same flags; train.py:92-158 + model.py:81-101,116-142), the last row of SURVEY 8f.

Every optimisation step runs in libataxxzero.so (csrc/az_train.cu, bound by ataxxzero_b200/trainer.py): the network of
model.py (conv3x3 + batch-norm + ReLU tower, residual blocks, policy / value heads) forward and backward on the tensor
cores, the reference's loss -- mean soft-label cross-entropy over the 833 logits + mean squared value error + 1e-4 * sum
of l2_loss over every trainable variable -- and ``MomentumOptimizer(lr, 0.9)``.  The games are packed once and kept on the
device; every step draws its samples' descriptions in one NumPy call (train_data.draw_arrays) and ``az_trainer_step_picks`` runs
sample extraction (the ``az_samples_extract`` kernel) and the step back to back, instead of one Python-built sample at a time
(train.py:121-128).  There is
no PyTorch and no CPU path here: without a B200 the script stops (tests/torch_train_reference.py holds the fp32 PyTorch
restatement the step is checked against).
Like the reference, only conv / FC weights and the batch-norm MOVING statistics are written to the ``.npy``: the learned
batch-norm gamma / beta are dropped on save (model.py:173-183) -- kept, because the inference side
(reference and ours) assumes gamma = 1, beta = 0.
"""
import argparse
import random

import numpy as np

from .. import model as azmodel


def make_minibatch_fn(entries, ctx):
    """``size, rng -> (features, policies, values)``: picks drawn like get_sample_from_entries (train.py:43-52), tensors built by
    az_samples_extract on the GPU."""
    from .. import train_data
    packed = train_data.pack_entries(entries)
    return lambda size, rng: train_data.minibatch(ctx, packed, size, rng)


def build_parser():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--games", metavar="PATH", required=True, nargs="+", help="Path to .json self-play games files.")
    parser.add_argument("--old-path", metavar="PATH", help="Path for input network.")
    parser.add_argument("--new-path", metavar="PATH", required=True, help="Path for output network.")
    parser.add_argument("--steps", metavar="COUNT", type=int, default=1000, help="Training steps.")
    parser.add_argument("--minibatch-size", metavar="COUNT", type=int, default=512, help="Minibatch size.")
    parser.add_argument("--learning-rate", metavar="LR", type=float, default=0.001, help="Learning rate.")
    parser.add_argument("--device", metavar="N", type=int, default=0, help="GPU index.")
    parser.add_argument("--filters", type=int, default=azmodel.Network.FILTERS)
    parser.add_argument("--blocks", type=int, default=azmodel.Network.BLOCK_COUNT)
    return parser


def main(argv=None):
    from .. import Context, train_data, trainer
    args = build_parser().parse_args(argv)
    print("Arguments:", args)
    if args.filters != azmodel.Network.FILTERS:
        # libataxxzero.so is built for the reference's width (model.py:16), on the training side as on the self-play side
        raise SystemExit("--filters %d: libataxxzero.so is built for %d filters" % (args.filters, azmodel.Network.FILTERS))
    random.seed(123456789)                                   # train.py:103: shuffle the loaded games deterministically
    entries = train_data.load_entries(args.games, shuffle=True, rng=random)
    print("Found %i games with %i plies." % (len(entries), sum(len(e["moves"]) for e in entries)))
    test_entries, train_entries = entries[:10], entries[10:]
    ctx = Context(device=args.device)                        # raises without a GPU: there is no CPU training path
    if args.old_path:
        print("Loading old model.")
        network = azmodel.Network.load(args.old_path)
    else:
        print("WARNING: Not loading a previous model!")
        network = azmodel.Network.random_init(seed=random.getrandbits(31), filters=args.filters, blocks=args.blocks)
    net = trainer.Trainer(ctx, network, max_batch=max(args.minibatch_size, 2))
    rng = random.Random(123456789)
    val_batch = make_minibatch_fn(test_entries or train_entries, ctx)(min(2048, 64 * max(len(test_entries), 1)), rng)
    # the training games live on the device for the whole run; a step ships only the description of its samples
    train_games = train_data.pack_entries(train_entries or test_entries)
    net.set_games(train_games)
    np_rng = np.random.default_rng(rng.getrandbits(63))
    print("Model dimensions: %i filters, %i blocks, %i parameters." % (network.filters, network.blocks, network.total_parameters))
    print("=== BEGINNING TRAINING ===")
    history = []

    def report(step):
        p_loss, v_loss = net.losses(*val_batch)
        print("Step: %4i -- loss: %.6f  (policy: %.6f  value: %.6f)" % (step, p_loss + v_loss, p_loss, v_loss))
        history.append((p_loss, v_loss))

    for step in range(args.steps):
        if step % 100 == 0:
            report(step)
        net.train_picks(*train_data.draw_arrays(train_games, args.minibatch_size, np_rng), learning_rate=args.learning_rate)
    p_loss, v_loss = net.losses(*val_batch)
    history.append((p_loss, v_loss))
    azmodel.save_model(net.network(), args.new_path)
    net.close()
    ctx.close()
    return history


if __name__ == "__main__":
    main()

"""perft.py of the reference (perft.py:5-26: divide at the root, depth 4 = 1 + 3) on the GPU.

Prints the same lines -- ``Move: <move> Size: <n>`` per root move and ``Total: <n>`` -- for any depth / FEN; the
subtrees are counted by the batched device perft (az_perft_batch).  Root moves come in C++ movegen order."""
import argparse
import time

from .. import Context, ataxx_rules, rules


def divide(ctx, board, depth):
    """[(reference move, leaf count below it)] for every legal move of ``board``."""
    moves = [m for m in board.legal_moves() if m != "pass"]
    children = []
    for m in moves:
        b = board.copy()
        b.move(m)
        children.append(b.to_position())
    counts = rules.perft_batch(ctx, children, depth - 1) if children else []
    return list(zip(moves, (int(c) for c in counts)))


def main(argv=None):
    ap = argparse.ArgumentParser(description="perft from a position, divided by root move (reference perft.py runs depth 4)")
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--fen", default="x5o/7/7/7/7/7/o5x x", help="start position (perft.py:18 uses AtaxxState.initial())")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    board = ataxx_rules.AtaxxState.from_fen(args.fen)
    from .. import dist
    rank, local_rank, world = dist.init()
    if world > 1:                              # torchrun: root-split over the ranks, one all-reduce of the count
        with Context(device=local_rank) as ctx:
            t0 = time.perf_counter()
            total = dist.perft_sharded(ctx, board.to_position(), args.depth)
            dt = time.perf_counter() - t0
        if rank == 0:
            print("Total:", total)
            print("(%d GPUs, %.3f s incl. communicator start-up, %.1f Mnodes/s)" % (world, dt, total / dt / 1e6))
        import torch.distributed as torch_dist
        torch_dist.destroy_process_group()
        return total
    with Context(device=args.device) as ctx:
        t0 = time.perf_counter()
        total = 0
        for move, n in divide(ctx, board, args.depth):
            print("Move:", move, "Size:", n)
            total += n
        dt = time.perf_counter() - t0
    print("Total:", total)
    print("(%.3f s, %.1f Mnodes/s)" % (dt, total / dt / 1e6))
    return total


if __name__ == "__main__":
    main()

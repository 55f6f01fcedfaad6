"""One process per GPU: how self-play is spread over the GPUs of a box.

Games are independent units (the reference already splits them over threads and ``--parallel-games-processes``
subprocesses, looper.py:33-41), so rank r simply owns games ``[r*G/W, (r+1)*G/W)``, a replica of the weights, the RNG
stream ``seed + rank`` and its own output file ``...-<rank>.json``.  There is NO collective on the data path.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only after the run: one all-reduce of the counters
and, if a single file is wanted, a gather of the record bytes to rank 0.
"""
import os


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def shard_games(total_games, world, rank):
    """Games owned by `rank`: a contiguous range, sizes differing by at most one."""
    if not (0 <= rank < world) or total_games < 0:
        raise ValueError("bad shard request: games=%d world=%d rank=%d" % (total_games, world, rank))
    base, extra = divmod(total_games, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def rank_output_path(path, rank, world):
    """``model-001.json`` -> ``model-001-<rank>.json`` (looper.py:72-74 naming); unchanged for a single process."""
    if world == 1:
        return path
    root, ext = os.path.splitext(path)
    return "%s-%d%s" % (root, rank, ext)


def rank_seed(seed, rank):
    return (int(seed) + rank) & (2**64 - 1)


def init(backend=None):
    """Join the process group torchrun describes; returns (rank, local_rank, world).  No-op for one process."""
    rank, local_rank, world = env_rank()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kwargs = {}
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
                kwargs["device_id"] = torch.device("cuda", local_rank)
            dist.init_process_group(backend, **kwargs)
    return rank, local_rank, world


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_stats(stats, max_keys=("max_depth", "seconds", "ms")):
    """Sum (or max, for `max_keys`) a dict of numbers over all ranks; every rank gets the result."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    keys = sorted(stats)
    sums = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=_device())
    maxs = sums.clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    return {k: (maxs[i].item() if k in max_keys else sums[i].item()) for i, k in enumerate(keys)}


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_records(local_path, merged_path):
    """Concatenate every rank's JSON-lines file into `merged_path` on rank 0 (size exchange + byte gather)."""
    import torch
    import torch.distributed as dist
    data = open(local_path, "rb").read() if os.path.exists(local_path) else b""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        with open(merged_path, "ab") as f:
            f.write(data)
        return len(data)
    dev, world, rank = _device(), dist.get_world_size(), dist.get_rank()
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = len(data)
    dist.all_reduce(sizes)
    cap = int(sizes.max().item())
    mine = torch.zeros(max(cap, 1), dtype=torch.uint8, device=dev)
    if data:
        mine[:len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    total = 0
    if rank == 0:
        with open(merged_path, "ab") as f:
            for r in range(world):
                n = int(sizes[r].item())
                f.write(parts[r][:n].cpu().numpy().tobytes())
                total += n
    return total


def perft_sharded(ctx, root, depth):
    """perft over all ranks: the depth-2 frontier (256 sub-trees from the opening) is dealt round-robin to the
    ranks, each counts its share on its own GPU, one all-reduce(sum) of the uint64 total.  Any world size, also 1."""
    import numpy as np
    from . import rules
    rank, _, world = env_rank()
    if depth < 3:
        return rules.perft(ctx, root, depth)
    frontier = [root]
    for _ in range(2):
        moves = rules.movegen_batch(ctx, frontier)
        parents = [p for p, mv in zip(frontier, moves) for _ in mv]
        flat = [m for mv in moves for m in mv]
        if not flat:
            return 0
        frontier = rules.array_to_positions(rules.makemove_batch(ctx, parents, flat))
    mine = frontier[rank::world]
    total = int(rules.perft_batch(ctx, mine, depth - 2).sum()) if mine else 0
    # two 32-bit halves as float64 sums stay exact (each < 2^53) on both NCCL and gloo
    parts = allreduce_stats({"lo": float(total & 0xffffffff), "hi": float(total >> 32)})
    return (int(parts["hi"]) << 32) + int(parts["lo"])

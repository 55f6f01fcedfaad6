"""Multi-process host logic on CPU: world_size 2 over gloo.  Games shard with no data-path collective;
the only exchanges are the end-of-run counter all-reduce and the optional record gather."""
import json
import os
import socket
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_shard_games_partitions_exactly():
    from ataxxzero_b200 import dist
    for total in (0, 1, 7, 2048, 16384, 16385):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_games(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (a, n), (b, _) in zip(spans, spans[1:]):
                assert a + n == b
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    assert dist.shard_games(16384, 8, 3) == (6144, 2048)
    assert dist.rank_output_path("games/model-001.json", 3, 8) == "games/model-001-3.json"
    assert dist.rank_output_path("games/model-001.json", 0, 1) == "games/model-001.json"
    assert dist.rank_seed(5, 2) == 7


WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    from ataxxzero_b200 import dist
    rank, local_rank, world = dist.init(backend="gloo")
    first, count = dist.shard_games(4097, world, rank)
    out = dist.rank_output_path(os.path.join(sys.argv[1], "model-001.json"), rank, world)
    with open(out, "w") as f:                       # each rank writes its own games, nothing is exchanged meanwhile
        for g in range(first, first + min(count, 3)):
            f.write(json.dumps({"boards": [], "dists": [], "moves": [], "result": 1 + g %% 2, "game": g}) + "\\n")
    stats = dist.allreduce_stats({"positions": 100 * (rank + 1), "evals": count, "max_depth": 10 + rank, "seconds": 1.5 + rank})
    merged = os.path.join(sys.argv[1], "merged.json")
    dist.barrier()
    dist.gather_records(out, merged)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"stats": stats, "lines": open(merged).read().splitlines()}))
""") % ROOT


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    got = json.loads(line)
    assert got["stats"] == {"positions": 300.0, "evals": 4097.0, "max_depth": 11.0, "seconds": 2.5}
    games = [json.loads(l)["game"] for l in got["lines"]]
    assert games == [0, 1, 2, 2049, 2050, 2051]                # rank 0's shard, then rank 1's (4097 = 2049 + 2048)
    assert os.path.exists(tmp_path / "model-001-0.json") and os.path.exists(tmp_path / "model-001-1.json")

#!/usr/bin/env python
"""bench.py -- self-play positions/s of the B200-native hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU pipeline

Workload (config.workload): the per-GPU shard of BASELINE configs[3] -- 2048 concurrent self-play
games x 800 visits/move, random-init model-001 net (bf16 tensor-core mode), Dirichlet noise on,
moves sampled ~ visits, every game slot started from a seeded random-playout position (ply ~U[0,120])
so the pool is in a steady-state mix of game phases from the first timed tick.

A "step" is TICKS_PER_STEP ticks; one tick = one pass of the hot path over the whole pool: tree kernel
(consume evaluations, PUCT select/expand/backup, move selection + records) then the net kernel over
the <= 2048 requested leaves.  `value` = recorded plies of all ranks / device time (CUDA events on the
library's stream, max over ranks), nothing crosses PCIe in that region.  `e2e` = the same through the
public API with host inputs and outputs: every step re-uploads the weights from host memory and
az_selfplay_ticks() drains finished games to a JSON-lines file (records D2H + serialisation timed).
Games are independent, so ranks share nothing (weak scaling); NCCL only sums the counters at the end.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "selfplay_positions_per_sec"
UNIT = "positions/s"
GAMES_PER_GPU = 2048
VISITS = 800
TICKS_PER_STEP = 1024
FLOP_PER_EVAL = 347.49e6          # dense FLOPs of one forward pass


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in open(self.path).read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# the reference's CPU pipeline (oracle/_ref/self_play_client.so + torch-CPU restatement of model.py)
# ----------------------------------------------------------------------------------------------
class ReferencePipeline:
    """accelerated_generate_games.py:24-83 verbatim, except that `engine.sess.run` (TensorFlow, not
    installable here) is the fp32 torch-CPU (oneDNN) restatement of model.py
    (oracle/net_torch.py, checked against the NumPy restatement).  All host cores: the client runs
    2*buffer_size worker threads, torch's intra-op pool uses the rest."""

    def __init__(self, visits, buffer_size=128):
        import ctypes
        import numpy as np
        import torch
        from oracle import cpu as ocpu, net_numpy, net_torch
        torch.set_num_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1; the reference arm gets every core
        self.np, self.ctypes = np, ctypes
        self.kind = "reference" if os.path.exists(ocpu.REF_CLIENT_SO) else "port"
        self.visits, self.B = visits, buffer_size
        conv, bn = net_numpy.init_weights(seed=0)
        self.net = net_torch.TorchNet(conv, bn)            # fp32 torch-CPU (oneDNN) stand-in for TF's sess.run
        self.ocpu = ocpu
        self.evals_per_position = None
        if self.kind == "reference":
            self.dll = ctypes.CDLL(ocpu.REF_CLIENT_SO)
            self.dll.get_workload.restype = ctypes.c_int
            self.out = tempfile.mktemp(suffix=".json")
            self.buffers = [np.zeros((buffer_size, 7, 7, 4), dtype=np.float32) for _ in (0, 1)]
            self.keep = []
            # the client announces itself on stdout (self_play_client.cpp:686-689); stdout is reserved for the JSON line
            # (and every finished game, :640): for the lifetime of the pipeline fd 1 points at stderr
            sys.stdout.flush()
            self.saved_stdout = os.dup(1)
            os.dup2(2, 1)
            self.dll.launch_threads(self.out.encode(), ctypes.c_int(visits), ctypes.c_void_p(self.buffers[0].ctypes.data),
                                    ctypes.c_void_p(self.buffers[1].ctypes.data), ctypes.c_int(buffer_size),
                                    ctypes.c_int(2 * buffer_size))

    def measure_evals_per_position(self):
        """Average evaluations the REFERENCE's own search core needs per played move at this visit count
        (tree reuse carries visits over): a short greedy game through oracle/_ref/libref.so."""
        if self.evals_per_position is None:
            if self.ocpu.Reference.available():
                ref, orc = self.ocpu.Reference(), self.ocpu.Oracle()
                plies, _, evals = ref.selfplay_greedy(self.ocpu.START_FEN, self.visits, 12, orc.probe_eval_ptr)
                self.evals_per_position = evals / max(len(plies), 1)
            else:
                orc = self.ocpu.Oracle()
                tree = orc.tree(orc.set_board(self.ocpu.START_FEN), "probe")
                for _ in range(12):
                    tree.search(self.visits)
                    d = tree.dist()
                    tree.play(max(d, key=lambda e: e[1])[0])
                self.evals_per_position = tree.evals / 12.0
        return self.evals_per_position

    def run(self, seconds, evaluator=None):
        """Returns evaluations completed in ~`seconds` of wall time.  `evaluator(features) -> (policy, value)` replaces
        the torch-CPU net (used for the host-ceiling / GPU-evaluator variants of the baseline)."""
        np, ctypes = self.np, self.ctypes
        t0 = time.perf_counter()
        evals = 0
        forward = evaluator or self.net.forward
        if self.kind == "reference":
            while time.perf_counter() - t0 < seconds:
                i = self.dll.get_workload()
                policy, value = forward(self.buffers[i])
                policy = np.ascontiguousarray(policy, dtype=np.float32)
                value = np.ascontiguousarray(value, dtype=np.float32)
                self.keep = (self.keep + [(policy, value)])[-6:]
                self.dll.complete_workload(ctypes.c_int(i), ctypes.c_void_p(policy.ctypes.data), ctypes.c_void_p(value.ctypes.data))
                evals += self.B
        else:       # port: our C restatement of the search, one game at a time, same NumPy net
            orc = self.ocpu.Oracle()

            def ev(f):
                p, v = self.net.forward(f.reshape(1, 7, 7, 4))
                return p.reshape(-1), float(v[0, 0])
            tree = orc.tree(orc.set_board(self.ocpu.START_FEN), py_eval=ev)
            while time.perf_counter() - t0 < seconds:
                tree.step()
            evals = tree.evals
        return evals, time.perf_counter() - t0

    def close(self):
        if self.kind == "reference":
            self.dll.shutdown()
            self.ctypes.CDLL(None).fflush(None)
            os.dup2(self.saved_stdout, 1)
            os.close(self.saved_stdout)
            try:
                os.unlink(self.out)
            except OSError:
                pass


def reference_sample(seconds, visits=VISITS, gpu_evaluator=None):
    pipe = ReferencePipeline(visits)
    variants = {}
    try:
        epp = pipe.measure_evals_per_position()
        pipe.run(min(2.0, seconds / 4))                     # warm the thread pool / BLAS
        evals, dt = pipe.run(seconds)
        if pipe.kind == "reference":
            # BASELINE.md's two other views of the reference: its host side with a free evaluator (the ceiling of the
            # thread pool + hash-map trees), and the reference client fed by OUR GPU net through its own legacy ABI
            zeros = (pipe.np.zeros((pipe.B, 7, 7, 17), dtype=pipe.np.float32), pipe.np.zeros((pipe.B, 1), dtype=pipe.np.float32))
            e0, t0 = pipe.run(4.0, evaluator=lambda f: zeros)
            variants["zero_cost_evaluator_host_ceiling"] = {"leaf_evals_per_s": e0 / t0, "positions_per_s": e0 / t0 / epp}
            if gpu_evaluator is not None:
                e1, t1 = pipe.run(4.0, evaluator=gpu_evaluator)
                variants["reference_client_with_our_gpu_net"] = {"leaf_evals_per_s": e1 / t1, "positions_per_s": e1 / t1 / epp}
    finally:
        pipe.close()
    cores = os.cpu_count() or 1
    return {"value": evals / dt / epp, "unit": UNIT, "cores": cores, "kind": pipe.kind, "variants": variants,
            "sample": "%.0f s of %s + torch-CPU fp32 net (TensorFlow absent), %d-visit searches, buffer 128 / 256 threads: "
                      "%.0f leaf-evals/s / %.0f evals per played move (measured on the reference search core)"
                      % (dt, "oracle/_ref/self_play_client.so" if pipe.kind == "reference" else "oracle C port", visits,
                         evals / dt, epp)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, w = max(args.steps, 1), max(args.warmup, 0)
    step_s = max(2.0, min(10.0, 150.0 / (k + w)))
    pipe = ReferencePipeline(VISITS)
    try:
        epp = pipe.measure_evals_per_position()
        for _ in range(w):
            pipe.run(step_s)
        evals, dt = 0, 0.0
        for _ in range(k):
            e, t = pipe.run(step_s)
            evals, dt = evals + e, dt + t
    finally:
        pipe.close()
    value = evals / dt / epp
    cores = os.cpu_count() or 1
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w,
            "ms_per_step": dt / k * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "selfplay %d-visit MCTS, reference CPU pipeline (buffer 128 / 256 threads)" % VISITS,
                       "visits": VISITS, "buffer_size": 128, "step": "%.1f s sample" % step_s},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": pipe.kind,
                             "sample": "%d x %.1f s; %.0f leaf-evals/s / %.0f evals per played move" % (k, step_s, evals / dt, epp)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def synthetic_roots(ctx, n, seed):
    """n start positions from seeded random playouts (ply ~U[0,120]) played with the GPU rule kernels."""
    import random
    import numpy as np
    from ataxxzero_b200 import rules
    rng = random.Random(seed)
    arr = rules.positions_array([rules.set_board(rules.START_FEN)] * n)
    target = np.array([rng.randrange(0, 121) for _ in range(n)])
    for ply in range(120):
        res = rules.result_batch(ctx, arr)
        idx = np.nonzero((target > ply) & (res == 0))[0]
        if len(idx) == 0:
            break
        lists = rules.movegen_batch(ctx, arr[idx])
        keep, moves = [], []
        for i, mv in zip(idx, lists):
            if not mv:
                continue
            m = rng.choice(mv)
            keep.append(i)
            moves.append(m)
        if not keep:
            break
        nxt = rules.makemove_batch(ctx, arr[keep], moves)
        ok = rules.result_batch(ctx, nxt) == 0          # never start a game slot on a finished position
        keep = np.array(keep)
        arr[keep[ok]] = nxt[ok]
    return arr


def run_ours(args):
    import numpy as np
    import torch
    import ataxxzero_b200 as az
    from ataxxzero_b200 import model, net, rules, search

    from ataxxzero_b200 import dist as azdist
    rank, local_rank, world = azdist.env_rank()
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL may print its version banner on stdout while the communicator comes up; stdout carries only the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            azdist.init("nccl")               # counters only: the games themselves never cross ranks
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()

    def allreduce(x, op="sum"):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
        return float(t.item())

    k, w = max(args.steps, 1), max(args.warmup, 3)
    games, visits, ticks = args.games, args.visits, args.ticks_per_step
    ctx = az.Context(device=local_rank, seed=azdist.rank_seed(1000, rank))
    network = model.Network.random_init(seed=0)          # "model-001": the reference's init distributions
    packed = network.packed()
    net.load_weights(ctx, network)
    pool = search.Pool(ctx, games, visits, eval_mode=search.EVAL_BF16, noise=True, auto_play=True, seed=1000 + rank)
    pool.set_roots(synthetic_roots(ctx, games, seed=rank))
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    for _ in range(w):
        pool.selfplay_ticks(ticks)

    # ---- device-resident timed region ----
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    ctx.sync()
    s0 = pool.stats()
    sampler.start()
    ev0.record(stream)
    for _ in range(k):
        pool.selfplay_ticks(ticks)
    ev1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    s1 = pool.stats()
    ms = allreduce(ev0.elapsed_time(ev1), "max")
    d = {key: s1[key] - s0[key] for key in s1}
    positions = allreduce(d["positions"])
    evals = allreduce(d["evals"])
    steps = allreduce(d["steps"])
    value = positions / (ms * 1e-3)

    # ---- roofline of the dominant kernel (net), from CUDA events around its launches ----
    pk, pk_kind = peaks()
    net_s = max(d["net_seconds"], 1e-9)
    achieved = d["evals"] * FLOP_PER_EVAL / net_s / 1e12
    peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0)))
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_net_tc"]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "kernel": "k_net_tc (bf16 tcgen05 tower)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "dram__bytes_read+write per launch from profiles/ (ncu --set full); the weights stream from L2",
                "flop_per_launch": d["evals"] / max(d["ticks"], 1) * FLOP_PER_EVAL,
                "peak_kind": "%s bf16_tflops_sustained (kernel timed inside a long step)" % pk_kind,
                "net_share_of_step": d["net_seconds"] / max(d["net_seconds"] + d["tree_seconds"], 1e-9),
                "tree_ms_per_tick": d["tree_seconds"] / max(d["ticks"], 1) * 1e3,
                "net_ms_per_tick": d["net_seconds"] / max(d["ticks"], 1) * 1e3}

    # ---- end to end: host weights in, JSON game records out, every step ----
    out_path = azdist.rank_output_path(os.path.join(tempfile.gettempdir(), "az_bench_model-001.json"), rank, max(world, 2))
    if os.path.exists(out_path):
        os.unlink(out_path)
    barrier()
    ctx.sync()
    e0 = pool.stats()
    t0 = time.perf_counter()
    pinned = torch.from_numpy(packed).pin_memory()       # the step's input lives in pinned host memory
    pinned_np = pinned.numpy()
    for _ in range(k):
        net.load_packed(ctx, pinned_np, network.filters, network.blocks)   # H2D: the weights of the .npy, every step
        pool.selfplay_ticks(ticks, out_path)             # D2H: finished games -> JSON lines on the host
    ctx.sync()
    e2e_s = allreduce(time.perf_counter() - t0, "max")
    e1 = pool.stats()
    barrier()
    e2e_positions = allreduce(e1["positions"] - e0["positions"])
    record_bytes = e1["record_bytes"] - e0["record_bytes"]
    try:
        json_bytes = os.path.getsize(out_path)
        os.unlink(out_path)
    except OSError:
        json_bytes = 0
    e2e = {"value": e2e_positions / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(packed.nbytes),
           "d2h_bytes_per_step": int(record_bytes / k), "json_bytes_per_step": int(json_bytes / k),
           "api": "net.load_packed (pinned host weights) + Pool.selfplay_ticks(ticks, path) -> az_net_load / az_selfplay_ticks"}

    # ---- BASELINE configs[4]: single-tree search (replicas only, rank 0) + leaf-batch latency of the net kernel ----
    single = {}
    if rank == 0 and not args.no_single_tree:
        golden = os.path.join(ROOT, "tests", "golden", "mcts_golden.json")
        fen = json.load(open(golden))["midgame_fen"] if os.path.exists(golden) else rules.START_FEN
        with search.Pool(ctx, 1, 64, eval_mode=search.EVAL_BF16, node_capacity=args.single_tree_visits + 64, steps_per_tick=64) as tree:
            tree.set_root(0, rules.set_board(fen))
            tree.run()                                   # warm-up: 64 visits
            tree.set_visits(args.single_tree_visits)
            t0 = time.perf_counter()
            tree.run()
            dt = time.perf_counter() - t0
            st = tree.stats()
            single = {"fen": fen, "visits": args.single_tree_visits, "visits_per_s": (args.single_tree_visits - 64) / dt,
                      "mode": "bit-exact sequential PUCT (one leaf per tick, no virtual loss)", "ticks": st["ticks"]}
        lat = {}
        feats = np.zeros((128, 7, 7, 4), dtype=np.float32)
        feats[..., 0] = 1.0
        for b in (1, 8, 32, 128):
            samples = []
            for _ in range(30):
                t0 = time.perf_counter()
                net.forward(ctx, feats[:b], net.BF16)
                samples.append((time.perf_counter() - t0) * 1e6)
            samples.sort()
            lat["batch%d" % b] = {"p50_us": samples[len(samples) // 2], "p99_us": samples[-1]}
        single["leaf_batch_latency_host_to_host"] = lat

    # ---- BASELINE configs[0]: uniformly random play, 2000 games, on the device (records copied back to the host) ----
    start = rules.set_board(rules.OPEN_FEN)
    rules.random_playouts(ctx, start, 2000, 400, seed=1)
    t0 = time.perf_counter()
    _, n_plies, _ = rules.random_playouts(ctx, start, 2000, 400, seed=2)
    random_play = {"games": 2000, "positions": int(n_plies.sum()), "positions_per_s": float(n_plies.sum()) / (time.perf_counter() - t0)}

    # ---- secondary metric of BASELINE.json: perft Mnodes/s (device time incl. frontier expansion) ----
    perft = {}
    for depth in (7, 8):
        p = rules.set_board(rules.OPEN_FEN)
        rules.perft(ctx, p, depth)
        t0 = time.perf_counter()
        nodes = rules.perft(ctx, p, depth)
        dt = time.perf_counter() - t0
        perft["depth%d" % depth] = {"nodes": nodes, "mnodes_per_s": nodes / dt / 1e6}
    try:        # integer-ALU roofline of the perft walk kernel, from the committed ncu capture (profiles/)
        perft["k_walk_alu_pipe_pct_of_peak_ncu"] = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_walk"]["alu_pipe_pct_of_peak_active"]
    except Exception:
        pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": k, "warmup": w,
            "ms_per_step": ms / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "selfplay %d games/GPU x %d visits (BASELINE configs[3] per-GPU shard), random-init model-001"
                                   % (games, visits),
                       "games_per_gpu": games, "visits": visits, "ticks_per_step": ticks, "noise": True,
                       "start": "seeded random-playout positions, ply~U[0,120]",
                       "l2": "working set (node pool + records, several GB per GPU) is far larger than the 126 MB L2; no flush needed",
                       "parallelism": "games sharded across ranks, no data-path collective"},
            "e2e": e2e, "roofline": roofline, "clocks": clocks, "gpu_launches": int(d["kernel_launches"]),
            "extra": {"leaf_evals_per_s": evals / (ms * 1e-3), "mcts_steps_per_s": steps / (ms * 1e-3),
                      "evals_per_position": evals / max(positions, 1), "max_depth": s1["max_depth"], "perft": perft,
                      "single_tree": single, "random_play": random_play}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pool.close()
        try:
            line["cpu_baseline"] = reference_sample(args.cpu_seconds, visits,
                                                    gpu_evaluator=lambda f: net.forward(ctx, f, net.BF16))
            # second metric of BASELINE.json: the reference's own movegen/makemove (oracle/_ref/perft_ref) on the host cores
            from oracle import cpu as ocpu
            if os.path.exists(ocpu.REF_PERFT):
                cores = os.cpu_count() or 1
                n1, s1 = ocpu.ref_perft(ocpu.OPEN_FEN, 6, 1)
                nN, sN = ocpu.ref_perft(ocpu.OPEN_FEN, 7, cores)
                line["cpu_baseline"]["perft"] = {"kind": "reference", "depth6_1thread_mnodes_per_s": n1 / s1 / 1e6,
                                                 "depth7_%dthreads_mnodes_per_s" % cores: nN / sN / 1e6, "nodes": [n1, nN]}
        except Exception as exc:        # the baseline is reported, never the product path
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
                                    "sample": "failed: %r" % (exc,)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--visits", type=int, default=VISITS)
    ap.add_argument("--ticks-per-step", type=int, default=TICKS_PER_STEP)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-tree", action="store_true")
    ap.add_argument("--single-tree-visits", type=int, default=4000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

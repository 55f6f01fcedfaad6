"""Game-generation half of the reference's looper.py (looper.py:21-74): same PREFIX layout
(``PREFIX/models/model-%03i.npy``, ``PREFIX/games/model-%03i-%i.json``), same flags, same stop rule (count the lines
of the games files every few seconds, SIGTERM the generators when ``--game-count`` is reached).

One generator process per GPU (``--gpus N`` replaces ``--parallel-games-processes``; both are accepted): games are
independent, so the processes share nothing and there is no collective on the data path.  Training (train.py) is out of
scope: when the next model file is missing after generation, the loop stops and says so, unless
``--train-command`` names a program to run (it receives the reference's train.py arguments).
"""
import argparse
import atexit
import os
import signal
import subprocess
import sys
import time


def count_games(paths):
    total = 0
    for path in paths:
        if os.path.exists(path):
            with open(path) as f:
                total += sum(1 for line in f if line.strip())
    return total


def index_to_model_path(args, i):
    return os.path.join(args.prefix, "models", "model-%03i.npy" % i)


def index_to_games_paths(args, i):
    return [os.path.join(args.prefix, "games", "model-%03i-%i.json" % (i, p)) for p in range(args.processes)]


def generate_games(args, model_number):
    paths = index_to_games_paths(args, model_number)
    for path in paths:
        open(path, "a").close()
    if count_games(paths) >= args.game_count:
        print("Enough games to start with!")
        return
    if args.processes > max(args.gpus, 1):
        print("WARNING: %d generator processes share %d GPU(s); each allocates its own node pool." % (args.processes, max(args.gpus, 1)))
    procs = [subprocess.Popen([sys.executable, "-m", "ataxxzero_b200.cli.accelerated_generate_games",
                               "--network", index_to_model_path(args, model_number), "--output-games", path,
                               "--visits", str(args.visits), "--buffer-size", str(args.buffer_size),
                               "--device", str(rank % max(args.gpus, 1))], close_fds=True)
             for rank, path in enumerate(paths)]

    def reap():
        for proc in procs:
            if proc.poll() is None:
                proc.kill()
    atexit.register(reap)
    while True:
        n = count_games(paths)
        print("Game count:", n)
        if n >= args.game_count:
            break
        if all(p.poll() is not None for p in procs):
            # every generator died (bad model file, out of memory, ...) before the target was reached: the reference would
            # block here for ever; training on too few games would be worse, so stop with the children's exit codes
            atexit.unregister(reap)
            raise SystemExit("generators exited with codes %r after %d of %d games" % ([p.returncode for p in procs], n, args.game_count))
        time.sleep(args.poll_seconds)
    for proc in procs:
        if proc.poll() is None:
            os.kill(proc.pid, signal.SIGTERM)       # the generators exit cleanly on SIGTERM
    deadline = time.time() + 30
    for proc in procs:
        try:
            proc.wait(timeout=max(0.1, deadline - time.time()))
        except subprocess.TimeoutExpired:
            proc.kill()
    atexit.unregister(reap)
    print("Exiting.")


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--prefix", metavar="PATH", default=".")
    parser.add_argument("--visits", metavar="N", type=int, default=400)
    parser.add_argument("--game-count", metavar="N", type=int, default=500)
    parser.add_argument("--training-steps-const", metavar="N", type=int, default=200)
    parser.add_argument("--training-steps-linear", metavar="N", type=int, default=50)
    parser.add_argument("--training-window", metavar="N", type=int, default=10)
    parser.add_argument("--training-window-exclude", metavar="N", type=int, default=3)
    parser.add_argument("--parallel-games-processes", metavar="N", type=int, default=None)
    parser.add_argument("--gpus", metavar="N", type=int, default=1, help="Generator processes, one per GPU.")
    parser.add_argument("--buffer-size", metavar="N", type=int, default=1024, help="Each generator plays 2*N games concurrently.")
    parser.add_argument("--poll-seconds", metavar="S", type=float, default=10.0)
    parser.add_argument("--iterations", metavar="N", type=int, default=None, help="Stop after N generation rounds.")
    parser.add_argument("--train-command", metavar="CMD", default=None,
                        help="Program run as CMD --steps S --games PATHS... --old-path OLD --new-path NEW (train.py's arguments).")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    args.processes = args.parallel_games_processes or args.gpus
    print("Arguments:", args)
    current, done = 1, 0
    while args.iterations is None or done < args.iterations:
        start = time.time()
        old_model, new_model = index_to_model_path(args, current), index_to_model_path(args, current + 1)
        if os.path.exists(new_model):
            print("Model already exists, skipping:", new_model)
            current += 1
            continue
        print("=========================== Doing data generation for:", old_model)
        generate_games(args, current)
        done += 1
        low = min(current, max(args.training_window_exclude + 1, current - args.training_window + 1))
        games_paths = sum((index_to_games_paths(args, i) for i in range(low, current + 1)), [])
        steps = args.training_steps_const + args.training_steps_linear * (current - low + 1)
        if args.train_command is None:
            print("=========================== Training is out of scope here; next model expected at", new_model)
            print("Game paths:", games_paths, "Steps:", steps)
            break
        print("=========================== Doing training:", old_model, "->", new_model)
        subprocess.check_call(args.train_command.split() + ["--steps", str(steps), "--games"] + games_paths +
                              ["--old-path", old_model, "--new-path", new_model], close_fds=True)
        print("Total seconds for iteration:", time.time() - start)
        current += 1


if __name__ == "__main__":
    main()

"""accelerated_generate_games.py of the reference (same flags), with the whole loop on the GPU.

The reference feeds TensorFlow from 2*buffer_size C++ worker threads through two shared buffers
(accelerated_generate_games.py:24-83).  Here the 2*buffer_size games live in a device-resident pool and the
network is the tensor-core kernel in the same library, so there is nothing to feed: one call runs until SIGTERM /
SIGINT (how looper.py stops it, looper.py:58-64) or until --game-count / --max-seconds is reached.  Records are
appended to --output-games as the reference's JSON lines."""
import argparse
import signal
import sys
import time

from .. import Context, net, search


def build_parser():
    parser = argparse.ArgumentParser(description="Generates self-play games into the .json format on a B200.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--network", metavar="PATH", required=True, help="Path of the model to load.")
    parser.add_argument("--output-games", metavar="PATH", required=True,
                        help="Path to write .json games to. Writes in append mode, so it won't overwrite existing games.")
    parser.add_argument("--visits", metavar="N", type=int, default=100,
                        help="At each move in the self-play games perform MCTS until the root node has N visits.")
    parser.add_argument("--buffer-size", metavar="N", type=int, default=128,
                        help="Plays 2*N games concurrently (the reference launches 2*N threads for an N-sample buffer).")
    # additions (the reference runs until it is killed and always uses GPU 0)
    parser.add_argument("--device", metavar="N", type=int, default=0, help="GPU index.")
    parser.add_argument("--game-count", metavar="N", type=int, default=0, help="Stop after N finished games (0 = run until signalled).")
    parser.add_argument("--max-seconds", metavar="S", type=float, default=0.0, help="Stop after S seconds (0 = no limit).")
    parser.add_argument("--fp32", action="store_true", help="Evaluate with the fp32 reference-accurate kernel instead of bf16 tensor cores.")
    parser.add_argument("--seed", metavar="N", type=int, default=None, help="RNG seed (default: from the clock, like the reference).")
    parser.add_argument("--one-random-move", action="store_true",
                        help="The reference's compile-time ONE_RANDOM_MOVE variant (self_play_client.cpp:24): one uniformly random ply per "
                             "game, greedy play after it, records carry \"random_ply\".")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    print("Arguments:", args)
    seed = int(time.time_ns()) & (2**63 - 1) if args.seed is None else args.seed
    stop = {"flag": False}

    def handler(signum, frame):
        print("Exiting cleanly...", end=" ")
        sys.stdout.flush()
        stop["flag"] = True
    signal.signal(signal.SIGTERM, handler)
    signal.signal(signal.SIGINT, handler)

    with Context(device=args.device, seed=seed) as ctx:
        net.load_weights(ctx, args.network)
        pool = search.Pool(ctx, 2 * args.buffer_size, args.visits, eval_mode=search.EVAL_FP32 if args.fp32 else search.EVAL_BF16,
                           noise=True, auto_play=True, seed=seed, one_random_move=args.one_random_move)
        start, games, last = time.time(), 0, None
        while not stop["flag"]:
            # a few seconds per call so that signals are honoured promptly
            stats = pool.selfplay(args.output_games, max_seconds=2.0)
            games = stats["games_finished"]
            elapsed = time.time() - start
            if last is None or elapsed - last >= 10.0:
                last = elapsed
                print("Rate: %.3fk evals/s  (Total: %ik)  games: %i  positions/s: %.1f" % (
                    stats["evals"] / elapsed * 1e-3, stats["evals"] * 1e-3, games, stats["positions"] / elapsed))
            if args.game_count and games >= args.game_count:
                break
            if args.max_seconds and elapsed >= args.max_seconds:
                break
        pool.close()
    print("all threads shutdown." if stop["flag"] else "Done generating games.")
    return games


if __name__ == "__main__":
    main()

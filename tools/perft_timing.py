import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataxxzero_b200 as az
from ataxxzero_b200 import rules
ctx = az.Context(0)
for fen in (rules.OPEN_FEN, rules.START_FEN):
    p = rules.set_board(fen)
    for d in (5, 6, 7, 8):
        rules.perft(ctx, p, d)
        t0 = time.perf_counter(); n = rules.perft(ctx, p, d); dt = time.perf_counter() - t0
        print(fen, d, n, "%.3f ms" % (dt * 1e3), "%.1f Mnodes/s" % (n / dt / 1e6), rules.perft_last_stats(ctx), flush=True)

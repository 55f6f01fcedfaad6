"""perft(7) from the opening position, for ncu (k_walk is the kernel that matters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ataxxzero_b200 as az
from ataxxzero_b200 import rules
ctx = az.Context(0)
p = rules.set_board(rules.OPEN_FEN)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 7
print(d, rules.perft(ctx, p, d), rules.perft_last_stats(ctx))

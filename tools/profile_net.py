"""Standalone net forward for ncu: B boards, a few launches (bf16 tcgen05 tower)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, _native
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = az.Context(0)
net.load_weights(ctx, model.Network.random_init(seed=0))
lib = _native.lib()
feats = (torch.rand(B, 196, device="cuda") < 0.3).float()
logits = torch.empty(B, 833, device="cuda"); values = torch.empty(B, device="cuda")
torch.cuda.synchronize()
for _ in range(reps):
    _native.check(lib.az_net_forward_dev(ctx.handle, C.c_void_p(feats.data_ptr()), B, net.BF16, C.c_void_p(logits.data_ptr()), C.c_void_p(values.data_ptr())))
ctx.sync()
print("ok", float(values.abs().sum()))

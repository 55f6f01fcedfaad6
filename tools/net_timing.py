import sys, time, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ataxxzero_b200 as az
from ataxxzero_b200 import model, net, _native
ctx = az.Context(0)
network = model.Network.random_init(seed=0)
net.load_weights(ctx, network)
lib = _native.lib()
dev = torch.device("cuda:0")
for B in (256, 2048, 4096, 16384):
    feats = torch.rand(B, 196, device=dev).round().float()
    logits = torch.empty(B, 833, device=dev); values = torch.empty(B, device=dev)
    torch.cuda.synchronize()
    for mode in (net.BF16, net.FP32):
        if mode == net.FP32 and B > 4096: continue
        for it in range(3):
            t0 = time.perf_counter()
            _native.check(lib.az_net_forward_dev(ctx.handle, C.c_void_p(feats.data_ptr()), B, mode, C.c_void_p(logits.data_ptr()), C.c_void_p(values.data_ptr())))
            ctx.sync(); dt = time.perf_counter() - t0
        print("B=%d mode=%d %.3f ms  %.1f kpos/s  %.1f TFLOP/s" % (B, mode, dt * 1e3, B / dt / 1e3, B * 347.49e6 / dt / 1e12), flush=True)

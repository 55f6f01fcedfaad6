"""Evaluation contract of the reference on the GPU: ``features [B,7,7,4] -> (policy logits
[B,7,7,17] f32, value [B,1] f32)`` exactly as ``sess.run([policy_output, value_output], ...)``
returns them (accelerated_generate_games.py:57-67, engine.py:184-190), plus the RPC wire form
of gpu_server.py:52-56 / rpc_client.py:16-21 (196 int8 in, 3332 float32 bytes + a float out)."""
import ctypes as C

import numpy as np

from . import _native
from ._native import AZ_FEATURES, AZ_LOGITS, AzError, check, lib
from .model import Network

FP32 = 0     # AZ_NET_FP32: CUDA-core fp32, reference-accurate
BF16 = 1     # AZ_NET_BF16: tcgen05 tensor cores, fp32 accumulate
F16 = 3      # AZ_NET_F16: the same kernel with IEEE-half operands (3 more mantissa bits)

_vp = C.c_void_p
_native.register("az_net_load", C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_int])
_native.register("az_net_param_count", C.c_size_t, [C.c_int, C.c_int])
_native.register("az_net_forward", C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp])
_native.register("az_net_forward_i8", C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp])
_native.register("az_net_forward_sym8", C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp])
_native.register("az_net_forward_dev", C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp])
_native.register("az_net_forward_pos_dev", C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp])


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def load_weights(ctx, network):
    """Upload a ``model.Network`` (or a path to a reference ``.npy``) to the context's GPU."""
    if isinstance(network, str):
        network = Network.load(network)
    packed = network.packed()
    expect = lib().az_net_param_count(network.filters, network.blocks)
    if packed.size != expect:
        raise AzError(-1, "packed weight count %d != %d" % (packed.size, expect))
    check(lib().az_net_load(ctx.handle, _ptr(packed), packed.size, network.filters, network.blocks))
    return network


def load_packed(ctx, packed, filters=128, blocks=12):
    """Upload an already packed float32 parameter vector (``Network.packed()``), e.g. one kept in pinned host memory."""
    packed = np.ascontiguousarray(packed, dtype=np.float32)
    check(lib().az_net_load(ctx.handle, _ptr(packed), packed.size, int(filters), int(blocks)))


def forward(ctx, features, mode=BF16):
    """features: float32/int8 array [B,7,7,4] (or [B,196]) -> (logits [B,7,7,17], values [B,1])."""
    feats = np.asarray(features)
    n = feats.shape[0] if feats.ndim > 1 else 1
    if feats.size != n * AZ_FEATURES:
        raise AzError(-1, "features must have 196 entries per board, got shape %r" % (feats.shape,))
    logits = np.zeros((max(n, 1), 7, 7, 17), dtype=np.float32)
    values = np.zeros((max(n, 1), 1), dtype=np.float32)
    if feats.dtype == np.int8:
        feats = np.ascontiguousarray(feats)
        check(lib().az_net_forward_i8(ctx.handle, _ptr(feats), n, mode, _ptr(logits), _ptr(values)))
    else:
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        check(lib().az_net_forward(ctx.handle, _ptr(feats), n, mode, _ptr(logits), _ptr(values)))
    return logits[:n], values[:n]


def apply_symmetry(tensor, symmetry):
    """nn_evals.py:7-15."""
    assert 0 <= symmetry < 8
    if symmetry & 1:
        tensor = tensor[::-1, :]
    if symmetry & 2:
        tensor = tensor[:, ::-1]
    if symmetry & 4:
        tensor = np.moveaxis(tensor, 0, 1)
    return tensor


inverse_symmetry = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 6, 6: 5, 7: 7}      # nn_evals.py:27


def evaluate_symmetric(ctx, features, mode=BF16):
    """nn_evals.evaluate (nn_evals.py:48-62) for a batch: features [B,7,7,4] -> (mean policy [B,7,7,17], mean value [B])
    over the 8 dihedral images of every position, expanded / reduced on the GPU (az_net_forward_sym8)."""
    feats = np.ascontiguousarray(features, dtype=np.float32).reshape(-1, 7, 7, 4)
    n = len(feats)
    logits = np.zeros((max(n, 1), 7, 7, 17), dtype=np.float32)
    values = np.zeros(max(n, 1), dtype=np.float32)
    check(lib().az_net_forward_sym8(ctx.handle, _ptr(feats), n, mode, _ptr(logits), _ptr(values)))
    return logits[:n], values[:n]


def network_rpc(ctx, feature_bytes, mode=BF16):
    """gpu_server.py:78-82 ``network(feature_string)``: 196 int8 bytes -> (3332 float32 bytes, float)."""
    if len(feature_bytes) != AZ_FEATURES:
        raise AzError(-1, "feature string must be 196 bytes (gpu_server.py:36)")
    feats = np.frombuffer(feature_bytes, dtype=np.int8).reshape(1, 7, 7, 4)
    logits, values = forward(ctx, feats, mode)
    return logits.astype(np.float32).tobytes(), float(values[0, 0])

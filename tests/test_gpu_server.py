"""The batched evaluation service (gpu_server.py contract): wire format, batching rule (<= 16 positions or 10 ms),
concurrent clients, error paths -- with a stub evaluator on CPU, and with the GPU network under -m gpu."""
import socket
import threading
import time

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stub(batch):
    """logits[b, x, y, k] = sum(features[b]) + k ; value = (#ones in plane 1) / 49"""
    b = len(batch)
    base = batch.reshape(b, -1).sum(axis=1).astype(np.float32)
    logits = np.zeros((b, 7, 7, 17), dtype=np.float32) + base[:, None, None, None] + np.arange(17, dtype=np.float32)
    values = batch[..., 1].reshape(b, -1).sum(axis=1, keepdims=True).astype(np.float32) / 49.0
    return logits, values


def _serve(evaluate):
    from ataxxzero_b200 import gpu_server
    server = gpu_server.NetworkServer(_free_port(), evaluate)
    thread = threading.Thread(target=server.serve_forever, daemon=True)
    thread.start()
    return server, server.server_address[1]


def test_wire_contract_and_batching():
    from ataxxzero_b200 import gpu_server
    server, port = _serve(_stub)
    try:
        feats = np.zeros((7, 7, 4), dtype=np.int8)
        feats[..., 0] = 1
        feats[0, 0, 1] = feats[6, 6, 1] = 1
        client = gpu_server.RPCClient("127.0.0.1", port)
        posterior, value = client.call("network", feats.tobytes())
        assert len(posterior) == 7 * 7 * 17 * 4 and isinstance(value, float)          # gpu_server.py:52-56
        got = np.frombuffer(posterior, dtype=np.float32).reshape(7, 7, 17)
        assert np.array_equal(got[3, 3], 51.0 + np.arange(17)) and abs(value - 2 / 49) < 1e-7
        with pytest.raises(RuntimeError):
            client.call("network", b"too short")
        with pytest.raises(RuntimeError):
            client.call("no_such_method", feats.tobytes())
        # 40 concurrent clients: marshalled into batches of at most 16 (MARSHALL_COUNT), each gets its own row
        results = {}

        def worker(i):
            f = feats.copy()
            f[i % 7, i // 7, 2] = 1                     # a distinguishable position
            f[1, 1, 1] = i % 2
            c = gpu_server.RPCClient("127.0.0.1", port)
            results[i] = (f, c.call("network", f.tobytes()))
            c.close()
        before = len(server.processor.batch_sizes)
        threads = [threading.Thread(target=worker, args=(i,)) for i in range(40)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=30)
        assert len(results) == 40
        for i, (f, (posterior, value)) in results.items():
            want_l, want_v = _stub(f[None])
            assert np.array_equal(np.frombuffer(posterior, dtype=np.float32), want_l.reshape(-1)) and abs(value - want_v[0, 0]) < 1e-7
        sizes = server.processor.batch_sizes[before:]
        assert sum(sizes) == 40 and max(sizes) <= 16 and len(sizes) < 40          # really batched
        # a lone request is answered after at most ~MAXIMUM_WAIT_TIME
        t0 = time.time()
        client.call("network", feats.tobytes())
        assert time.time() - t0 < 0.5
        client.close()
    finally:
        server.shutdown()
        server.server_close()


@pytest.mark.gpu
def test_service_with_gpu_network_and_rpc_evaluator(tmp_path, ctx):
    from ataxxzero_b200 import ataxx_rules, engine, gpu_server, model, net, rpc_client
    network = model.Network.random_init(seed=0)
    path = str(tmp_path / "model-001.npy")
    network.save(path)
    server, port = _serve(gpu_server.gpu_evaluator(path))
    try:
        rpc_client.setup_rpc(port)
        board = ataxx_rules.AtaxxState.initial()
        rpc_client.RPCEvaluator(temperature=0.0).populate(board)
        assert set(board.evaluations.posterior) == set(board.legal_moves()) and abs(sum(board.evaluations.posterior.values()) - 1) < 1e-4
        net.load_weights(ctx, network)
        logits, values = net.forward(ctx, engine.board_to_features(board)[None], net.BF16)
        posterior, value = rpc_client.rpc_connection.call("network", engine.board_to_features(board).tobytes())
        assert np.array_equal(np.frombuffer(posterior, dtype=np.float32), logits.reshape(-1)) and value == float(values[0, 0])
    finally:
        server.shutdown()
        server.server_close()

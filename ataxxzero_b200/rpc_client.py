"""Client side of the evaluation service (``ataxxzero_b200.gpu_server``), with the surface of the reference's
``rpc_client.py`` (rpc_client.py:11-56): ``setup_rpc(port)``, ``evaluate(board, temperature)`` ->
``(posterior dict over legal moves, value)`` and ``RPCEvaluator(temperature).populate(board)``.

The wire call is ``network(196 int8 feature bytes) -> (3332 float32 logit bytes, float)``; turning logits into a
posterior over legal moves is shared with the local evaluator (``engine.posterior_from_logits``)."""
import numpy as np

from . import engine, gpu_server

rpc_connection = None


def setup_rpc(port=6000, host="127.0.0.1"):
    global rpc_connection
    rpc_connection = gpu_server.RPCClient(host, port)


def evaluate(board, temperature):
    planes = engine.board_to_features(board).astype(np.int8, copy=False)
    logit_bytes, value = rpc_connection.call("network", planes.tobytes())
    logits = np.frombuffer(logit_bytes, dtype=np.float32).reshape(7, 7, 17)
    return engine.posterior_from_logits(board, logits, temperature), value


class RPCEvaluator(engine.NNEvaluator):
    """``NNEvaluator`` whose network lives behind the service: same ``populate`` / adjudication, no local cache."""

    def __init__(self, temperature=0.0):
        super().__init__(temperature)

    def __contains__(self, board):
        return False

    def evaluate(self, input_board):
        posterior, value = evaluate(input_board, self.temperature)
        self.cache = {engine.NNEvaluator.board_key(input_board): engine.NNEvaluator.Entry(input_board, value, posterior, False)}

    def add_to_queue(self, board):
        pass

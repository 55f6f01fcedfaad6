"""Client side of the evaluation service, with the surface of the reference's ``rpc_client.py`` (rpc_client.py:11-56):
``setup_rpc(port)``, ``evaluate(board, temperature)``, ``RPCEvaluator(temperature).populate(board)``."""
import numpy as np

from . import engine, gpu_server

rpc_connection = None


def setup_rpc(port=6000, host="127.0.0.1"):
    global rpc_connection
    rpc_connection = gpu_server.RPCClient(host, port)


def evaluate(board, temperature):
    features = engine.board_to_features(board)
    assert features.dtype == np.int8
    feature_string = features.tobytes()
    assert len(feature_string) == gpu_server.FEATURE_BYTES
    posterior, value = rpc_connection.call("network", feature_string)
    raw_posterior = np.frombuffer(posterior, dtype=np.float32).reshape((7, 7, 17))
    if temperature:
        raw_posterior = raw_posterior + np.random.randn(7, 7, 17) * temperature
    softmax_posterior = engine.softmax(raw_posterior)
    posterior = {move: float(engine.get_move_score(softmax_posterior, move)) for move in board.legal_moves()}
    denominator = sum(posterior.values()) + 1e-6
    return {move: prob / denominator for move, prob in posterior.items()}, value


class RPCEvaluator:
    def __init__(self, temperature=0.0):
        self.temperature = temperature
        self.cache = {}

    def populate(self, board):
        if getattr(board, "evaluations", None) is not None:
            return
        posterior, value = evaluate(board, self.temperature)
        entry = engine.NNEvaluator.Entry(board=board, value=value, posterior=posterior, game_over=False)
        result = board.result()
        if result is not None:
            entry.value = 1.0 if result == board.to_move else -1.0
            entry.game_over = True
        board.evaluations = entry

    def add_to_queue(self, board):
        pass

"""The reference's ``engine.py`` search interface on the GPU library.

Same module surface (engine.py:19-611): ``initialize_model``, ``setup_evaluator``, ``board_to_features``,
``softmax``, ``get_move_score``, ``add_move_to_heatmap``, ``encode_move_as_heatmap``, ``sample_by_weight``,
``add_dirichlet_noise_to_posterior``, ``NNEvaluator``, ``MCTSEdge``, ``MCTSNode``, ``MCTS``, ``MCTSEngine``.

What changed underneath: the tree is a device-resident PUCT tree (``search.Pool`` with one game) and the network is
the tensor-core kernel, so ``MCTS.step()`` is one tree-kernel + one net-kernel launch and ``MCTSEngine.genmove``
advances the search in bursts instead of one Python-level step at a time.  ``MCTSNode`` / ``MCTSEdge`` objects are
read-only snapshots of the root (that is all the reference's callers -- ``genmove``, ``uai_interface``,
``generate_games`` -- ever look at).  Search semantics are those of the reference's C++ core
(cpp/self_play_client.cpp:310-473: c_puct = 1, first-play value 0, no virtual loss, tree reuse on ``play``), which the
Python original shares up to its ``+1e-6`` prior renormalisation and dict-order tie-breaks.
"""
import logging
import random
import time

import numpy as np

from . import ataxx_rules, model, net, rules, search
from ._native import Context

RED = "\x1b[91m"
ENDC = "\x1b[0m"
DIRICHLET_ALPHA = 0.15
DIRICHLET_WEIGHT = 0.25

initialized = False
context = None              # the GPU context the model lives on
network = None              # model.Network (weights as loaded from the .npy)
global_evaluator = None
eval_mode = net.BF16        # net.FP32 for the reference-accurate CUDA-core path


def initialize_model(path, device=0, seed=None):
    """engine.py:19-27: load ``path`` (reference ``.npy``) onto the GPU.  ``path`` may also be a model.Network."""
    global network, context, initialized
    assert not initialized
    context = Context(device=device, seed=random.getrandbits(63) if seed is None else seed)
    network = net.load_weights(context, path)
    initialized = True


def setup_evaluator(use_rpc=False, temperature=0.0):
    """engine.py:29-38.  ``use_rpc``: evaluate through a running ``ataxxzero_b200.gpu_server`` (port 6000) instead of a
    local context; that evaluator serves NNEvaluator-style ``populate`` calls, the device-resident MCTS always uses the
    local network."""
    global global_evaluator
    if use_rpc:
        print("Using RPC evaluator.")
        from . import rpc_client
        rpc_client.setup_rpc()
        global_evaluator = rpc_client.RPCEvaluator(temperature=temperature)
    else:
        global_evaluator = NNEvaluator(temperature=temperature)


def sample_by_weight(weights):
    assert abs(sum(weights.values()) - 1) < 1e-6, "Distribution not normalized: %r" % (weights,)
    x = random.random()
    for outcome, weight in weights.items():
        if x <= weight:
            return outcome
        x -= weight
    return next(iter(weights.keys()))


def softmax(logits):
    e_x = np.exp(logits - np.max(logits))
    return e_x / e_x.sum()


def board_to_features(board):
    """engine.py:53-73: int8 [7,7,4] planes (ones, side to move, opponent, blockers), indexed [x][y]."""
    features = np.zeros((model.BOARD_SIZE, model.BOARD_SIZE, model.Network.INPUT_FEATURE_COUNT), dtype=np.int8)
    features[:, :, 0] = 1
    blocked = getattr(board, "blocked", ataxx_rules.BLOCKED_CELLS)
    for y in range(model.BOARD_SIZE):
        for x in range(model.BOARD_SIZE):
            piece = board[x, y]
            if piece:
                features[x, y, 1 if piece == board.to_move else 2] = 1
            if (x, y) in blocked:
                features[x, y, 3] = 1
    return features


position_delta_layers = {delta: i for i, delta in enumerate(ataxx_rules.FAR_NEIGHBOR_OFFSETS)}
assert len(position_delta_layers) == 16


def _plane(move):
    start, end = move
    if start == "c":
        return end[0], end[1], model.MOVE_TYPES - 1
    return end[0], end[1], position_delta_layers[(end[0] - start[0], end[1] - start[1])]


def add_move_to_heatmap(heatmap, move, coef=1):
    heatmap[_plane(move)] += coef


def encode_move_as_heatmap(move):
    heatmap = np.zeros((model.BOARD_SIZE, model.BOARD_SIZE, model.MOVE_TYPES), dtype=np.int8)
    add_move_to_heatmap(heatmap, move)
    return heatmap


def get_move_score(softmaxed_posterior, move):
    assert softmaxed_posterior.shape == (7, 7, 17)
    if move == "pass":
        return 1.0
    return softmaxed_posterior[_plane(move)]


def add_dirichlet_noise_to_posterior(posterior, alpha, weight):
    noise = np.random.dirichlet([alpha] * len(posterior))
    return {move: (1.0 - weight) * prob + weight * n for (move, prob), n in zip(posterior.items(), noise)}


def posterior_from_logits(board, raw, temperature=0.0):
    """Raw policy logits [7,7,17] -> prior over the board's legal moves (engine.py:194-203 / rpc_client.py:22-29):
    optional Gaussian logit noise, softmax, gather per move, renormalise with the reference's +1e-6."""
    if temperature:
        raw = raw + np.random.randn(*raw.shape) * temperature
    sm = softmax(raw)
    posterior = {move: float(get_move_score(sm, move)) for move in board.legal_moves()}
    denominator = sum(posterior.values()) + 1e-6
    return {move: p / denominator for move, p in posterior.items()}


class NNEvaluator:
    """engine.py:121-233: evaluation cache in front of the network.  ``evaluate`` runs the GPU net."""
    MAXIMUM_CACHE_ENTRIES = 200000

    class Entry:
        __slots__ = ["board", "value", "posterior", "noisy_posterior", "game_over"]

        def __init__(self, board, value, posterior, game_over):
            self.board, self.value, self.posterior, self.game_over = board, value, posterior, game_over
            self.noisy_posterior = None

        def populate_noisy_posterior(self):
            if self.noisy_posterior is None:
                self.noisy_posterior = add_dirichlet_noise_to_posterior(self.posterior, DIRICHLET_ALPHA, DIRICHLET_WEIGHT)

    def __init__(self, temperature=0.0):
        self.temperature = temperature
        self.cache = {}
        self.board_queue = []
        self.ensemble_sizes = []

    @staticmethod
    def board_key(b):
        return (b.to_move, tuple(b.board))

    def __contains__(self, board):
        return NNEvaluator.board_key(board) in self.cache

    def add_to_queue(self, board):
        if board not in self:
            self.board_queue.append(board)

    def evaluate(self, input_board):
        ensemble = [input_board] + [b for b in self.board_queue if b not in self]
        self.board_queue = []
        self.ensemble_sizes.append(len(ensemble))
        features = np.stack([board_to_features(b) for b in ensemble])
        posteriors, values = net.forward(context, features, eval_mode)
        for board, raw, (value,) in zip(ensemble, posteriors, values):
            posterior = posterior_from_logits(board, raw, self.temperature)
            self.cache[NNEvaluator.board_key(board)] = NNEvaluator.Entry(board, float(value), posterior, False)

    def populate(self, board):
        if getattr(board, "evaluations", None) is not None:
            return
        if board not in self:
            self.evaluate(board)
        entry = self.cache[NNEvaluator.board_key(board)]
        result = board.result()
        if result is not None:
            entry.value = 1.0 if result == board.to_move else -1.0
            entry.game_over = True
        board.evaluations = entry
        if len(self.cache) > NNEvaluator.MAXIMUM_CACHE_ENTRIES:
            self.cache = {}


class MCTSEdge:
    """Snapshot of one edge out of the root (engine.py:235-259)."""

    def __init__(self, move, child_node=None, parent_node=None, edge_visits=0, edge_total_score=0.0):
        self.move, self.child_node, self.parent_node = move, child_node, parent_node
        self.edge_visits, self.edge_total_score = edge_visits, edge_total_score

    def get_edge_score(self):
        return self.edge_total_score / self.edge_visits

    def __str__(self):
        from .cli import uai_interface
        return "<%s v=%i s=%.5f>" % (uai_interface.uai_encode_move(self.move), self.edge_visits,
                                     self.get_edge_score() if self.edge_visits else 0.0)


class MCTSNode:
    """Snapshot of a tree node (engine.py:261-302): board, all_edge_visits, outgoing_edges, posterior."""

    def __init__(self, board, parent=None):
        self.board, self.parent = board, parent
        self.all_edge_visits = 0
        self.outgoing_edges = {}
        self.posterior = {}

    def total_action_score(self, move):
        u = MCTS.exploration_parameter * self.posterior[move] * (1.0 + self.all_edge_visits) ** 0.5
        edge = self.outgoing_edges.get(move)
        if edge is None:
            return u
        return (edge.get_edge_score() if edge.edge_visits > 0 else 0.0) + u / (1.0 + edge.edge_visits)

    def select_action(self, use_dirichlet_noise=False):
        if not self.posterior or self.board.result() is not None:
            return None
        return max(self.posterior, key=self.total_action_score)


class MCTS:
    """engine.py:320-447 on a device-resident tree (one game of a ``search.Pool``)."""
    exploration_parameter = 1.0
    NODE_CAPACITY = 1 << 17           # tree slots kept on the GPU (9 KB each)
    SPECULATE = 4                     # children evaluated ahead per consumed node (0: one leaf per net round trip)

    def __init__(self, root_board, use_dirichlet_noise=False, visits=None):
        assert initialized, "call engine.initialize_model(path) first"
        self.use_dirichlet_noise = use_dirichlet_noise
        self.board = root_board.copy()
        # speculate: the device twin of NNEvaluator.add_to_queue (engine.py:166-175,387-392) -- likely children of every new
        # node are evaluated in the same batch and linked from a cache when the search reaches them
        self.pool = search.Pool(context, 1, visits or 1, eval_mode=eval_mode, noise=use_dirichlet_noise, auto_play=False,
                                node_capacity=self.NODE_CAPACITY, steps_per_tick=64,
                                speculate=self.SPECULATE if eval_mode != search.EVAL_FP32 else 0)
        self.pool.set_root(0, self.board.to_position())
        self._target = 0
        self._snapshot = None

    def close(self):
        self.pool.close()

    # ---- search ----
    def search(self, visits):
        """Advance until the root has ``visits`` visits (the C++ loop at self_play_client.cpp:522)."""
        visits = min(int(visits), self.NODE_CAPACITY - 2)
        if visits > self._target:
            self._target = visits
            self.pool.set_visits(visits)
        self.pool.run()
        self._snapshot = None

    def step(self):
        """One MCTS step (engine.py:356-424); returns the first edge out of the root that the step went through."""
        before = self.root_node
        if not before.posterior and before.all_edge_visits == 0 and before.board.result() is not None:
            return None
        self.search(before.all_edge_visits + 1)
        after = self.root_node
        for move, edge in after.outgoing_edges.items():
            old = before.outgoing_edges.get(move)
            if edge.edge_visits != (old.edge_visits if old else 0):
                return edge
        return None

    @property
    def root_node(self):
        if self._snapshot is None:
            r = self.pool.root(0)
            node = MCTSNode(self.board)
            node.all_edge_visits = r["root_visits"]
            for mv, n, w, p in zip(r["moves"], r["visits"], r["total_score"], r["prior"]):
                move = rules.to_reference_move(mv)
                node.posterior[move] = float(p)
                if n > 0:
                    node.outgoing_edges[move] = MCTSEdge(move, None, node, int(n), float(w))
            self._snapshot = node
        return self._snapshot

    def select_principal_variation(self, best=False):
        """``best=True``: the most-visited line as a list of edges (engine.py:331-336).  The PUCT descent itself
        (``best=False``) happens inside the tree kernel; here it returns the root's PUCT choice only."""
        root = self.root_node
        if not best:
            move = root.select_action(self.use_dirichlet_noise)
            edge = root.outgoing_edges.get(move)
            return root, move, [edge] if edge else []
        pv = [MCTSEdge(rules.to_reference_move(mv), None, None, n, 0.0) for mv, n in self.pool.principal_variation(0)]
        return root, (pv[-1].move if pv else None), pv

    def play(self, player, move, print_variation_count=True):
        assert self.board.to_move == player, "Bad play direction for MCTS!"
        if print_variation_count:
            edge = self.root_node.outgoing_edges.get(move)
            logging.debug("Traversing to variation with %i visits." % edge.edge_visits if edge else "Completely unexpected variation!")
        self.pool.play(0, rules.from_reference_move(move))
        self.board.move(move)
        self._target = 0
        self._snapshot = None


class MCTSEngine:
    """engine.py:449-572."""
    VISITS = 10000000
    MAX_STEPS = 10000000
    TIME_SAFETY_MARGIN = 0.1
    IMPORTANCE_FACTOR = {1: 0.1, 2: 0.2, 3: 0.3, 4: 0.4, 5: 0.5, 6: 0.6, 7: 0.7, 8: 0.8, 9: 0.9}
    BURST = 64                       # MCTS steps per host round trip while thinking

    def __init__(self):
        self.state = ataxx_rules.AtaxxState.initial()
        self.mcts = MCTS(self.state)
        self.plies_played = 0

    def set_state(self, new_board):
        new_board = new_board.copy()
        # reuse the subtree when the new position is two plies below the current one (engine.py:467-479)
        if self.state.result() is None:
            for m1 in self.state.legal_moves():
                mid = self.state.copy()
                mid.move(m1)
                if m1 == "pass" or mid.result() is not None:
                    continue
                for m2 in mid.legal_moves():
                    if m2 == "pass":
                        continue
                    after = mid.copy()
                    after.move(m2)
                    if after == new_board:
                        self.mcts.play(self.state.to_move, m1, print_variation_count=False)
                        self.mcts.play(mid.to_move, m2, print_variation_count=False)
                        self.state = new_board
                        self.plies_played += 2
                        return
        logging.debug(RED + "Failed to match a subtree." + ENDC)
        self.state = new_board
        self.mcts.close()
        self.mcts = MCTS(self.state)

    def genmove(self, time_to_think, early_out=True, use_weighted_exponent=None):
        start_time = time.time()
        if self.state.result() is not None or self.state.legal_moves() == ["pass"]:
            return "pass"
        limit = min(self.MAX_STEPS, MCTS.NODE_CAPACITY - 2)
        steps0 = self.mcts.root_node.all_edge_visits
        total_steps = 0
        while total_steps < limit:
            remaining = time_to_think - (time.time() - start_time)
            if remaining <= 0.0 and total_steps > 0:
                break
            root = self.mcts.root_node
            top = sorted((e.edge_visits for e in root.outgoing_edges.values()), reverse=True)[:2]
            if top and top[0] >= self.VISITS:
                break
            if early_out and len(top) == 2 and total_steps > 0:
                rate = total_steps / max(time.time() - start_time, 1e-6)
                if top[1] + remaining * rate < top[0]:
                    logging.debug("Early out; cannot catch up in %f seconds." % (remaining,))
                    break
            burst = min(self.BURST, limit - total_steps)
            self.mcts.search(root.all_edge_visits + burst)
            done = self.mcts.root_node.all_edge_visits - steps0
            if done == total_steps:          # no progress: terminal root or exhausted tree
                break
            total_steps = done
        logging.debug("Completed %i steps." % total_steps)
        self.print_principal_variation()
        print("info speed %f nps" % (total_steps / max(time.time() - start_time, 1e-9),))
        edges = self.mcts.root_node.outgoing_edges
        if not edges:
            return "pass"
        if not use_weighted_exponent:
            return max(edges.values(), key=lambda e: e.edge_visits).move
        return self.sample_with_exponential_weight(use_weighted_exponent)

    def sample_with_exponential_weight(self, exponent):
        root = self.mcts.root_node
        max_visits = max(e.edge_visits for e in root.outgoing_edges.values())
        weights = {m: (e.edge_visits / float(root.all_edge_visits)) ** exponent
                   for m, e in root.outgoing_edges.items() if e.edge_visits >= max_visits * 0.5}
        norm = 1.0 / sum(weights.values())
        return sample_by_weight({m: w * norm for m, w in weights.items()})

    def genmove_with_time_control(self, our_time, our_increment):
        moves_remaining = 20.0
        time_budget = (our_time + our_increment * moves_remaining) / moves_remaining
        # the reference reads a non-existent `state.fullmove_number` here; plies // 2 + 1 is what it meant
        time_budget *= self.IMPORTANCE_FACTOR.get(self.plies_played // 2 + 1, 1.3) * 1.5
        time_budget = max(0.0, min(time_budget, 0.5 * our_time - self.TIME_SAFETY_MARGIN))
        logging.debug("Budgeting %.2fms for this move." % (time_budget * 1e3,))
        return self.genmove(time_budget)

    def print_principal_variation(self):
        from .cli import uai_interface
        _, _, pv = self.mcts.select_principal_variation(best=True)
        logging.debug("PV [%2i]: %s" % (len(pv), " ".join(uai_interface.uai_encode_move(e.move) for e in pv)))

"""Agent -- mirror of code/utils/agent.py:7-147 (root search driver).  The epoch / depth loops, the Q cache and the
in-tree policy (MCTSAgent, agent.py:149-223, and MCTSMPGameRunner, mp_game_runner.py:79-115) run as CUDA kernels
inside the Engine; this class keeps the reference's constructor, `make_moves(games, ids)`, `records` / `values`
and `clear()`."""
import numpy as np
import torch


class Agent:

    def __init__(self, nnet, softmax_base=100, training=False, max_MCTS_depth=8, max_MCTS_breadth=128):
        self.nnet = nnet
        self.softmax_base = softmax_base
        self.training = training
        self.max_MCTS_depth = max_MCTS_depth
        self.max_MCTS_breadth = max_MCTS_breadth
        self._engine = None
        # record data for training (agent.py:21-23)
        if training:
            self.records = []
            self.values = []

    def make_moves(self, games, ids):
        eng = getattr(games, "engine", None)
        if eng is None:
            raise TypeError("Agent.make_moves needs the games of an engine-backed MPGameRunner (no CPU path)")
        self._engine = eng
        value_fn = native = None
        if self.nnet is not None and not getattr(self.nnet, "is_stub", False):
            if getattr(self.nnet, "backend", "native") == "torch":
                value_fn = self.nnet.v_device          # explicit PyTorch reference network: Python-driven loop
            else:
                native = self.nnet._get_native()       # product path: the whole root turn in one native call
        q, mv = eng.search(value_fn=value_fn, net=native)
        qh = q.cpu().numpy()
        mvh = mv.cpu().numpy()
        moves = [int(mvh[g, s]) for g, s in ids]
        if any(m > 2 for m in moves):      # a live snake without a searched row: never hand an arbitrary direction to the tic
            from .._lib import AszError
            raise AszError("search returned no move for a live snake (ids do not match the engine's live snakes)")
        if self.training:
            # agent.py:93-97: root states and their Q rows (snapshots; the reference stores aliases)
            planes, rows = eng.encode_rows()
            row_of = {(int(r) // 8, int(r) % 8): i for i, r in enumerate(rows)}
            ph = planes.cpu().numpy()
            for g, s in ids:
                self.records.append(ph[row_of[(g, s)]])
                self.values.append(qh[g, s].copy())
        return moves

    # agent.py:114-122 (host restatement for callers that use it directly)
    def softermax(self, z):
        z = np.asarray(z, dtype=np.float32)
        with np.errstate(divide="ignore"):
            normalized = np.power(np.float32(self.softmax_base), np.arctanh(z)).astype(np.float32)
        sigma = np.float32(0)
        for n in normalized:
            sigma = np.float32(sigma + n)
        if sigma == 0.0:
            return np.array([1.0 / 3.0] * 3, dtype=np.float32)
        return normalized / sigma

    # agent.py:124-137
    def argmaxs(self, Z):
        out = [-1] * len(Z)
        for i in range(len(Z)):
            if Z[i][0] > Z[i][1]:
                out[i] = 0 if Z[i][0] > Z[i][2] else 2
            else:
                out[i] = 1 if Z[i][1] > Z[i][2] else 2
        return out

    # agent.py:140-147
    def clear(self):
        if self._engine is not None:
            self._engine.search_clear()
        if self.training:
            self.records = []
            self.values = []


class StubNet:
    """deterministic value function evaluated inside the engine (value from the plane key + obstacle mask)."""
    is_stub = True

"""MPGameRunner -- mirror of code/utils/mp_game_runner.py:5-77 over a device-resident Engine.

Same constructor, `run(Alice)` return value and log attributes as the reference.  All game_cnt games advance in one
fused tic(+encode) launch per root turn; the agent's search runs on the same device between tics."""
from time import time

import numpy as np
import torch

from ..engine import Engine
from .. import _lib
from .game import Game, frames_text, replay_frames


class GameDict(dict):
    """dict {game_id: Game view} of the live games, carrying the engine for engine-aware agents."""
    engine = None


class MPGameRunner:

    def __init__(self, height=11, width=11, snake_cnt=4, health_dec=1, game_cnt=1, seed=0, verbose=True, device=None,
                 table_log2=0):
        if height != width:
            raise ValueError("the value network needs square boards (rot90 of the plane, game.py:257)")
        self.height, self.width, self.snake_cnt, self.health_dec, self.game_cnt = height, width, snake_cnt, health_dec, game_cnt
        self.seed, self.verbose, self.device, self.table_log2 = seed, verbose, device, table_log2
        self.engine = None
        self.games = GameDict()
        # log (mp_game_runner.py:14-20)
        self.wall_collision = 0
        self.body_collision = 0
        self.head_collision = 0
        self.starvation = 0
        self.food_eaten = 0
        self.game_length = 0

    def _make_engine(self, Alice):
        kw = dict(side=self.height, snakes=self.snake_cnt, health_dec=self.health_dec, games=self.game_cnt, seed=self.seed,
                  device=self.device, table_log2=self.table_log2)
        if hasattr(Alice, "max_MCTS_breadth"):
            kw.update(max_depth=Alice.max_MCTS_depth, max_breadth=Alice.max_MCTS_breadth,
                      softmax_base=float(Alice.softmax_base), training=bool(Alice.training))
        self.engine = Engine(**kw)
        self.engine.reset()
        self.games = GameDict({i: Game(self.engine, i) for i in range(self.game_cnt)})
        self.games.engine = self.engine

    def live_ids(self):
        """[(game_id, snake_id)] of every live snake of every live game, game order then snake order
        (mp_game_runner.py:40-42)."""
        alive = self.engine.alive_mask().cpu().numpy()
        return [(int(g), int(s)) for g, s in zip(*np.nonzero(alive))]

    # Alice is the agent
    def run(self, Alice, max_turns=None):
        """mp_game_runner.py:23-77.  max_turns (extension, default None = the reference's behaviour) stops after that many
        root turns and leaves the remaining games in self.games so that a later call continues them."""
        t0 = time()
        if self.engine is None:
            self._make_engine(Alice)
        eng, games = self.engine, self.games
        turn = 0
        if not hasattr(self, "_rewards"):
            self._rewards = [None] * self.game_cnt
        rewards = self._rewards
        while games and (max_turns is None or turn < max_turns):
            turn += 1
            if self.verbose:
                if len(games) == 1:
                    print("Running the root game. On turn", str(turn) + "...")
                else:
                    print("Concurrently running", len(games), "root games. On turn", str(turn) + "...")
            ids = self.live_ids()
            moves = Alice.make_moves(games, ids)
            actions = np.ones((self.game_cnt, 8), np.uint8)
            for (g, s), m in zip(ids, moves):
                actions[g, s] = m
            show = self.game_cnt == 1                      # mp_game_runner.py:26: a single game is recorded in replay.rep
            pre = eng.get_state(0) if show else None
            eng.step(actions=torch.from_numpy(actions).to(eng.device), spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=False)
            if show:                                       # game.py:140-141,194-195: two frames per tic
                with open("replay.rep", "a") as f:
                    f.write(frames_text(replay_frames(pre, [m for (g, s), m in zip(ids, moves)], eng.get_state(0), self.width)))
            ended = eng.ended.cpu().numpy()
            rw = eng.rewards.cpu().numpy()
            for g in np.nonzero(ended)[0]:
                g = int(g)
                if g in games:
                    rewards[g] = [None if r == 0 else float(r) for r in rw[g, :self.snake_cnt]]
                    del games[g]
            if self.verbose:
                print("Root game turn", str(turn), "finished. Total time spent:", time() - t0, end="\n\n")
        tot = eng.totals()
        # mp_game_runner.py:56-61,71-76: per-game averages
        self.wall_collision = tot["wall"] / self.game_cnt
        self.body_collision = tot["body"] / self.game_cnt
        self.head_collision = tot["head"] / self.game_cnt
        self.starvation = tot["starve"] / self.game_cnt
        self.food_eaten = tot["food_eaten"] / self.game_cnt
        self.game_length = tot["game_length"] / self.game_cnt
        return rewards

"""pit MPGameRunner -- mirror of code/utils/pit_mp_game_runner.py:3-63 (two agents, argmax policy, early exit).
Uses the same fused tic+encode kernel; the planes of all live snakes are produced by one launch per turn."""
import numpy as np
import torch

from ..engine import Engine
from .. import _lib


class MPGameRunner:

    def __init__(self, height=11, width=11, snake_cnt=4, health_dec=1, game_cnt=1, seed=0, device=None):
        self.height, self.width, self.snake_cnt, self.health_dec, self.game_cnt = height, width, snake_cnt, health_dec, game_cnt
        self.engine = Engine(side=height, snakes=snake_cnt, health_dec=health_dec, games=game_cnt, seed=seed, device=device)
        self.engine.reset()

    # Alice and Bob are agents using different nets
    def run(self, Alice, Bob, Alice_snake_cnt=None, spawn_trace=None, move_log=None):
        """pit_mp_game_runner.py:14-63.  Extensions for trace replay (tests): spawn_trace[turn] = int32 [game_cnt] food cell
        spawned in that turn's tic (-1 none) instead of the engine's RNG; move_log (a list) receives per turn a uint8
        [game_cnt, 8] array of the moves played (255 = no move)."""
        eng = self.engine
        if Alice_snake_cnt is None:
            Alice_snake_cnt = self.snake_cnt // 2
        winners = [None] * self.game_cnt
        running = np.ones(self.game_cnt, bool)
        eng.step(tic=False, encode=True)
        turn = -1
        while running.any():
            turn += 1
            planes, rows = eng.encode_rows(refresh=False)
            rows = rows.astype(np.int64)
            g, s = rows // 8, rows % 8
            keep = running[g]
            a_idx = np.nonzero(keep & (s < Alice_snake_cnt))[0]     # pit_mp_game_runner.py:30-33
            b_idx = np.nonzero(keep & (s >= Alice_snake_cnt))[0]
            moves_a = Alice.make_moves(planes[torch.from_numpy(a_idx).to(planes.device)], [(int(g[i]), int(s[i])) for i in a_idx])
            moves_b = Bob.make_moves(planes[torch.from_numpy(b_idx).to(planes.device)], [(int(g[i]), int(s[i])) for i in b_idx])
            actions = np.ones((self.game_cnt, 8), np.uint8)
            actions[g[a_idx], s[a_idx]] = np.asarray(moves_a, np.uint8)
            actions[g[b_idx], s[b_idx]] = np.asarray(moves_b, np.uint8)
            if move_log is not None:
                played = np.full((self.game_cnt, 8), 255, np.uint8)
                played[g[a_idx], s[a_idx]] = np.asarray(moves_a, np.uint8)
                played[g[b_idx], s[b_idx]] = np.asarray(moves_b, np.uint8)
                move_log.append(played)
            if spawn_trace is not None:
                cells = torch.from_numpy(np.maximum(np.asarray(spawn_trace[turn], np.int32), -1)).to(eng.device)
                eng.step(actions=torch.from_numpy(actions).to(eng.device), spawn_cells=cells, spawn_mode=_lib.SPAWN_REPLAY,
                         tic=True, encode=True)
            else:
                eng.step(actions=torch.from_numpy(actions).to(eng.device), spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=True)
            ended = eng.ended.cpu().numpy().astype(bool)
            rw = eng.rewards.cpu().numpy()
            alive = eng.alive_mask().cpu().numpy()
            for gi in np.nonzero(running)[0]:
                if ended[gi]:                                        # pit_mp_game_runner.py:44-48
                    w = np.nonzero(rw[gi, :self.snake_cnt] == 1)[0]
                    if len(w):
                        winners[gi] = int(w[0])
                    running[gi] = False
                else:                                                # :49-60 the team with snakes left wins
                    live = np.nonzero(alive[gi])[0]
                    A = (live < Alice_snake_cnt).any()
                    B = (live >= Alice_snake_cnt).any()
                    if not A or not B:
                        winners[gi] = int(live[0])
                        running[gi] = False
        return winners

"""Game views -- mirror of the attributes of code/utils/game.py that the runners and agents touch.

The games themselves live in HBM inside an Engine (one warp steps one game); a `Game` here is a light view of game
`id` that fetches state on demand.  Batch work goes through the runner/agent classes, not through these views."""
import numpy as np


class Snake:
    """game.py:302-314 fields as seen by callers (id, health, length, head position, body)."""

    def __init__(self, ID, health, length, head, body):
        self.id, self.health, self.length, self.head_position, self.body = ID, health, length, head, body


class Game:
    """Two ways to get one:
      * `Game(engine, i)`: a view of game i of a multi-game Engine (what MPGameRunner.games holds);
      * `Game(ID, height, width, snake_cnt, health_dec=1, food_spawn_chance=0.15)`: the reference's own constructor
        (game.py:13): a standalone game backed by a private one-game Engine, for callers that step a game by hand.
    `tic` and `subgame` (game.py:87, 266) work on both; on a view `tic` steps only that game (through a one-game scratch
    engine), which is a convenience path -- lockstep work belongs to the runners."""

    def __init__(self, engine_or_id, ID_or_height=None, width=None, snake_cnt=4, health_dec=1, food_spawn_chance=0.15, seed=None):
        from ..engine import Engine
        if isinstance(engine_or_id, Engine):
            self.engine, self.id, self._index, self._own = engine_or_id, ID_or_height, ID_or_height, False
        else:
            height = ID_or_height
            if height != width:
                raise ValueError("boards must be square (rot90 of the plane, game.py:257)")
            self.id, self._index, self._own = engine_or_id, 0, True
            self.engine = Engine(side=height, snakes=snake_cnt, health_dec=health_dec, food_chance=food_spawn_chance, games=1,
                                 seed=int(engine_or_id) if seed is None else seed)
            self.engine.reset()
        self.height = self.width = self.engine.side
        self.snake_cnt = self.engine.S
        self.health_dec = self.engine.cfg.health_dec
        self.food_spawn_chance = float(self.engine.cfg.food_chance)

    # ---- game.py:87-205 -------------------------------------------------------------------------------------------
    def tic(self, moves, show=False):
        """moves[i] in {0, 1, 2} for the i-th LIVE snake (ascending id).  Returns 0, or the rewards list when the game ended."""
        import torch
        from .. import _lib
        d = self._dump()
        live = [s for s in range(self.snake_cnt) if d["snake"][s][0]]
        if len(moves) != len(live):
            raise ValueError("tic needs one move per live snake (%d), got %d" % (len(live), len(moves)))
        eng, idx = self.engine, self._index
        if not self._own and eng.G > 1:
            eng, idx = self._scratch(), 0
            eng.set_state(0, d, episode=int(d["counters"][6]))
        actions = np.ones((eng.G, 8), np.uint8)
        for s, m in zip(live, moves):
            actions[idx, s] = int(m)
        eng.step(actions=torch.from_numpy(actions).to(eng.device), spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=False)
        after = eng.get_state(idx)
        if show:                                           # game.py:140-141,194-195: two frames per tic into replay.rep
            with open("replay.rep", "a") as f:
                f.write(frames_text(replay_frames(d, list(moves), after, self.width)))
        if eng is not self.engine:
            self.engine.set_state(self._index, after, episode=int(after["counters"][6]))
            if int(after["counters"][7]):
                self._finish_in_parent(after)
        if int(after["counters"][7]):
            return [None if r == 0 else float(r) for r in after["snake"][:, 5]]
        return 0

    def _scratch(self):
        """one-game engine with this game's rules, shared by all views of the same parent engine"""
        from ..engine import Engine
        p = self.engine
        if getattr(p, "_scratch_engine", None) is None:
            p._scratch_engine = Engine(side=p.side, snakes=p.S, health_dec=p.cfg.health_dec, food_chance=float(p.cfg.food_chance),
                                       games=1, seed=int(p.cfg.seed) ^ 0x5ca7c4, device=p.device)
        return p._scratch_engine

    def _finish_in_parent(self, after):
        # asz_set_state marks a game with <= 1 live snake as finished, which is exactly an ended game; nothing else to do
        return None

    # ---- game.py:266-276 ------------------------------------------------------------------------------------------
    def subgame(self, ID):
        """a standalone copy that never spawns food (food_spawn_chance 0.0), counters restarted, rewards kept"""
        d = self._dump()
        g = Game(ID, self.height, self.width, self.snake_cnt, self.health_dec, 0.0)
        c = np.zeros(8, np.int32)
        g.engine.set_state(0, dict(snake=d["snake"], owner=d["owner"], dist=d["dist"], food=d["food"], counters=c))
        return g

    def _dump(self):
        return self.engine.get_state(self._index)

    @property
    def snakes(self):
        """live snakes in ascending id (game.py:37,191)."""
        d = self._dump()
        W = self.width
        out = []
        for s in range(self.snake_cnt):
            alive, health, length, _, head, _ = d["snake"][s]
            if alive:
                cells = np.nonzero(d["owner"] == s)[0]
                order = cells[np.argsort(-d["dist"][cells])]
                out.append(Snake(s, int(health), int(length), (int(head) // W, int(head) % W),
                                 [(int(c) // W, int(c) % W) for c in order]))
        return out

    @property
    def food(self):
        d = self._dump()
        W = self.width
        return {(int(c) // W, int(c) % W) for c in np.nonzero(d["food"])[0]}

    @property
    def rewards(self):
        """game.py:20: list of S entries None / -1.0 / 1.0"""
        d = self._dump()
        return [None if r == 0 else float(r) for r in d["snake"][:, 5]]

    @property
    def last_moves(self):
        d = self._dump()
        return {i: int(d["snake"][i][3]) for i in range(self.snake_cnt)}

    def counters(self):
        c = self._dump()["counters"]
        return dict(wall_collision=int(c[0]), body_collision=int(c[1]), head_collision=int(c[2]), starvation=int(c[3]),
                    food_eaten=int(c[4]), game_length=int(c[5]))

    def get_ids(self):
        """game.py:76-77"""
        d = self._dump()
        return [(self.id, s) for s in range(self.snake_cnt) if d["snake"][s][0]]

    def get_states(self):
        """game.py:68-69: planes of the live snakes, in live-list order (encode kernel, copied to the host)."""
        return self.engine.states_of(self._index)


# ---- replay.rep (game.py:140-141, 194-195, 281-300; read by player.py:63-79) -----------------------------------------
# The reference draws two text frames per tic when a single game is run (mp_game_runner.py:26,52): one after the snakes
# have moved, eaten and the food has spawned but before the dead are removed, one after.  The device tic is one launch,
# so the first frame is rebuilt on the host from the state before the tic, the moves and the food after it.
def _draw(width, height, food_cells, heads, bodies):
    """heads: [(length, id, cell or None)] in live-list order; bodies: [(id, cells)] in live-list order (game.py:281-300)."""
    board = [[0] * width for _ in range(height)]
    for c in food_cells:
        board[c // width][c % width] = 9
    for _, sid, cell in sorted(heads, key=lambda h: h[0]):       # stable: ties keep the live-list order
        if cell is not None:
            board[cell // width][cell % width] = -(sid + 1)
    for sid, cells in bodies:
        for c in cells:
            board[c // width][c % width] = sid + 1
    return board


def replay_frames(pre, moves, post, width, height=None):
    """(frame before removal, frame after removal) of one tic as lists of rows.
    pre / post: state dumps (Engine.get_state) before and after the tic; moves: relative move of every snake alive in
    `pre`, in ascending snake id (the live-list order of game.py:87-92)."""
    height = width if height is None else height
    snake0, owner0, dist0 = np.asarray(pre["snake"]), np.asarray(pre["owner"]), np.asarray(pre["dist"])
    food0 = np.asarray(pre["food"])
    live = [i for i in range(len(snake0)) if snake0[i][0]]
    eaten, heads, bodies = set(), [], []
    for k, i in enumerate(live):
        last, head, length = int(snake0[i][3]), int(snake0[i][4]), int(snake0[i][2])
        d = (int(moves[k]) + last - 1) % 4                                   # game.py:92
        y, x = head // width, head % width
        y += -1 if d == 0 else 1 if d == 2 else 0                            # game.py:330-342
        x += 1 if d == 1 else -1 if d == 3 else 0
        cell = y * width + x if 0 <= y < height and 0 <= x < width else None
        if cell is not None and food0[cell] and cell not in eaten:           # first come, first served (game.py:121-127)
            eaten.add(cell)
            length += 1
        heads.append((length, i, cell))
        # Snake.move pops the tail (a stamp of distance 1 leaves its cell); everything else, the old head included, is body
        bodies.append((i, [int(c) for c in np.nonzero((owner0 == i) & (dist0 >= 2))[0]]))
    food1 = [int(c) for c in np.nonzero(np.asarray(post["food"]))[0]]        # eaten and spawned before the first frame
    first = _draw(width, height, food1, heads, bodies)
    snake1, owner1 = np.asarray(post["snake"]), np.asarray(post["owner"])
    heads, bodies = [], []
    for i in range(len(snake1)):
        if snake1[i][0]:
            h = int(snake1[i][4])
            heads.append((int(snake1[i][2]), i, h))
            bodies.append((i, [int(c) for c in np.nonzero(owner1 == i)[0] if int(c) != h]))
    second = _draw(width, height, food1, heads, bodies)
    return first, second


def frames_text(frames):
    """the bytes Game.draw appends to replay.rep for these frames (game.py:296-300)"""
    return "".join("".join(str(row) + "\n" for row in board) + "\n" for board in frames)

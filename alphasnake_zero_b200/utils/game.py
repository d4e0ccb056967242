"""Game views -- mirror of the attributes of code/utils/game.py that the runners and agents touch.

The games themselves live in HBM inside an Engine (one warp steps one game); a `Game` here is a light view of game
`id` that fetches state on demand.  Batch work goes through the runner/agent classes, not through these views."""
import numpy as np


class Snake:
    """game.py:302-314 fields as seen by callers (id, health, length, head position, body)."""

    def __init__(self, ID, health, length, head, body):
        self.id, self.health, self.length, self.head_position, self.body = ID, health, length, head, body


class Game:

    def __init__(self, engine, ID):
        self.engine = engine
        self.id = ID
        self.height = self.width = engine.side
        self.snake_cnt = engine.S

    def _dump(self):
        return self.engine.get_state(self.id)

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
        return self.engine.states_of(self.id)

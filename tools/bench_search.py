"""Search kernels alone (stub value function, no network): throughput and the algorithmic bytes of the table operations
(64 B probe + 24 B per ancestor backed up + 8 B path push per node visit).
  python tools/bench_search.py [games] [breadth] [turns]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphasnake_zero_b200 import _lib  # noqa: E402
from alphasnake_zero_b200.engine import Engine  # noqa: E402


def main():
    games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    breadth = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    turns = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    eng = Engine(side=11, snakes=4, health_dec=1, games=games, seed=3, max_depth=8, max_breadth=breadth, softmax_base=2.0, training=True)
    eng.reset()
    for _ in range(32):
        eng.step(spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=False, auto_reset=True, random_actions=True)

    def turn():
        q, mv = eng.search(value_fn=None)
        act = torch.where(mv < 3, mv, torch.ones_like(mv))
        eng.step(actions=act, spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=False, auto_reset=True)
    turn()
    torch.cuda.synchronize()
    s0 = eng.search_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(turns):
        turn()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    s1 = eng.search_stats()
    d = {k: s1[k] - s0[k] for k in ("evals", "node_visits", "subgames", "subgame_tics")}
    # every node visit backs its estimate up all earlier (slot, move) pairs of the same snake in the same sub-game: the mean
    # path length is (visits per (sub-game, snake) - 1) / 2; bounded above by subgame_tics / subgames
    mean_depth = d["subgame_tics"] / max(d["subgames"], 1)
    alg = d["node_visits"] * (64 + 8) + d["node_visits"] * 24 * max(mean_depth - 1, 0) / 2 + d["evals"] * 5292
    print(json.dumps({"games": games, "breadth": breadth, "turns": turns, "seconds": dt, "sims_per_sec": d["subgames"] / dt,
                      "node_visits_per_sec": d["node_visits"] / dt, "evals_per_sec": d["evals"] / dt,
                      "subgame_tics_per_sim": mean_depth, "algorithmic_bytes": alg, "algorithmic_GBps": alg / dt / 1e9,
                      "note": "algorithmic bytes = 72 B per node visit (probe + path push) + 24 B per backed-up ancestor + 5,292 B per "
                              "miss plane written into the evaluation batch"}))


if __name__ == "__main__":
    main()

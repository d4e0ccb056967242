"""A CI-sized pass over every kernel of the engine, meant to be run under compute-sanitizer (one tool per gpurun call):
  compute-sanitizer --tool memcheck python tools/sanitize_run.py
  compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alphasnake_zero_b200 import _lib  # noqa: E402
from alphasnake_zero_b200.engine import Engine  # noqa: E402
from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet  # noqa: E402
from alphasnake_zero_b200.net import NativeNet  # noqa: E402


def main():
    for side, S in ((11, 4), (7, 4), (19, 8)):
        eng = Engine(side=side, snakes=S, health_dec=1, games=96, seed=1, max_depth=4, max_breadth=8, softmax_base=2.0, training=True,
                     table_log2=14)
        eng.reset()
        for _ in range(6):
            eng.step(spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=True, auto_reset=True, random_actions=True, keys=True)
        net = AlphaNNet(input_shape=(2 * side - 1, 2 * side - 1, 3), seed=0, backend="native")
        net._native = NativeNet(net.weights, "cuda", chunk_images=64)
        for variant in (3, 2, 1):
            _lib.check(_lib.lib().asz_net_set_variant(net._native.h, variant))
            q, mv = eng.search(net=net._native)
            act = torch.where(mv < 3, mv, torch.ones_like(mv))
            eng.step(actions=act, spawn_mode=_lib.SPAWN_NATIVE, tic=True, encode=False, auto_reset=True)
        q, mv = eng.search(value_fn=None)
        torch.cuda.synchronize()
        st = eng.search_stats()
        assert st["evals"] > 0 and np.isfinite(q.cpu().numpy()).all()
        eng.close()
        print("side %d ok: %d evals" % (side, st["evals"]))


if __name__ == "__main__":
    main()

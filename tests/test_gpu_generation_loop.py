"""One generation of the reference's training loop (alpha_snake_zero_trainer.py:52-91) and one pit (pit.py:28-35) on the
drop-in classes, with tiny settings: self-play with the search agent -> log counters -> sampled records, mirrored
(:93-100) -> nnet.train -> save / reload -> the new generation plays the old one."""
import os
from random import sample, seed as rseed

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_generation_and_a_pit(tmp_path, monkeypatch):
    import torch
    from alphasnake_zero_b200.utils.agent import Agent
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    from alphasnake_zero_b200.utils import pit_agent, pit_mp_game_runner
    monkeypatch.chdir(tmp_path)
    rseed(0); torch.manual_seed(0)
    nnet = AlphaNNet(input_shape=(21, 21, 3), seed=3)                     # train.py:32
    nnet = nnet.copy_and_compile()                                        # alpha_snake_zero_trainer.py:34
    Alice = Agent(nnet, 2, True, 4, 16)                                   # :52 (depth 4, breadth 16 instead of 8, 128)
    gr = MPGameRunner(11, 11, 4, 9, 48, verbose=False)                    # :53
    rewards = gr.run(Alice)                                               # :54, to completion
    assert len(rewards) == 48 and all(r is not None and len(r) == 4 for r in rewards)
    for r in rewards:                                                     # at most one winner, everybody else lost
        assert sorted(x for x in r if x is not None)[:-1].count(1.0) == 0 and r.count(1.0) <= 1 and set(r) <= {1.0, -1.0}
    log_list = [gr.wall_collision, gr.body_collision, gr.head_collision, gr.starvation, gr.food_eaten, gr.game_length]
    assert all(np.isfinite(x) and x >= 0 for x in log_list) and gr.game_length > 1
    deaths = gr.wall_collision + gr.body_collision + gr.head_collision + gr.starvation
    assert 3.0 <= deaths <= 4.0                                           # 3 or 4 snakes die per game
    n = len(Alice.records)
    assert n == len(Alice.values) and n > 48 * 4
    batch_size = min(256, n)
    idx = sample(range(n), batch_size)                                    # :71
    X = [Alice.records[i] for i in idx]
    V = [Alice.values[i] for i in idx]
    assert X[0].shape == (21, 21, 3) and X[0].dtype == np.float32 and V[0].shape == (3,)
    Alice.clear()                                                         # :75
    assert len(Alice.records) == 0
    X += list(np.flip(X, axis=2))                                         # mirror_states, :93-97
    V += list(np.flip(V, axis=1))                                         # mirror_values, :99-100
    before = nnet.v(X[:32])
    new = nnet.copy_and_compile(learning_rate=1e-3)                       # :79
    new.train(X, V, epochs=2, batch_size=batch_size)                      # :81
    new = new.copy_and_compile()                                          # :83
    after = new.v(X[:32])
    assert np.isfinite(after).all() and np.abs(after - before).max() > 1e-4        # the weights moved
    new.save("Test1")                                                     # :91 -> models/Test1 (npz instead of h5)
    loaded = AlphaNNet(model_name="models/Test1")
    np.testing.assert_allclose(loaded.v(X[:32]), after, rtol=0, atol=1e-6)
    # pit.py:28-35: the new generation (snake ids 0..1) against the old one, 2 vs 2
    A, B = pit_agent.Agent(loaded), pit_agent.Agent(nnet)
    winners = pit_mp_game_runner.MPGameRunner(11, 11, 4, 9, 24).run(A, B, 2)
    assert len(winners) == 24 and all(w is None or 0 <= w < 4 for w in winners)
    assert sum(w is not None for w in winners) > 0
